"""Next-token selection for LibraForCausalLM.generate (N1): the logits processors / warpers and the sampling step of the
reference's `greedy_search` and `sample` loops (libra/models/libra/modeling_libra_utils.py:61-328, :330-635).

The reference keeps Q = vision_codebook_num planes of ids `[Q, B, T]` and logits `[Q, B, V]`; it calls
`logits_processor(input_ids, next_token_logits)` and `logits_warper(...)` on those 3-D tensors (:241, :538-539), takes
`argmax(-1)` (:263) or, when sampling, `softmax` + one `torch.multinomial` PER PLANE in plane order (:559-564), and
replaces the tokens of finished samples by `pad_token_id` (:270-274, :571-575).  HF's stock processors are written for 2-D
`[batch, vocab]` scores (RepetitionPenalty gathers along dim 1, TopP scatters along dim 1), so here every processor is
applied plane by plane on `(input_ids[q], scores[q])` -- the natural reading, and identical for the shape-agnostic ones
(temperature, top-k).  Everything stays on the device: no `.item()`, no host branch on tensor values, so the selection can
be captured in the CUDA graph of the one-token step.

Built-in policy (the usual `generate` keywords): repetition_penalty -> temperature -> top_k -> top_p, each following the
transformers implementation (checked against transformers' own classes in tests/test_generation_cpu.py); arbitrary
`logits_processor` / `logits_warper` callables `(input_ids [B,T], scores [B,V]) -> scores [B,V]` (e.g. a
`transformers.LogitsProcessorList`) run before / after the built-in ones as in HF's `_get_logits_processor` ordering.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence

import torch

FILTER = -float("inf")


@dataclass
class SelectionPolicy:
    do_sample: bool = False
    temperature: float = 1.0
    top_k: int = 0                       # 0 = off (transformers' default 50 applies only when the caller passes it)
    top_p: float = 1.0
    repetition_penalty: float = 1.0
    min_tokens_to_keep: int = 1
    logits_processor: Sequence[Callable] = field(default_factory=list)
    logits_warper: Sequence[Callable] = field(default_factory=list)
    generator: Optional[torch.Generator] = None

    def __post_init__(self):
        if self.temperature is None or float(self.temperature) <= 0.0:
            raise ValueError(f"temperature must be a strictly positive float, got {self.temperature}")
        if self.top_k is None:
            self.top_k = 0
        if int(self.top_k) < 0:
            raise ValueError(f"top_k must be >= 0, got {self.top_k}")
        if self.top_p is None:
            self.top_p = 1.0
        if not (0.0 <= float(self.top_p) <= 1.0):
            raise ValueError(f"top_p must be in [0, 1], got {self.top_p}")
        if self.repetition_penalty is None or float(self.repetition_penalty) <= 0.0:
            raise ValueError(f"repetition_penalty must be a strictly positive float, got {self.repetition_penalty}")
        self.logits_processor = _as_list(self.logits_processor)
        self.logits_warper = _as_list(self.logits_warper)

    @property
    def warps(self) -> bool:
        return self.do_sample and (self.temperature != 1.0 or self.top_k > 0 or self.top_p < 1.0 or len(self.logits_warper) > 0)


def _as_list(x) -> List[Callable]:
    if x is None:
        return []
    if isinstance(x, (list, tuple)):            # transformers.LogitsProcessorList is a list subclass
        return list(x)
    if callable(x):
        return [x]
    return list(x)


# ---------------------------------------------------------------------------------------------- built-in processors (2-D)
def repetition_penalty_(input_ids: torch.Tensor, scores: torch.Tensor, penalty: float) -> torch.Tensor:
    """transformers.RepetitionPenaltyLogitsProcessor: scores of ids already in the sequence are divided (positive) or
    multiplied (negative) by the penalty."""
    seen = torch.gather(scores, 1, input_ids)
    seen = torch.where(seen < 0, seen * penalty, seen / penalty)
    return scores.scatter(1, input_ids, seen)


def temperature_(scores: torch.Tensor, temperature: float) -> torch.Tensor:
    return scores / temperature


def top_k_(scores: torch.Tensor, top_k: int, min_tokens_to_keep: int = 1) -> torch.Tensor:
    """transformers.TopKLogitsWarper: everything below the k-th largest score is filtered."""
    k = min(max(int(top_k), int(min_tokens_to_keep)), scores.size(-1))
    kth = torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(scores < kth, FILTER)


def top_p_(scores: torch.Tensor, top_p: float, min_tokens_to_keep: int = 1) -> torch.Tensor:
    """transformers.TopPLogitsWarper: ascending sort, drop the prefix whose cumulative probability is <= 1 - top_p, always
    keep the last min_tokens_to_keep entries."""
    sorted_logits, sorted_indices = torch.sort(scores, descending=False)
    cumulative = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove = cumulative <= (1 - top_p)
    remove[..., -min_tokens_to_keep:] = False
    remove = remove.scatter(1, sorted_indices, remove)
    return scores.masked_fill(remove, FILTER)


# ---------------------------------------------------------------------------------------------- the selection step
def process(policy: SelectionPolicy, input_ids: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
    """input_ids [Q,B,T], logits [Q,B,V] (any float dtype) -> fp32 scores [Q,B,V] after processors (always) and warpers
    (only when sampling, as in the reference: greedy_search applies `logits_processor` only)."""
    planes = []
    for q in range(logits.shape[0]):
        ids, s = input_ids[q], logits[q].float()
        for proc in policy.logits_processor:
            s = proc(ids, s)
        if policy.repetition_penalty != 1.0:
            s = repetition_penalty_(ids, s, float(policy.repetition_penalty))
        if policy.do_sample:
            if policy.temperature != 1.0:
                s = temperature_(s, float(policy.temperature))
            if policy.top_k > 0:
                s = top_k_(s, policy.top_k, policy.min_tokens_to_keep)
            if policy.top_p < 1.0:
                s = top_p_(s, float(policy.top_p), policy.min_tokens_to_keep)
            for warp in policy.logits_warper:
                s = warp(ids, s)
        planes.append(s)
    return torch.stack(planes)


def select(policy: SelectionPolicy, scores: torch.Tensor) -> torch.Tensor:
    """scores [Q,B,V] -> next tokens [Q,B]: argmax, or one multinomial draw per plane in plane order (:559-564)."""
    if not policy.do_sample:
        return scores.argmax(dim=-1)
    probs = torch.softmax(scores, dim=-1)
    return torch.stack([torch.multinomial(p, num_samples=1, generator=policy.generator).squeeze(1) for p in probs])


def next_tokens(policy: SelectionPolicy, input_ids: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
    return select(policy, process(policy, input_ids, logits))


def finish_(nxt: torch.Tensor, done: torch.Tensor, eos_ids: Optional[torch.Tensor], pad_token_id: Optional[int]):
    """Finished-sample bookkeeping of the reference loops (:266-287, :567-588), plane by plane IN ORDER: a finished sample's
    token becomes pad_token_id, and a sample finishes as soon as one plane emits an EOS id -- so a later plane of the same
    step already sees it as finished.  nxt [Q,B] and done [B] (bool) are updated in place; no host synchronisation."""
    if eos_ids is None:
        return nxt, done
    for q in range(nxt.shape[0]):
        row = torch.where(done, torch.full_like(nxt[q], pad_token_id), nxt[q])
        nxt[q] = row
        done.logical_or_((row[:, None] == eos_ids[None, :]).any(dim=1))
    return nxt, done


@dataclass
class GenerateOutput:
    """What `generate(return_dict_in_generate=True)` returns (GenerateDecoderOnlyOutput's fields that exist on this path)."""
    sequences: torch.Tensor                          # [Q, B, T + n]
    scores: Optional[tuple] = None                   # processed scores [Q, B, V] per generated step (output_scores=True)
    past_key_values: Optional[object] = None
