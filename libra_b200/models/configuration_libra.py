"""LibraConfig -- same fields and defaults as the reference (libra/models/libra/configuration_libra.py:9-58,
libra/models/llama/configuration_llama.py:84-118), so reference checkpoints' config.json load unchanged."""
from __future__ import annotations

from transformers.configuration_utils import PretrainedConfig


class LibraConfig(PretrainedConfig):
    model_type = "libra"
    keys_to_ignore_at_inference = ["past_key_values"]

    def __init__(
        self,
        # language part (LlamaConfig defaults)
        vocab_size=32000,
        hidden_size=4096,
        intermediate_size=11008,
        num_hidden_layers=32,
        num_attention_heads=32,
        hidden_act="silu",
        max_position_embeddings=2048,
        initializer_range=0.02,
        rms_norm_eps=1e-6,
        use_cache=True,
        pad_token_id=0,
        bos_token_id=1,
        eos_token_id=2,
        tie_word_embeddings=False,
        # vision part
        vision_down_ratio=4,
        vision_vocab_size=514,
        vision_codebook_num=2,
        max_vision_token_length=578,
        newline_token_id=13,
        vision_embd_pdrop=0.0,
        vision_resid_pdrop=0.0,
        contiguous_signal_size=2048,
        image_feature_resolution=24,
        vision_prediction_mode="1d",
        use_bridge=True,
        bridge_rank=8,
        concat_signals=True,
        norm_signals=True,
        addition_mode=False,
        use_vision_position_embedding=False,
        unified_head=False,
        use_2d_rope=False,
        resid_pdrop=0.0,
        attn_pdrop=0.0,
        embd_pdrop=0.0,
        **kwargs,
    ):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.intermediate_size = intermediate_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.hidden_act = hidden_act
        self.max_position_embeddings = max_position_embeddings
        self.initializer_range = initializer_range
        self.rms_norm_eps = rms_norm_eps
        self.use_cache = use_cache
        self.vision_down_ratio = vision_down_ratio
        self.vision_vocab_size = vision_vocab_size
        self.vision_codebook_num = vision_codebook_num
        self.max_vision_token_length = max_vision_token_length
        self.newline_token_id = newline_token_id
        self.vision_embd_pdrop = vision_embd_pdrop
        self.vision_resid_pdrop = vision_resid_pdrop
        self.contiguous_signal_size = contiguous_signal_size
        self.image_feature_resolution = image_feature_resolution
        self.vision_prediction_mode = vision_prediction_mode
        self.use_bridge = use_bridge
        self.bridge_rank = bridge_rank
        self.concat_signals = concat_signals
        self.norm_signals = norm_signals
        self.addition_mode = addition_mode
        self.use_vision_position_embedding = use_vision_position_embedding
        self.unified_head = unified_head
        self.use_2d_rope = use_2d_rope
        self.resid_pdrop = resid_pdrop
        self.attn_pdrop = attn_pdrop
        self.embd_pdrop = embd_pdrop
        super().__init__(pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id,
                         tie_word_embeddings=tie_word_embeddings, **kwargs)

    def unsupported_branches(self):
        """Config branches the reference implements but ships disabled ('Libra decoder' row);
        the CUDA path refuses them loudly instead of silently computing something else."""
        bad = []
        if self.addition_mode: bad.append("addition_mode")
        if self.use_2d_rope: bad.append("use_2d_rope")
        if self.unified_head: bad.append("unified_head")
        if self.use_vision_position_embedding: bad.append("use_vision_position_embedding")
        if self.vision_prediction_mode != "1d": bad.append(f"vision_prediction_mode={self.vision_prediction_mode}")
        if not self.use_bridge: bad.append("use_bridge=False")
        if not (self.concat_signals and self.norm_signals): bad.append("concat_signals/norm_signals != True")
        for k in ("resid_pdrop", "attn_pdrop", "embd_pdrop", "vision_embd_pdrop", "vision_resid_pdrop"):
            if getattr(self, k) != 0.0: bad.append(f"{k}>0")
        if self.vision_codebook_num != 2: bad.append("vision_codebook_num != 2")
        return bad
