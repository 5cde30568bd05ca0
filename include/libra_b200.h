/* libra_b200 -- C ABI of the sm_100a kernel library (liblibra_b200.so).
 *
 * The reference (YifanXu74/Libra) is pure Python/PyTorch: it has no native
 * interface to mirror.  Each entry point below replaces the PyTorch-eager code
 * of one reference function on the training hot path (file:line cited per
 * function, paths relative to the reference root); libra_b200/ops.py binds
 * them with ctypes and INTEGRATION.md shows the binding a reference maintainer
 * would add.
 *
 * Conventions
 *  - plain C types and raw device pointers only; the caller (PyTorch) owns all
 *    memory including workspaces;
 *  - every function only enqueues work on `stream` (a cudaStream_t passed as
 *    void*): it never synchronises, allocates device memory or throws;
 *  - returns LB_OK (0) or a negative LB_E* code; the message for the calling
 *    thread is available through lb_last_error();
 *  - tensors are row-major and dense unless a leading dimension is given;
 *    activations and weights are bf16 (uint16 storage), statistics fp32;
 *  - "sorted rows": the decoder keeps its [N, C] activations permuted so that
 *    language tokens come first and vision tokens second (the reference's
 *    modality routing, modeling_libra.py:111-147, then needs no gather/scatter);
 *    `flag[r]` (uint8) is 1 for a vision row;
 *  - device: the current CUDA device of the calling thread, which must be
 *    compute capability 10.x (no fallback path).
 */
#ifndef LIBRA_B200_H
#define LIBRA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB_OK 0
#define LB_EINVAL (-1)   /* bad shape / argument */
#define LB_EALIGN (-2)   /* pointer or leading dimension not aligned as required */
#define LB_EDTYPE (-3)   /* unsupported dtype code */
#define LB_ELAUNCH (-4)  /* CUDA launch / runtime failure (text in lb_last_error) */
#define LB_EARCH (-5)    /* device is not sm_100 */
#define LB_EDRIVER (-6)  /* driver entry point (tensor-map encode) unavailable */

#define LB_DT_BF16 0
#define LB_DT_F32 1

int lb_version(void);
/* copies the calling thread's last error text (NUL terminated) into buf */
int lb_last_error(char* buf, int n);
/* LB_OK when the current device can run this library */
int lb_device_check(void);
int lb_sm_count(void);
/* Programmatic dependent launch for the launch-bound one-token decode chain (N1: rmsnorm -> skinny GEMMs -> attention
 * operand prologue -> decode attention -> combine -> ...; reference path modeling_libra.py:343-361, 437-491 at q_len 1).
 * on != 0: those kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization, start their prologue (barrier
 * init, TMEM allocation, WEIGHT streaming into shared memory) while the previous kernel of the stream drains and wait
 * (griddepcontrol.wait) before touching activations; captured into CUDA graphs as programmatic edges.  Results are
 * bit-identical to serial launches.  Process-wide switch; returns the previous value. */
int lb_set_pdl(int on);

/* ---- A13 LlamaRMSNorm routed by modality --------------------------------
 * libra/models/llama/modeling_llama.py:127-132, routed at
 * libra/models/libra/modeling_libra.py:463,479,817 (and :559,642 un-routed).
 * y = w[flag] * x * rsqrt(mean(x^2)+eps), fp32 internally, one cast.
 * flag may be NULL (every row uses w_lang).  rstd [rows] fp32 is saved for bwd. */
int lb_rmsnorm_fwd(const void* x, const void* w_lang, const void* w_vis, const uint8_t* flag, void* y, float* rstd,
                   int64_t rows, int cols, float eps, void* stream);
/* dx; dw_lang/dw_vis [cols] fp32 are ACCUMULATED into (caller zeroes).  partial is a
 * workspace of lb_rmsnorm_bwd_workspace(rows, cols) bytes.  residual_grad
 * (optional, may be NULL) is added to dx (fuses the residual branch). */
int64_t lb_rmsnorm_bwd_workspace(int64_t rows, int cols);
int lb_rmsnorm_bwd(const void* dy, const void* x, const void* w_lang, const void* w_vis, const uint8_t* flag,
                   const float* rstd, const void* residual_grad, void* dx, float* dw_lang, float* dw_vis, void* partial,
                   int64_t rows, int cols, void* stream);

/* ---- A1/A4 nn.LayerNorm (CLIP) ------------------------------------------
 * libra/models/clip/modeling_clip.py:386-388,866 (eps from config, affine). */
int lb_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows,
                     int cols, float eps, void* stream);
int64_t lb_layernorm_bwd_workspace(int64_t rows, int cols);
int lb_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                     float* dw, float* db, void* partial, int64_t rows, int cols, void* stream);

/* ---- A15 SwiGLU product  silu(gate) * up ---------------------------------
 * libra/models/libra/modeling_libra.py:232-233.  ld_* are row pitches in elements. */
int lb_swiglu_fwd(const void* gate, const void* up, void* out, int64_t rows, int cols, int64_t ld_gate, int64_t ld_up,
                  int64_t ld_out, void* stream);
int lb_swiglu_bwd(const void* dout, const void* gate, const void* up, void* dgate, void* dup, int64_t rows, int cols,
                  int64_t ld_dout, int64_t ld_gate, int64_t ld_up, int64_t ld_dgate, int64_t ld_dup, void* stream);

/* ---- A3 quick_gelu with bias:  y = g(x+b), g(z) = z*sigmoid(1.702 z) ------
 * libra/models/clip/modeling_clip.py:374-378 (fc1 bias + activation). bias may be NULL. */
int lb_bias_quick_gelu_fwd(const void* x, const void* bias, void* y, int64_t rows, int cols, void* stream);
/* dx = dy * g'(x+b) */
int lb_bias_quick_gelu_bwd(const void* dy, const void* x, const void* bias, void* dx, int64_t rows, int cols,
                           void* stream);

/* ---- row permutation (sorted <-> original token order) --------------------
 * dst[r] = src[index[r]] for r in [0,rows); rows of `cols` bf16. */
int lb_gather_rows(const void* src, const int32_t* index, void* dst, int64_t rows, int cols, void* stream);

/* ---- A17 embeddings ------------------------------------------------------
 * libra/models/libra/modeling_libra.py:625-661.  Writes, for sorted row r:
 *  language: out[r, :] = embed[ids0[r]]                         (cols = hidden)
 *  vision  : cat[i, :] = [vemb0[ids0[i]] | vemb1[ids1[i]] | signal[signal_row[i]]]  (2*half + signal cols)
 * ids are int64 per (sorted) row, vision ids already offset-subtracted by the caller; signal is the
 * [B*T, signal_cols] tensor in original token order (NULL = zeros, modeling_libra.py:647-653) and
 * signal_row[i] the original token index of vision row i (NULL = identity). */
int lb_embed_lang_fwd(const int64_t* ids, const void* table, void* out, int64_t rows, int cols, void* stream);
int lb_embed_vision_cat_fwd(const int64_t* ids0, const int64_t* ids1, const void* table0, const void* table1,
                            const void* signal, const int32_t* signal_row, void* out, int64_t rows, int half,
                            int signal_cols, void* stream);
/* scatter-add of row gradients into an fp32 table gradient: dtable[ids[r], :] += dy[r, col0:col0+cols] */
int lb_embed_bwd(const int64_t* ids, const void* dy, int64_t ld_dy, int col0, float* dtable, int64_t rows, int cols,
                 void* stream);

/* ---- A7 LFQ sign-quantise + index pack (bit exact) ------------------------
 * libra/models/libra/taming/modules/quantization/lookup_free_quantization.py:185-208,
 * ImageTokenizer.encode offsets/BOI/EOI libra/models/libra/image_tokenizer.py:75-95.
 * h: [n_img*tokens, num_codebooks*bits] (bf16 or fp32 by dtype); writes
 * ids[q, img, 1+t] = offset + sum_d [h>0] 2^(bits-1-d), ids[q,img,0]=boi, ids[q,img,tokens+1]=eoi
 * into an int64 [num_codebooks, n_img, tokens+2] tensor. */
int lb_lfq_pack(const void* h, int dtype, int64_t n_img, int tokens, int num_codebooks, int bits, int64_t offset,
                int64_t boi, int64_t eoi, int64_t* ids, void* stream);
/* inverse: codes[..., q*bits+d] = +-1 (bf16/fp32) from ids (without BOI/EOI, offset removed): indices_to_codes :129-158 */
int lb_lfq_unpack(const int64_t* idx, int64_t n, int num_codebooks, int bits, void* codes, int dtype, void* stream);

/* ---- A12 + A10(bridge operands) attention prologue ------------------------
 * libra/models/libra/modeling_libra.py:320-340 (RoPE on q and both key variants, variant selection per key modality),
 * :282-286 (value variants).  Inputs in sorted rows [N, H*D]; kc = k + kb and vc = v + vb are the bridged ("cross")
 * tensors (the rank-r bridge products are plain GEMMs with beta = 1 done by the caller; NULL = no bridge).
 * Outputs in ORIGINAL token order [B*T, H*D]:
 *   Q    = rope(q)
 *   Kfv  = rope(lang j ? kc : k)    Vfv = lang j ? vc : v      (what VISION queries see)
 *   Kfl  = rope(vis  j ? kc : k)    Vfl = vis  j ? vc : v      (what LANGUAGE queries see)
 * sorted_of[bt] = sorted row of original token bt; pos[bt] = rotary position; cos/sin: fp32 tables [n_pos, D/2];
 * flag_sorted[r] = 1 for vision rows.  kv_row (int32 [n_tokens], NULL = identity): row of Kfv/Kfl/Vfv/Vfl that token bt's
 * key/value operands go to -- the decode step writes them straight into the token's slot of the KV cache
 * ([B*capacity, H*D] views, kv_row[b] = b*capacity + length). */
int lb_attn_prep_fwd(const void* q, const void* k, const void* kc, const void* v, const void* vc,
                     const uint8_t* flag_sorted, const int32_t* sorted_of, const int32_t* pos, const float* cos_t,
                     const float* sin_t, void* Q, void* Kfv, void* Kfl, void* Vfv, void* Vfl, int64_t n_tokens, int heads,
                     int head_dim, const int32_t* kv_row, void* stream);
/* The same prologue with the rank-r bridge products folded in (modeling_libra.py:310-319: kc = k + B_k(A_k x), vc likewise):
 * tk / tv [N, rank] = x A_k^T / x A_v^T in sorted rows, B* [H*D, rank] the bridges' second factors for language / vision
 * tokens, rank a multiple of 8.  kc = bf16(k + bf16(tk . B_k^T)) -- the rounding sequence of the GEMM epilogue.  Meant for
 * the one-token decode step (N1), where a separate rank-8 GEMM launch costs more than its arithmetic; training keeps the
 * GEMM (its backward needs kc / vc's producers anyway). */
int lb_attn_prep_fwd_bridge(const void* q, const void* k, const void* v, const void* tk, const void* tv, const void* Bk_lang,
                            const void* Bk_vis, const void* Bv_lang, const void* Bv_vis, int rank, const uint8_t* flag_sorted,
                            const int32_t* sorted_of, const int32_t* pos, const float* cos_t, const float* sin_t, void* Q, void* Kfv,
                            void* Kfl, void* Vfv, void* Vfl, int64_t n_tokens, int heads, int head_dim, const int32_t* kv_row,
                            void* stream);
/* adjoint: from dQ,dKfv,dKfl,dVfv,dVfl (original order) to dq,dk,dv,dkb,dvb (sorted rows, [N,H*D]) */
int lb_attn_prep_bwd(const void* dQ, const void* dKfv, const void* dKfl, const void* dVfv, const void* dVfl,
                     const uint8_t* flag_sorted, const int32_t* sorted_of, const int32_t* pos, const float* cos_t,
                     const float* sin_t, void* dq, void* dk, void* dv, void* dkb, void* dvb, int64_t n_tokens, int heads,
                     int head_dim, void* stream);

/* ---- A10 bridge attention core (tcgen05 flash attention) ------------------
 * libra/models/libra/modeling_libra.py:363-397 + attn_with_bridge :267-296; replaces the disabled
 * utils/llama_flash_attn_monkey_patch.py:75-178.  Also A2 (CLIPAttention core, modeling_clip.py:309-349)
 * with causal=0, one variant, head_dim 64.
 *
 * Q,K*,V*: [B*T, H*D] bf16, original token order.  work: int32 [n_work,4] = {batch, q_tile, variant, 0}
 * (variant 0: rows with qflag==0 attend K0/V0; variant 1: rows with qflag==1 attend K1/V1; a 128-row
 * q tile that holds both modalities appears twice).  qflag[B*T] (may be NULL when only variant 0 exists).
 * kv_start/kv_end [B]: valid key range per sample (padding), NULL = [0,T).
 * out_row[B*T] (may be NULL): destination row of each token in O (sorted-row scatter).
 * lse [B,H,T] fp32 (natural log, of the scaled scores).  scale multiplies Q.K^T. */
int lb_attn_fwd(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const uint8_t* qflag,
                const int32_t* work, int n_work, const int32_t* kv_start, const int32_t* kv_end,
                const int32_t* out_row, void* O, float* lse, int batch, int seqlen, int heads, int head_dim, int causal,
                float scale, void* stream);

/* Same operation and work list as lb_attn_fwd, persistent streaming kernel (csrc/attn_fwd_stream.cu): one CTA per SM
 * walks its share of the (work item, head) list as one stream of score tiles; S and P are double-buffered in TMEM (QK^T
 * runs two tiles ahead of PV), two softmax warpgroups take alternate tiles, a separate warpgroup writes O out.
 * List position L of (work item w, head h): heads run in groups of head_group, inside group g (gl = heads in the group)
 * L = g*head_group*n_work + w*gl + (h - g*head_group).
 * plan_items [n_work*heads] / plan_off [n_cta+1] (int32, device): the list positions each of the n_cta CTAs handles, in
 * order (host-side balanced split, libra_b200/schedule.py: stream_plan); max_cta_items = the longest per-CTA list of the
 * plan.  Both NULL: a static snake split over one CTA per SM (n_cta, max_cta_items ignored).  A CTA holds at most
 * lb_attn_fwd_stream_max_cta_items() decoded items (LB_EINVAL beyond; use lb_attn_fwd then).
 * head_group <= 0: library default (LB_ATTN_HEAD_GROUP, 8). */
int lb_attn_fwd_stream(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const uint8_t* qflag,
                       const int32_t* work, int n_work, const int32_t* plan_items, const int32_t* plan_off, int n_cta,
                       int max_cta_items, int head_group, const int32_t* kv_start, const int32_t* kv_end,
                       const int32_t* out_row, void* O, float* lse, int batch, int seqlen, int heads, int head_dim,
                       int causal, float scale, void* stream);
int lb_attn_fwd_stream_max_cta_items(void);
/* diagnostics: every CTA logs {smid, items, tiles, clock64 at entry, first Q landed, exit, -, -} into buf
 * ([number of SMs][8] int64, device memory).  NULL = off */
int lb_attn_fwd_stream_set_cta_log(void* buf);
/* diagnostics: CTA 0 writes clock64 stamps into buf ([64][32] int64, device; one row per kv tile: slots 0-9 tcgen05 thread, 10-17
 * softmax thread 0 of the tile's warpgroup).  NULL = off */
int lb_attn_fwd_stream_set_trace(void* buf);

/* diagnostics: CTA (0,0) of subsequent lb_attn_fwd launches writes clock64 stamps into buf ([64][8] int64, device
 * memory; slots: MMA K-ready / QK-issued / P-seen / PV-issued, softmax S-seen / max-done / exchanged / P-arrived). NULL = off */
int lb_attn_fwd_set_trace(void* buf);

/* delta[b,h,t] = sum_d dO*O per head, plus dO gathered to original order.
 * O_rows/dO_rows are indexed by row_of[bt] (NULL = identity). */
int lb_attn_bwd_prepare(const void* O, const void* dO, const int32_t* row_of, void* dO_orig, float* delta, int batch,
                        int seqlen, int heads, int head_dim, void* stream);

/* dQ (original order).  Same work list semantics as forward. */
int lb_attn_bwd_dq(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                   const float* lse, const float* delta, const uint8_t* qflag, const int32_t* work, int n_work,
                   const int32_t* kv_start, const int32_t* kv_end, void* dQ, int batch, int seqlen, int heads,
                   int head_dim, int causal, float scale, void* stream);
/* Same operation and work list as lb_attn_bwd_dq, persistent streaming kernel (csrc/attn_bwd_dq_stream.cu): one CTA per SM
 * walks its share of the (work item, head) list as one stream of 128x64 tiles; S and dP double-buffered in TMEM and issued one
 * tile ahead of the dS computation, K ring of 4 / V ring of 3, two compute warpgroups on alternate tiles.  plan_* / n_cta /
 * max_cta_items / head_group: as for lb_attn_fwd_stream (the forward's plan is the balanced split for this kernel too). */
int lb_attn_bwd_dq_stream(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                          const float* lse, const float* delta, const uint8_t* qflag, const int32_t* work, int n_work,
                          const int32_t* plan_items, const int32_t* plan_off, int n_cta, int max_cta_items, int head_group,
                          const int32_t* kv_start, const int32_t* kv_end, void* dQ, int batch, int seqlen, int heads,
                          int head_dim, int causal, float scale, void* stream);
int lb_attn_bwd_dq_stream_max_cta_items(void);
/* diagnostics: CTA 0 writes clock64 stamps into buf ([64][32] int64, device; see csrc/attn_bwd_dq_stream.cu).  NULL = off */
int lb_attn_bwd_dq_stream_set_trace(void* buf);
/* dK0,dV0 (gradient w.r.t. the variant-0 operands, from qflag==0 query rows) and dK1,dV1.
 * work_kv: int32 [n_work,4] = {batch, kv_tile, variant, first_q_tile}; qtile_has: uint8 [B,2,ceil(T/128)]
 * (1 when the q tile holds rows of that modality; NULL = visit every tile).  Rows of kv tiles without a
 * work item are not written: the caller zero-fills dK*,dV*. */
int lb_attn_bwd_dkv(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                    const float* lse, const float* delta, const uint8_t* qflag, const uint8_t* qtile_has,
                    const int32_t* work_kv, int n_work, const int32_t* kv_start, const int32_t* kv_end, void* dK0,
                    void* dV0, void* dK1, void* dV1, int batch, int seqlen, int heads, int head_dim, int causal,
                    float scale, void* stream);
/* diagnostics: every CTA of subsequent lb_attn_bwd_dkv launches logs {q tiles, clock64 at entry, tile list ready, K/V
 * landed, last MMA issued, all MMAs done, exit, -} into buf ([n_work*heads][8] int64, device memory).  NULL = off */
int lb_attn_bwd_dkv_set_cta_log(void* buf);

/* ---- N1: KV-cached decode step of the bridge attention ---------------------
 * libra/models/libra/modeling_libra.py:343-397 with past_key_value, q_len = 1 (prepare_inputs_for_generation :1190-1231).
 * q [B, H*D] = rope(q) of the new tokens; the cache holds the operand tensors of lb_attn_prep_fwd for every position so
 * far, new token included: K0/V0 = Kfl/Vfl (what language queries see), K1/V1 = Kfv/Vfv (what vision queries see), each
 * [B, capacity, H*D] bf16, positions [0, kv_len) filled.  qflag[b] = 1 if sample b's new token is a vision token (NULL:
 * all language).  kv_start/kv_end [B]: the sample's visible key range (left padding / shorter samples; NULL: [0, kv_len)).
 * out [B, H*D] bf16, sample b at row out_row[b] (NULL: b).  The keys are split n_split ways; workspace = fp32
 * [lb_attn_decode_workspace_floats(...)] owned by the caller.  HBM-bound: 2 * (kv_end-kv_start) * H*D * 2 bytes read
 * per sample. */
int lb_attn_decode_workspace_floats(int batch, int heads, int head_dim, int n_split);
int lb_attn_decode(const void* q, const void* K0, const void* V0, const void* K1, const void* V1, const uint8_t* qflag,
                   const int32_t* kv_start, const int32_t* kv_end, const int32_t* out_row, float* workspace, void* out,
                   int batch, int heads, int head_dim, int capacity, int kv_len, int n_split, float scale, void* stream);

/* ---- tcgen05 GEMM ---------------------------------------------------------
 * C[M,N] = op(A) . op(B) (+ C when accumulate), bf16 inputs, fp32 accumulation in TMEM.
 *  trans_a = 0: A stored [M,K] (K contiguous);  1: A stored [K,M]
 *  trans_b = 0: B stored [N,K] (K contiguous, the nn.Linear weight layout);  1: B stored [K,N]
 * out_dtype LB_DT_BF16 / LB_DT_F32.  bias (bf16 [N], may be NULL) is added before `act`
 * (0 none, 1 quick_gelu).  Leading dimensions in elements, multiples of 8. */
int lb_gemm_bf16(const void* A, const void* B, void* C, const void* bias, int64_t M, int64_t N, int64_t K, int64_t lda,
                 int64_t ldb, int64_t ldc, int trans_a, int trans_b, int out_dtype, int accumulate, int act,
                 void* stream);

/* Same operation and work list as lb_attn_bwd_dkv, persistent streaming kernel (csrc/attn_bwd_dkv_stream.cu): one CTA per SM
 * walks its share of the (work item, head) list; K/V of the next item are loaded while the current item finishes, dK/dV leave
 * through a dedicated staging tile and TMA stores, scores and gradients have their own issuing threads.  plan_* / n_cta /
 * max_cta_items / head_group as for lb_attn_fwd_stream, over the work_kv list (libra_b200/schedule.py: stream_plan(which="kv")).
 * Every kv tile that has a work item is written whole (rows beyond seqlen clipped); rows without one are not touched.
 * lb_attn_bwd_dkv_stream_supported(): 0 when the platform's shared-memory window does not start 1024-byte aligned (the kernel's
 * 224 KB footprint has no room for alignment slack) -- use lb_attn_bwd_dkv then. */
int lb_attn_bwd_dkv_stream(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                           const float* lse, const float* delta, const uint8_t* qflag, const uint8_t* qtile_has,
                           const int32_t* work_kv, int n_work, const int32_t* plan_items, const int32_t* plan_off, int n_cta,
                           int max_cta_items, int head_group, const int32_t* kv_start, const int32_t* kv_end, void* dK0, void* dV0,
                           void* dK1, void* dV1, int batch, int seqlen, int heads, int head_dim, int causal, float scale,
                           void* stream);
int lb_attn_bwd_dkv_stream_max_cta_items(void);
int lb_attn_bwd_dkv_stream_supported(void);

/* ---- grouped persistent tcgen05 GEMM (the decoder's / ViT's / heads' dense products) ----------------------
 * Replaces every nn.Linear / F.linear product the reference sends to cuBLAS:
 *   LlamaAttention / LlamaMLP projections     libra/models/llama/modeling_llama.py:185-201
 *   LibraLinear chain F.linear(F.linear(x,A),B)  libra/models/libra/modeling_libra.py:192-199 (routed by :129-147, :310-319)
 *   SwiGLU product silu(gate x) * up x        modeling_libra.py:232-233
 *   CLIP q/k/v/out/fc1/fc2 (+bias, quick_gelu)  libra/models/clip/modeling_clip.py:279-282, 371-378
 *   lm_head / vision heads                    modeling_libra.py:1018-1052
 * and their dgrad / wgrad products in backward (all four operand layouts, no transposes materialised).
 * Up to 16 problems run in ONE persistent launch (CTA pairs, tcgen05 cta_group::2, 256 x 256 tiles, fp32 accumulation
 * in TMEM, TMA-store epilogue); for problem i
 *     C_i[M,N] = epi( op(A_i) . op(B_i) [+ bias_i] ) [+ D_i]          all bf16, row pitches ld* in elements (x8)
 *   trans_a = 0: A stored [M,K];  1: A stored [K,M]        trans_b = 0: B stored [N,K] (nn.Linear weight);  1: [K,N]
 *   D (optional, may alias C): addend, added after C was rounded to bf16 (`residual + linear(x)` of eager PyTorch;
 *     with D == C this is beta = 1 accumulation into a gradient buffer)
 *   epilogue LB_EPI_NONE; LB_EPI_QGELU: C = quick_gelu(.), G (optional) receives the pre-activation;
 *     LB_EPI_SWIGLU: B = gate weight, B2 = up weight (both [N,K]); C = silu(x B^T) * (x B2^T), G / U (optional)
 *     receive the gate / up pre-activations
 *   wait_on = j (< i, -1 none): A_i is C_j (same M): tiles of problem i start once the row block of C_j they read is
 *     complete -- the LibraLinear chain in one launch.
 *   workspace: lb_gemm_grouped_workspace_bytes() bytes of device memory private to the stream (zeroed by the call): the
 *     launch's tile counter (tiles are claimed dynamically, so CTAs the hardware cannot place while another kernel holds SMs
 *     -- an overlapped NCCL reduction -- cost throughput in proportion, not a second wave) and the chain counters.
 * Problems with M == 0 or N == 0 are skipped (an empty modality segment).
 *   flags & LB_GEMM_ACCUMULATE_PREV: this entry is one more product A_i . B_i summed (in the fp32 accumulator, before
 *     the epilogue) into the PREVIOUS entry's problem -- same M, N and operand layouts, own K; C / D / bias / epilogue
 *     are taken from the first entry of the chain.  dx = dq.Wq + dk.Wk + dv.Wv of a fan-out is one problem of three
 *     segments instead of three beta = 1 passes over dx.
 *   A dependent whose A is read transposed (or has another M) waits for the whole producer instead of one row block.
 *   When N is not a multiple of 8 the 16-byte unit holding the last columns of each row is written whole (zeros in
 *     the padding), so ldc must cover N rounded up to 8. */
#define LB_GEMM_ACCUMULATE_PREV 1
#define LB_EPI_NONE 0
#define LB_EPI_QGELU 1
#define LB_EPI_SWIGLU 2
typedef struct lb_gemm_problem {
    const void* A;
    const void* B;
    void* C;
    const void* D;
    const void* bias;
    const void* B2;
    void* G;
    void* U;
    int64_t M, N, K;
    int64_t lda, ldb, ldc, ldd;
    int32_t trans_a, trans_b, epilogue, wait_on;
    const float* alpha;   /* optional device fp32 scalar: C = epi(alpha * (A.B) + bias) + D */
    int64_t flags;        /* LB_GEMM_* bits */
} lb_gemm_problem;
int lb_gemm_grouped_workspace_bytes(const lb_gemm_problem* problems, int n);
int lb_gemm_grouped(const lb_gemm_problem* problems, int n, void* workspace, int64_t workspace_bytes, void* stream);
/* Skinny form for the one-token decode step (csrc/gemm_skinny.cu): the same problems with M <= 32 rows, restricted to
 * plain x W^T products (no transposes, chains, alpha, G/U outputs, accumulation segments; epilogue NONE (+bias, +D) or SWIGLU
 * (+D)), up to 8 per launch.  Swap-AB tcgen05 tiles of 128 output features x (16|32) tokens, split-K over enough CTAs that
 * every SM streams weights; partial sums are added in split order by the last CTA of a tile (deterministic).
 * workspace: lb_gemm_skinny_workspace_bytes() bytes = [64 KB tile counters | partials]; the counter region must be ZERO before
 * the first launch (every launch leaves it zero), so one workspace serves launches with different problem lists. */
int64_t lb_gemm_skinny_workspace_bytes(const lb_gemm_problem* problems, int n);
int lb_gemm_skinny(const lb_gemm_problem* problems, int n, void* workspace, int64_t workspace_bytes, void* stream);
/* debug hook (scripts/gemm_skinny_trace.py): device buffer of [work units][12] int64 that the following lb_gemm_skinny launches
 * fill with %globaltimer stamps per CTA (start, prologue done, dependency resolved, first stage landed, accumulator complete,
 * partial published, output written, end); NULL switches it off. */
int lb_gemm_skinny_set_trace(void* buf);
/* encoded TMA descriptors are cached by (pointer, shape, pitch, box): counters for the "no per-call encode" check */
int lb_gemm_tmap_cache_stats(int64_t* hits, int64_t* misses);

/* ---- A1 patch embedding (im2col-free, TMA-staged) ----------------------------
 * libra/models/clip/modeling_clip.py:193-228.  pixels [B,3,S,S] bf16 (NCHW), class_emb [C], pos_emb [(S/14)^2+1, C];
 * writes emb [B, (S/14)^2+1, C] = cat(cls, conv14x14/s14(pixels)) + pos.  The conv weight [C,3,14,14] is K-packed
 * once into [C, 768] (3 channels x 4 kernel-row groups x 64, zero padded) by lb_patch_embed_pack_weight. */
int lb_patch_embed_pack_weight(const void* weight, void* packed, int channels_out, int patch, void* stream);
int lb_patch_embed_fwd(const void* pixels, const void* weight_packed, const void* class_emb, const void* pos_emb, void* emb,
                       int batch, int image_size, int patch, int channels_out, void* stream);

/* ---- A18 routed heads: fused cross-entropy over a logits block -------------
 * libra/models/libra/modeling_libra.py:1159-1174 restricted to the finite vocabulary range of the row's modality
 * (the -inf placeholders of :1020-1052 contribute exp(-inf)=0).  logits [rows, vocab] bf16 (ld elements),
 * labels int64 (already shifted; -100 = ignore; label index relative to this block's vocabulary range).
 * Writes per-row loss (fp32, 0 where ignored) and overwrites logits with d(loss_sum)/dlogits * grad_scale. */
int lb_cross_entropy_fwd_bwd(void* logits, int64_t ld, const int64_t* labels, float* row_loss, int64_t rows, int vocab,
                             float grad_scale, void* stream);

/* ---- fused AdamW over flat bf16 buffers (torch.optim.AdamW semantics, reference recipe trainer.py:38-85 /
 * libra_pretrain.yaml:81-91 uses AdamW; fp32 maths, bf16 storage).  n elements (multiple of 8); step counts from 1. */
int lb_adamw_bf16(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, void* stream);

/* same with the gradient multiplied by *grad_scale (device fp32 scalar, may be NULL): gradient clipping folded into
 * the optimizer pass (trainer.py / libra_pretrain.yaml:116 max_grad_norm 1.0); and with weight decay switched off inside
 * `nodecay_ranges` (device int64 [n_nodecay][2], sorted, disjoint [lo, hi) in units of 8 elements; may be NULL): the
 * reference's decay exclusions (norm weights, biases; trainer.py:27-37) without splitting the launch. */
int lb_adamw_bf16_scaled(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, const float* grad_scale,
                         const int64_t* nodecay_ranges, int n_nodecay, void* stream);
/* global L2 norm of a flat bf16 gradient buffer and the clip factor, on the device, deterministic (two stages, no atomics):
 * out2[0] = ||g||, out2[1] = min(1, max_norm / (||g|| + 1e-6)) -- torch.nn.utils.clip_grad_norm_ semantics.
 * workspace: >= 64 floats (one partial per CTA; more floats = more CTAs, 2368 saturates a B200). */
int lb_grad_clip_scale(const void* grad, int64_t n, float max_norm, float* workspace, int workspace_floats, float* out2,
                       void* stream);

/* ---- N2: vision-tokenizer decode (ids -> pixels), the kernels around the grouped GEMM (csrc/vqdec.cu) ---------------
 * Reference: ImageTokenizer.decode libra/models/libra/image_tokenizer.py:97-124; LFQ.indices_to_codes
 * taming/modules/quantization/lookup_free_quantization.py:129-158; taming Decoder taming/modules/diffusionmodules/model.py
 * :34-41 (Normalize = GroupNorm 32, eps 1e-6), :29-31 (swish), :44-60 (Upsample), :85-141 (ResnetBlock), :141-230 (AttnBlock),
 * :474-588 (Decoder).  Activations are NHWC bf16; "padded" = the padded-row layout row(b,y,x) = (b*(H+2)+y+1)*(W+2)+x+1 in
 * which a 3x3 / pad 1 convolution is ONE lb_gemm_grouped problem of nine K segments over nine row shifts of its input (the
 * caller keeps W+3 guard rows before and after the buffer); pointers are to row 0 of that layout.  channels % 8 == 0. */
/* ids [Q, tokens] (int64 token ids; `offset` is subtracted) -> codes [tokens, ld] bf16: column q*bits+d = +-1 by bit
 * (bits-1-d) of code q (most significant first), columns >= Q*bits zero (K padding for the GEMM that follows) */
int lb_vq_codes(const int64_t* ids, int64_t offset, int num_codebooks, int64_t tokens, int bits, void* codes, int ld, void* stream);
/* y = GroupNorm(groups, eps, gamma, beta)(x) [then swish]; x compact or padded (interior read), y compact or padded (zero
 * borders written).  workspace: batch * lb_vq_groupnorm_chunks(H, W) * 2 * channels floats.  Deterministic (no atomics). */
int lb_vq_groupnorm_chunks(int height, int width);
int lb_vq_groupnorm(const void* x, const void* gamma, const void* beta, void* y, float* workspace, int batch, int height, int width,
                    int channels, int groups, float eps, int swish, int in_padded, int out_padded, void* stream);
/* nearest-neighbour resize into the padded layout (zero borders written); src_y [out_height] / src_x [out_width]: source
 * row / column of every output row / column (the host evaluates torch's index formula once per shape) */
int lb_vq_upsample_nearest(const void* x, void* y, const int32_t* src_y, const int32_t* src_x, int batch, int height, int width,
                           int channels, int out_height, int out_width, int in_padded, void* stream);
/* compact [B*H*W, channels] (row pitch ldx) -> padded with zero borders; addend (padded, may be NULL) is added on the interior */
int lb_vq_pad(const void* x, int64_t ldx, const void* addend, void* y, int batch, int height, int width, int channels, void* stream);
/* padded NHWC [.., channels] -> NCHW [B, channels_out, H, W] (the first channels_out channels) */
int lb_vq_to_nchw(const void* x, void* y, int batch, int height, int width, int channels, int channels_out, void* stream);
/* in place over bf16 rows: x = softmax(bf16(x * scale)) (fp32 maths; AttnBlock's w_ * c**-0.5 then softmax, :207-209) */
int lb_softmax_rows(void* x, int64_t rows, int cols, int64_t ld, float scale, void* stream);

/* ---- diagnostics: single-tile tcgen05 probes (tests/test_umma_probe.py) ---- */
int lb_probe_umma(int mode, const void* A, const void* B, float* D, int K, void* stream);

/* ---- N4 (second half): CLIP image preprocessing, raw uint8 RGB -> pixel_values ----------------------------------
 * libra/data/processors/libra_processor.py:44-60 (Expand2Square), :65-111 (LibraEvalImageProcessor / LibraImageProcessor);
 * libra/models/clip/image_processing_clip.py:124-217, 296-337 (resize shortest edge -> `size` with Pillow's antialiased
 * BICUBIC, center crop `crop`, rescale, normalise).  Bit-exact with Pillow's 8-bit two-pass resampler (22-bit fixed point,
 * clamp after each pass) and with the reference's float32 rounding sequence.
 * images: DEVICE buffer of packed HWC uint8 RGB images, image i at byte offsets[i], readable 16 bytes past the end of its
 * last image (the word-load kernels over-read the last row's tail); offsets / heights / widths / pad_rgb /
 * mean / std are HOST arrays.  pad_to_square != 0: paste on a square canvas of colour pad_rgb first.  out: DEVICE
 * [n_images, 3, crop, crop] LB_DT_F32 or LB_DT_BF16; out_u8 (optional, may be NULL): the uint8 image after resize + crop,
 * [n_images, crop, crop, 3].  workspace: lb_clip_preprocess_workspace(...) bytes of device memory. */
int64_t lb_clip_preprocess_workspace(const int32_t* heights, const int32_t* widths, int n_images, int size, int crop,
                                     int pad_to_square);
/* HOST only (no device needed): the fixed-point resampling table the kernels use for output indices [first_out, first_out +
 * n_out) of an axis resized in_size -> out_size -- Pillow's precompute_coeffs + normalize_coeffs_8bpc for BICUBIC.  Writes
 * n_out * ksize coefficients (if coeffs != NULL) and n_out * 2 bounds (first input index, taps); returns ksize (> 0) or an
 * error code (< 0). */
int lb_clip_resample_coeffs(int in_size, int out_size, int first_out, int n_out, int32_t* coeffs, int coeffs_capacity,
                            int32_t* bounds);
int lb_clip_preprocess(const uint8_t* images, const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n_images,
                       int size, int crop, int pad_to_square, const uint8_t* pad_rgb, const float* mean, const float* std,
                       double rescale_factor, void* out, int out_dtype, uint8_t* out_u8, void* workspace, int64_t workspace_bytes,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LIBRA_B200_H */
