/*
 * vct.h -- C ABI of libvct_b200.so: the B200 (sm_100a) kernels behind the
 * Video-Captioning-Transformer hot path.
 *
 * The reference (Kamino666/Video-Captioning-Transformer) has no FFI / plugin interface: its
 * arithmetic is reached through PyTorch modules.  This header is the
 * boundary a maintainer would bind instead; every entry point names the reference call it
 * replaces (file:line under /root/reference, or torch/... for the un-vendored PyTorch code the
 * reference calls into).  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the name says host
 *   - caller allocates every output and workspace; nothing is allocated inside
 *   - asynchronous on the given CUDA stream (a cudaStream_t passed as void*); no host sync
 *   - returns 0 on success, <0 on error (VCT_ERR_*); text via vct_last_error() (thread-local)
 *   - row-major, batch-first; "ld" arguments are row strides in ELEMENTS
 *   - dtype codes: VCT_F32 = 0, VCT_BF16 = 1 (storage type; arithmetic/accumulation is fp32)
 *   - masks: uint8, 1 = ignore (torch.bool True), like the reference's padding masks
 *   - dropout: counter-based (Philox4x32-7, the minimum round count the Philox authors qualify: passes BigCrush).  `rng_state` is a device array of two uint64
 *     {seed, step}; `site` identifies the dropout call site; the mask is a pure function of
 *     (seed, step, site, element index), so backward regenerates it.  drop_p <= 0 disables.
 */
#ifndef VCT_B200_H
#define VCT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCT_F32 0
#define VCT_BF16 1

#define VCT_ERR_INVALID (-1) /* bad argument / unsupported shape */
#define VCT_ERR_CUDA (-2)    /* CUDA runtime error (launch, attribute, driver entry point) */

typedef void* vct_stream_t; /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------------- */
int vct_version(void);
const char* vct_last_error(void);
/* sm count / compute capability of the current device; fails unless it is sm_100. */
int vct_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- per-step state ----------------------------------------------------------------------
 * rng_state: {seed, step}; adam_hyper: float[8] = {lr, beta1, beta2, eps, weight_decay, step,
 * 1-beta1^step, 1-beta2^step}.  One launch: step += 1 in both, bias corrections recomputed.
 * Replaces the Python-side bookkeeping of torch.optim.Adam.step (train.py:126) and of the
 * global torch RNG that nn.Dropout draws from; being a kernel it is CUDA-graph capturable. */
int vct_step_tick(unsigned long long* rng_state, float* adam_hyper, vct_stream_t stream);

/* ---- GEMM with fused epilogue --------------------------------------------------------------
 * C[M,N] = epilogue( sum_k opA[m,k] * opB[n,k] )
 *   a_trans = 0: A is [M,K] (K contiguous);  1: A is stored [K,M] (M contiguous)
 *   b_trans = 0: B is [N,K] (K contiguous) -- the nn.Linear weight layout;  1: B stored [K,N]
 * epilogue, in order: + bias[n]; + row_table[(m % row_period), n]; activation; + addend[m,n];
 * store C (c_dtype) and optionally C2.
 *   act = VCT_ACT_NONE
 *   act = VCT_ACT_GELU_FWD : C  = z (pre-activation), C2 = dropout(gelu(z))   [FFN linear1]
 *   act = VCT_ACT_GELU_BWD : C  = acc * gelu'(aux[m,n]) * dropmask[m,n]        [FFN backward]
 *   act = VCT_ACT_GELU_FWD_F / VCT_ACT_MUL_AUX: the same pair with the backward factor gelu'(z) * dropmask computed
 *         once in the forward epilogue (where z, the CDF and the density are at hand) and stored instead of z, so that
 *         the backward epilogue is one multiply (the training plans use this pair)
 * Replaces nn.Linear / F.linear calls: model/MMEncoder.py:246 (unify), model/CapDecoder.py:55
 * (generator), torch/nn/functional.py _in_projection_packed + out_proj inside
 * multi_head_attention_forward, torch/nn/modules/transformer.py:980-982,1197-1199 (FFN), and the
 * autograd dgrad / wgrad of each.
 * impl: VCT_GEMM_SIMT (fp32 FFMA tiles; any dtype mix; the exact path), VCT_GEMM_TCGEN05
 * (bf16 operands, TMA + tcgen05.mma, fp32 accumulation in TMEM), or VCT_GEMM_TCGEN05_X3 / _X6: the
 * reference-precision tensor-core path -- fp32 operands are split into 2 / 3 bf16 pieces inside the call
 * (into split_ws) and the 3 / 6 leading cross terms are accumulated in fp32 by ONE tcgen05 launch with
 * K' = terms * K (csrc/gemm_split.cu); X6 reproduces fp32 products to ~2^-23, X3 to ~2^-16. */
#define VCT_ACT_NONE 0
#define VCT_ACT_GELU_FWD 1
#define VCT_ACT_GELU_BWD 2
#define VCT_ACT_GELU_FWD_F 3 /* C = f = gelu'(z) * dropmask (the factor backward needs), C2 = dropout(gelu(z)) */
#define VCT_ACT_MUL_AUX 4    /* C = acc * aux[m,n] (+ addend): FFN backward with the saved factor f, no erf / RNG */
#define VCT_GEMM_SIMT 0
#define VCT_GEMM_TCGEN05 1
#define VCT_GEMM_TCGEN05_X3 2
#define VCT_GEMM_TCGEN05_X6 3

typedef struct {
    int M, N, K;
    const void* A; int a_dtype; long long lda; int a_trans;
    const void* B; int b_dtype; long long ldb; int b_trans;
    void* C; int c_dtype; long long ldc;
    void* C2; int c2_dtype; long long ldc2;          /* optional (NULL) */
    const float* bias;                               /* [N] or NULL */
    const float* row_table; int row_period;          /* fp32 [row_period, N] or NULL */
    const float* addend; long long ld_addend;        /* fp32 [M, N] or NULL (may alias C if c_dtype is F32) */
    int act;
    const void* aux; int aux_dtype; long long ld_aux;/* z for VCT_ACT_GELU_BWD */
    float drop_p; const unsigned long long* rng_state; unsigned int site;
    int impl;
    /* optional split-K workspace (fp32, caller-allocated): lets the tcgen05 path split a long K (e.g. the
     * generator dgrad, K = vocab) over several CTAs and reduce deterministically in a second kernel */
    void* splitk_ws; long long splitk_ws_floats;
    /* VCT_GEMM_TCGEN05_X3 / _X6 only: caller-allocated, 256-byte aligned scratch of at least
     * vct_gemm_split_workspace_bytes(M, N, K, a_trans, b_trans, terms) bytes for the bf16 pieces of A and B */
    void* split_ws; long long split_ws_bytes;
} vct_gemm_args;

int vct_gemm(const vct_gemm_args* args, vct_stream_t stream);
/* `count` independent GEMMs (an array of vct_gemm_args).  A group of weight-gradient GEMMs -- tcgen05, bf16 operands stored
 * [K, M] and [K, N] (a_trans = b_trans = 1), fp32 C, no epilogue extras, count <= 8: the dW = dY^T X products of one
 * transformer layer (autograd of the nn.Linear calls listed above) -- runs as ONE persistent launch over the tiles of all
 * problems; any other group is executed problem by problem with vct_gemm. */
int vct_gemm_grouped(const vct_gemm_args* args, int count, vct_stream_t stream);
long long vct_gemm_split_workspace_bytes(int M, int N, int K, int a_trans, int b_trans, int terms);
/* The operand decomposition on its own: src fp32 [rows, cols] -> dst bf16 with the `terms` (3 or 6) pieces of side
 * 0 (A: h h m | h h m m h l) or 1 (B: h m h | h m h m l h) laid along the contraction axis (1: columns, dst [rows,
 * terms * Kp]; 0: rows, dst [terms * Kp, ld_dst]; Kp = K rounded up to 8, padding zero-filled). */
int vct_split_bf16(const float* src, long long ld_src, int rows, int cols, int axis, int side, int terms, void* dst,
                   long long ld_dst, vct_stream_t stream);

/* Tuning hook for the tcgen05 path (tools/gemm_sweep.py): force the tile width (block_n = 64 / 128 / 256;
 * 0 restores the built-in cost model), the split-K factor (0/1 = none) and the kernel flavour
 * (ring = 0: one CTA per SM with the full operand ring, 1: half ring so two CTAs share an SM,
 * -1: persistent kernel, block_n 256 only).  Process-wide; not used on the product path. */
int vct_gemm_tune(int block_n, int splits, int ring);
/* Debug: CTA (0,0,0) of every following tcgen05 tile kernel writes clock64 timestamps of its pipeline phases
 * into dev_buf (>= 128 int64; NULL turns tracing off): [0] entry, [1] setup done, [2] after griddepcontrol.wait,
 * [3] last MMA committed, [4] accumulator visible to the epilogue, [5] TMEM drained to smem, [6] stores issued,
 * [7] exit; [16+kb] TMA of k-block kb issued, [56+kb] operands of k-block kb landed. */
int vct_gemm_trace(void* dev_buf);

/* ---- frame staging -----------------------------------------------------------------------
 * feats fp32 [B,T,Din] -> out [B*(T+1), Din]: row 0 of each batch element = mean over ALL T
 * frames (padded ones included), rows 1..T = the frames.  Because unify is linear,
 * unify(mean) == mean(unify): this lets one GEMM produce the global token and the frame tokens
 * (model/MMEncoder.py:246-250, GlobalAggregation avg :183-197). */
int vct_prep_frames(const float* feats, void* out, int out_dtype, int B, int T, int Din, vct_stream_t stream);

/* ---- attention core ------------------------------------------------------------------------
 * softmax(scale * Q K^T + mask) V per (batch, head), one warp per (b,h), warp-shuffle softmax.
 * element (b, i, h, c) of q is q[b*q_bs + i*q_ld + h*dh + c]; same for k, v (Lk rows), o.
 * key_pad: uint8 [B, Lk] or NULL; causal: key j > query i masked (requires Lq == Lk).
 * probs: optional fp32 [B,H,Lq,Lk] post-softmax, pre-dropout probabilities (need_weights).
 * Replaces F.scaled_dot_product_attention / the softmax path of
 * torch/nn/functional.py multi_head_attention_forward incl. mask merge (:6607-6620) and
 * attention-probability dropout.  Limits: Lq, Lk <= 1024, dh <= 128, dh % 4 == 0.
 * Lq, Lk <= 64 (every shipped config): the whole head lives in one CTA / one tcgen05 tile.  Longer sequences (the
 * reference truncates neither captions nor frames, train.py / dataloader.py) run the tiled kernels: 16-row tiles,
 * the other operand streamed in 32-row blocks; their backward needs the `row_stats` workspace and does not produce
 * `dbias` (vct_colsum over dq | dk | dv does).  The dropout mask of probability (row r, key j) is element
 * r * 64 + j of `site` for Lk <= 64 and element r * 8 * ceil(Lk / 8) + j otherwise (vct_dropout_mask). */
typedef struct {
    int B, H, Lq, Lk, dh;
    int dtype;                          /* storage type of q, k, v, o and their gradients */
    const void* q; long long q_ld;
    const void* k; long long k_ld;
    const void* v; long long v_ld;
    void* o; long long o_ld;
    const unsigned char* key_pad;
    int causal;
    float scale;
    float drop_p; const unsigned long long* rng_state; unsigned int site;
    float* probs;
    /* backward only */
    const void* d_o; long long do_ld;
    void* dq; long long dq_ld;
    void* dk; long long dk_ld;
    void* dv; long long dv_ld;
    /* batch strides in elements; 0 = dense default (Lq * ld for q/o/dq/d_o, Lk * ld for k/v/dk/dv).
     * Non-default strides let incremental decoding attend over a [B, Lmax, 3d] K/V cache. */
    long long q_bs, k_bs, v_bs, o_bs, do_bs, dq_bs, dk_bs, dv_bs;
    /* backward only, optional: gradient of the packed in-projection bias (nn.MultiheadAttention.in_proj_bias),
     * i.e. the column sums of dq | dk | dv over all (batch, position) rows, fp32 [3 * H * dh] laid out q | k | v.
     * dbias_partials: fp32 workspace [B, 3 * H * dh]; dbias_counters: H zero-initialised uint32 (self-resetting).
     * Deterministic (per-CTA partials, fixed-order final sum by the last CTA of each head). */
    float* dbias; float* dbias_partials; unsigned int* dbias_counters;
    /* backward only, required when Lq > 64 or Lk > 64: fp32 workspace [B * H * Lq, 4], 16-byte aligned (row max,
     * 1 / row sum and sum_j P dP of every probability row, handed from the query-tile to the key-tile kernel). */
    float* row_stats;
} vct_attn_args;

int vct_attn_fwd(const vct_attn_args* args, vct_stream_t stream);
int vct_attn_bwd(const vct_attn_args* args, vct_stream_t stream);

/* ---- the three attention flavours of the reference (projection + core) ----------------------
 * x      : [B*L, d]   (x_dtype)            query-side input rows
 * mem    : [B*Lk, d]  cross only; NULL for self-attention
 * w_in   : [3d, d] packed q,k,v (w_dtype), b_in fp32 [3d]   (nn.MultiheadAttention.in_proj_*)
 * qkv    : workspace / saved-for-backward [B*L, 3d] (self) ; q [B*L, d] + kv [B*Lk, 2d] (cross)
 * o      : [B*L, d] attention output BEFORE out_proj
 * enc self  : key-padding mask [B,L]   (torch/nn/modules/transformer.py:961-978)
 * dec self  : causal + key-padding     (torch/nn/modules/transformer.py:1158-1175)
 * dec cross : NO mask      (torch/nn/modules/transformer.py:1177-1195); if kv_ready
 *             is nonzero the K/V projection of `mem` is taken from `kv` as is (decode reuses it). */
typedef struct {
    int B, L, Lk, d, H;
    int dtype;                           /* storage type of x, mem, w_in, qkv/q/kv, o */
    const void* x; const void* mem;
    const void* w_in; const float* b_in;
    void* qkv; void* kv; int kv_ready;
    void* o;
    const unsigned char* key_pad;
    float drop_p; const unsigned long long* rng_state; unsigned int site;
    float* probs;
    int gemm_impl;
    void* split_ws; long long split_ws_bytes;   /* for gemm_impl = VCT_GEMM_TCGEN05_X3 / _X6 (see vct_gemm_args) */
} vct_mha_args;

int vct_attn_enc_self_fwd(const vct_mha_args* args, vct_stream_t stream);
int vct_attn_dec_self_fwd(const vct_mha_args* args, vct_stream_t stream);
int vct_attn_dec_cross_fwd(const vct_mha_args* args, vct_stream_t stream);

/* ---- residual + dropout + LayerNorm ---------------------------------------------------------
 * s = x + dropout(r)   (x may be NULL: s = r, no dropout)   ;  y = LN(s) * gamma + beta, eps 1e-5
 * writes y (fp32), optionally y_c (y_c_dtype copy for the next GEMM), s_out (fp32 pre-LN sum,
 * may alias r), mean/rstd [R] for backward.  One warp per row.  d % 4 == 0, d <= 1024.
 * Replaces x = norm(x + dropout(sublayer(x))) of torch/nn/modules/transformer.py:946-959,
 * 1131-1156 and the final nn.LayerNorm of both stacks (model/MMEncoder.py:238,
 * model/CapDecoder.py:20). */
int vct_ln_residual_fwd(const float* x, const float* r, const float* gamma, const float* beta,
                        float* y, void* y_c, int y_c_dtype, float* s_out, float* mean, float* rstd,
                        int R, int d, float drop_p, const unsigned long long* rng_state, unsigned int site,
                        vct_stream_t stream);

/* backward: dy fp32 [R,d] -> ds fp32 (gradient wrt s: goes to the residual input as is) and
 * dr_c (dr_dtype) = ds * dropmask (gradient wrt the branch output r; GEMM operand).
 * Column sums: dgamma += sum dy*xhat, dbeta += sum dy, dbias_r = sum dr (bias gradient of the
 * linear that produced r; NULL to skip).  `partials` is a fp32 workspace of
 * vct_ln_bwd_workspace_floats(R, d) floats; `counter` is unused (kept for ABI stability, may be NULL).
 * dgamma/dbeta/dbias_r are OVERWRITTEN (deterministic two-kernel reduction, no atomics).
 * If dgamma, dbeta and dbias_r are ALL NULL only the per-CTA partials are produced and the caller
 * finishes with vct_ln_bwd_reduce (same R, d) -- e.g. on a side stream, off the critical path of backward. */
long long vct_ln_bwd_workspace_floats(int R, int d);
int vct_ln_bwd_reduce(const float* partials, int R, int d, float* dgamma, float* dbeta, float* dbias_r,
                      vct_stream_t stream);
int vct_ln_residual_bwd(const float* dy, const float* s, const float* mean, const float* rstd, const float* gamma,
                        float* ds, void* dr_c, int dr_dtype, float* dgamma, float* dbeta, float* dbias_r,
                        float* partials, unsigned int* counter,
                        int R, int d, float drop_p, const unsigned long long* rng_state, unsigned int site,
                        vct_stream_t stream);

/* ---- eval-mode encoder fast path ------------------------------------------------------------
 * x[r, :] = 0 for every row r with mask[r] != 0 (x fp32 [R, d], d % 4 == 0).  Under eval() + no_grad with a
 * key-padding mask -- what val_epoch (train.py:151-168) and eval.py:140 run -- nn.TransformerEncoder takes torch's
 * nested-tensor path (torch/nn/modules/transformer.py:452-548), which zero-fills the padded positions before the
 * final LayerNorm: memory rows of padded frames become norm.bias.  The eval plan calls this on the last
 * encoder layer's output right before vct_ln_residual_fwd. */
int vct_zero_rows(float* x, const unsigned char* mask, int R, int d, vct_stream_t stream);

/* ---- input staging ----------------------------------------------------------------------------
 * One launch moves a step's DEVICE-resident batch into the buffers the launch plans read (what the reference does with
 * `.to(device)` per modality, train.py:120-121, plus CapPreprocessor's mask, model/CapPreprocessor.py:35):
 *   feats_src fp32 [B,T,Din] -> feats_dst (16-byte aligned, Din % 4 == 0); vid_src bytes [B,T] (NULL: nothing padded)
 *   -> vid_dst [B,T+1] whose column 0 (the global token) is never padded (model/MMEncoder.py:252-260);
 *   ids_src int64 [B,S1] -> ids_dst; tok_dst [B,S1-1] = tok_src (explicit key-padding mask) or ids[:, :-1] == pad_id.
 * feats_src NULL skips the feature half, ids_src NULL the token half. */
int vct_stage_inputs(const float* feats_src, float* feats_dst, const unsigned char* vid_src, unsigned char* vid_dst,
                     int B, int T, int Din, const long long* ids_src, long long* ids_dst,
                     const unsigned char* tok_src, unsigned char* tok_dst, int S1, long long pad_id, vct_stream_t stream);

/* ---- token embedding + positional table ------------------------------------------------------
 * x[b,s,:] = dropout(E[ids[b*ids_ld + s], :] + pos[s, :])   (no sqrt(d) scaling)
 * model/CapDecoder.py:48 + model/Embedding.py:23-25.  ids int64.  Writes x fp32 and optional x_c.
 * pos_offset: position of column 0 (incremental decode embeds one new column at a time). */
int vct_embed_fwd(const long long* ids, long long ids_ld, const float* E, const float* pos,
                  float* x, void* x_c, int x_c_dtype, int B, int S, int d, int V, int pos_offset,
                  float drop_p, const unsigned long long* rng_state, unsigned int site, vct_stream_t stream);
/* dE[ids] += dropmask * dx  (dE pre-zeroed; rows with id == pad_id receive nothing: padding_idx) */
int vct_embed_bwd(const long long* ids, long long ids_ld, const float* dx, float* dE, int B, int S, int d, int V,
                  int pad_id, float drop_p, const unsigned long long* rng_state, unsigned int site,
                  vct_stream_t stream);

/* The same gradient in its sparse form: rows[b*S + s, :] = dropmask * dx[b*S + s, :] (fp32 [B*S, d]), no scatter.  The
 * data-parallel trainer all-gathers these rows + the ids of every rank (B*S*d*4 bytes per rank, < 4 MB) and scatters them
 * locally with vct_embed_bwd (drop_p = 0) instead of all-reducing the dense [V, d] table gradient that DDP exchanges
 * (train.py:218; 94 MB with at most B*S non-zero rows). */
int vct_embed_bwd_rows(const float* dx, float* rows, int B, int S, int d, float drop_p,
                       const unsigned long long* rng_state, unsigned int site, vct_stream_t stream);
/* Deterministic scatter: dE[id, :] = sum over the tokens (b, s) with ids[b*ids_ld + s] == id of rows[b*S + s, :], summed in a
 * FIXED order (tokens sorted by id, then by index), rows of dE that no token maps to are left untouched, pad / out-of-range
 * ids are skipped.  Every data-parallel rank runs it on the same gathered rows and obtains bit-identical table gradients
 * (replicas must not drift; an atomicAdd scatter sums in a run-dependent order).  keys_ws: uint32 workspace [B*S].
 * Limits: B*S <= 16384, V <= 32768, d % 4 == 0, d <= 1024. */
int vct_embed_bwd_det(const long long* ids, long long ids_ld, const float* rows, float* dE, int B, int S, int d, int V,
                      int pad_id, unsigned int* keys_ws, vct_stream_t stream);
/* The two halves of vct_embed_bwd_det: the sort depends on the ids only, so the trainer runs it at the START of a step
 * (off the tail of backward) and only the segment sums once the gradient rows have been gathered. */
int vct_embed_sort(const long long* ids, long long ids_ld, int B, int S, int V, int pad_id, unsigned int* keys_ws,
                   vct_stream_t stream);
int vct_embed_segment_sum(const unsigned int* keys, const float* rows, float* dE, int n, int d, vct_stream_t stream);
/* stamp[id] = current step (rng_state[1]) for every non-pad token id: row `id` of the table is touched by this step iff
 * stamp[id] == step.  uint32 [V], never cleared. */
int vct_embed_mark(const long long* ids, long long ids_ld, int B, int S, int V, int pad_id, unsigned int* stamp,
                   const unsigned long long* rng_state, vct_stream_t stream);
/* dE[ids[b*ids_ld + s], :] = 0 for every (b, s): re-zeroes exactly the rows vct_embed_bwd scattered into, once the
 * optimizer has consumed them (replaces a memset of the whole table gradient per step; optimizer.zero_grad, train.py:124). */
int vct_embed_zero(const long long* ids, long long ids_ld, float* dE, int B, int S, int d, int V, vct_stream_t stream);

/* ---- SCE loss (model/loss.py:69-92; closed form SURVEY Q9) -----------------------------------
 * logits fp32 [N, ld_logits] (N = B*S), label of row (b,s) = ids[b*ids_ld + s + 1].
 * loss = alpha * mean_{label != pad}(lse - z_y) + beta * mean_all( A * sum_{c != y} clamp(p_c,1e-7,1) )
 * (alpha == 1: plain cross entropy, model/CapDecoder.py:28-30).
 * One CTA per row; the row is staged in shared memory once (bulk async copy) and all passes run
 * from there.  If loss_out != NULL writes the scalar loss (row_parts: fp32 workspace [N,2],
 * counter: zero-initialised uint32, self-resetting).  If dlogits != NULL writes
 * d loss / d logits * (*upstream or 1) in dl_dtype with row stride ld_dl (columns V..ld_dl-1 are
 * zeroed so the buffer can feed a GEMM whose K is padded). */
int vct_sce(const float* logits, long long ld_logits, const long long* ids, long long ids_ld,
            int B, int S, int V, float alpha, float beta, int pad_id,
            float* loss_out, float* row_parts, unsigned int* counter,
            void* dlogits, int dl_dtype, long long ld_dl, const float* upstream,
            vct_stream_t stream);

/* The same with the logits stored in logits_dtype (VCT_F32 or VCT_BF16): the training plans let the generator GEMM write
 * bf16 logits, so no fp32 [B*S, V] buffer exists on that path.  Rows of up to 32768 columns with ld % 8 == 0 and
 * ld_dl == ld_logits run on a register-resident kernel (no shared-memory staging). */
int vct_sce_typed(const void* logits, int logits_dtype, long long ld_logits, const long long* ids, long long ids_ld,
                  int B, int S, int V, float alpha, float beta, int pad_id,
                  float* loss_out, float* row_parts, unsigned int* counter,
                  void* dlogits, int dl_dtype, long long ld_dl, const float* upstream,
                  vct_stream_t stream);

/* ---- column sums (bias gradients): out[n] = sum_m X[m,n]; deterministic ----------------------
 * partials: fp32 workspace vct_colsum_workspace_floats(M, N); counter as above. */
long long vct_colsum_workspace_floats(int M, int N);
int vct_colsum(const void* X, int dtype, long long ld, int M, int N, float* out, float* partials,
               unsigned int* counter, vct_stream_t stream);

/* ---- Adam over the flat parameter arena (torch.optim.Adam semantics, train.py:22-31,126) -----
 * p, m, v: fp32 [n]; g: gradients in g_dtype (VCT_F32, or VCT_BF16 when the data-parallel trainer exchanged the
 * gradient bucket in bf16); hyper: the float[8] of vct_step_tick (already ticked for this step);
 * grad_scale multiplies g first (1/world_size after a SUM all-reduce).  If p_c != NULL also
 * writes the bf16 shadow copy the tensor-core GEMMs read next step. */
int vct_adam(float* p, const void* g, int g_dtype, float* m, float* v, void* p_c, long long n, const float* hyper,
             float grad_scale, vct_stream_t stream);

/* Adam on the rows of a [R, d] table (the embedding) selected by the step stamps of vct_embed_mark: touched = 0 updates the
 * rows this step does NOT touch -- their gradient is identically zero and is not read; torch.optim.Adam still decays m, v
 * and moves p along m for them -- which needs nothing from this step's backward and is issued at the start of the step;
 * touched = 1 updates the (at most B*S) rows that received gradient, at the end.  Together they equal vct_adam over the
 * table.  g may be NULL when touched = 0. */
int vct_adam_rows(float* p, const float* g, float* m, float* v, void* p_c, int R, int d, const float* hyper, float grad_scale,
                  const unsigned int* stamp, const unsigned long long* rng_state, int touched, vct_stream_t stream);

/* fp32 -> dtype copy (initial bf16 shadow of the parameters, input staging) */
int vct_cast(const float* src, void* dst, int dst_dtype, long long n, vct_stream_t stream);

/* ---- greedy decode step tail (model/MMT4Caption.py:165-172) ----------------------------------
 * next[b] = argmax_v logits[b, :V] (lowest index on ties, like torch.max); ys[b, t] = next[b];
 * ended[b] |= (next[b] == end_id); *n_ended = number of ended rows (device int32). */
int vct_argmax_append(const float* logits, long long ld_logits, int B, int V, long long* ys, long long ys_ld, int t,
                      int end_id, int* ended, int* n_ended, vct_stream_t stream);

/* ---- data-parallel gradient exchange over NVLink peer memory (csrc/peer_comm.cu) ---------------
 * Replaces DistributedDataParallel's bucketed NCCL all-reduce (train.py:218).  One process per GPU: every rank creates
 * a communication region (cudaMalloc), publishes its cudaIpc handle (64 bytes, exchanged by the host over any
 * transport), and maps all peers with vct_comm_connect.  The collectives are plain kernels on `stream` (CUDA-graph
 * capturable) that read the peers' regions over NVLink and synchronise through flag words in peer memory.  All ranks
 * must issue the same sequence of calls per channel (0..3) with the same arguments; `ctas` (grid size, <= 64) must be the
 * same on every rank.  world <= 8.
 *   vct_peer_allreduce_bf16: in-place SUM over ranks of n_elems bf16 values at byte_off of every region (two-shot, push
 *       model: every rank sends its values of chunk p into rank p's staging area at stage_off -- world * ceil(n_elems / 8 /
 *       world) * 16 bytes, disjoint from the data --, reduces its own chunk in fp32 in rank order with one rounding, and
 *       writes the result into every region): identical results on all ranks.  n_elems % 8 == 0, offsets % 16 == 0.
 *   vct_peer_allgather: the range is [world][slot_bytes]; rank r has written slot r of its own region; afterwards every
 *       region holds every slot.  slot_bytes % 16 == 0.
 *   vct_comm_status: 0, or 1 after a barrier timed out (~20 s: a peer died or the ranks diverged); synchronises.
 *   vct_comm_connect_in_process: ranks that live in ONE process (tests): peer_bases[p] = vct_comm_base(handle of rank p). */
int vct_comm_create(int rank, int world, long long bytes, int ctas, void** handle_out);
void* vct_comm_base(void* handle);
int vct_comm_ipc_handle(void* handle, unsigned char* out64);
int vct_comm_connect(void* handle, const unsigned char* all_handles /* world x 64 bytes */);
int vct_comm_connect_in_process(void* handle, void* const* peer_bases, const int* peer_devices);
int vct_peer_allreduce_bf16(void* handle, long long byte_off, long long n_elems, long long stage_off, int channel,
                            vct_stream_t stream);
int vct_peer_allgather(void* handle, long long byte_off, long long slot_bytes, int channel, vct_stream_t stream);
int vct_comm_status(void* handle);
int vct_comm_destroy(void* handle);

/* ---- debug: the keep mask (1 = kept) of `n` elements of a dropout site ----------------------- */
int vct_dropout_mask(unsigned char* out, long long n, float drop_p, const unsigned long long* rng_state,
                     unsigned int site, vct_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VCT_B200_H */
