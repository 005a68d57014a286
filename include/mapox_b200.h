/* mapox_b200.h - C ABI of libmapox_b200.so: the B200 (sm_100a) hot path of mapox_trainer's
 * transformer-over-time policy (reference: gabe00122/jaxrl, paths below are relative to its repo root).
 *
 * Conventions:
 *  - every pointer is a DEVICE pointer to caller-owned, row-major, contiguous memory unless a stride is given;
 *  - bf16 tensors are passed as raw 16-bit words (mpx_bf16); bools arrive as 1-byte PRED (uint8_t);
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t); nothing synchronises, nothing is allocated
 *    persistently, all calls are re-entrant and CUDA-graph capturable;
 *  - return value 0 = MPX_OK, negative = error; mpx_last_error() returns a thread-local message;
 *  - there is no CPU fallback: on a machine without an sm_100 device every compute call returns MPX_ERR_CUDA.
 *
 * The reference has no FFI of its own (it is pure JAX); each entry point cites the reference function whose
 * device work it replaces.  INTEGRATION.md shows the jax.ffi / ctypes binding for each.
 */
#ifndef MAPOX_B200_H_
#define MAPOX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mpx_stream_t;
typedef uint16_t mpx_bf16;

enum { MPX_OK = 0, MPX_ERR_INVALID = -1, MPX_ERR_CUDA = -2, MPX_ERR_UNSUPPORTED = -3 };

const char* mpx_last_error(void);
int mpx_version(void);
/* Debug/tuning knobs (tests only). key 0: swap LBO/SBO in MN-major UMMA descriptors.
 * key 1: decode attention implementation (0 = mma.sync online softmax, default; 1 = SIMT exact two-pass;
 * 2 = mma.sync exact two-pass).  key 2: training attention (0 = tcgen05/TMEM forward and backward when q,k,v are packed
 * and Nkv == 1 - forward with the four heads of a GQA group per CTA when Nq == 4 - default; 1 = mma.sync everywhere;
 * 2 = tcgen05 forward with one head per CTA, mma.sync backward).  key 3: mpx_glu_fused cluster size (0 = auto; 1, 2 or 4 = force).
 * key 4: mpx_value_head_fused cluster size, same values.  key 6: programmatic dependent launch, one bit per kernel
 * family (1 = GEMM / elementwise / value head, 2 = observation encoder, 4 = decode attention, 8 = fused GLU,
 * 16 = policy head; default 12).  No key changes results beyond the documented choice of an equivalent kernel. */
int mpx_debug_set(int key, int value);
/* Bring-up aid: device buffer of >= 32 uint64 slots in which CTA 0 of the fused rollout kernels records %globaltimer at
 * its phase boundaries (NULL switches it off, the default). */
int mpx_debug_set_stamps(void* device_buffer);

/* ---------------------------------------------------------------------------------------------- GAE
 * Rollout.calculate_advantage, mapox_trainer/rollout.py:128-167 (reverse lax.scan + global normalisation).
 * rewards (B,T) f32, values (B,T+1) f32, next_terminated (B,T) u8 -> advantages, targets (B,T) f32 (un-normalised)
 * and stats[0..1] = (sum adv, sum adv^2) in f64 for the normalisation; a multi-GPU caller all-reduces `stats`
 * before mpx_adv_normalize.  workspace >= mpx_gae_workspace_bytes(B,T). */
size_t mpx_gae_workspace_bytes(int B, int T);
int mpx_gae(const float* rewards, const float* values, const uint8_t* next_terminated, int B, int T,
            double discount, double gae_lambda, float* advantages, float* targets, double* stats,
            void* workspace, mpx_stream_t stream);
/* rollout.py:159-162: adv = (adv - mean) / (std + 1e-8), population std, over `count` elements in total. */
int mpx_adv_normalize(float* advantages, long long n, const double* stats, double count, mpx_stream_t stream);

/* ---------------------------------------------------------------------------------------------- HL-Gauss
 * HlGaussValue.get_value, mapox_trainer/values.py:43-45. logits (n,NL) f32, centers (NL) f32 -> value (n) f32. */
int mpx_hlgauss_value(const float* logits, long long n, int n_logits, const float* centers, float* value,
                      mpx_stream_t stream);
/* HlGaussValue.get_loss, values.py:47-63, and its gradient: loss[0] = mean CE; dlogits (n,NL) (nullable) =
 * grad_scale * d loss / d logits.  supports (NL+1) f32 are the bin edges (values.py:12-19). */
size_t mpx_hlgauss_loss_workspace_bytes(long long n);
int mpx_hlgauss_loss(const float* logits, const float* targets, long long n, int n_logits, const float* supports,
                     float vmin, float vmax, float sigma, float grad_scale, float* loss, float* dlogits,
                     mpx_bf16* dlogits_bf16, int ld_bf16, void* workspace, mpx_stream_t stream);
/* dlogits_bf16 (nullable): the same gradient as bf16 rows of pitch ld_bf16 >= n_logits, zero padded - the
 * operand layout the tensor-core backward of the value head consumes. */

/* ---------------------------------------------------------------------------------------------- dense (tcgen05)
 * nnx.Linear / nnx.LinearGeneral with dtype=bf16 (attention.py:53-92, feed_forward.py:59-71, network.py:228-285):
 * y = bf16(x @ wt^T) [+ bias -> bf16] [gelu -> bf16] [+ residual -> bf16].  TMA-fed persistent tcgen05 GEMM,
 * fp32 accumulation in TMEM.  wt is the TRANSPOSED flax kernel: (N,K), row n = output feature n.
 * The same entry point computes the data gradient of a Linear: dx = linear(dy, wt = kernel as stored (in,out)). */
typedef struct {
  const mpx_bf16* x;        long long ldx;    /* (M,K) */
  const mpx_bf16* wt;       long long ldw;    /* (N,K) */
  const mpx_bf16* bias;                       /* (N) or NULL */
  const mpx_bf16* residual; long long ldr;    /* (M,N) or NULL */
  mpx_bf16* y;              long long ldy;    /* (M,N) or NULL */
  float* y_f32;             long long ldy32;  /* optional raw fp32 accumulator output (M,N) */
  int M, N, K;
  int act;                                    /* 0 none, 1 gelu (tanh approximation, jax.nn.gelu default),
                                               * 2 gelu TRANSPOSE for data gradients: y = bf16(bf16(x @ wt^T) * gelu'(z))
                                               *   with z = the saved pre-activation passed in `residual` (no add) */
  mpx_bf16* y_pre;                            /* optional (M,N), pitch ldy: bf16 pre-activation saved for backward */
} mpx_linear_args;
int mpx_linear(const mpx_linear_args* a, mpx_stream_t stream);

/* Weight gradient of a Linear: dw (K,N) f32 += x^T (K,M) @ dy (M,N)   (contraction over the M tokens; both
 * operands are read MN-major straight from their row-major activations - no transposes). dw must be zeroed
 * (or hold a running sum) by the caller. */
typedef struct {
  const mpx_bf16* x;  long long ldx;   /* (M,K) */
  const mpx_bf16* dy; long long ldy;   /* (M,N) */
  float* dw;          long long lddw;  /* (K,N) fp32, accumulated (+=) */
  int M, N, K;
  void* workspace; size_t workspace_bytes;   /* optional: >= mpx_linear_wgrad_workspace_bytes(M,N,K) makes the split-K
                                                reduction a deterministic two-pass sum; otherwise red.add atomics */
} mpx_wgrad_args;
size_t mpx_linear_wgrad_workspace_bytes(int M, int N, int K);
int mpx_linear_wgrad(const mpx_wgrad_args* a, mpx_stream_t stream);

/* q/k/v projections + RoPE (+ KV-cache ring write): AttentionBlock.__call__ attention.py:121-134 and
 * update_kv_cache :108-114, apply_rope positional_embeddings.py:7-27.
 * wt_qkv rows: [q heads (Nq*D) | k heads (Nkv*D) | v heads (Nkv*D)], each row = H input features.
 * Row r uses position pos[r % pos_len] (decode: pos_len = M = B; train: pos_len = T).
 * If cache_S > 0, k/v are the cache bases (B,S,Nkv,D) and the rows are written at ring slot pos[0] % cache_S
 * (one scalar for the whole batch, exactly as attention.py:109). head_dim must be 32.
 * norm_scale (H) f32, when not NULL (H == 128 only), fuses the block's pre-norm (network.py:160, history_norm) in front:
 * the rows of x are RMS-normalised in shared memory before they enter the tensor cores. */
typedef struct {
  const mpx_bf16* x; long long ldx;          /* (M,H) */
  const mpx_bf16* wt_qkv;                    /* ((Nq+2Nkv)*D, H) */
  const int32_t* pos; int pos_len;
  const float* rope_timescale;               /* (D/2) f32: max_wavelength ** (2i/D) */
  mpx_bf16* q; long long q_row_stride;       /* (M, Nq*D) with row pitch q_row_stride elements */
  mpx_bf16* k; long long k_row_stride;       /* elements between consecutive rows of k */
  mpx_bf16* v; long long v_row_stride;
  int cache_S;
  int M, H, Nq, Nkv, D;
  const float* norm_scale; float norm_eps;   /* optional fused pre-norm */
} mpx_qkv_rope_args;
int mpx_qkv_rope(const mpx_qkv_rope_args* a, mpx_stream_t stream);

/* GLU up/gate projections + activation: GLUBlock.__call__ feed_forward.py:73-77:
 * h = bf16( bf16(gelu(bf16(x@Wup))) * bf16(x@Wgate) ).  wt_ug is (2F,K) packed in 256-row tiles:
 * rows [256j, 256j+128) = Wup^T rows 128j.., rows [256j+128, 256j+256) = Wgate^T rows 128j.. (F % 128 == 0).
 * u_save/g_save (nullable, (M,F)) keep the bf16 pre-activations for the backward pass. */
typedef struct {
  const mpx_bf16* x; long long ldx;   /* (M,K) */
  const mpx_bf16* wt_ug;              /* (2F,K) tile-interleaved */
  mpx_bf16* h;                        /* (M,F) */
  mpx_bf16* u_save; mpx_bf16* g_save; /* (M,F) or NULL */
  int M, F, K;
} mpx_glu_up_args;
int mpx_glu_up(const mpx_glu_up_args* a, mpx_stream_t stream);

/* Whole GLU block + residual in one kernel (feed_forward.py:73-78 and the residual add of network.py:167-170):
 * out = bf16(residual + bf16(h @ Wdown)), h = bf16(bf16(gelu(bf16(x@Wup))) * bf16(x@Wgate)); h stays on chip.
 * wt_ug64 is (2F,H) packed in 128-row tiles [64 up rows | 64 gate rows]; wt_d is Wdown^T (H,F).  H must be 128.
 * norm_scale (H) f32, when not NULL, applies RMSNorm(x; scale, eps) to the rows first (network.py:167: ffn_norm), so
 * the caller passes the un-normalised residual stream as both x and residual.
 * wt_o (H,H) = Wo^T, when not NULL (requires norm_scale), folds the attention output projection and its residual add
 * (network.py:163-165) in front: x is then the attention output and the kernel computes
 * x1 = residual + x @ Wo, out = x1 + GLU(RMSNorm(x1)); x1 never leaves shared memory.
 * h_out (M,F), when not NULL, also receives h (bulk stores from the operand tiles): the training forward keeps it for
 * dWdown = h^T dy, while u and g are recomputed by mpx_glu_bwd_fused. */
int mpx_glu_fused(const mpx_bf16* x, const mpx_bf16* wt_ug64, const mpx_bf16* wt_d, const mpx_bf16* residual,
                  mpx_bf16* out, int M, int H, int F, const float* norm_scale, float eps, const mpx_bf16* wt_o,
                  mpx_bf16* h_out, mpx_stream_t stream);

/* The same block for the LAST layer of a rollout step with the policy head folded in (network.py:321-333 + sample_actions,
 * network.py:345-351 / train.py:97-98): after out = x1 + GLU(RMSNorm(x1)) the kernel applies the output RMSNorm to the
 * finished rows, forms the tied fp32 action logits (masked), draws the Gumbel arg-max action (Philox counters as
 * mpx_gumbel_noise / mpx_policy_head_sample) and its log-probability - bit-identical to mpx_glu_fused followed by
 * mpx_policy_head_sample, one launch less per step.  Needs wt_o (the projection prologue); the fused epilogue runs when
 * the launch is a cluster of four (M <= 32 * 148 / 4 rows per SM wave) and A <= 8, otherwise the call issues the two
 * kernels itself.  logits_out (M,A), xf_out (M,H), action_mask (M,A), gumbel_in (M,A), stream_offset may be NULL. */
typedef struct {
  const float* norm_scale; float eps;             /* output_norm.scale (H) */
  const float* table; int A;                      /* action_embedder.embedding_table (A,H) f32 */
  const uint8_t* action_mask;
  const float* gumbel_in;
  unsigned long long seed, stream_id;
  const unsigned long long* stream_offset;
  float* logits_out;
  mpx_bf16* xf_out;
  int32_t* action;                                /* (M) */
  float* log_prob;                                /* (M) */
} mpx_policy_head_args;
int mpx_glu_fused_head(const mpx_bf16* x, const mpx_bf16* wt_ug64, const mpx_bf16* wt_d, const mpx_bf16* residual,
                       mpx_bf16* out, int M, int H, int F, const float* norm_scale, float eps, const mpx_bf16* wt_o,
                       const mpx_policy_head_args* head, mpx_stream_t stream);

/* Backward of the same block in one kernel (jax.grad of feed_forward.py:73-78 inside train.py:180-230), for a forward
 * that kept only xn = RMSNorm(x1): u, g, a = gelu(u) and dh = dy @ Wdown^T are recomputed on chip and never stored.
 * Outputs dug (M,2F) = [du | dg], the operand of dWup/dWgate = xn^T dug, and dxn (M,H) = du @ Wup^T + dg @ Wgate^T
 * (h for dWdown = h^T dy comes from the forward's h_out).  Same weight layouts as mpx_glu_fused. */
int mpx_glu_bwd_fused(const mpx_bf16* xn, const mpx_bf16* dy, const mpx_bf16* wt_ug64, const mpx_bf16* wt_d,
                      mpx_bf16* dug, mpx_bf16* dxn, int M, int H, int F, mpx_stream_t stream);

/* Backward of the value path in one kernel (network.py:335-341 + values.py:34,40-41 under nnx.grad, train.py:236-238):
 * given xf (M,H) = the normalised trunk output and dvl (M,64) bf16 = d loss / d value logits (columns >= NL zero), it
 * recomputes z = bf16(bf16(xf @ Wvm) + b_vm) on chip and writes pv (M,Vh) = bf16(gelu(z)) and dz (M,Vh) =
 * bf16(bf16(dvl @ Wvh^T) * gelu'(z)) (operands of dWvh = pv^T dvl and dWvm = xf^T dz), dxf (M,H) = bf16(dz @ Wvm^T),
 * and db_vm (Vh) fp32 += column sums of dz (fixed order).  The forward therefore keeps nothing of the value MLP
 * (mpx_value_head_fused with vlogits).  wt_vm (Vh,H) = Wvm^T and w_vh_pad (Vh,64) = Wvh zero-padded to 64 columns, bf16
 * (mpx_prep_weights).  H must be 128, Vh % 64 == 0, Vh <= 1024.  workspace: mpx_value_bwd_fused_workspace_bytes(Vh). */
size_t mpx_value_bwd_fused_workspace_bytes(int Vh);
int mpx_value_bwd_fused(const mpx_bf16* xf, const mpx_bf16* dvl, const mpx_bf16* wt_vm, const mpx_bf16* b_vm,
                        const mpx_bf16* w_vh_pad, mpx_bf16* pv, mpx_bf16* dz, mpx_bf16* dxf, float* db_vm, int M, int H,
                        int Vh, void* workspace, size_t workspace_bytes, mpx_stream_t stream);

/* Whole value path of a rollout step in one kernel (network.py:321,335-341 + values.py:40-45):
 * value = sum softmax(f32(pv) @ Wvh + b_vh) * centers, pv = bf16(gelu(bf16(bf16(xf @ Wvm) + b_vm))),
 * xf = RMSNorm(x; norm_scale, eps) when norm_scale != NULL, else xf = x (rows already normalised).
 * wt_vm is Wvm^T (Vh,H) bf16, wt_vh_hilo the (128,Vh) bf16 [hi rows 0..63 | lo rows 64..127] split of Wvh^T
 * (see mpx_prep_weights), b_vm (Vh) bf16, b_vh/centers (NL) f32.  value (M) f32; vlogits (M,NL) f32 is optional.
 * H must be 128, Vh a multiple of 64, NL <= 64. */
int mpx_value_head_fused(const mpx_bf16* x, const float* norm_scale, float eps, const mpx_bf16* wt_vm,
                         const mpx_bf16* b_vm, const mpx_bf16* wt_vh_hilo, const float* b_vh, const float* centers,
                         float* value, float* vlogits, int M, int H, int Vh, int NL, mpx_stream_t stream);

/* ---------------------------------------------------------------------------------------------- decode attention
 * Decode branch of AttentionBlock.__call__, attention.py:136-150 (jax.nn.dot_product_attention with
 * key_value_seq_lengths): one RoPE'd query token per agent against that agent's KV ring cache (already holding
 * the new token, see mpx_qkv_rope).  kv_len = min(pos[0]+1, S) for every row.  head_dim must be 32. */
typedef struct {
  const mpx_bf16* q;        /* (B, Nq*D) */
  const mpx_bf16* kcache;   /* (B, S, Nkv, D) */
  const mpx_bf16* vcache;   /* (B, S, Nkv, D) */
  const int32_t* pos;       /* (B) int32 on device; element 0 is the shared step index */
  mpx_bf16* out;            /* (B, Nq*D) */
  int B, S, Nq, Nkv, D;
} mpx_attn_decode_args;
int mpx_attn_decode(const mpx_attn_decode_args* a, mpx_stream_t stream);

/* Attention half of a block for one rollout step in one kernel (network.py:158-165, attention.py:96-150):
 * history_norm -> q/k/v projection -> RoPE -> ring write at pos[0] % S -> the attention above, with the projection
 * running under the start of the cache stream.  Replaces mpx_rmsnorm_fwd + mpx_qkv_rope + mpx_attn_decode for
 * H in {128, 256}, Nq == 4, Nkv == 1, D == 32 (MPX_ERR_INVALID otherwise).  x (B,H) is the un-normalised residual
 * stream (8-byte aligned, 16 at H == 256); wt_qkv ((Nq+2*Nkv)*D, H) as in mpx_qkv_rope, 16-byte aligned; kcache/vcache
 * (B,S,1,D) are updated in place; out (B, Nq*D). */
typedef struct {
  const mpx_bf16* x; const float* norm_scale; float eps;
  const mpx_bf16* wt_qkv;
  const int32_t* pos; const float* rope_timescale;
  mpx_bf16* kcache; mpx_bf16* vcache;
  mpx_bf16* out;
  int B, S, H, Nq, Nkv, D;
  int flags;   /* bit 0: every ring slot except pos[0] % S was written by kernels that had COMPLETED before this launch
                * could begin (in a rollout: earlier decode steps, with at least one fully serialised launch in between),
                * so under programmatic dependent launch their stream may start before griddepcontrol.wait.  0 = safe. */
} mpx_attn_decode_fused_args;
int mpx_attn_decode_fused(const mpx_attn_decode_fused_args* a, mpx_stream_t stream);

/* ---------------------------------------------------------------------------------------------- row-wise ops
 * flax nnx.RMSNorm (eps 1e-6) as used by TransformerBlock (network.py:158-172) and output_norm (:321):
 * y = bf16(x32 * (rsqrt(mean(x32^2)+eps) * scale)).  x,y (M,H) bf16, scale (H) f32; H in {128,256,384,512}. */
int mpx_rmsnorm_fwd(const mpx_bf16* x, const float* scale, float eps, mpx_bf16* y, long long M, int H,
                    mpx_stream_t stream);
/* Its transpose: dx = bf16(d/dx) [+ dres, bf16 add: the cotangent of the residual branch]; dscale (H) f32 is
 * accumulated atomically (nullable). */
int mpx_rmsnorm_bwd(const mpx_bf16* dy, const mpx_bf16* x, const float* scale, float eps, const mpx_bf16* dres,
                    mpx_bf16* dx, float* dscale, long long M, int H, mpx_stream_t stream);
/* Transpose of the GLU activation (feed_forward.py:75-77): dug (M,2F) = [du | dg] from dh, u, g (M,F). */
int mpx_glu_bwd(const mpx_bf16* dh, const mpx_bf16* u, const mpx_bf16* g, mpx_bf16* dug, long long M, int F,
                mpx_stream_t stream);
/* dz = bf16(dy * gelu'(z)) (value_mlp / conv activations, network.py:338, observation.py:93). */
int mpx_gelu_bwd(const mpx_bf16* dy, const mpx_bf16* z, mpx_bf16* dz, long long n, mpx_stream_t stream);
/* apply_rope (positional_embeddings.py:7-27) in place on heads [0,nheads) of x (M,ld); inverse=1 applies the
 * transpose (used for dq, dk).  Row r uses pos[r % pos_len]. */
int mpx_rope(mpx_bf16* x, long long ld, const int32_t* pos, int pos_len, const float* rope_timescale, long long M,
             int nheads, int head_dim, int inverse, mpx_stream_t stream);
/* out (N) f32 += column sums of x (M,ld) bf16 (bias gradients). */
int mpx_colsum(const mpx_bf16* x, long long ld, long long M, int N, float* out, mpx_stream_t stream);
/* out = bf16(a + b) (sum of two bf16 cotangents). */
int mpx_add_bf16(const mpx_bf16* a, const mpx_bf16* b, mpx_bf16* out, long long n, mpx_stream_t stream);

/* ---------------------------------------------------------------------------------------------- training attention
 * Training branch of AttentionBlock.__call__, attention.py:151-169 (causal jax.nn.dot_product_attention) over
 * full trajectory windows.  q (B*T, Nq*D), k/v (B*T, Nkv*D) bf16 already rotated (mpx_qkv_rope), given with row
 * strides so they may be column slices of one (B*T, (Nq+2Nkv)*D) buffer.  o (B*T, Nq*D) bf16;
 * lse (B*T, Nq) f32 is saved for the backward pass (log2 units of the scaled scores).  T % 64 == 0, D == 32. */
typedef struct {
  const mpx_bf16* q; long long ldq;
  const mpx_bf16* k; long long ldk;
  const mpx_bf16* v; long long ldv;
  mpx_bf16* o;       long long ldo;
  float* lse;
  int B, T, Nq, Nkv, D;
} mpx_attn_train_args;
int mpx_attn_train_fwd(const mpx_attn_train_args* a, mpx_stream_t stream);

/* Its transpose (what nnx.grad, train.py:236-238, derives): given dout = d loss / d o, writes the cotangents of the
 * *un-rotated* projections (inverse RoPE applied): dq (B*T, Nq*D), dk, dv (B*T, Nkv*D) bf16 with row strides.
 * delta (B*T, Nq) f32 and dq_acc (B*T, Nq*D) f32 are caller-provided scratch. */
typedef struct {
  const mpx_bf16* q; long long ldq;
  const mpx_bf16* k; long long ldk;
  const mpx_bf16* v; long long ldv;
  const mpx_bf16* o; long long ldo;
  const mpx_bf16* dout; long long lddo;
  const float* lse;
  float* delta;
  float* dq_acc;
  mpx_bf16* dq; long long lddq;
  mpx_bf16* dk; long long lddk;
  mpx_bf16* dv; long long lddv;
  const int32_t* pos; int pos_len;
  const float* rope_timescale;
  int B, T, Nq, Nkv, D;
} mpx_attn_train_bwd_args;
int mpx_attn_train_bwd(const mpx_attn_train_bwd_args* a, mpx_stream_t stream);

/* nnx.Linear with dtype=None on f32(bf16 activations) (the HL-Gauss head, values.py:34/40-41): y (M,N) f32 =
 * f32(x) @ W + bias with the fp32 weight carried as a bf16 hi/lo pair (wt_hilo (128,K): rows [0,64) = bf16(W^T),
 * rows [64,128) = bf16(W^T - hi); rows >= N of each half zero).  N <= 64. */
int mpx_linear_f32w(const mpx_bf16* x, long long ldx, const mpx_bf16* wt_hilo, const float* bias, float* y,
                    long long ldy, int M, int N, int K, mpx_stream_t stream);

/* ---------------------------------------------------------------------------------------------- embeds and heads
 * network.py:303-309: x = obs_emb + reward_encoder(reward) + action_embedder.encode(last_action) [+ task embed],
 * every term and every sum rounded to bf16.  Tables/kernels are the fp32 parameters. task_ids/task_table nullable. */
int mpx_embed_combine(const mpx_bf16* obs_emb, const float* reward, const int32_t* last_action,
                      const int32_t* task_ids, const float* w_reward, const float* b_reward,
                      const float* action_table, int A, const float* task_table, int n_tasks, mpx_bf16* x,
                      long long M, int H, mpx_stream_t stream);
/* Its transpose: accumulates (+=) into the fp32 gradients of reward_encoder.{kernel,bias} and the embedding tables. */
int mpx_embed_bwd(const mpx_bf16* dx, const float* reward, const int32_t* last_action, const int32_t* task_ids,
                  int A, int n_tasks, float* d_w_reward, float* d_b_reward, float* d_action_table,
                  float* d_task_table, long long M, int H, mpx_stream_t stream);
/* Rollout-step observation encoder + input embedding in one kernel (observation.py:84-96 + network.py:303-309) for the
 * standard geometry: (M,11,11,K) uint8 view, one-hot sizes[K] with prod(sizes[k]+1) <= 48, conv 3x3/2 -> 16, gelu,
 * conv 3x3/1 -> 32, gelu, conv 3x3/1 -> H = 128, then x = obs_emb + reward_encoder(reward) + embed(last_action)
 * (+ embed(task)).  The kernel reads the encoder weights from an IMAGE (mpx_obs_fused_image_bytes() bytes, 16-byte
 * aligned) that mpx_obs_fused_prep builds - once per weight update - from w1 (9*C,16) bf16 rows = (tap, one-hot
 * channel), wt2 (32,144) and wt3 (128,288) (the transposed kernels, K index = tap*cin + c, as mpx_prep_weights lays
 * them out for mpx_linear) and the bf16 biases.  Embedding tables are fp32 as in mpx_embed_combine.  obs_emb (M,H) is
 * an optional copy of the encoder output before the embedding sum. */
size_t mpx_obs_fused_image_bytes(void);
int mpx_obs_fused_prep(const int* sizes, int K, const mpx_bf16* w1, const mpx_bf16* b1, const mpx_bf16* wt2,
                       const mpx_bf16* b2, const mpx_bf16* wt3, const mpx_bf16* b3, void* image, mpx_stream_t stream);
int mpx_obs_embed_fused(const uint8_t* obs, const int* sizes, int K, const void* image, const float* reward,
                        const int32_t* last_action, const int32_t* task_ids, const float* w_reward, const float* b_reward,
                        const float* action_table, int A, const float* task_table, int n_tasks, mpx_bf16* x,
                        mpx_bf16* obs_emb, long long M, int IH, int IW, int H, mpx_stream_t stream);
/* Training-side observation encoder (observation.py:84-96 + network.py:303-309 as nnx.grad(ppo_loss) evaluates and
 * differentiates them, train.py:180-192,236-238) for the geometry, image and embedding arguments of
 * mpx_obs_embed_fused; one tile = 128 tokens, no im2col matrix exists anywhere.
 *   mpx_obs_train_supported(sizes, K) = 1 when the joint table of these one-hot sizes fits next to the operand tiles in
 *     shared memory (prod(sizes[k]+1) <= 39); mpx_obs_train_tiles(M) = ceil(M / 128).
 *   mpx_obs_train_fwd: x (M,H) [+ obs_emb (M,H), nullable] and what the backward needs: a2 (M,288) = gelu(z2) row-major
 *     (operand of the conv3 weight gradient) and, in layouts private to these kernels (TILE layouts, buffers sized for
 *     mpx_obs_train_tiles(M) whole tiles): z1t / a1t [tile][25 positions][2][128 rows][8 ch] = the bf16 pre-activation
 *     (+bias) of conv1 and its gelu, z2t [tile][9][4][128][8] = the pre-activation of conv2.
 *   mpx_obs_train_bwd: data gradients given dx0 (M,H) = d loss / d (encoder output): dz2t (tile layout, operand of
 *     mpx_obs_conv2_wgrad), dz1 (M*25,16) row-major (operand of mpx_conv1_onehot_wgrad); db1 (16) / db2 (32) fp32 +=
 *     the conv1 / conv2 bias gradients (fixed summation order).  The conv3 weight / bias gradients are
 *     mpx_linear_wgrad(a2, dx0) / mpx_colsum(dx0).
 *   mpx_obs_conv2_wgrad: dw (144,32) fp32 += sum_p a1[p + tap]^T dz2[p] (kernel layout (kh,kw,cin,cout) flattened).
 * workspace: mpx_obs_train_bwd_workspace_bytes() bytes, shared by the two backward calls of a stream. */
int mpx_obs_train_supported(const int* sizes, int K);
long long mpx_obs_train_tiles(long long M);
int mpx_obs_train_fwd(const uint8_t* obs, const int* sizes, int K, const void* image, const float* reward,
                      const int32_t* last_action, const int32_t* task_ids, const float* w_reward, const float* b_reward,
                      const float* action_table, int A, const float* task_table, int n_tasks, mpx_bf16* z1t, mpx_bf16* a1t,
                      mpx_bf16* z2t, mpx_bf16* a2, mpx_bf16* x, mpx_bf16* obs_emb, long long M, int IH, int IW, int H,
                      mpx_stream_t stream);
size_t mpx_obs_train_bwd_workspace_bytes(void);
int mpx_obs_train_bwd(const mpx_bf16* dx0, const mpx_bf16* z1t, const mpx_bf16* z2t, const void* image, mpx_bf16* dz2t,
                      mpx_bf16* dz1, float* db1, float* db2, long long M, void* workspace, size_t workspace_bytes,
                      mpx_stream_t stream);
int mpx_obs_conv2_wgrad(const mpx_bf16* a1t, const mpx_bf16* dz2t, float* dw, long long M, void* workspace,
                        size_t workspace_bytes, mpx_stream_t stream);
/* Weight gradient of the first observation conv without materialising the one-hot im2col matrix:
 * dw ((kh*kw*C), cout) f32 += onehot_im2col(obs)^T @ dz, dz (M*OH*OW, cout) bf16 contiguous.  The one-hot operand
 * tiles are generated in shared memory from the uint8 observation.  kh*kw*C <= 128, kh*kw*K <= 32, cout <= 64. */
size_t mpx_conv1_onehot_wgrad_workspace_bytes(long long M, int IH, int IW, int kh, int kw, int sh, int sw);
int mpx_conv1_onehot_wgrad(const uint8_t* obs, const mpx_bf16* dz, float* dw, long long M, int IH, int IW, int K,
                           const int* sizes, int kh, int kw, int sh, int sw, int cout, void* workspace,
                           size_t workspace_bytes, mpx_stream_t stream);
/* FlattenedObsEncoder (observation.py:98-145), first half: one_hot(obs) -> Linear(C,4) per cell as a table lookup.
 * obs (M, cells) uint8, w (C,4) / bias (4) fp32 master params, out (M, KP) bf16 with KP = round_up(4*cells, 8)
 * (pad columns zeroed); the dense half is an mpx_linear on out.  mpx_flat_embed_bwd is the transpose:
 * dw (C,4) += per-class sums of dflat (M,KP), dbias (4) += column sums. */
int mpx_flat_embed_fwd(const uint8_t* obs, const float* w, const float* bias, mpx_bf16* out, long long M, int cells, int C,
                       int KP, mpx_stream_t stream);
int mpx_flat_embed_bwd(const uint8_t* obs, const mpx_bf16* dflat, float* dw, float* dbias, long long M, int cells, int C,
                       int KP, mpx_stream_t stream);
/* network.py:323-331: tied action logits f32(x) @ table^T (M,A) f32, masked entries = finfo(f32).min. A <= 32. */
int mpx_policy_logits(const mpx_bf16* x, const float* table, const uint8_t* action_mask, float* logits, long long M,
                      int H, int A, mpx_stream_t stream);
/* distrax.Categorical(logits).sample / log_prob (train.py:97-98) as Gumbel arg-max with the noise as an input:
 * action = argmax(logits + gumbel) (first maximum), log_prob = log_softmax(logits)[action]. */
int mpx_policy_sample(const float* logits, const float* gumbel, int32_t* action, float* log_prob, long long M, int A,
                      mpx_stream_t stream);
/* The rollout step's whole policy path in one kernel (network.py:321-333, train.py:97-98): xf = RMSNorm(x) when
 * norm_scale != NULL (else x is used as is), logits = f32(xf) @ table^T, mask, action = argmax(logits + gumbel),
 * log_prob = log_softmax(logits)[action].  gumbel_in (M,A) overrides the built-in Philox draw (which is the one
 * mpx_gumbel_noise(seed, stream_id + *stream_offset) produces for an (M,A) buffer).  logits_out (M,A) and xf_out (M,H)
 * are optional outputs. */
int mpx_policy_head_sample(const mpx_bf16* x, const float* norm_scale, float eps, const float* table,
                           const uint8_t* action_mask, const float* gumbel_in, unsigned long long seed,
                           unsigned long long stream_id, const unsigned long long* stream_offset, float* logits_out,
                           mpx_bf16* xf_out, int32_t* action, float* log_prob, long long M, int H, int A,
                           mpx_stream_t stream);
/* Counter-based Philox-4x32-10 Gumbel(0,1) noise (the reference draws it from JAX's threefry; the noise source is
 * outside the parity contract, the arg-max is inside it). */
int mpx_gumbel_noise(float* out, long long n, unsigned long long seed, unsigned long long stream_id,
                     const unsigned long long* stream_offset /* device, nullable: added to stream_id */,
                     mpx_stream_t stream);
/* Policy part of ppo_loss, train.py:193-212, and its gradient: losses[0] = actor_loss, losses[1] = entropy_loss
 * (means over M tokens); dlogits (M,A) (nullable) = d(actor_loss + entropy_coef*entropy_loss)/dlogits. */
size_t mpx_ppo_policy_loss_workspace_bytes(long long M);
int mpx_ppo_policy_loss(const float* logits, const int32_t* actions, const float* old_log_prob,
                        const float* advantages, float clip_eps, float entropy_coef,
                        const float* entropy_coef_dev /* device scalar, nullable: overrides entropy_coef */,
                        float* losses, float* dlogits, void* workspace, long long M, int A, mpx_stream_t stream);
/* Transpose of mpx_policy_logits: dx = bf16(dlogits @ table) [+ dx_other, bf16 add]; dtable += dlogits^T @ f32(x). */
int mpx_policy_head_bwd(const float* dlogits, const mpx_bf16* x, const float* table, const mpx_bf16* dx_other,
                        mpx_bf16* dx, float* dtable, long long M, int H, int A, mpx_stream_t stream);

/* ---------------------------------------------------------------------------------------------- obs encoder support
 * GridCnnObsEncoder (observation.py:84-96) lowered to GEMMs.  obs (M,IH,IW,K) u8 -> one-hot im2col rows
 * (M*OH*OW, KP) bf16, column (ky*kw+kx)*C + channel, zero padded to KP (multiple of 8). */
int mpx_obs_onehot_im2col(const uint8_t* obs, mpx_bf16* out, long long M, int IH, int IW, int K,
                          const int* onehot_sizes /* host array [K] */, int kh, int kw, int sh, int sw, int KP,
                          mpx_stream_t stream);
/* First conv layer evaluated directly on the uint8 observation (gather-sum of kernel rows selected by the one-hot,
 * fp32 accumulate == the GEMM on the one-hot rows): a_out (M*OH*OW, cout) = [gelu](bf16(conv) + bias), z_out
 * (nullable) = the bf16 pre-activation.  w is the bf16 kernel as stored, (kh*kw*C, cout); cout % 8 == 0. */
int mpx_conv1_onehot_fwd(const uint8_t* obs, const mpx_bf16* w, const mpx_bf16* bias, mpx_bf16* z_out, mpx_bf16* a_out,
                         long long M, int IH, int IW, int K, const int* onehot_sizes, int kh, int kw, int sh, int sw,
                         int cout, int apply_gelu, mpx_stream_t stream);
/* NHWC im2col / its transpose for the hidden conv layers (C % 8 == 0, VALID padding).  mpx_col2im's optional z
 * (M,IH,IW,C) folds the activation transpose of the layer below into the same pass: da = bf16(bf16(sum) * gelu'(z)). */
int mpx_im2col(const mpx_bf16* a, mpx_bf16* x, long long M, int IH, int IW, int C, int kh, int kw, int sh, int sw,
               mpx_stream_t stream);
int mpx_col2im(const mpx_bf16* dx, mpx_bf16* da, long long M, int IH, int IW, int C, int kh, int kw, int sh, int sw,
               const mpx_bf16* z, mpx_stream_t stream);

/* ---------------------------------------------------------------------------------------------- optimizer
 * optimizer.py:12-23: optax.adamw(linear_schedule(lr0, 0, total_steps)) then clip_by_global_norm(max_norm) on the
 * update, over flat fp32 buffers of n elements.  `step` is a device int32 counter (read, then incremented).
 * grads are multiplied by grad_scale first (1/world_size after a sum all-reduce). */
size_t mpx_adamw_workspace_bytes(long long n);
int mpx_adamw_step(float* params, const float* grads, float* mu, float* nu, float* update_scratch, long long n,
                   int* step, float lr0, float total_steps, float b1, float b2, float eps, float weight_decay,
                   float max_norm, float grad_scale, float* update_norm_out, void* workspace, mpx_stream_t stream);
/* optimizer.py:25-37: optax.contrib.muon (nesterov momentum beta, 5 Newton-Schulz steps in fp32, sqrt(max(1,out/in))
 * scaling, weight decay, linear LR schedule) for the 2-D matrices described by `mats`, optax.adamw(nesterov=True)
 * for elements [0, n_adam) of the flat buffers, then clip_by_global_norm(max_norm) on the combined update.
 * mats: DEVICE array of nmat records {int64 off; int32 rows, cols, r, c, transposed, pad; int64 xoff, aoff}
 * (r <= c is the orientation Newton-Schulz runs in; xoff/aoff are offsets into the workspaces).
 * ns_ws: fp32 scratch of 2*x_total + 2*a_total + nmat floats (x_total = sum r*c, a_total = sum r*r). */
size_t mpx_muon_workspace_bytes(long long n, int nmat);
int mpx_muon_step(float* params, const float* grads, float* mu, float* nu, float* update_scratch, long long n,
                  long long n_adam, const void* mats, int nmat, long long x_total, long long a_total, int max_r,
                  int max_c, float* ns_ws, int* step, float lr0, float total_steps, float adam_b1, float adam_b2,
                  float eps, float weight_decay, float muon_beta, int ns_steps, float max_norm, float grad_scale,
                  float* update_norm_out, void* workspace, mpx_stream_t stream);
/* The same step in two enqueue phases (phases = 1 | 2; 3 = mpx_muon_step): 1 = the Muon-labelled matrices only
 * (momentum, Newton-Schulz, their update and squared-norm partials - touches neither params nor the step counter),
 * 2 = the AdamW-labelled leaves, the global-norm clip over both partitions, the parameter write, step += 1.  The
 * matrices' gradients (layer weights, optimizer.py:51-69) are complete before the observation-encoder backward
 * starts, so a caller runs phase 1 (after its own all-reduce of grads[n_adam:]) on a second stream under that
 * backward and joins before phase 2.  Same arguments in both calls. */
int mpx_muon_step_phased(float* params, const float* grads, float* mu, float* nu, float* update_scratch, long long n,
                         long long n_adam, const void* mats, int nmat, long long x_total, long long a_total,
                         int max_r, int max_c, float* ns_ws, int* step, float lr0, float total_steps, float adam_b1,
                         float adam_b2, float eps, float weight_decay, float muon_beta, int ns_steps, float max_norm,
                         float grad_scale, float* update_norm_out, void* workspace, int phases, mpx_stream_t stream);
/* Sharded Muon (reduce-scatter -> Newton-Schulz on 1/world of the matrices -> all-gather; new: the reference is single
 * device): after phase 1 on the LOCAL matrices (mpx_muon_step_phased, phases = 1, local table) and an all-gather of
 * update_scratch[n_adam:], this call runs the AdamW-labelled leaves [0, n_adam), the global-norm clip over the whole
 * update (a fixed-order sum of the gathered data: bit-identical on every rank), the parameter write and step += 1.
 * workspace >= mpx_muon_workspace_bytes(n, 1). */
int mpx_muon_apply_sharded(float* params, const float* grads, float* mu, float* nu, float* update_scratch, long long n,
                           long long n_adam, int* step, float lr0, float total_steps, float adam_b1, float adam_b2,
                           float eps, float weight_decay, float max_norm, float grad_scale, float* update_norm_out,
                           void* workspace, mpx_stream_t stream);
/* fp32 master weights -> bf16 GEMM operand copies. descs is a DEVICE array; each entry casts (and optionally
 * transposes) a (rows, cols) fp32 block with pitch src_ld into a bf16 destination with pitch dst_ld. */
typedef struct {
  const float* src; long long src_ld;
  mpx_bf16* dst;    long long dst_ld;
  int rows, cols;
  int transpose;   /* bit 0: transpose; bit 1: store bf16(w - f32(bf16(w))) (lo half of a hi/lo split) */
  int _pad;
} mpx_prep_desc;
int mpx_prep_weights(const mpx_prep_desc* descs_device, int ndesc, mpx_stream_t stream);
/* Rollout buffers (rollout.py:11-30) are written time-major (T,B,item) by the decode loop; this produces the
 * batch-major rows training reads, applying the row permutation of rollout.py:169-181 (perm nullable):
 * dst[b][t] = src[t][perm[b]]. */
int mpx_transpose_tb(const void* src, void* dst, const int32_t* perm, int T, int B, int item_bytes,
                     mpx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAPOX_B200_H_ */
