// XLA typed-FFI adapter: exposes the C ABI of include/mapox_b200.h as jax.ffi custom-call targets - one handler per
// exported compute entry point (tests/test_xla_ffi_source.py checks that list against the header).
//
// NOT BUILT BY DEFAULT and UNVERIFIED against a real jaxlib in this environment: jaxlib (and therefore
// xla/ffi/api/ffi.h) is not installed in the build image or on the GPU box.  What IS checked here: the file compiles
// with `g++ -fsyntax-only` against tests/xla_ffi_stub/ (a minimal stand-in for the Bind()/Buffer API that also
// static_asserts every handler signature against its binding), i.e. every call below type-checks against the real
// prototypes of include/mapox_b200.h.  Build, where JAX exists, with
//   g++ -O2 -std=c++17 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -I../../../include \
//       xla_ffi.cc -L../.. -lmapox_b200 -Wl,-rpath,'$ORIGIN' -o ../../libmapox_b200_xla.so
// and register from Python with jaxrl_b200/jax_ffi.py.
//
// Conventions: XLA owns every buffer; operands and results are dense row-major, so every leading
// dimension below is the natural one; scalars arrive as attributes; scratch memory is an extra U8 result sized by the
// Python side with the library's *_workspace_bytes functions; in-place entry points (KV rings, accumulated weight
// gradients, optimizer state) take the buffer as an operand that jax.ffi.ffi_call aliases to the result
// (`input_output_aliases`, listed next to each handler as "aliases: operand -> result").  Handlers are stateless,
// enqueue on XLA's stream only, never synchronise, and are command-buffer compatible so XLA can capture them into CUDA
// graphs (the rollout fori_loop body, mapox_trainer/train.py:128-140).
#include <cstdint>
#include <string>

#include "mapox_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using Stream = ffi::PlatformStream<mpx_stream_t>;
using BF = ffi::Buffer<ffi::BF16>;
using F32 = ffi::Buffer<ffi::F32>;
using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
using U8 = ffi::Buffer<ffi::U8>;
using U64 = ffi::Buffer<ffi::U64>;
using PRED = ffi::Buffer<ffi::PRED>;
using RBF = ffi::ResultBuffer<ffi::BF16>;
using RF32 = ffi::ResultBuffer<ffi::F32>;
using RF64 = ffi::ResultBuffer<ffi::F64>;
using RS32 = ffi::ResultBuffer<ffi::S32>;
using RU8 = ffi::ResultBuffer<ffi::U8>;

#define MPX_TRAITS {xla::ffi::Traits::kCmdBufferCompatible}

static ffi::Error Fail(const char* what) {
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + mpx_last_error());
}
#define MPX_CALL(expr, name) \
  do {                       \
    if ((expr) != 0) return Fail(name); \
    return ffi::Error::Success();       \
  } while (0)

template <typename B> static const mpx_bf16* bf(B& b) { return reinterpret_cast<const mpx_bf16*>(b.untyped_data()); }
template <typename R> static mpx_bf16* rbf(R& r) { return reinterpret_cast<mpx_bf16*>(r->untyped_data()); }
template <typename B> static const uint8_t* u8(B& b) { return reinterpret_cast<const uint8_t*>(b.untyped_data()); }
template <typename B> static int last_dim(B& b) { auto d = b.dimensions(); return static_cast<int>(d[d.size() - 1]); }
template <typename B> static long long rows_of(B& b) { return static_cast<long long>(b.element_count()) / last_dim(b); }

// ================================================================================================ GAE
// Rollout.calculate_advantage (rollout.py:128-157): rewards (B,T) f32, values (B,T+1) f32, next_terminated (B,T) pred
static ffi::Error GaeImpl(mpx_stream_t stream, F32 rewards, F32 values, PRED next_terminated, double discount,
                          double gae_lambda, RF32 advantages, RF32 targets, RF64 stats, RU8 workspace) {
  const auto dims = rewards.dimensions();
  const int B = static_cast<int>(dims[0]), T = static_cast<int>(dims[1]);
  MPX_CALL(mpx_gae(rewards.typed_data(), values.typed_data(), u8(next_terminated), B, T, discount, gae_lambda,
                   advantages->typed_data(), targets->typed_data(), stats->typed_data(), workspace->untyped_data(), stream),
           "mpx_gae");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGae, GaeImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<PRED>()
                                  .Attr<double>("discount").Attr<double>("gae_lambda")
                                  .Ret<F32>().Ret<F32>().Ret<F64>().Ret<U8>(),
                              MPX_TRAITS);

// rollout.py:159-162.  aliases: advantages(0) -> 0
static ffi::Error AdvNormalizeImpl(mpx_stream_t stream, F32 adv_in, F64 stats, double count, RF32 adv) {
  (void)adv_in;
  MPX_CALL(mpx_adv_normalize(adv->typed_data(), static_cast<long long>(adv->element_count()), stats.typed_data(), count,
                             stream), "mpx_adv_normalize");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAdvNormalize, AdvNormalizeImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F64>().Attr<double>("count").Ret<F32>(), MPX_TRAITS);

// ================================================================================================ HL-Gauss
// HlGaussValue.get_value (values.py:43-45)
static ffi::Error HlGaussValueImpl(mpx_stream_t stream, F32 logits, F32 centers, RF32 value) {
  MPX_CALL(mpx_hlgauss_value(logits.typed_data(), rows_of(logits), last_dim(logits), centers.typed_data(),
                             value->typed_data(), stream), "mpx_hlgauss_value");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxHlGaussValue, HlGaussValueImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<F32>(), MPX_TRAITS);

// HlGaussValue.get_loss (values.py:47-63) forward + gradient in one pass (the custom_vjp keeps dlogits as residual).
// ld_bf16 = last dimension of dlogits_bf16 (0-sized buffer => not produced).
static ffi::Error HlGaussLossImpl(mpx_stream_t stream, F32 logits, F32 targets, F32 supports, float vmin, float vmax,
                                  float sigma, float grad_scale, RF32 loss, RF32 dlogits, RBF dlogits_bf16, RU8 workspace) {
  const int nl = last_dim(logits);
  const int ld = dlogits_bf16->element_count() ? last_dim(*dlogits_bf16) : 0;
  MPX_CALL(mpx_hlgauss_loss(logits.typed_data(), targets.typed_data(), rows_of(logits), nl, supports.typed_data(), vmin, vmax,
                            sigma, grad_scale, loss->typed_data(), dlogits->element_count() ? dlogits->typed_data() : nullptr,
                            ld ? rbf(dlogits_bf16) : nullptr, ld, workspace->untyped_data(), stream), "mpx_hlgauss_loss");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxHlGaussLoss, HlGaussLossImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<float>("vmin").Attr<float>("vmax").Attr<float>("sigma").Attr<float>("grad_scale")
                                  .Ret<F32>().Ret<F32>().Ret<BF>().Ret<U8>(),
                              MPX_TRAITS);

// ================================================================================================ dense
// nnx.Linear / LinearGeneral (attention.py:53-92, feed_forward.py:59-71, network.py:228-285) and, with the kernel as
// stored, its data gradient.  bias / residual / y_pre are optional: pass 0-element buffers to leave them out.
static ffi::Error LinearImpl(mpx_stream_t stream, BF x, BF wt, BF bias, BF residual, int32_t act, RBF y, RBF y_pre) {
  mpx_linear_args a{};
  a.K = last_dim(x); a.M = static_cast<int>(rows_of(x)); a.N = static_cast<int>(wt.dimensions()[0]);
  a.x = bf(x); a.ldx = a.K;
  a.wt = bf(wt); a.ldw = a.K;
  a.bias = bias.element_count() ? bf(bias) : nullptr;
  a.residual = residual.element_count() ? bf(residual) : nullptr; a.ldr = a.N;
  a.y = rbf(y); a.ldy = a.N;
  a.y_f32 = nullptr; a.ldy32 = a.N;
  a.act = act;
  a.y_pre = y_pre->element_count() ? rbf(y_pre) : nullptr;
  MPX_CALL(mpx_linear(&a, stream), "mpx_linear");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxLinear, LinearImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Attr<int32_t>("act")
                                  .Ret<BF>().Ret<BF>(),
                              MPX_TRAITS);

// dW (K,N) f32 += x^T dy.  aliases: dw_in(2) -> 0 (pass zeros for a fresh gradient)
static ffi::Error LinearWgradImpl(mpx_stream_t stream, BF x, BF dy, F32 dw_in, RF32 dw, RU8 workspace) {
  (void)dw_in;
  mpx_wgrad_args a{};
  a.K = last_dim(x); a.N = last_dim(dy); a.M = static_cast<int>(rows_of(x));
  a.x = bf(x); a.ldx = a.K;
  a.dy = bf(dy); a.ldy = a.N;
  a.dw = dw->typed_data(); a.lddw = a.N;
  a.workspace = workspace->untyped_data(); a.workspace_bytes = workspace->element_count();
  MPX_CALL(mpx_linear_wgrad(&a, stream), "mpx_linear_wgrad");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxLinearWgrad, LinearWgradImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<F32>().Ret<F32>().Ret<U8>(), MPX_TRAITS);

// fp32 Linear of the HL-Gauss head (values.py:34,40-41) with the weight as a bf16 hi/lo pair
static ffi::Error LinearF32wImpl(mpx_stream_t stream, BF x, BF wt_hilo, F32 bias, RF32 y) {
  const int K = last_dim(x), N = last_dim(*y);
  MPX_CALL(mpx_linear_f32w(bf(x), K, bf(wt_hilo), bias.typed_data(), y->typed_data(), N, static_cast<int>(rows_of(x)), N, K,
                           stream), "mpx_linear_f32w");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxLinearF32w, LinearF32wImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<F32>().Ret<F32>(), MPX_TRAITS);

// ================================================================================================ q/k/v + RoPE
// Decode: attention.py:121-134 + :108-114.  aliases: kcache_in(4) -> 1, vcache_in(5) -> 2 (the ring write is in place,
// like the donated dynamic_update_slice of the reference).  norm_scale: 0 elements => no fused pre-norm.
static ffi::Error QkvRopeDecodeImpl(mpx_stream_t stream, BF x, BF wt_qkv, S32 seq_pos, F32 timescale, BF kcache_in,
                                    BF vcache_in, F32 norm_scale, int32_t num_heads, float eps, RBF q, RBF kcache,
                                    RBF vcache) {
  (void)kcache_in; (void)vcache_in;
  const auto kd = kcache->dimensions();
  mpx_qkv_rope_args a{};
  a.M = static_cast<int>(kd[0]); a.H = last_dim(x);
  a.Nq = num_heads; a.Nkv = static_cast<int>(kd[2]); a.D = static_cast<int>(kd[3]);
  a.x = bf(x); a.ldx = a.H;
  a.wt_qkv = bf(wt_qkv);
  a.pos = seq_pos.typed_data(); a.pos_len = a.M;
  a.rope_timescale = timescale.typed_data();
  a.q = rbf(q); a.q_row_stride = static_cast<long long>(a.Nq) * a.D;
  a.k = rbf(kcache); a.v = rbf(vcache);
  a.cache_S = static_cast<int>(kd[1]);
  a.k_row_stride = a.v_row_stride = static_cast<long long>(a.cache_S) * a.Nkv * a.D;
  a.norm_scale = norm_scale.element_count() ? norm_scale.typed_data() : nullptr; a.norm_eps = eps;
  MPX_CALL(mpx_qkv_rope(&a, stream), "mpx_qkv_rope");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxQkvRopeDecode, QkvRopeDecodeImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<S32>().Arg<F32>().Arg<BF>().Arg<BF>()
                                  .Arg<F32>().Attr<int32_t>("num_heads").Attr<float>("eps").Ret<BF>().Ret<BF>().Ret<BF>(),
                              MPX_TRAITS);

// Training: q/k/v of a full window into ONE (B*T, (Nq+2Nkv)*D) buffer (the layout the attention kernels read).
static ffi::Error QkvRopeTrainImpl(mpx_stream_t stream, BF x, BF wt_qkv, S32 positions, F32 timescale, int32_t num_heads,
                                   int32_t num_kv_heads, int32_t head_dim, RBF qkv) {
  mpx_qkv_rope_args a{};
  a.H = last_dim(x); a.M = static_cast<int>(rows_of(x));
  a.Nq = num_heads; a.Nkv = num_kv_heads; a.D = head_dim;
  const long long ld = static_cast<long long>(a.Nq + 2 * a.Nkv) * a.D;
  a.x = bf(x); a.ldx = a.H;
  a.wt_qkv = bf(wt_qkv);
  a.pos = positions.typed_data(); a.pos_len = static_cast<int>(positions.element_count());
  a.rope_timescale = timescale.typed_data();
  a.q = rbf(qkv); a.k = a.q + static_cast<long long>(a.Nq) * a.D; a.v = a.k + static_cast<long long>(a.Nkv) * a.D;
  a.q_row_stride = a.k_row_stride = a.v_row_stride = ld;
  a.cache_S = 0; a.norm_scale = nullptr; a.norm_eps = 0.f;
  MPX_CALL(mpx_qkv_rope(&a, stream), "mpx_qkv_rope");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxQkvRopeTrain, QkvRopeTrainImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<S32>().Arg<F32>()
                                  .Attr<int32_t>("num_heads").Attr<int32_t>("num_kv_heads").Attr<int32_t>("head_dim").Ret<BF>(),
                              MPX_TRAITS);

// apply_rope (positional_embeddings.py:7-27) / its transpose, in place.  aliases: x_in(0) -> 0
static ffi::Error RopeImpl(mpx_stream_t stream, BF x_in, S32 positions, F32 timescale, int32_t num_heads, int32_t head_dim,
                           int32_t inverse, RBF x) {
  (void)x_in;
  const long long ld = static_cast<long long>(num_heads) * head_dim;
  MPX_CALL(mpx_rope(rbf(x), ld, positions.typed_data(), static_cast<int>(positions.element_count()), timescale.typed_data(),
                    static_cast<long long>(x->element_count()) / ld, num_heads, head_dim, inverse, stream), "mpx_rope");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxRope, RopeImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<S32>().Arg<F32>().Attr<int32_t>("num_heads")
                                  .Attr<int32_t>("head_dim").Attr<int32_t>("inverse").Ret<BF>(),
                              MPX_TRAITS);

// ================================================================================================ decode attention
// AttentionBlock.__call__ decode branch (attention.py:136-150): q (B,1,Nq,D), caches (B,S,Nkv,D), seq_pos (B,1).
static ffi::Error AttnDecodeImpl(mpx_stream_t stream, BF q, BF kcache, BF vcache, S32 seq_pos, RBF out) {
  const auto kd = kcache.dimensions();
  const auto qd = q.dimensions();
  mpx_attn_decode_args a{};
  a.q = bf(q); a.kcache = bf(kcache); a.vcache = bf(vcache);
  a.pos = seq_pos.typed_data();
  a.out = rbf(out);
  a.B = static_cast<int>(kd[0]); a.S = static_cast<int>(kd[1]); a.Nkv = static_cast<int>(kd[2]);
  a.D = static_cast<int>(kd[3]); a.Nq = static_cast<int>(qd[qd.size() - 2]);
  MPX_CALL(mpx_attn_decode(&a, stream), "mpx_attn_decode");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAttnDecode, AttnDecodeImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<S32>().Ret<BF>(), MPX_TRAITS);

// Attention half of a block for one rollout step (network.py:158-165 + attention.py:96-150): history_norm, q/k/v, RoPE,
// ring write, attention.  aliases: kcache_in(5) -> 1, vcache_in(6) -> 2
static ffi::Error AttnBlockDecodeImpl(mpx_stream_t stream, BF x, F32 norm_scale, BF wt_qkv, S32 seq_pos, F32 timescale,
                                      BF kcache_in, BF vcache_in, int32_t num_heads, float eps, RBF out, RBF kcache,
                                      RBF vcache) {
  (void)kcache_in; (void)vcache_in;
  const auto kd = kcache->dimensions();
  mpx_attn_decode_fused_args a{};
  a.x = bf(x);
  a.norm_scale = norm_scale.typed_data(); a.eps = eps;
  a.wt_qkv = bf(wt_qkv);
  a.pos = seq_pos.typed_data(); a.rope_timescale = timescale.typed_data();
  a.kcache = rbf(kcache); a.vcache = rbf(vcache);
  a.out = rbf(out);
  a.B = static_cast<int>(kd[0]); a.S = static_cast<int>(kd[1]); a.Nkv = static_cast<int>(kd[2]);
  a.D = static_cast<int>(kd[3]); a.Nq = num_heads; a.H = last_dim(x);
  a.flags = 0;
  MPX_CALL(mpx_attn_decode_fused(&a, stream), "mpx_attn_decode_fused");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAttnBlockDecode, AttnBlockDecodeImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Arg<BF>().Arg<S32>().Arg<F32>().Arg<BF>()
                                  .Arg<BF>().Attr<int32_t>("num_heads").Attr<float>("eps").Ret<BF>().Ret<BF>().Ret<BF>(),
                              MPX_TRAITS);

// ================================================================================================ training attention
// AttentionBlock.__call__ train branch (attention.py:151-169) on the packed qkv (B,T,(Nq+2Nkv)*D) of MpxQkvRopeTrain.
static ffi::Error AttnTrainFwdImpl(mpx_stream_t stream, BF qkv, int32_t num_heads, int32_t num_kv_heads, int32_t head_dim,
                                   RBF o, RF32 lse) {
  const auto d = qkv.dimensions();
  mpx_attn_train_args a{};
  a.B = static_cast<int>(d[0]); a.T = static_cast<int>(d[1]);
  a.Nq = num_heads; a.Nkv = num_kv_heads; a.D = head_dim;
  const long long ld = static_cast<long long>(a.Nq + 2 * a.Nkv) * a.D;
  a.q = bf(qkv); a.k = a.q + static_cast<long long>(a.Nq) * a.D; a.v = a.k + static_cast<long long>(a.Nkv) * a.D;
  a.ldq = a.ldk = a.ldv = ld;
  a.o = rbf(o); a.ldo = static_cast<long long>(a.Nq) * a.D;
  a.lse = lse->typed_data();
  MPX_CALL(mpx_attn_train_fwd(&a, stream), "mpx_attn_train_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAttnTrainFwd, AttnTrainFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Attr<int32_t>("num_heads").Attr<int32_t>("num_kv_heads")
                                  .Attr<int32_t>("head_dim").Ret<BF>().Ret<F32>(),
                              MPX_TRAITS);

// Its transpose (the bwd rule of the custom_vjp): dqkv holds the cotangents of the UN-rotated projections.
static ffi::Error AttnTrainBwdImpl(mpx_stream_t stream, BF qkv, BF o, BF dout, F32 lse, S32 positions, F32 timescale,
                                   int32_t num_heads, int32_t num_kv_heads, int32_t head_dim, RBF dqkv, RF32 delta,
                                   RF32 dq_acc) {
  const auto d = qkv.dimensions();
  mpx_attn_train_bwd_args a{};
  a.B = static_cast<int>(d[0]); a.T = static_cast<int>(d[1]);
  a.Nq = num_heads; a.Nkv = num_kv_heads; a.D = head_dim;
  const long long qd = static_cast<long long>(a.Nq) * a.D, kd = static_cast<long long>(a.Nkv) * a.D, ld = qd + 2 * kd;
  a.q = bf(qkv); a.k = a.q + qd; a.v = a.k + kd;
  a.ldq = a.ldk = a.ldv = ld;
  a.o = bf(o); a.ldo = qd;
  a.dout = bf(dout); a.lddo = qd;
  a.lse = lse.typed_data();
  a.delta = delta->typed_data(); a.dq_acc = dq_acc->typed_data();
  a.dq = rbf(dqkv); a.dk = a.dq + qd; a.dv = a.dk + kd;
  a.lddq = a.lddk = a.lddv = ld;
  a.pos = positions.typed_data(); a.pos_len = static_cast<int>(positions.element_count());
  a.rope_timescale = timescale.typed_data();
  MPX_CALL(mpx_attn_train_bwd(&a, stream), "mpx_attn_train_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAttnTrainBwd, AttnTrainBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<F32>().Arg<S32>().Arg<F32>()
                                  .Attr<int32_t>("num_heads").Attr<int32_t>("num_kv_heads").Attr<int32_t>("head_dim")
                                  .Ret<BF>().Ret<F32>().Ret<F32>(),
                              MPX_TRAITS);

// ================================================================================================ GLU
// Up/gate projections + activation (feed_forward.py:73-77), general H.  u_save / g_save: 0 elements => not kept.
static ffi::Error GluUpImpl(mpx_stream_t stream, BF x, BF wt_ug, RBF h, RBF u_save, RBF g_save) {
  mpx_glu_up_args a{};
  a.K = last_dim(x); a.M = static_cast<int>(rows_of(x)); a.F = last_dim(*h);
  a.x = bf(x); a.ldx = a.K;
  a.wt_ug = bf(wt_ug);
  a.h = rbf(h);
  a.u_save = u_save->element_count() ? rbf(u_save) : nullptr;
  a.g_save = g_save->element_count() ? rbf(g_save) : nullptr;
  MPX_CALL(mpx_glu_up(&a, stream), "mpx_glu_up");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGluUp, GluUpImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Ret<BF>().Ret<BF>().Ret<BF>(), MPX_TRAITS);

// Whole feed-forward half of a block in one kernel (network.py:163-170 + feed_forward.py:73-78), H = 128.
// norm_scale / wt_o: 0 elements => absent.  h_out: 0 elements => h stays on chip (rollout); kept for training.
static ffi::Error GluFusedImpl(mpx_stream_t stream, BF x, BF wt_ug64, BF wt_d, BF residual, F32 norm_scale, BF wt_o, float eps,
                               RBF out, RBF h_out) {
  const int H = last_dim(x);
  const int F = static_cast<int>(wt_ug64.dimensions()[0]) / 2;
  MPX_CALL(mpx_glu_fused(bf(x), bf(wt_ug64), bf(wt_d), bf(residual), rbf(out), static_cast<int>(rows_of(x)), H, F,
                         norm_scale.element_count() ? norm_scale.typed_data() : nullptr, eps,
                         wt_o.element_count() ? bf(wt_o) : nullptr, h_out->element_count() ? rbf(h_out) : nullptr, stream),
           "mpx_glu_fused");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGluFused, GluFusedImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<F32>().Arg<BF>()
                                  .Attr<float>("eps").Ret<BF>().Ret<BF>(),
                              MPX_TRAITS);

// The last layer's feed-forward half with the policy head folded in (network.py:163-170 + 321-333, train.py:97-98): one
// launch instead of MpxGluFused + MpxPolicyHeadSample, bit-identical results.  action_mask / gumbel / stream_offset /
// logits / xf: 0 elements => absent.
static ffi::Error GluFusedHeadImpl(mpx_stream_t stream, BF x, BF wt_ug64, BF wt_d, BF residual, F32 norm_scale, BF wt_o,
                                   F32 head_norm_scale, F32 table, PRED action_mask, F32 gumbel, U64 stream_offset, float eps,
                                   float head_eps, int64_t seed, int64_t stream_id, RBF out, RS32 action, RF32 log_prob,
                                   RF32 logits, RBF xf) {
  const int H = last_dim(x);
  const int F = static_cast<int>(wt_ug64.dimensions()[0]) / 2;
  mpx_policy_head_args a{};
  a.norm_scale = head_norm_scale.typed_data(); a.eps = head_eps;
  a.table = table.typed_data(); a.A = static_cast<int>(table.dimensions()[0]);
  a.action_mask = action_mask.element_count() ? u8(action_mask) : nullptr;
  a.gumbel_in = gumbel.element_count() ? gumbel.typed_data() : nullptr;
  a.seed = static_cast<unsigned long long>(seed); a.stream_id = static_cast<unsigned long long>(stream_id);
  a.stream_offset = stream_offset.element_count()
                        ? reinterpret_cast<const unsigned long long*>(stream_offset.untyped_data()) : nullptr;
  a.logits_out = logits->element_count() ? logits->typed_data() : nullptr;
  a.xf_out = xf->element_count() ? rbf(xf) : nullptr;
  a.action = action->typed_data();
  a.log_prob = log_prob->typed_data();
  MPX_CALL(mpx_glu_fused_head(bf(x), bf(wt_ug64), bf(wt_d), bf(residual), rbf(out), static_cast<int>(rows_of(x)), H, F,
                              norm_scale.typed_data(), eps, bf(wt_o), &a, stream),
           "mpx_glu_fused_head");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGluFusedHead, GluFusedHeadImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<F32>().Arg<BF>()
                                  .Arg<F32>().Arg<F32>().Arg<PRED>().Arg<F32>().Arg<U64>()
                                  .Attr<float>("eps").Attr<float>("head_eps").Attr<int64_t>("seed").Attr<int64_t>("stream_id")
                                  .Ret<BF>().Ret<S32>().Ret<F32>().Ret<F32>().Ret<BF>(),
                              MPX_TRAITS);

// Its backward in one kernel (u, g, dh recomputed on chip): dug (M,2F) = [du | dg], dxn (M,H)
static ffi::Error GluBwdFusedImpl(mpx_stream_t stream, BF xn, BF dy, BF wt_ug64, BF wt_d, RBF dug, RBF dxn) {
  const int H = last_dim(xn);
  const int F = static_cast<int>(wt_ug64.dimensions()[0]) / 2;
  MPX_CALL(mpx_glu_bwd_fused(bf(xn), bf(dy), bf(wt_ug64), bf(wt_d), rbf(dug), rbf(dxn), static_cast<int>(rows_of(xn)), H, F,
                             stream), "mpx_glu_bwd_fused");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGluBwdFused, GluBwdFusedImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Ret<BF>().Ret<BF>(),
                              MPX_TRAITS);

// Transpose of the GLU activation for the unfused path (feed_forward.py:75-77)
static ffi::Error GluBwdImpl(mpx_stream_t stream, BF dh, BF u, BF g, RBF dug) {
  MPX_CALL(mpx_glu_bwd(bf(dh), bf(u), bf(g), rbf(dug), rows_of(dh), last_dim(dh), stream), "mpx_glu_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGluBwd, GluBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Ret<BF>(), MPX_TRAITS);

// ================================================================================================ row-wise
static ffi::Error RmsNormFwdImpl(mpx_stream_t stream, BF x, F32 scale, float eps, RBF y) {
  MPX_CALL(mpx_rmsnorm_fwd(bf(x), scale.typed_data(), eps, rbf(y), rows_of(x), last_dim(x), stream), "mpx_rmsnorm_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxRmsNormFwd, RmsNormFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Attr<float>("eps").Ret<BF>(), MPX_TRAITS);

// dres: 0 elements => no residual cotangent.  aliases: dscale_in(4) -> 1 (accumulated)
static ffi::Error RmsNormBwdImpl(mpx_stream_t stream, BF dy, BF x, F32 scale, BF dres, F32 dscale_in, float eps, RBF dx,
                                 RF32 dscale) {
  (void)dscale_in;
  MPX_CALL(mpx_rmsnorm_bwd(bf(dy), bf(x), scale.typed_data(), eps, dres.element_count() ? bf(dres) : nullptr, rbf(dx),
                           dscale->typed_data(), rows_of(x), last_dim(x), stream), "mpx_rmsnorm_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxRmsNormBwd, RmsNormBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<F32>().Arg<BF>().Arg<F32>()
                                  .Attr<float>("eps").Ret<BF>().Ret<F32>(),
                              MPX_TRAITS);

static ffi::Error GeluBwdImpl(mpx_stream_t stream, BF dy, BF z, RBF dz) {
  MPX_CALL(mpx_gelu_bwd(bf(dy), bf(z), rbf(dz), static_cast<long long>(dy.element_count()), stream), "mpx_gelu_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGeluBwd, GeluBwdImpl, ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Ret<BF>(), MPX_TRAITS);

// aliases: out_in(1) -> 0 (accumulated)
static ffi::Error ColSumImpl(mpx_stream_t stream, BF x, F32 out_in, RF32 out) {
  (void)out_in;
  const int N = last_dim(x);
  MPX_CALL(mpx_colsum(bf(x), N, rows_of(x), N, out->typed_data(), stream), "mpx_colsum");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxColSum, ColSumImpl, ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Ret<F32>(), MPX_TRAITS);

static ffi::Error AddBf16Impl(mpx_stream_t stream, BF a, BF b, RBF out) {
  MPX_CALL(mpx_add_bf16(bf(a), bf(b), rbf(out), static_cast<long long>(a.element_count()), stream), "mpx_add_bf16");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAddBf16, AddBf16Impl, ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Ret<BF>(), MPX_TRAITS);

// ================================================================================================ embeds and heads
// network.py:303-309.  task_ids / task_table: 0 elements => no task embedding.
static ffi::Error EmbedCombineImpl(mpx_stream_t stream, BF obs_emb, F32 reward, S32 last_action, S32 task_ids, F32 w_reward,
                                   F32 b_reward, F32 action_table, F32 task_table, RBF x) {
  const int H = last_dim(obs_emb);
  const bool tasks = task_ids.element_count() && task_table.element_count();
  MPX_CALL(mpx_embed_combine(bf(obs_emb), reward.typed_data(), last_action.typed_data(),
                             tasks ? task_ids.typed_data() : nullptr, w_reward.typed_data(), b_reward.typed_data(),
                             action_table.typed_data(), static_cast<int>(action_table.dimensions()[0]),
                             tasks ? task_table.typed_data() : nullptr,
                             tasks ? static_cast<int>(task_table.dimensions()[0]) : 0, rbf(x), rows_of(obs_emb), H, stream),
           "mpx_embed_combine");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxEmbedCombine, EmbedCombineImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Ret<BF>(),
                              MPX_TRAITS);

// Its transpose.  aliases: d_w_reward_in(4) -> 0, d_b_reward_in(5) -> 1, d_action_table_in(6) -> 2, d_task_table_in(7) -> 3
static ffi::Error EmbedBwdImpl(mpx_stream_t stream, BF dx, F32 reward, S32 last_action, S32 task_ids, F32 dwr_in, F32 dbr_in,
                               F32 dat_in, F32 dtt_in, RF32 d_w_reward, RF32 d_b_reward, RF32 d_action_table,
                               RF32 d_task_table) {
  (void)dwr_in; (void)dbr_in; (void)dat_in; (void)dtt_in;
  const bool tasks = task_ids.element_count() && d_task_table->element_count();
  MPX_CALL(mpx_embed_bwd(bf(dx), reward.typed_data(), last_action.typed_data(), tasks ? task_ids.typed_data() : nullptr,
                         static_cast<int>(d_action_table->dimensions()[0]),
                         tasks ? static_cast<int>(d_task_table->dimensions()[0]) : 0, d_w_reward->typed_data(),
                         d_b_reward->typed_data(), d_action_table->typed_data(),
                         tasks ? d_task_table->typed_data() : nullptr, rows_of(dx), last_dim(dx), stream), "mpx_embed_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxEmbedBwd, EmbedBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>(),
                              MPX_TRAITS);

// Weight image of the fused observation encoder (once per weight update).  sizes: one-hot sizes, a HOST attribute.
static ffi::Error ObsFusedPrepImpl(mpx_stream_t stream, BF w1, BF b1, BF wt2, BF b2, BF wt3, BF b3,
                                   ffi::Span<const int32_t> sizes, RU8 image) {
  int sz[8] = {0};
  const int K = static_cast<int>(sizes.size());
  if (K > 8) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MpxObsFusedPrep: at most 8 observation components");
  for (int i = 0; i < K; ++i) sz[i] = sizes[i];
  MPX_CALL(mpx_obs_fused_prep(sz, K, bf(w1), bf(b1), bf(wt2), bf(b2), bf(wt3), bf(b3), image->untyped_data(), stream),
           "mpx_obs_fused_prep");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxObsFusedPrep, ObsFusedPrepImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>()
                                  .Attr<ffi::Span<const int32_t>>("sizes").Ret<U8>(),
                              MPX_TRAITS);

// Rollout-step observation encoder + input embedding sum (observation.py:84-96 + network.py:303-309), obs (M,IH,IW,K) u8
static ffi::Error ObsEmbedFusedImpl(mpx_stream_t stream, U8 obs, U8 image, F32 reward, S32 last_action, S32 task_ids,
                                    F32 w_reward, F32 b_reward, F32 action_table, F32 task_table,
                                    ffi::Span<const int32_t> sizes, RBF x) {
  int sz[8] = {0};
  const int K = static_cast<int>(sizes.size());
  if (K > 8) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MpxObsEmbedFused: at most 8 observation components");
  for (int i = 0; i < K; ++i) sz[i] = sizes[i];
  const auto od = obs.dimensions();
  const int n = static_cast<int>(od.size());
  const int IH = static_cast<int>(od[n - 3]), IW = static_cast<int>(od[n - 2]);
  const long long M = static_cast<long long>(obs.element_count()) / (static_cast<long long>(IH) * IW * K);
  const bool tasks = task_ids.element_count() && task_table.element_count();
  MPX_CALL(mpx_obs_embed_fused(obs.typed_data(), sz, K, image.untyped_data(), reward.typed_data(), last_action.typed_data(),
                               tasks ? task_ids.typed_data() : nullptr, w_reward.typed_data(), b_reward.typed_data(),
                               action_table.typed_data(), static_cast<int>(action_table.dimensions()[0]),
                               tasks ? task_table.typed_data() : nullptr,
                               tasks ? static_cast<int>(task_table.dimensions()[0]) : 0, rbf(x), nullptr, M, IH, IW,
                               last_dim(*x), stream), "mpx_obs_embed_fused");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxObsEmbedFused, ObsEmbedFusedImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Arg<U8>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Attr<ffi::Span<const int32_t>>("sizes").Ret<BF>(),
                              MPX_TRAITS);

// Training-side observation encoder in one kernel (observation.py:84-96 + network.py:303-309 under nnx.grad): x plus the
// saved activations - z1t / a1t / z2t in the kernels' private tile layouts (flat, sized by the Python side from
// mpx_obs_train_tiles) and a2 (M,288) row-major
static ffi::Error ObsTrainFwdImpl(mpx_stream_t stream, U8 obs, U8 image, F32 reward, S32 last_action, S32 task_ids,
                                  F32 w_reward, F32 b_reward, F32 action_table, F32 task_table,
                                  ffi::Span<const int32_t> sizes, RBF x, RBF z1t, RBF a1t, RBF z2t, RBF a2) {
  int sz[8] = {0};
  const int K = static_cast<int>(sizes.size());
  if (K > 8) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MpxObsTrainFwd: at most 8 observation components");
  for (int i = 0; i < K; ++i) sz[i] = sizes[i];
  const auto od = obs.dimensions();
  const int n = static_cast<int>(od.size());
  const int IH = static_cast<int>(od[n - 3]), IW = static_cast<int>(od[n - 2]);
  const long long M = static_cast<long long>(obs.element_count()) / (static_cast<long long>(IH) * IW * K);
  const bool tasks = task_ids.element_count() && task_table.element_count();
  MPX_CALL(mpx_obs_train_fwd(obs.typed_data(), sz, K, image.untyped_data(), reward.typed_data(), last_action.typed_data(),
                             tasks ? task_ids.typed_data() : nullptr, w_reward.typed_data(), b_reward.typed_data(),
                             action_table.typed_data(), static_cast<int>(action_table.dimensions()[0]),
                             tasks ? task_table.typed_data() : nullptr,
                             tasks ? static_cast<int>(task_table.dimensions()[0]) : 0, rbf(z1t), rbf(a1t), rbf(z2t), rbf(a2),
                             rbf(x), nullptr, M, IH, IW, last_dim(*x), stream), "mpx_obs_train_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxObsTrainFwd, ObsTrainFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Arg<U8>().Arg<F32>().Arg<S32>().Arg<S32>().Arg<F32>()
                                  .Arg<F32>().Arg<F32>().Arg<F32>().Attr<ffi::Span<const int32_t>>("sizes").Ret<BF>()
                                  .Ret<BF>().Ret<BF>().Ret<BF>().Ret<BF>(),
                              MPX_TRAITS);

// Its data gradients: dz2t (tile layout), dz1 (M*25,16); db1 / db2 accumulate.  aliases: db1_in(4) -> 2, db2_in(5) -> 3
static ffi::Error ObsTrainBwdImpl(mpx_stream_t stream, BF dx0, BF z1t, BF z2t, U8 image, F32 db1_in, F32 db2_in, RBF dz2t,
                                  RBF dz1, RF32 db1, RF32 db2, RU8 workspace) {
  (void)db1_in; (void)db2_in;
  MPX_CALL(mpx_obs_train_bwd(bf(dx0), bf(z1t), bf(z2t), image.untyped_data(), rbf(dz2t), rbf(dz1), db1->typed_data(),
                             db2->typed_data(), rows_of(dx0), workspace->untyped_data(), workspace->element_count(), stream),
           "mpx_obs_train_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxObsTrainBwd, ObsTrainBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<U8>().Arg<F32>().Arg<F32>()
                                  .Ret<BF>().Ret<BF>().Ret<F32>().Ret<F32>().Ret<U8>(),
                              MPX_TRAITS);

// Tap-wise conv2 weight gradient from the position tiles, M = token count.  aliases: dw_in(2) -> 0 (accumulated)
static ffi::Error ObsConv2WgradImpl(mpx_stream_t stream, BF a1t, BF dz2t, F32 dw_in, int64_t M, RF32 dw, RU8 workspace) {
  (void)dw_in;
  MPX_CALL(mpx_obs_conv2_wgrad(bf(a1t), bf(dz2t), dw->typed_data(), static_cast<long long>(M), workspace->untyped_data(),
                               workspace->element_count(), stream), "mpx_obs_conv2_wgrad");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxObsConv2Wgrad, ObsConv2WgradImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<F32>().Attr<int64_t>("M").Ret<F32>()
                                  .Ret<U8>(),
                              MPX_TRAITS);

// conv1 straight from the uint8 observation: a_out (M*OH*OW, cout), z_out (0 elements => not kept)
static ffi::Error Conv1OnehotFwdImpl(mpx_stream_t stream, U8 obs, BF w, BF bias, ffi::Span<const int32_t> sizes, int32_t kh,
                                     int32_t kw, int32_t sh, int32_t sw, int32_t apply_gelu, RBF a_out, RBF z_out) {
  int sz[8] = {0};
  const int K = static_cast<int>(sizes.size());
  if (K > 8) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MpxConv1OnehotFwd: at most 8 observation components");
  for (int i = 0; i < K; ++i) sz[i] = sizes[i];
  const auto od = obs.dimensions();
  const int n = static_cast<int>(od.size());
  const int IH = static_cast<int>(od[n - 3]), IW = static_cast<int>(od[n - 2]);
  const long long M = static_cast<long long>(obs.element_count()) / (static_cast<long long>(IH) * IW * K);
  MPX_CALL(mpx_conv1_onehot_fwd(obs.typed_data(), bf(w), bf(bias), z_out->element_count() ? rbf(z_out) : nullptr, rbf(a_out), M,
                                IH, IW, K, sz, kh, kw, sh, sw, last_dim(w), apply_gelu, stream), "mpx_conv1_onehot_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxConv1OnehotFwd, Conv1OnehotFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Arg<BF>().Arg<BF>()
                                  .Attr<ffi::Span<const int32_t>>("sizes").Attr<int32_t>("kh").Attr<int32_t>("kw")
                                  .Attr<int32_t>("sh").Attr<int32_t>("sw").Attr<int32_t>("apply_gelu").Ret<BF>().Ret<BF>(),
                              MPX_TRAITS);

// conv1 weight gradient with the one-hot operand generated in shared memory.  aliases: dw_in(2) -> 0
static ffi::Error Conv1OnehotWgradImpl(mpx_stream_t stream, U8 obs, BF dz, F32 dw_in, ffi::Span<const int32_t> sizes, int32_t kh,
                                       int32_t kw, int32_t sh, int32_t sw, RF32 dw, RU8 workspace) {
  (void)dw_in;
  int sz[8] = {0};
  const int K = static_cast<int>(sizes.size());
  if (K > 8) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MpxConv1OnehotWgrad: at most 8 observation components");
  for (int i = 0; i < K; ++i) sz[i] = sizes[i];
  const auto od = obs.dimensions();
  const int n = static_cast<int>(od.size());
  const int IH = static_cast<int>(od[n - 3]), IW = static_cast<int>(od[n - 2]);
  const long long M = static_cast<long long>(obs.element_count()) / (static_cast<long long>(IH) * IW * K);
  MPX_CALL(mpx_conv1_onehot_wgrad(obs.typed_data(), bf(dz), dw->typed_data(), M, IH, IW, K, sz, kh, kw, sh, sw, last_dim(dz),
                                  workspace->untyped_data(), workspace->element_count(), stream), "mpx_conv1_onehot_wgrad");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxConv1OnehotWgrad, Conv1OnehotWgradImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Arg<BF>().Arg<F32>()
                                  .Attr<ffi::Span<const int32_t>>("sizes").Attr<int32_t>("kh").Attr<int32_t>("kw")
                                  .Attr<int32_t>("sh").Attr<int32_t>("sw").Ret<F32>().Ret<U8>(),
                              MPX_TRAITS);

// one-hot im2col rows of the first conv (only the generic weight-gradient path needs them)
static ffi::Error ObsOnehotIm2colImpl(mpx_stream_t stream, U8 obs, ffi::Span<const int32_t> sizes, int32_t kh, int32_t kw,
                                      int32_t sh, int32_t sw, RBF out) {
  int sz[8] = {0};
  const int K = static_cast<int>(sizes.size());
  if (K > 8) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "MpxObsOnehotIm2col: at most 8 observation components");
  for (int i = 0; i < K; ++i) sz[i] = sizes[i];
  const auto od = obs.dimensions();
  const int n = static_cast<int>(od.size());
  const int IH = static_cast<int>(od[n - 3]), IW = static_cast<int>(od[n - 2]);
  const long long M = static_cast<long long>(obs.element_count()) / (static_cast<long long>(IH) * IW * K);
  MPX_CALL(mpx_obs_onehot_im2col(obs.typed_data(), rbf(out), M, IH, IW, K, sz, kh, kw, sh, sw, last_dim(*out), stream),
           "mpx_obs_onehot_im2col");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxObsOnehotIm2col, ObsOnehotIm2colImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Attr<ffi::Span<const int32_t>>("sizes")
                                  .Attr<int32_t>("kh").Attr<int32_t>("kw").Attr<int32_t>("sh").Attr<int32_t>("sw").Ret<BF>(),
                              MPX_TRAITS);

// NHWC im2col / col2im of the hidden conv layers; a (M,IH,IW,C)
static ffi::Error Im2colImpl(mpx_stream_t stream, BF a, int32_t kh, int32_t kw, int32_t sh, int32_t sw, RBF x) {
  const auto d = a.dimensions();
  const int n = static_cast<int>(d.size());
  const int IH = static_cast<int>(d[n - 3]), IW = static_cast<int>(d[n - 2]), C = static_cast<int>(d[n - 1]);
  MPX_CALL(mpx_im2col(bf(a), rbf(x), static_cast<long long>(a.element_count()) / (static_cast<long long>(IH) * IW * C), IH, IW, C,
                      kh, kw, sh, sw, stream), "mpx_im2col");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxIm2col, Im2colImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Attr<int32_t>("kh").Attr<int32_t>("kw")
                                  .Attr<int32_t>("sh").Attr<int32_t>("sw").Ret<BF>(),
                              MPX_TRAITS);

// z (the saved pre-activation of the layer below): 0 elements => plain col2im; da has the (M,IH,IW,C) shape
static ffi::Error Col2imImpl(mpx_stream_t stream, BF dx, BF z, int32_t kh, int32_t kw, int32_t sh, int32_t sw, RBF da) {
  const auto d = da->dimensions();
  const int n = static_cast<int>(d.size());
  const int IH = static_cast<int>(d[n - 3]), IW = static_cast<int>(d[n - 2]), C = static_cast<int>(d[n - 1]);
  MPX_CALL(mpx_col2im(bf(dx), rbf(da), static_cast<long long>(da->element_count()) / (static_cast<long long>(IH) * IW * C), IH, IW,
                      C, kh, kw, sh, sw, z.element_count() ? bf(z) : nullptr, stream), "mpx_col2im");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxCol2im, Col2imImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Attr<int32_t>("kh").Attr<int32_t>("kw")
                                  .Attr<int32_t>("sh").Attr<int32_t>("sw").Ret<BF>(),
                              MPX_TRAITS);

// FlattenedObsEncoder (observation.py:98-145): embedding lookup half; obs (M,cells) u8
static ffi::Error FlatEmbedFwdImpl(mpx_stream_t stream, U8 obs, F32 w, F32 bias, int32_t cells, RBF out) {
  MPX_CALL(mpx_flat_embed_fwd(obs.typed_data(), w.typed_data(), bias.typed_data(), rbf(out),
                              static_cast<long long>(obs.element_count()) / cells, cells, static_cast<int>(w.dimensions()[0]),
                              last_dim(*out), stream), "mpx_flat_embed_fwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxFlatEmbedFwd, FlatEmbedFwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Arg<F32>().Arg<F32>().Attr<int32_t>("cells").Ret<BF>(),
                              MPX_TRAITS);

// aliases: dw_in(2) -> 0, dbias_in(3) -> 1
static ffi::Error FlatEmbedBwdImpl(mpx_stream_t stream, U8 obs, BF dflat, F32 dw_in, F32 dbias_in, int32_t cells, RF32 dw,
                                   RF32 dbias) {
  (void)dw_in; (void)dbias_in;
  MPX_CALL(mpx_flat_embed_bwd(obs.typed_data(), bf(dflat), dw->typed_data(), dbias->typed_data(),
                              static_cast<long long>(obs.element_count()) / cells, cells, static_cast<int>(dw->dimensions()[0]),
                              last_dim(dflat), stream), "mpx_flat_embed_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxFlatEmbedBwd, FlatEmbedBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Arg<BF>().Arg<F32>().Arg<F32>().Attr<int32_t>("cells")
                                  .Ret<F32>().Ret<F32>(),
                              MPX_TRAITS);

// network.py:323-331: tied logits + mask.  action_mask: 0 elements => unmasked.
static ffi::Error PolicyLogitsImpl(mpx_stream_t stream, BF x, F32 table, PRED action_mask, RF32 logits) {
  MPX_CALL(mpx_policy_logits(bf(x), table.typed_data(), action_mask.element_count() ? u8(action_mask) : nullptr,
                             logits->typed_data(), rows_of(x), last_dim(x), static_cast<int>(table.dimensions()[0]), stream),
           "mpx_policy_logits");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxPolicyLogits, PolicyLogitsImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Arg<PRED>().Ret<F32>(), MPX_TRAITS);

// distrax.Categorical(logits).sample / log_prob (train.py:97-98) with the Gumbel noise as an input (e.g. from
// jax.random.gumbel: same logits + same noise => same index)
static ffi::Error PolicySampleImpl(mpx_stream_t stream, F32 logits, F32 gumbel, RS32 action, RF32 log_prob) {
  MPX_CALL(mpx_policy_sample(logits.typed_data(), gumbel.typed_data(), action->typed_data(), log_prob->typed_data(),
                             rows_of(logits), last_dim(logits), stream), "mpx_policy_sample");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxPolicySample, PolicySampleImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Ret<S32>().Ret<F32>(), MPX_TRAITS);

// Whole policy path of a rollout step (network.py:321-333, train.py:97-98).  gumbel / action_mask / norm_scale /
// stream_offset: 0 elements => absent (built-in Philox noise keyed by seed, stream_id).
static ffi::Error PolicyHeadSampleImpl(mpx_stream_t stream, BF x, F32 norm_scale, F32 table, PRED action_mask, F32 gumbel,
                                       U64 stream_offset, float eps, int64_t seed, int64_t stream_id, RS32 action,
                                       RF32 log_prob, RF32 logits) {
  MPX_CALL(mpx_policy_head_sample(bf(x), norm_scale.element_count() ? norm_scale.typed_data() : nullptr, eps,
                                  table.typed_data(), action_mask.element_count() ? u8(action_mask) : nullptr,
                                  gumbel.element_count() ? gumbel.typed_data() : nullptr,
                                  static_cast<unsigned long long>(seed), static_cast<unsigned long long>(stream_id),
                                  stream_offset.element_count()
                                      ? reinterpret_cast<const unsigned long long*>(stream_offset.untyped_data()) : nullptr,
                                  logits->element_count() ? logits->typed_data() : nullptr, nullptr, action->typed_data(),
                                  log_prob->typed_data(), rows_of(x), last_dim(x), static_cast<int>(table.dimensions()[0]),
                                  stream), "mpx_policy_head_sample");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxPolicyHeadSample, PolicyHeadSampleImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Arg<F32>().Arg<PRED>().Arg<F32>().Arg<U64>()
                                  .Attr<float>("eps").Attr<int64_t>("seed").Attr<int64_t>("stream_id")
                                  .Ret<S32>().Ret<F32>().Ret<F32>(),
                              MPX_TRAITS);

static ffi::Error GumbelNoiseImpl(mpx_stream_t stream, U64 stream_offset, int64_t seed, int64_t stream_id, RF32 out) {
  MPX_CALL(mpx_gumbel_noise(out->typed_data(), static_cast<long long>(out->element_count()),
                            static_cast<unsigned long long>(seed), static_cast<unsigned long long>(stream_id),
                            stream_offset.element_count()
                                ? reinterpret_cast<const unsigned long long*>(stream_offset.untyped_data()) : nullptr,
                            stream), "mpx_gumbel_noise");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGumbelNoise, GumbelNoiseImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U64>().Attr<int64_t>("seed").Attr<int64_t>("stream_id").Ret<F32>(),
                              MPX_TRAITS);

// Policy part of ppo_loss (train.py:193-212) forward + gradient.  entropy_coef_dev: 0 elements => the attribute is used.
static ffi::Error PpoPolicyLossImpl(mpx_stream_t stream, F32 logits, S32 actions, F32 old_log_prob, F32 advantages,
                                    F32 entropy_coef_dev, float clip_eps, float entropy_coef, RF32 losses, RF32 dlogits,
                                    RU8 workspace) {
  MPX_CALL(mpx_ppo_policy_loss(logits.typed_data(), actions.typed_data(), old_log_prob.typed_data(), advantages.typed_data(),
                               clip_eps, entropy_coef,
                               entropy_coef_dev.element_count() ? entropy_coef_dev.typed_data() : nullptr,
                               losses->typed_data(), dlogits->element_count() ? dlogits->typed_data() : nullptr,
                               workspace->untyped_data(), rows_of(logits), last_dim(logits), stream), "mpx_ppo_policy_loss");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxPpoPolicyLoss, PpoPolicyLossImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<S32>().Arg<F32>().Arg<F32>().Arg<F32>()
                                  .Attr<float>("clip_eps").Attr<float>("entropy_coef").Ret<F32>().Ret<F32>().Ret<U8>(),
                              MPX_TRAITS);

// Transpose of the tied policy head.  dx_other: 0 elements => none.  aliases: dtable_in(4) -> 1
static ffi::Error PolicyHeadBwdImpl(mpx_stream_t stream, F32 dlogits, BF x, F32 table, BF dx_other, F32 dtable_in, RBF dx,
                                    RF32 dtable) {
  (void)dtable_in;
  MPX_CALL(mpx_policy_head_bwd(dlogits.typed_data(), bf(x), table.typed_data(), dx_other.element_count() ? bf(dx_other) : nullptr,
                               rbf(dx), dtable->typed_data(), rows_of(x), last_dim(x), static_cast<int>(table.dimensions()[0]),
                               stream), "mpx_policy_head_bwd");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxPolicyHeadBwd, PolicyHeadBwdImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<BF>().Arg<F32>().Arg<BF>().Arg<F32>().Ret<BF>().Ret<F32>(),
                              MPX_TRAITS);

// Backward of the value MLP + fp32 head in one kernel (network.py:335-341, values.py:34,40-41 under nnx.grad): pv, dz
// (weight-gradient operands), d xf; the value-MLP bias gradient accumulates.  aliases: db_in(5) -> 3
static ffi::Error ValueBwdFusedImpl(mpx_stream_t stream, BF xf, BF dvl, BF wt_vm, BF b_vm, BF w_vh_pad, F32 db_in, RBF pv,
                                    RBF dz, RBF dxf, RF32 db, RU8 workspace) {
  (void)db_in;
  MPX_CALL(mpx_value_bwd_fused(bf(xf), bf(dvl), bf(wt_vm), bf(b_vm), bf(w_vh_pad), rbf(pv), rbf(dz), rbf(dxf), db->typed_data(),
                               static_cast<int>(rows_of(xf)), last_dim(xf), static_cast<int>(wt_vm.dimensions()[0]),
                               workspace->untyped_data(), workspace->element_count(), stream), "mpx_value_bwd_fused");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxValueBwdFused, ValueBwdFusedImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<F32>()
                                  .Ret<BF>().Ret<BF>().Ret<BF>().Ret<F32>().Ret<U8>(),
                              MPX_TRAITS);

// Whole value path of a rollout step (network.py:321,335-341 + values.py:40-45).  norm_scale: 0 elements => x is xf.
static ffi::Error ValueHeadFusedImpl(mpx_stream_t stream, BF x, F32 norm_scale, BF wt_vm, BF b_vm, BF wt_vh_hilo, F32 b_vh,
                                     F32 centers, float eps, RF32 value, RF32 vlogits) {
  MPX_CALL(mpx_value_head_fused(bf(x), norm_scale.element_count() ? norm_scale.typed_data() : nullptr, eps, bf(wt_vm), bf(b_vm),
                                bf(wt_vh_hilo), b_vh.typed_data(), centers.typed_data(), value->typed_data(),
                                vlogits->element_count() ? vlogits->typed_data() : nullptr, static_cast<int>(rows_of(x)),
                                last_dim(x), static_cast<int>(wt_vm.dimensions()[0]), static_cast<int>(centers.element_count()),
                                stream), "mpx_value_head_fused");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxValueHeadFused, ValueHeadFusedImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<BF>().Arg<F32>().Arg<BF>().Arg<BF>().Arg<BF>().Arg<F32>().Arg<F32>()
                                  .Attr<float>("eps").Ret<F32>().Ret<F32>(),
                              MPX_TRAITS);

// ================================================================================================ optimizer / staging
// optimizer.py:12-23 on flat fp32 buffers.  aliases: params(0) -> 0, mu(2) -> 1, nu(3) -> 2, step(4) -> 3
static ffi::Error AdamwStepImpl(mpx_stream_t stream, F32 params_in, F32 grads, F32 mu_in, F32 nu_in, S32 step_in, float lr0,
                                float total_steps, float b1, float b2, float eps, float weight_decay, float max_norm,
                                float grad_scale, RF32 params, RF32 mu, RF32 nu, RS32 step, RF32 update_norm, RF32 scratch,
                                RU8 workspace) {
  (void)params_in; (void)mu_in; (void)nu_in; (void)step_in;
  MPX_CALL(mpx_adamw_step(params->typed_data(), grads.typed_data(), mu->typed_data(), nu->typed_data(), scratch->typed_data(),
                          static_cast<long long>(params->element_count()), step->typed_data(), lr0, total_steps, b1, b2, eps,
                          weight_decay, max_norm, grad_scale, update_norm->typed_data(), workspace->untyped_data(), stream),
           "mpx_adamw_step");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAdamwStep, AdamwStepImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>()
                                  .Attr<float>("lr0").Attr<float>("total_steps").Attr<float>("b1").Attr<float>("b2")
                                  .Attr<float>("eps").Attr<float>("weight_decay").Attr<float>("max_norm").Attr<float>("grad_scale")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>().Ret<F32>().Ret<F32>().Ret<U8>(),
                              MPX_TRAITS);

// optimizer.py:25-37 (muon partition + adamw leaves + post-update clip); phases = 3 is the whole step, 1 / 2 its two
// enqueue phases.  mats: device table of matrix records.  aliases: params(0) -> 0, mu(2) -> 1, nu(3) -> 2, step(4) -> 3
static ffi::Error MuonStepImpl(mpx_stream_t stream, F32 params_in, F32 grads, F32 mu_in, F32 nu_in, S32 step_in, U8 mats,
                               int64_t n_adam, int32_t nmat, int64_t x_total, int64_t a_total, int32_t max_r, int32_t max_c,
                               float lr0, float total_steps, float adam_b1, float adam_b2, float eps, float weight_decay,
                               float muon_beta, int32_t ns_steps, float max_norm, float grad_scale, int32_t phases, RF32 params,
                               RF32 mu, RF32 nu, RS32 step, RF32 update_norm, RF32 scratch, RF32 ns_ws, RU8 workspace) {
  (void)params_in; (void)mu_in; (void)nu_in; (void)step_in;
  MPX_CALL(mpx_muon_step_phased(params->typed_data(), grads.typed_data(), mu->typed_data(), nu->typed_data(),
                                scratch->typed_data(), static_cast<long long>(params->element_count()), n_adam,
                                mats.untyped_data(), nmat, x_total, a_total, max_r, max_c, ns_ws->typed_data(),
                                step->typed_data(), lr0, total_steps, adam_b1, adam_b2, eps, weight_decay, muon_beta, ns_steps,
                                max_norm, grad_scale, update_norm->typed_data(), workspace->untyped_data(), phases, stream),
           "mpx_muon_step_phased");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxMuonStep, MuonStepImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<U8>()
                                  .Attr<int64_t>("n_adam").Attr<int32_t>("nmat").Attr<int64_t>("x_total").Attr<int64_t>("a_total")
                                  .Attr<int32_t>("max_r").Attr<int32_t>("max_c").Attr<float>("lr0").Attr<float>("total_steps")
                                  .Attr<float>("adam_b1").Attr<float>("adam_b2").Attr<float>("eps").Attr<float>("weight_decay")
                                  .Attr<float>("muon_beta").Attr<int32_t>("ns_steps").Attr<float>("max_norm")
                                  .Attr<float>("grad_scale").Attr<int32_t>("phases")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<U8>(),
                              MPX_TRAITS);

// Second half of the sharded Muon step (after the all-gather of the Muon partition of `scratch`).
// aliases: params(0) -> 0, mu(2) -> 1, nu(3) -> 2, step(4) -> 3, scratch(5) -> 5
static ffi::Error MuonApplyShardedImpl(mpx_stream_t stream, F32 params_in, F32 grads, F32 mu_in, F32 nu_in, S32 step_in,
                                       F32 scratch_in, int64_t n_adam, float lr0, float total_steps, float adam_b1,
                                       float adam_b2, float eps, float weight_decay, float max_norm, float grad_scale,
                                       RF32 params, RF32 mu, RF32 nu, RS32 step, RF32 update_norm, RF32 scratch,
                                       RU8 workspace) {
  (void)params_in; (void)mu_in; (void)nu_in; (void)step_in; (void)scratch_in;
  MPX_CALL(mpx_muon_apply_sharded(params->typed_data(), grads.typed_data(), mu->typed_data(), nu->typed_data(),
                                  scratch->typed_data(), static_cast<long long>(params->element_count()), n_adam,
                                  step->typed_data(), lr0, total_steps, adam_b1, adam_b2, eps, weight_decay, max_norm,
                                  grad_scale, update_norm->typed_data(), workspace->untyped_data(), stream),
           "mpx_muon_apply_sharded");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxMuonApplySharded, MuonApplyShardedImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<S32>().Arg<F32>()
                                  .Attr<int64_t>("n_adam").Attr<float>("lr0").Attr<float>("total_steps")
                                  .Attr<float>("adam_b1").Attr<float>("adam_b2").Attr<float>("eps").Attr<float>("weight_decay")
                                  .Attr<float>("max_norm").Attr<float>("grad_scale")
                                  .Ret<F32>().Ret<F32>().Ret<F32>().Ret<S32>().Ret<F32>().Ret<F32>().Ret<U8>(),
                              MPX_TRAITS);

// fp32 master weights -> bf16 GEMM operand copies.  The descriptor table holds DEVICE pointers into XLA buffers and is
// therefore only valid while those buffers do not move: the JAX side builds it per call from the operand addresses, or
// (simpler) stages weights with jnp casts/transposes and skips this entry point.  Bound for completeness.
static ffi::Error PrepWeightsImpl(mpx_stream_t stream, U8 descs, int32_t ndesc, RU8 token) {
  (void)token;
  MPX_CALL(mpx_prep_weights(reinterpret_cast<const mpx_prep_desc*>(descs.untyped_data()), ndesc, stream), "mpx_prep_weights");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxPrepWeights, PrepWeightsImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Attr<int32_t>("ndesc").Ret<U8>(), MPX_TRAITS);

// Time-major rollout rows -> batch-major training rows with the minibatch permutation (rollout.py:93-126,169-181).
// src (T,B,item...) as bytes; perm: 0 elements => identity.
static ffi::Error TransposeTbImpl(mpx_stream_t stream, U8 src, S32 perm, int32_t T, int32_t B, RU8 dst) {
  const int item = static_cast<int>(static_cast<long long>(src.element_count()) / (static_cast<long long>(T) * B));
  MPX_CALL(mpx_transpose_tb(src.untyped_data(), dst->untyped_data(), perm.element_count() ? perm.typed_data() : nullptr, T, B,
                            item, stream), "mpx_transpose_tb");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxTransposeTb, TransposeTbImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<U8>().Arg<S32>().Attr<int32_t>("T").Attr<int32_t>("B").Ret<U8>(),
                              MPX_TRAITS);
