// XLA typed-FFI adapter: exposes the C ABI of include/mapox_b200.h as jax.ffi custom-call targets.
//
// NOT BUILT BY DEFAULT and UNVERIFIED in this environment: jaxlib (and therefore xla/ffi/api/ffi.h) is not
// installed in the build image or on the GPU box.  Build, where JAX exists, with
//   g++ -O2 -fPIC -shared -I$(python -c "import jax; print(jax.ffi.include_dir())") -I../../../include \
//       xla_ffi.cc -L../.. -lmapox_b200 -Wl,-rpath,'$ORIGIN' -o ../../libmapox_b200_xla.so
// and register from Python as INTEGRATION.md shows.  Handlers are stateless, enqueue on XLA's stream only, never
// synchronise, and are marked command-buffer compatible so XLA can capture them into CUDA graphs (the rollout
// fori_loop body, mapox_trainer/train.py:128-140).
#include <cstdint>

#include "mapox_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using Stream = ffi::PlatformStream<mpx_stream_t>;

static ffi::Error Fail(const char* what) {
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + ": " + mpx_last_error());
}

// ---------------------------------------------------------------------------------------------- GAE
// Rollout.calculate_advantage (rollout.py:128-157): rewards (B,T) f32, values (B,T+1) f32, next_terminated (B,T) pred
static ffi::Error GaeImpl(mpx_stream_t stream, ffi::Buffer<ffi::F32> rewards, ffi::Buffer<ffi::F32> values,
                          ffi::Buffer<ffi::PRED> next_terminated, double discount, double gae_lambda,
                          ffi::ResultBuffer<ffi::F32> advantages, ffi::ResultBuffer<ffi::F32> targets,
                          ffi::ResultBuffer<ffi::F64> stats, ffi::ResultBuffer<ffi::U8> workspace) {
  const auto dims = rewards.dimensions();
  const int B = static_cast<int>(dims[0]), T = static_cast<int>(dims[1]);
  if (mpx_gae(rewards.typed_data(), values.typed_data(),
              reinterpret_cast<const uint8_t*>(next_terminated.untyped_data()), B, T, discount, gae_lambda,
              advantages->typed_data(), targets->typed_data(), stats->typed_data(), workspace->untyped_data(), stream))
    return Fail("mpx_gae");
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxGae, GaeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::PRED>>()
                                  .Attr<double>("discount")
                                  .Attr<double>("gae_lambda")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::U8>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});

// ---------------------------------------------------------------------------------------------- decode attention
// AttentionBlock.__call__ decode branch (attention.py:136-150).  q (B,1,Nq,D) bf16, caches (B,S,Nkv,D) bf16,
// seq_pos (B,1) s32 -> out (B,1,Nq,D) bf16.
static ffi::Error AttnDecodeImpl(mpx_stream_t stream, ffi::Buffer<ffi::BF16> q, ffi::Buffer<ffi::BF16> kcache,
                                 ffi::Buffer<ffi::BF16> vcache, ffi::Buffer<ffi::S32> seq_pos,
                                 ffi::ResultBuffer<ffi::BF16> out) {
  const auto kd = kcache.dimensions();
  const auto qd = q.dimensions();
  mpx_attn_decode_args a;
  a.q = reinterpret_cast<const mpx_bf16*>(q.untyped_data());
  a.kcache = reinterpret_cast<const mpx_bf16*>(kcache.untyped_data());
  a.vcache = reinterpret_cast<const mpx_bf16*>(vcache.untyped_data());
  a.pos = seq_pos.typed_data();
  a.out = reinterpret_cast<mpx_bf16*>(out->untyped_data());
  a.B = static_cast<int>(kd[0]); a.S = static_cast<int>(kd[1]); a.Nkv = static_cast<int>(kd[2]);
  a.D = static_cast<int>(kd[3]); a.Nq = static_cast<int>(qd[qd.size() - 2]);
  if (mpx_attn_decode(&a, stream)) return Fail("mpx_attn_decode");
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAttnDecode, AttnDecodeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});

// ---------------------------------------------------------------------------------------------- q/k/v + RoPE + cache write
// attention.py:121-134 + :108-114.  The caches are aliased in -> out (input_output_aliases) so the ring write is
// in place, exactly like the donated dynamic_update_slice of the reference.
static ffi::Error QkvRopeDecodeImpl(mpx_stream_t stream, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::BF16> wt_qkv,
                                    ffi::Buffer<ffi::S32> seq_pos, ffi::Buffer<ffi::F32> timescale,
                                    ffi::Buffer<ffi::BF16> kcache_in, ffi::Buffer<ffi::BF16> vcache_in,
                                    int32_t num_heads, ffi::ResultBuffer<ffi::BF16> q,
                                    ffi::ResultBuffer<ffi::BF16> kcache, ffi::ResultBuffer<ffi::BF16> vcache) {
  (void)kcache_in; (void)vcache_in;   // aliased with the results
  const auto kd = kcache->dimensions();
  const auto xd = x.dimensions();
  mpx_qkv_rope_args a{};
  a.M = static_cast<int>(kd[0]); a.H = static_cast<int>(xd[xd.size() - 1]);
  a.Nq = num_heads; a.Nkv = static_cast<int>(kd[2]); a.D = static_cast<int>(kd[3]);
  a.x = reinterpret_cast<const mpx_bf16*>(x.untyped_data()); a.ldx = a.H;
  a.wt_qkv = reinterpret_cast<const mpx_bf16*>(wt_qkv.untyped_data());
  a.pos = seq_pos.typed_data(); a.pos_len = a.M;
  a.rope_timescale = timescale.typed_data();
  a.q = reinterpret_cast<mpx_bf16*>(q->untyped_data()); a.q_row_stride = (long long)a.Nq * a.D;
  a.k = reinterpret_cast<mpx_bf16*>(kcache->untyped_data());
  a.v = reinterpret_cast<mpx_bf16*>(vcache->untyped_data());
  a.cache_S = static_cast<int>(kd[1]);
  a.k_row_stride = a.v_row_stride = (long long)a.cache_S * a.Nkv * a.D;
  if (mpx_qkv_rope(&a, stream)) return Fail("mpx_qkv_rope");
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxQkvRopeDecode, QkvRopeDecodeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("num_heads")
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});

// ---------------------------------------------------------------------------------------------- attention half, one kernel
// network.py:158-165 + attention.py:96-150 for one rollout step: history_norm, q/k/v, RoPE, ring write, attention.
// x is the un-normalised residual stream; the caches are aliased in -> out as above.
static ffi::Error AttnBlockDecodeImpl(mpx_stream_t stream, ffi::Buffer<ffi::BF16> x, ffi::Buffer<ffi::F32> norm_scale,
                                      ffi::Buffer<ffi::BF16> wt_qkv, ffi::Buffer<ffi::S32> seq_pos,
                                      ffi::Buffer<ffi::F32> timescale, ffi::Buffer<ffi::BF16> kcache_in,
                                      ffi::Buffer<ffi::BF16> vcache_in, int32_t num_heads, float eps,
                                      ffi::ResultBuffer<ffi::BF16> out, ffi::ResultBuffer<ffi::BF16> kcache,
                                      ffi::ResultBuffer<ffi::BF16> vcache) {
  (void)kcache_in; (void)vcache_in;
  const auto kd = kcache->dimensions();
  const auto xd = x.dimensions();
  mpx_attn_decode_fused_args a{};
  a.x = reinterpret_cast<const mpx_bf16*>(x.untyped_data());
  a.norm_scale = norm_scale.typed_data(); a.eps = eps;
  a.wt_qkv = reinterpret_cast<const mpx_bf16*>(wt_qkv.untyped_data());
  a.pos = seq_pos.typed_data(); a.rope_timescale = timescale.typed_data();
  a.kcache = reinterpret_cast<mpx_bf16*>(kcache->untyped_data());
  a.vcache = reinterpret_cast<mpx_bf16*>(vcache->untyped_data());
  a.out = reinterpret_cast<mpx_bf16*>(out->untyped_data());
  a.B = static_cast<int>(kd[0]); a.S = static_cast<int>(kd[1]); a.Nkv = static_cast<int>(kd[2]);
  a.D = static_cast<int>(kd[3]); a.Nq = num_heads; a.H = static_cast<int>(xd[xd.size() - 1]);
  if (mpx_attn_decode_fused(&a, stream)) return Fail("mpx_attn_decode_fused");
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxAttnBlockDecode, AttnBlockDecodeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Arg<ffi::Buffer<ffi::BF16>>()
                                  .Attr<int32_t>("num_heads")
                                  .Attr<float>("eps")
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>()
                                  .Ret<ffi::Buffer<ffi::BF16>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});

// ---------------------------------------------------------------------------------------------- HL-Gauss
static ffi::Error HlGaussValueImpl(mpx_stream_t stream, ffi::Buffer<ffi::F32> logits, ffi::Buffer<ffi::F32> centers,
                                   ffi::ResultBuffer<ffi::F32> value) {
  const auto d = logits.dimensions();
  const int nl = static_cast<int>(d[d.size() - 1]);
  const long long n = static_cast<long long>(logits.element_count()) / nl;
  if (mpx_hlgauss_value(logits.typed_data(), n, nl, centers.typed_data(), value->typed_data(), stream))
    return Fail("mpx_hlgauss_value");
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(MpxHlGaussValue, HlGaussValueImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>(),
                              {xla::ffi::Traits::kCmdBufferCompatible});
