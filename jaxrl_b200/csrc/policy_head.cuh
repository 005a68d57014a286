// Row-level policy head for kernels that already hold a finished trunk-output row in registers (the fused GLU kernel of
// the last layer): output RMSNorm, tied fp32 action logits, mask, Philox Gumbel arg-max and its log-probability
// (reference network.py:321-333 + train.py:97-98).  Arithmetic, summation order and Philox counters are those of
// policy_head_sample_kernel (heads.cu) for H = 128: 8 lanes per row, lane `sub` owns features [16 sub, 16 sub + 16) and,
// after the xor-reduction, actions {sub, sub + 8, ...}; the two paths produce bit-identical actions, log-probabilities
// and logits (tests/test_gpu_kernels.py::test_glu_fused_head_matches_separate_kernels).
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

struct PolicyHeadDev {
  const float* norm_scale;        // (128) output_norm.scale
  float eps;
  const float* table;             // (A, 128) fp32 tied action embedding
  const uint8_t* mask;            // (M, A) or null
  const float* gumbel_in;         // (M, A) or null: noise supplied by the caller instead of Philox
  unsigned long long seed, stream_id;
  const unsigned long long* stream_off;
  float* logits_out;              // (M, A) or null
  __nv_bfloat16* xf_out;          // (M, 128) or null: the normalised row
  int32_t* action;                // (M)
  float* log_prob;                // (M)
  int A;
};

// Philox-4x32-10 counter RNG (the same function as in heads.cu): counter = (element index, stream id), key = seed.
__device__ __forceinline__ void ph_philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll 1
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// Called by all 8 lanes of a row group (lanes 8g .. 8g+7 of a converged warp) with xv = the lane's 16 features of the
// (bf16-valued) trunk output row.  h.A <= 8: after the reductions lane `sub` owns action `sub`.  The table must sit in
// shared memory: read from global memory / L2 the six dependent rounds of (4 loads, FMA chain, lane reduction) cost 3 us.
// `tab` / `scale`: the action table and the norm scale staged by the caller (shared memory): action a, lane sub, feature c at
// tab[a * astride + sub * pstride + c]; scale[f] plain.
__device__ __forceinline__ void policy_head_row16(float (&xv)[16], bool live, long long row, int sub, const PolicyHeadDev& h,
                                                  unsigned long long stream_id, const float* tab, int astride, int pstride,
                                                  const float* scale, unsigned long long* dbg = nullptr) {
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 16; ++e) ss = fmaf(xv[e], xv[e], ss);
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  ss += __shfl_xor_sync(0xffffffffu, ss, 4);
  const float inv = rsqrtf(ss / 128.0f + h.eps);
#pragma unroll
  for (int c = 0; c < 16; c += 8) {                      // nnx.RMSNorm: fp32 statistics, result rounded to bf16
    uint32_t w[4];
    const float4 s0 = *reinterpret_cast<const float4*>(scale + sub * 16 + c);
    const float4 s1 = *reinterpret_cast<const float4*>(scale + sub * 16 + c + 4);
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int e = 0; e < 8; e += 2) {
      w[e >> 1] = pack_bf16(xv[c + e] * inv * sc[e], xv[c + e + 1] * inv * sc[e + 1]);
      xv[c + e] = bf16_lo(w[e >> 1]);
      xv[c + e + 1] = bf16_hi(w[e >> 1]);
    }
    if (h.xf_out && live) *reinterpret_cast<uint4*>(h.xf_out + row * 128 + sub * 16 + c) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  dbg_stamp(dbg, 26);
  const int A = h.A;
  float logit = 0.f;                                     // the logit of action `sub` (after the reduction)
#pragma unroll 1
  for (int a = 0; a < A; ++a) {
    const float4* tr = reinterpret_cast<const float4*>(tab + a * astride + sub * pstride);
    const float4 t0 = tr[0], t1 = tr[1], t2 = tr[2], t3 = tr[3];
    float acc = 0.f;
    acc = fmaf(xv[0], t0.x, acc); acc = fmaf(xv[1], t0.y, acc); acc = fmaf(xv[2], t0.z, acc); acc = fmaf(xv[3], t0.w, acc);
    acc = fmaf(xv[4], t1.x, acc); acc = fmaf(xv[5], t1.y, acc); acc = fmaf(xv[6], t1.z, acc); acc = fmaf(xv[7], t1.w, acc);
    acc = fmaf(xv[8], t2.x, acc); acc = fmaf(xv[9], t2.y, acc); acc = fmaf(xv[10], t2.z, acc); acc = fmaf(xv[11], t2.w, acc);
    acc = fmaf(xv[12], t3.x, acc); acc = fmaf(xv[13], t3.y, acc); acc = fmaf(xv[14], t3.z, acc); acc = fmaf(xv[15], t3.w, acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (a == sub) logit = acc;
  }
  dbg_stamp(dbg, 27);
  // this lane's action: masked logit, noise, then the arg-max / max over the 8 lanes
  float best = -INFINITY, mx = -INFINITY, mine = -INFINITY;
  int bi = 0x7fffffff;
  if (sub < A && live) {
    float l = logit;
    if (h.mask && !h.mask[row * A + sub]) l = -3.4028234663852886e38f;
    mine = l;
    if (h.logits_out) h.logits_out[row * A + sub] = l;
    float gn;
    if (h.gumbel_in) {
      gn = h.gumbel_in[row * A + sub];
    } else {
      const unsigned long long e = (unsigned long long)row * A + sub;
      const unsigned long long i = e >> 2;
      uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
      ph_philox4x32_10(c, (uint32_t)h.seed, (uint32_t)(h.seed >> 32));
      const uint32_t r = (e & 3) == 0 ? c[0] : (e & 3) == 1 ? c[1] : (e & 3) == 2 ? c[2] : c[3];
      const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
      gn = -logf(-logf(u));
    }
    const float v = l + gn;
    if (v > best) { best = v; bi = sub; }
    mx = fmaxf(mx, l);
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  dbg_stamp(dbg, 28);
  float se = 0.f, lbi = 0.f;
  if (sub < A && live) {
    se = expf(mine - mx);
    if (sub == bi) lbi = mine;
  }
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    lbi += __shfl_xor_sync(0xffffffffu, lbi, o);
  }
  if (live && sub == 0) {
    h.action[row] = bi;
    h.log_prob[row] = (lbi - mx) - logf(se);
  }
}

}  // namespace mpx
