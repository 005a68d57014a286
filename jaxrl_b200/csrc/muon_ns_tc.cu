// Newton-Schulz contractions of the Muon step (optax.contrib.muon as chained by the reference, mapox_trainer/
// optimizer.py:25-37) on 5th-generation tensor cores with fp32-faithful operands.
//
// The iteration  A = X X^T ;  B = c1 A + c2 A A ;  X' = c0 X + B X  runs on fp32 matrices (optax keeps the parameter
// dtype).  Each fp32 operand element is split into a bf16 pair  x = hi + lo  (hi = bf16(x), lo = bf16(x - hi)) while it
// is staged into shared memory, and every product is three tcgen05.mma passes into one fp32 TMEM accumulator:
// hi.hi + hi.lo + lo.hi  (the dropped lo.lo term is 2^-16 relative - tighter than the TF32 that XLA's default matmul
// precision gives the reference on GPU).  One launch per contraction, one CTA per (matrix, 128 x 128 output tile), the
// whole K extent per CTA: no split-K, no atomics, no zeroing of A - the result is bit-reproducible.
//   mode 0: A (r,r)  = X X^T * inv^2      P = X rows m (K-major)   Q = X rows n (K-major)          K = c
//   mode 1: B (r,r)  = c2 A A + c1 A      P = A rows m             Q = A rows n (A is symmetric)   K = r
//   mode 2: X'(r,c)  = (B X + c0 X) * inv P = B rows m             Q = X[k][n] (MN-major operand)  K = r
// (inv = 1 / (||X||_F + eps) on the first iteration: the Frobenius normalisation is folded in.)
// 256 threads: all eight warps fetch fp32 tiles from global memory one K block (64) ahead into registers, split them and
// write the SWIZZLE_128B operand tiles (two stages of 4 x 16 KB); warp 0 then issues the 12 MMAs of the block; the eight
// warps share the epilogue (two per TMEM lane quarter, 64 columns each).
#include "common.cuh"
#include "muon.h"
#include "ptx.cuh"

namespace mpx {

constexpr int NT_THREADS = 256;
constexpr int NT_TILE = 16384;                       // [128 x 64] bf16
constexpr int NT_STAGE = 4 * NT_TILE;                // P hi, P lo, Q hi, Q lo
constexpr int NT_SMEM = 1024 + 2 * NT_STAGE + 64;

struct NsTcParams {
  const MuonMat* mats;
  const float* Xin;
  float* Xout;
  float* Abuf;
  float* Bbuf;
  const float* norm2;
  int first;
  float c0, c1, c2, eps;
};

// 4 consecutive floats at p[0..3]; elements at index >= nvalid read as zero.  vec: p is 16-byte aligned.
__device__ __forceinline__ float4 ns_ld4(const float* p, int nvalid, bool vec) {
  if (nvalid >= 4 && vec) return *reinterpret_cast<const float4*>(p);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid > 0) v.x = p[0];
  if (nvalid > 1) v.y = p[1];
  if (nvalid > 2) v.z = p[2];
  if (nvalid > 3) v.w = p[3];
  return v;
}

// 8 floats -> one 16-byte chunk of the hi tile and one of the lo tile
__device__ __forceinline__ void ns_split_store(uint32_t hi_addr, uint32_t lo_addr, const float4& a, const float4& b) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float h0 = bf16_round(v[2 * e]), h1 = bf16_round(v[2 * e + 1]);
    h[e] = pack_bf16(h0, h1);
    l[e] = pack_bf16(v[2 * e] - h0, v[2 * e + 1] - h1);
  }
  sts_v4(hi_addr, make_uint4(h[0], h[1], h[2], h[3]));
  sts_v4(lo_addr, make_uint4(l[0], l[1], l[2], l[3]));
}

template <int MODE>
__global__ void __launch_bounds__(NT_THREADS, 1) muon_ns_tc_kernel(const NsTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 2 * NT_STAGE);   // [2] MMAs reading the stage have completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 2);

  const MuonMat mt = p.mats[blockIdx.z];
  const int r = mt.r, c = mt.c;
  const int M = r, N = (MODE == 2) ? c : r, K = (MODE == 0) ? c : r;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
  if (m0 >= M || n0 >= N) return;                     // uniform per CTA (the grid is sized for the largest matrix)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tid = threadIdx.x;

  // operand sources (row-major fp32):  P[m][k] = Ps[m * ldp + k];  modes 0 / 1: Q[n][k] = Qs[n * ldq + k];  mode 2: Q[k][n]
  const float* Ps;  int ldp;
  const float* Qs;  int ldq;
  if (MODE == 0) { Ps = p.Xin + mt.xoff; ldp = c; Qs = Ps; ldq = c; }
  else if (MODE == 1) { Ps = p.Abuf + mt.aoff; ldp = r; Qs = Ps; ldq = r; }
  else { Ps = p.Bbuf + mt.aoff; ldp = r; Qs = p.Xin + mt.xoff; ldq = c; }
  const bool vecp = ((reinterpret_cast<uintptr_t>(Ps) & 15) == 0) && (ldp % 4 == 0);
  const bool vecq = ((reinterpret_cast<uintptr_t>(Qs) & 15) == 0) && (ldq % 4 == 0);

  if (tid == 0) {
    mbar_init(&done[0], 1);
    mbar_init(&done[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tacc = *tmem_slot;

  // ---- loader geometry
  // K-major tile [128 rows x 64 k]: thread -> (row = tid / 2, 32 consecutive k starting at 32 * (tid & 1))
  const int prow = tid >> 1, pk = (tid & 1) * 32;
  // MN-major tile [64 k x 128 n]:   thread -> (k = tid / 4, 32 consecutive n starting at 32 * (tid & 3))
  const int qk = tid >> 2, qn = (tid & 3) * 32;
  auto fetch_p = [&](int kb, float4 (&pr)[8]) {
    const int k0 = kb * 64, m = m0 + prow;
    const float* src = Ps + (long long)m * ldp + k0 + pk;
#pragma unroll
    for (int j = 0; j < 8; ++j) pr[j] = ns_ld4(src + 4 * j, m < M ? K - (k0 + pk + 4 * j) : 0, vecp);
  };
  auto fetch_q = [&](int kb, float4 (&qr)[8]) {
    const int k0 = kb * 64;
    if (MODE != 2) {
      const int n = n0 + prow;
      const float* src = Qs + (long long)n * ldq + k0 + pk;
#pragma unroll
      for (int j = 0; j < 8; ++j) qr[j] = ns_ld4(src + 4 * j, n < N ? K - (k0 + pk + 4 * j) : 0, vecq);
    } else {
      const int k = k0 + qk;
      const float* src = Qs + (long long)k * ldq + n0 + qn;
#pragma unroll
      for (int j = 0; j < 8; ++j) qr[j] = ns_ld4(src + 4 * j, k < K ? N - (n0 + qn + 4 * j) : 0, vecq);
    }
  };
  auto store_p = [&](int s, const float4 (&pr)[8]) {
    const uint32_t row = smem_u32(smem) + s * NT_STAGE + prow * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {                             // 16-byte chunk (pk / 8 + j) of the row, 128-byte swizzle
      const uint32_t off = (uint32_t)(((pk >> 3) + j) ^ (prow & 7)) << 4;
      ns_split_store(row + off, row + NT_TILE + off, pr[2 * j], pr[2 * j + 1]);
    }
  };
  auto store_q = [&](int s, const float4 (&qr)[8]) {
    const uint32_t base = smem_u32(smem) + s * NT_STAGE + 2 * NT_TILE;
    if (MODE != 2) {
      const uint32_t row = base + prow * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = (uint32_t)(((pk >> 3) + j) ^ (prow & 7)) << 4;
        ns_split_store(row + off, row + NT_TILE + off, qr[2 * j], qr[2 * j + 1]);
      }
    } else {
      // MN-major: two 64-column atoms of [64 k rows x 128 B]; chunk = 8 consecutive n
      const uint32_t row = base + (qn >> 6) * 8192 + qk * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = (uint32_t)((((qn & 63) >> 3) + j) ^ (qk & 7)) << 4;
        ns_split_store(row + off, row + NT_TILE + off, qr[2 * j], qr[2 * j + 1]);
      }
    }
  };
  constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, MODE == 2 ? 1 : 0);
  const int nkb = (K + 63) / 64;
  // diagonal tile of X X^T: both operands are the same rows of X - one set of tiles serves as P and Q
  const bool alias = (MODE == 0) && (m0 == n0);
  auto issue = [&](int kb) {                                  // the 12 MMAs of K block kb (stage kb & 1), warp 0
    const int s = kb & 1;
    tc_fence_after();
    const uint32_t ph = smem_u32(smem) + s * NT_STAGE, pl = ph + NT_TILE;
    const uint32_t qh = alias ? ph : ph + 2 * NT_TILE, ql = qh + NT_TILE;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t a_h = umma_smem_desc_sw128(ph + 32 * ks, 16, 1024), a_l = umma_smem_desc_sw128(pl + 32 * ks, 16, 1024);
      const uint64_t b_h = MODE == 2 ? umma_smem_desc_sw128(qh + 2048 * ks, 8192, 1024) : umma_smem_desc_sw128(qh + 32 * ks, 16, 1024);
      const uint64_t b_l = MODE == 2 ? umma_smem_desc_sw128(ql + 2048 * ks, 8192, 1024) : umma_smem_desc_sw128(ql + 32 * ks, 16, 1024);
      umma_bf16_1(tacc, a_h, b_h, idesc, (kb | ks) != 0 ? 1u : 0u);
      umma_bf16_1(tacc, a_h, b_l, idesc, 1u);
      umma_bf16_1(tacc, a_l, b_h, idesc, 1u);
    }
    umma_commit_1(&done[s]);
  };
  if (alias) {
    // long K (= c), one operand: the fp32 rows are fetched TWO K blocks ahead (two register sets)
    float4 pr[2][8];
    fetch_p(0, pr[0]);
    if (nkb > 1) fetch_p(1, pr[1]);
    for (int kb = 0; kb < nkb; kb += 2) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k = kb + u;
        if (k < nkb) {
          if (k >= 2) mbar_wait(&done[u], ((k >> 1) - 1) & 1);   // the MMAs of block k-2 have finished reading this stage
          store_p(u, pr[u]);
          if (k + 2 < nkb) fetch_p(k + 2, pr[u]);
          fence_proxy_async_smem();
          __syncthreads();
          if (warp == 0) issue(k);
        }
      }
    }
  } else {
    float4 pr[8], qr[8];
    fetch_p(0, pr);
    fetch_q(0, qr);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb & 1;
      if (kb >= 2) mbar_wait(&done[s], ((kb >> 1) - 1) & 1);
      store_p(s, pr);
      store_q(s, qr);
      if (kb + 1 < nkb) {                                     // in flight while the MMAs below run
        fetch_p(kb + 1, pr);
        fetch_q(kb + 1, qr);
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (warp == 0) issue(kb);
    }
  }
  mbar_wait(&done[(nkb - 1) & 1], ((nkb - 1) >> 1) & 1);      // commits complete in order: every MMA has landed
  tc_fence_after();

  // ---- epilogue: thread = row (TMEM lane), 64 columns per warp
  const int q4 = warp & 3, half = warp >> 2;
  const int m = m0 + q4 * 32 + lane;
  const float inv = p.first ? 1.0f / (sqrtf(p.norm2[blockIdx.z]) + p.eps) : 1.0f;
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    uint32_t v[32];
    tmem_ld_32x32(tacc + ((uint32_t)(q4 * 32) << 16) + 64 * half + 32 * cc, v);
    tmem_ld_wait();
    const int nb = n0 + 64 * half + 32 * cc;
    if (m < M) {
      float* dst;  const float* add;  int ld;
      if (MODE == 0) { dst = p.Abuf + mt.aoff; add = nullptr; ld = r; }
      else if (MODE == 1) { dst = p.Bbuf + mt.aoff; add = p.Abuf + mt.aoff; ld = r; }
      else { dst = p.Xout + mt.xoff; add = p.Xin + mt.xoff; ld = c; }
      const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) && (ld % 4 == 0) &&
                       (add == nullptr || (reinterpret_cast<uintptr_t>(add) & 15) == 0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = nb + 4 * j;
        if (n >= N) break;
        float o[4];
        const float4 ad = add ? ns_ld4(add + (long long)m * ld + n, N - n, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float a4[4] = {ad.x, ad.y, ad.z, ad.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float acc = __uint_as_float(v[4 * j + e]);
          if (MODE == 0) o[e] = acc * inv * inv;
          else if (MODE == 1) o[e] = p.c2 * acc + p.c1 * a4[e];
          else o[e] = (acc + p.c0 * a4[e]) * inv;
        }
        float* d = dst + (long long)m * ld + n;
        if (vec && n + 4 <= N) *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
        else
          for (int e = 0; e < 4 && n + e < N; ++e) d[e] = o[e];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tacc, 128);
}

template <int MODE>
static int ns_tc_launch_mode(const NsTcParams& p, dim3 grid, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(muon_ns_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(muon_ns_tc): %s", cudaGetErrorString(e));
      return MPX_ERR_CUDA;
    }
    attr_set = true;
  }
  muon_ns_tc_kernel<MODE><<<grid, NT_THREADS, NT_SMEM, s>>>(p);
  return MPX_OK;
}

// One Newton-Schulz iteration (three launches) for every matrix of the table; X' lands in xout.
int muon_ns_tc_iteration(const MuonMat* mats, int nmat, int max_r, int max_c, const float* xin, float* xout, float* Abuf,
                         float* Bbuf, const float* norm2, int first, float c0, float c1, float c2, float eps, cudaStream_t s) {
  NsTcParams p;
  p.mats = mats; p.Xin = xin; p.Xout = xout; p.Abuf = Abuf; p.Bbuf = Bbuf; p.norm2 = norm2;
  p.first = first; p.c0 = c0; p.c1 = c1; p.c2 = c2; p.eps = eps;
  const int tr = ceil_div(max_r, 128), tc = ceil_div(max_c, 128);
  int rc = ns_tc_launch_mode<0>(p, dim3(tr, tr, nmat), s);
  if (rc) return rc;
  p.first = 0;
  rc = ns_tc_launch_mode<1>(p, dim3(tr, tr, nmat), s);
  if (rc) return rc;
  p.first = first;
  return ns_tc_launch_mode<2>(p, dim3(tc, tr, nmat), s);
}

}  // namespace mpx
