// Policy / embedding glue around the trunk (reference mapox_trainer/model/network.py:300-343, train.py:97-100,
// :193-212).  All HBM-bound SIMT kernels:
//   embed_combine   x = obs_emb + reward_encoder(reward) + action_embedder.encode(last_action) [+ task]  (:303-309)
//   embed_bwd       its transpose (reward kernel/bias, action/task table gradients)
//   policy_logits   tied logits f32(x) . table^T, action mask -> finfo(f32).min                          (:323-331)
//   policy_sample   Gumbel arg-max + log-prob of the sample (distrax.Categorical, train.py:97-98)
//   gumbel_noise    counter-based (Philox-4x32-10) Gumbel(0,1) noise
//   ppo_policy_loss clipped-ratio actor loss + entropy bonus and d/dlogits                               (train.py:193-212)
//   policy_head_bwd cotangent of x and of the tied table from dlogits
#include "common.cuh"
#include "ptx.cuh"

namespace mpx {

constexpr int HD_WARPS = 8;
constexpr int MAX_A = 32;

// ------------------------------------------------------------------------------------------------ embed
__global__ void __launch_bounds__(256) embed_combine_kernel(
    const __nv_bfloat16* __restrict__ obs_emb, const float* __restrict__ reward, const int32_t* __restrict__ last_action,
    const int32_t* __restrict__ task_ids, const float* __restrict__ w_r, const float* __restrict__ b_r,
    const float* __restrict__ a_table, int A, const float* __restrict__ t_table, int n_tasks,
    __nv_bfloat16* __restrict__ x, long long M, int H) {
  const float sq = bf16_round(sqrtf((float)H));
  const unsigned total = (unsigned)(M * H);                           // host guarantees < 2^31
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned m = i / (unsigned)H;
    const int h = (int)(i % (unsigned)H);
    // reward_encoder: nnx.Linear(1,H) in bf16: bf16(bf16(r) * bf16(w)) then + bf16(b) -> bf16
    float re = bf16_round(bf16_round(reward[m]) * bf16_round(w_r[h]));
    re = bf16_round(re + bf16_round(b_r[h]));
    // Embedder.encode: jnp.take(mode=fill) wraps negatives once, zero-fills out of range; bf16(row) * bf16(sqrt(H))
    int a = last_action[m];
    if (a < 0) a += A;
    float ae = (a >= 0 && a < A) ? bf16_round(a_table[(long long)a * H + h]) : 0.f;
    ae = bf16_round(ae * sq);
    float v = bf16_round(__bfloat162float(obs_emb[i]) + re);
    v = bf16_round(v + ae);
    if (task_ids && t_table) {
      int tk = task_ids[m];
      if (tk < 0) tk += n_tasks;
      float te = (tk >= 0 && tk < n_tasks) ? bf16_round(t_table[(long long)tk * H + h]) : 0.f;
      v = bf16_round(v + bf16_round(te * sq));
    }
    x[i] = __float2bfloat16_rn(v);
  }
}

// Same arithmetic, eight consecutive features per thread (H % 8 == 0, 16-byte aligned rows): one 16-byte load / store
// of the activation row piece, the per-token scalars read once per thread, the fp32 parameter rows as float4 pairs.
__global__ void __launch_bounds__(256) embed_combine8_kernel(
    const __nv_bfloat16* __restrict__ obs_emb, const float* __restrict__ reward, const int32_t* __restrict__ last_action,
    const int32_t* __restrict__ task_ids, const float* __restrict__ w_r, const float* __restrict__ b_r,
    const float* __restrict__ a_table, int A, const float* __restrict__ t_table, int n_tasks,
    __nv_bfloat16* __restrict__ x, long long M, int H) {
  const float sq = bf16_round(sqrtf((float)H));
  const unsigned per_row = (unsigned)H >> 3;
  const unsigned total = (unsigned)(M * per_row);
  const unsigned stride = gridDim.x * blockDim.x;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const unsigned m = i / per_row;
    const int h = (int)(i % per_row) << 3;
    const uint4 ov = *reinterpret_cast<const uint4*>(obs_emb + (size_t)m * H + h);
    const float r = bf16_round(reward[m]);
    int a = last_action[m];
    if (a < 0) a += A;
    const bool a_ok = a >= 0 && a < A;
    int tk = -1;
    if (task_ids && t_table) {
      tk = task_ids[m];
      if (tk < 0) tk += n_tasks;
      if (tk >= n_tasks) tk = -1;
    }
    float o[8], wr[8], br[8], ar[8], tr[8];
    o[0] = bf16_lo(ov.x); o[1] = bf16_hi(ov.x); o[2] = bf16_lo(ov.y); o[3] = bf16_hi(ov.y);
    o[4] = bf16_lo(ov.z); o[5] = bf16_hi(ov.z); o[6] = bf16_lo(ov.w); o[7] = bf16_hi(ov.w);
    *reinterpret_cast<float4*>(wr) = *reinterpret_cast<const float4*>(w_r + h);
    *reinterpret_cast<float4*>(wr + 4) = *reinterpret_cast<const float4*>(w_r + h + 4);
    *reinterpret_cast<float4*>(br) = *reinterpret_cast<const float4*>(b_r + h);
    *reinterpret_cast<float4*>(br + 4) = *reinterpret_cast<const float4*>(b_r + h + 4);
    if (a_ok) {
      *reinterpret_cast<float4*>(ar) = *reinterpret_cast<const float4*>(a_table + (size_t)a * H + h);
      *reinterpret_cast<float4*>(ar + 4) = *reinterpret_cast<const float4*>(a_table + (size_t)a * H + h + 4);
    }
    if (tk >= 0) {
      *reinterpret_cast<float4*>(tr) = *reinterpret_cast<const float4*>(t_table + (size_t)tk * H + h);
      *reinterpret_cast<float4*>(tr + 4) = *reinterpret_cast<const float4*>(t_table + (size_t)tk * H + h + 4);
    }
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float re = bf16_round(r * bf16_round(wr[k]));
      re = bf16_round(re + bf16_round(br[k]));
      const float ae = bf16_round((a_ok ? bf16_round(ar[k]) : 0.f) * sq);
      float t = bf16_round(o[k] + re);
      t = bf16_round(t + ae);
      if (task_ids && t_table) t = bf16_round(t + bf16_round((tk >= 0 ? bf16_round(tr[k]) : 0.f) * sq));
      v[k] = t;
    }
    uint4 out;
    out.x = pack_bf16(v[0], v[1]); out.y = pack_bf16(v[2], v[3]);
    out.z = pack_bf16(v[4], v[5]); out.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(x + (size_t)m * H + h) = out;
  }
}

// Transpose of embed_combine.  Each block reduces `rows_per_block` token rows; thread = feature column.
// d_wr[h] += sum bf16(r) * dx ; d_br[h] += sum dx ; d_a_table[a][h] += sqrt(H) * dx ; same for tasks.
__global__ void embed_bwd_kernel(const __nv_bfloat16* __restrict__ dx, const float* __restrict__ reward,
                                 const int32_t* __restrict__ last_action, const int32_t* __restrict__ task_ids, int A,
                                 int n_tasks, float* __restrict__ d_wr, float* __restrict__ d_br,
                                 float* __restrict__ d_a_table, float* __restrict__ d_t_table, long long M, int H,
                                 int rows_per_block) {
  extern __shared__ float s_tab[];   // [A + n_tasks][H]
  const int ntab = A + (d_t_table ? n_tasks : 0);
  for (int i = threadIdx.x; i < ntab * H; i += blockDim.x) s_tab[i] = 0.f;
  __syncthreads();
  const float sq = bf16_round(sqrtf((float)H));
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + (long long)rows_per_block);
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float awr = 0.f, abr = 0.f;
    for (long long m0 = r0; m0 < r1; m0 += 8) {
      float gv[8], rv[8];
      int av[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {                          // independent loads first (memory-level parallelism)
        const long long m = m0 + u < r1 ? m0 + u : r1 - 1;
        gv[u] = m0 + u < r1 ? __bfloat162float(dx[m * H + h]) : 0.f;
        rv[u] = reward[m];
        av[u] = last_action[m];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float g = gv[u];
        awr += bf16_round(rv[u]) * g;
        abr += g;
        int a = av[u];
        if (a < 0) a += A;
        if (a >= 0 && a < A) s_tab[a * H + h] += g * sq;   // column h is owned by this thread: no race
        if (d_t_table && m0 + u < r1) {
          int tk = task_ids[m0 + u];
          if (tk < 0) tk += n_tasks;
          if (tk >= 0 && tk < n_tasks) s_tab[(A + tk) * H + h] += g * sq;
        }
      }
    }
    atomicAdd(d_wr + h, awr);
    atomicAdd(d_br + h, abr);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A * H; i += blockDim.x)
    if (s_tab[i] != 0.f) atomicAdd(d_a_table + i, s_tab[i]);
  if (d_t_table)
    for (int i = threadIdx.x; i < n_tasks * H; i += blockDim.x)
      if (s_tab[A * H + i] != 0.f) atomicAdd(d_t_table + i, s_tab[A * H + i]);
}

// H % 8 == 0: thread = (row group, 8 consecutive feature columns).  One 16-byte load per row piece, the per-token
// scalars once per 8 features, four rows in flight per thread; every row group accumulates the table rows in its own
// shared-memory copy (no races, no shared atomics), the copies meet at the end.
constexpr int EB_THREADS = 128;
__global__ void __launch_bounds__(EB_THREADS) embed_bwd8_kernel(
    const __nv_bfloat16* __restrict__ dx, const float* __restrict__ reward, const int32_t* __restrict__ last_action,
    const int32_t* __restrict__ task_ids, int A, int n_tasks, float* __restrict__ d_wr, float* __restrict__ d_br,
    float* __restrict__ d_a_table, float* __restrict__ d_t_table, long long M, int H, int rows_per_block) {
  extern __shared__ __align__(16) float s_tab8[];   // [nrg][ntab][H] then [nrg][2][H] for the reward kernel / bias
  const int ncg = H >> 3, nrg = EB_THREADS / ncg;
  const int ntab = A + (d_t_table ? n_tasks : 0);
  const int cg = threadIdx.x % ncg, rg = threadIdx.x / ncg;
  float* tab = s_tab8 + (size_t)rg * ntab * H + 8 * cg;
  for (int i = threadIdx.x; i < nrg * (ntab + 2) * H; i += EB_THREADS) s_tab8[i] = 0.f;
  __syncthreads();
  const float sq = bf16_round(sqrtf((float)H));
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(M, r0 + (long long)rows_per_block);
  float awr[8], abr[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) awr[k] = abr[k] = 0.f;
  if (rg < nrg) {
    for (long long m0 = r0 + rg; m0 < r1; m0 += 4 * nrg) {
      uint4 gv[4];
      float rv[4];
      int av[4], tv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                          // independent loads first
        const long long m = m0 + (long long)u * nrg;
        const bool ok = m < r1;
        gv[u] = ok ? *reinterpret_cast<const uint4*>(dx + m * H + 8 * cg) : make_uint4(0u, 0u, 0u, 0u);
        rv[u] = ok ? reward[m] : 0.f;
        av[u] = ok ? last_action[m] : A;                     // A = out of range: no table row
        tv[u] = (ok && d_t_table) ? task_ids[m] : n_tasks;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float g[8] = {bf16_lo(gv[u].x), bf16_hi(gv[u].x), bf16_lo(gv[u].y), bf16_hi(gv[u].y),
                            bf16_lo(gv[u].z), bf16_hi(gv[u].z), bf16_lo(gv[u].w), bf16_hi(gv[u].w)};
        const float rb = bf16_round(rv[u]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          awr[k] = fmaf(rb, g[k], awr[k]);
          abr[k] += g[k];
        }
        int a = av[u];
        if (a < 0) a += A;
        if (a >= 0 && a < A) {
          float4* t4 = reinterpret_cast<float4*>(tab + a * H);
          float4 t0 = t4[0], t1 = t4[1];
          t0.x = fmaf(g[0], sq, t0.x); t0.y = fmaf(g[1], sq, t0.y); t0.z = fmaf(g[2], sq, t0.z); t0.w = fmaf(g[3], sq, t0.w);
          t1.x = fmaf(g[4], sq, t1.x); t1.y = fmaf(g[5], sq, t1.y); t1.z = fmaf(g[6], sq, t1.z); t1.w = fmaf(g[7], sq, t1.w);
          t4[0] = t0; t4[1] = t1;
        }
        if (d_t_table) {
          int tk = tv[u];
          if (tk < 0) tk += n_tasks;
          if (tk >= 0 && tk < n_tasks) {
            float4* t4 = reinterpret_cast<float4*>(tab + (A + tk) * H);
            float4 t0 = t4[0], t1 = t4[1];
            t0.x = fmaf(g[0], sq, t0.x); t0.y = fmaf(g[1], sq, t0.y); t0.z = fmaf(g[2], sq, t0.z); t0.w = fmaf(g[3], sq, t0.w);
            t1.x = fmaf(g[4], sq, t1.x); t1.y = fmaf(g[5], sq, t1.y); t1.z = fmaf(g[6], sq, t1.z); t1.w = fmaf(g[7], sq, t1.w);
            t4[0] = t0; t4[1] = t1;
          }
        }
      }
    }
    float* wb = s_tab8 + (size_t)nrg * ntab * H + (size_t)rg * 2 * H + 8 * cg;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      wb[k] = awr[k];
      wb[H + k] = abr[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ntab * H; i += EB_THREADS) {
    float v = 0.f;
    for (int r = 0; r < nrg; ++r) v += s_tab8[(size_t)r * ntab * H + i];
    if (v != 0.f) atomicAdd((i < A * H ? d_a_table : d_t_table - A * H) + i, v);
  }
  for (int i = threadIdx.x; i < 2 * H; i += EB_THREADS) {
    float v = 0.f;
    for (int r = 0; r < nrg; ++r) v += s_tab8[(size_t)nrg * ntab * H + (size_t)r * 2 * H + i];
    atomicAdd((i < H ? d_wr : d_br - H) + i, v);
  }
}

// ------------------------------------------------------------------------------------------------ policy logits
// 8 lanes per token (4 tokens per warp): logits[a] = sum_h f32(x[h]) * table[a][h]; masked entries -> finfo(f32).min.
// Each lane owns H/8 consecutive features (16-byte loads); the A partial sums are reduced over the 8 lanes with 3
// xor-shuffles each.
template <int AMAX>
__global__ void __launch_bounds__(HD_WARPS * 32) policy_logits_kernel(
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ table, const uint8_t* __restrict__ mask,
    float* __restrict__ logits, long long M, int H, int A) {
  // [A][8 lane slots][per + 4]: the 4-float pad puts the 8 lanes of a token group on disjoint banks for the 16-byte
  // reads below (unpadded, per = 16 floats = 64 B makes slots 0,2,4,6 collide: 4-way conflicts on every read)
  extern __shared__ float s_table[];
  const int per = H / 8;                                               // features per lane (multiple of 8)
  const int pstride = per + 4, astride = 8 * pstride;
  for (int i = threadIdx.x; i < A * H; i += blockDim.x) {
    const int a = i / H, h = i % H;
    s_table[a * astride + (h / per) * pstride + (h % per)] = table[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7, rsel = lane >> 3;
  // two tokens per 8-lane group and iteration (rows base + rsel and base + 4 + rsel): every table float4 read from
  // shared memory feeds both (the kernel is bound by those reads)
  const long long stride = (long long)gridDim.x * HD_WARPS * 8;
  for (long long base = ((long long)blockIdx.x * HD_WARPS + (threadIdx.x >> 5)) * 8; base < M; base += stride) {
    const long long row0 = base + rsel, row1 = base + 4 + rsel;
    float p0[AMAX], p1[AMAX];
#pragma unroll
    for (int a = 0; a < AMAX; ++a) p0[a] = p1[a] = 0.f;
    for (int c = 0; c < per; c += 8) {
      const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
      const uint4 v0 = row0 < M ? *reinterpret_cast<const uint4*>(x + row0 * H + sub * per + c) : z4;
      const uint4 v1 = row1 < M ? *reinterpret_cast<const uint4*>(x + row1 * H + sub * per + c) : z4;
      const float xa[8] = {bf16_lo(v0.x), bf16_hi(v0.x), bf16_lo(v0.y), bf16_hi(v0.y),
                           bf16_lo(v0.z), bf16_hi(v0.z), bf16_lo(v0.w), bf16_hi(v0.w)};
      const float xb[8] = {bf16_lo(v1.x), bf16_hi(v1.x), bf16_lo(v1.y), bf16_hi(v1.y),
                           bf16_lo(v1.z), bf16_hi(v1.z), bf16_lo(v1.w), bf16_hi(v1.w)};
#pragma unroll
      for (int a = 0; a < AMAX; ++a) {
        if (a < A) {
          const float4* tr = reinterpret_cast<const float4*>(s_table + a * astride + sub * pstride + c);
          const float4 t0 = tr[0], t1 = tr[1];
          p0[a] = fmaf(xa[0], t0.x, p0[a]); p0[a] = fmaf(xa[1], t0.y, p0[a]);
          p0[a] = fmaf(xa[2], t0.z, p0[a]); p0[a] = fmaf(xa[3], t0.w, p0[a]);
          p0[a] = fmaf(xa[4], t1.x, p0[a]); p0[a] = fmaf(xa[5], t1.y, p0[a]);
          p0[a] = fmaf(xa[6], t1.z, p0[a]); p0[a] = fmaf(xa[7], t1.w, p0[a]);
          p1[a] = fmaf(xb[0], t0.x, p1[a]); p1[a] = fmaf(xb[1], t0.y, p1[a]);
          p1[a] = fmaf(xb[2], t0.z, p1[a]); p1[a] = fmaf(xb[3], t0.w, p1[a]);
          p1[a] = fmaf(xb[4], t1.x, p1[a]); p1[a] = fmaf(xb[5], t1.y, p1[a]);
          p1[a] = fmaf(xb[6], t1.z, p1[a]); p1[a] = fmaf(xb[7], t1.w, p1[a]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < AMAX; ++a) {
      if (a < A) {
        p0[a] += __shfl_xor_sync(0xffffffffu, p0[a], 1);
        p1[a] += __shfl_xor_sync(0xffffffffu, p1[a], 1);
        p0[a] += __shfl_xor_sync(0xffffffffu, p0[a], 2);
        p1[a] += __shfl_xor_sync(0xffffffffu, p1[a], 2);
        p0[a] += __shfl_xor_sync(0xffffffffu, p0[a], 4);
        p1[a] += __shfl_xor_sync(0xffffffffu, p1[a], 4);
      }
    }
#pragma unroll
    for (int a = 0; a < AMAX; ++a) {
      if (a < A && (a & 7) == sub) {
        if (row0 < M) {
          float v = p0[a];
          if (mask && !mask[row0 * A + a]) v = -3.4028234663852886e38f;
          logits[row0 * A + a] = v;
        }
        if (row1 < M) {
          float v = p1[a];
          if (mask && !mask[row1 * A + a]) v = -3.4028234663852886e38f;
          logits[row1 * A + a] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ sampling
// action = argmax_a(logits + gumbel) (first maximum, jnp.argmax), log_prob = log_softmax(logits)[action].
__global__ void policy_sample_kernel(const float* __restrict__ logits, const float* __restrict__ gumbel,
                                     int32_t* __restrict__ action, float* __restrict__ log_prob, long long M, int A) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* l = logits + m * A;
  const float* gn = gumbel + m * A;
  float best = -INFINITY, mx = -INFINITY;
  int bi = 0;
  for (int a = 0; a < A; ++a) {
    const float v = l[a] + gn[a];
    if (v > best) { best = v; bi = a; }
    mx = fmaxf(mx, l[a]);
  }
  float s = 0.f;
  for (int a = 0; a < A; ++a) s += expf(l[a] - mx);
  action[m] = bi;
  log_prob[m] = (l[bi] - mx) - logf(s);
}

// Philox-4x32-10 counter RNG -> Gumbel(0,1) = -log(-log(u)).  counter = (element index, stream id), key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
__global__ void gumbel_noise_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                    unsigned long long stream_id, const unsigned long long* __restrict__ stream_off) {
  if (stream_off) stream_id += stream_off[0];
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i * 4 < n; i += stride) {
    uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i * 4 + k < n) {
        const float u = ((float)(c[k] >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
        out[i * 4 + k] = -logf(-logf(u));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ fused policy head
// Rollout-step policy path in one kernel (network.py:321-333 + train.py:97-98): optional output RMSNorm, tied action
// logits, mask, Gumbel arg-max sample and its log-probability.  8 lanes per token as in policy_logits_kernel; after the
// xor-reduction every lane of the group holds all A logits, lane `sub` then owns actions {sub, sub+8, ...} for the noise
// draw (same Philox counters as gumbel_noise_kernel: element index = row*A + a), the arg-max and the log-sum-exp.
constexpr int PH_MAX_PER = 32;   // features per lane: H <= 256
__global__ void __launch_bounds__(HD_WARPS * 32) policy_head_sample_kernel(
    const __nv_bfloat16* __restrict__ x, const float* __restrict__ norm_scale, float eps, const float* __restrict__ table,
    const uint8_t* __restrict__ mask, const float* __restrict__ gumbel_in, unsigned long long seed,
    unsigned long long stream_id, const unsigned long long* __restrict__ stream_off, float* __restrict__ logits_out,
    __nv_bfloat16* __restrict__ xf_out, int32_t* __restrict__ action, float* __restrict__ log_prob, long long M, int H,
    int A) {
  extern __shared__ float s_table[];   // [A][H]
  griddep_launch();
  griddep_wait();
  const int per = H / 8;
  const int pstride = per + 4, astride = 8 * pstride;     // padded like policy_logits_kernel: conflict-free 16-byte reads
  for (int i = threadIdx.x; i < A * H; i += blockDim.x) {
    const int a = i / H, h = i % H;
    s_table[a * astride + (h / per) * pstride + (h % per)] = table[i];
  }
  __syncthreads();
  if (stream_off) stream_id += stream_off[0];
  const int lane = threadIdx.x & 31;
  const int sub = lane & 7, rsel = lane >> 3;
  const long long stride = (long long)gridDim.x * HD_WARPS * 4;
  for (long long base = ((long long)blockIdx.x * HD_WARPS + (threadIdx.x >> 5)) * 4; base < M; base += stride) {
    const long long row = base + rsel;
    const bool live = row < M;
    float xv[PH_MAX_PER];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < PH_MAX_PER; c += 8) {
      if (c < per) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (live) v = *reinterpret_cast<const uint4*>(x + row * H + sub * per + c);
        xv[c] = bf16_lo(v.x); xv[c + 1] = bf16_hi(v.x); xv[c + 2] = bf16_lo(v.y); xv[c + 3] = bf16_hi(v.y);
        xv[c + 4] = bf16_lo(v.z); xv[c + 5] = bf16_hi(v.z); xv[c + 6] = bf16_lo(v.w); xv[c + 7] = bf16_hi(v.w);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(xv[c + e], xv[c + e], ss);
      }
    }
    if (norm_scale) {     // nnx.RMSNorm: fp32 statistics, result rounded to bf16
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
      ss += __shfl_xor_sync(0xffffffffu, ss, 2);
      ss += __shfl_xor_sync(0xffffffffu, ss, 4);
      const float inv = rsqrtf(ss / (float)H + eps);
#pragma unroll
      for (int c = 0; c < PH_MAX_PER; c += 8) {
        if (c < per) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            w[e >> 1] = pack_bf16(xv[c + e] * inv * norm_scale[sub * per + c + e],
                                  xv[c + e + 1] * inv * norm_scale[sub * per + c + e + 1]);
            xv[c + e] = bf16_lo(w[e >> 1]);
            xv[c + e + 1] = bf16_hi(w[e >> 1]);
          }
          if (xf_out && live) *reinterpret_cast<uint4*>(xf_out + row * H + sub * per + c) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    float part[MAX_A];
#pragma unroll
    for (int a = 0; a < MAX_A; ++a) {
      part[a] = 0.f;
      if (a < A) {
        const float4* tr = reinterpret_cast<const float4*>(s_table + a * astride + sub * pstride);
#pragma unroll
        for (int c = 0; c < PH_MAX_PER; c += 4) {
          if (c < per) {
            const float4 tv = tr[c >> 2];
            part[a] = fmaf(xv[c], tv.x, part[a]); part[a] = fmaf(xv[c + 1], tv.y, part[a]);
            part[a] = fmaf(xv[c + 2], tv.z, part[a]); part[a] = fmaf(xv[c + 3], tv.w, part[a]);
          }
        }
        part[a] += __shfl_xor_sync(0xffffffffu, part[a], 1);
        part[a] += __shfl_xor_sync(0xffffffffu, part[a], 2);
        part[a] += __shfl_xor_sync(0xffffffffu, part[a], 4);
      }
    }
    // this lane's actions: masked logit, noise, local arg-max / max
    float best = -INFINITY, mx = -INFINITY, mine[MAX_A / 8];
    int bi = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < MAX_A / 8; ++q) {
      const int a = sub + 8 * q;
      mine[q] = -INFINITY;
      if (a < A && live) {
        float l = part[a];
        if (mask && !mask[row * A + a]) l = -3.4028234663852886e38f;
        mine[q] = l;
        if (logits_out) logits_out[row * A + a] = l;
        float gn;
        if (gumbel_in) {
          gn = gumbel_in[row * A + a];
        } else {
          const unsigned long long e = (unsigned long long)row * A + a;
          const unsigned long long i = e >> 2;
          uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)stream_id, (uint32_t)(stream_id >> 32)};
          philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
          const uint32_t r = (e & 3) == 0 ? c[0] : (e & 3) == 1 ? c[1] : (e & 3) == 2 ? c[2] : c[3];
          const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);
          gn = -logf(-logf(u));
        }
        const float v = l + gn;
        if (v > best) { best = v; bi = a; }      // ascending a within the lane: keeps the first maximum
        mx = fmaxf(mx, l);
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    float se = 0.f, lbi = 0.f;
#pragma unroll
    for (int q = 0; q < MAX_A / 8; ++q) {
      const int a = sub + 8 * q;
      if (a < A && live) {
        se += expf(mine[q] - mx);
        if (a == bi) lbi = mine[q];
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      se += __shfl_xor_sync(0xffffffffu, se, o);
      lbi += __shfl_xor_sync(0xffffffffu, lbi, o);
    }
    if (live && sub == 0) {
      action[row] = bi;
      log_prob[row] = (lbi - mx) - logf(se);
    }
  }
}

// ------------------------------------------------------------------------------------------------ PPO policy loss
// Per token (train.py:193-212, distrax.Categorical log_prob / entropy):
//   lp = log_softmax(logits); ratio = exp(lp[a] - old_lp); actor = -min(ratio*adv, clip(ratio,1-e,1+e)*adv)
//   entropy_loss = -H = sum p*lp;  total policy part = mean(actor) + ent_coef * mean(entropy_loss)
// dlogits = (1/M) * [ g * (onehot(a) - p) + ent_coef * p * (lp + H) ],  g = d actor / d lp[a]
// partial sums (actor, entropy_loss) per block -> fixed-order finalize (deterministic).
__global__ void __launch_bounds__(256) ppo_policy_loss_kernel(
    const float* __restrict__ logits, const int32_t* __restrict__ actions, const float* __restrict__ old_lp,
    const float* __restrict__ adv, float clip_eps, float ent_coef, const float* __restrict__ ent_coef_dev, float inv_m,
    float* __restrict__ dlogits, double* __restrict__ partials, long long M, int A) {
  __shared__ double s_a[256], s_e[256];
  if (ent_coef_dev) ent_coef = ent_coef_dev[0];
  double acc_a = 0.0, acc_e = 0.0;
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; m < M; m += stride) {
    const float* l = logits + m * A;
    float lp[MAX_A];
    float mx = -INFINITY;
    for (int a = 0; a < A; ++a) { lp[a] = l[a]; mx = fmaxf(mx, lp[a]); }
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += expf(lp[a] - mx);
    const float lse = logf(s);
    float negH = 0.f;
    for (int a = 0; a < A; ++a) {
      lp[a] = (lp[a] - mx) - lse;
      const float p = expf(lp[a]);
      negH += (p > 0.f) ? p * lp[a] : 0.f;
    }
    const int act = actions[m];
    const float ratio = expf(lp[act] - old_lp[m]);
    const float ad = adv[m];
    const float pg1 = ratio * ad;
    const float rc = fminf(fmaxf(ratio, 1.0f - clip_eps), 1.0f + clip_eps);
    const float pg2 = rc * ad;
    acc_a += (double)(-fminf(pg1, pg2));
    acc_e += (double)negH;
    // d(-min(pg1,pg2)) / d lp[act]
    const bool inside = (ratio > 1.0f - clip_eps) && (ratio < 1.0f + clip_eps);
    float g;
    if (pg1 < pg2) g = -pg1;                 // unclipped branch selected: d(ratio*adv)/dlp = ratio*adv
    else if (pg1 > pg2) g = inside ? -pg1 : 0.f;
    else g = inside ? -pg1 : -0.5f * pg1;    // tie: jnp.minimum splits the cotangent
    if (dlogits) {
      for (int a = 0; a < A; ++a) {
        const float p = expf(lp[a]);
        const float onehot = (a == act) ? 1.f : 0.f;
        const float dent = (p > 0.f) ? p * (lp[a] - negH) : 0.f;   // d(sum p lp)/dz = p (lp + H), H = -negH
        dlogits[m * A + a] = inv_m * (g * (onehot - p) + ent_coef * dent);
      }
    }
  }
  s_a[threadIdx.x] = acc_a;
  s_e[threadIdx.x] = acc_e;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_e[threadIdx.x] += s_e[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = s_a[0];
    partials[2 * blockIdx.x + 1] = s_e[0];
  }
}
__global__ void ppo_loss_finalize_kernel(const double* __restrict__ partials, int n, double inv_m,
                                         float* __restrict__ out /* [2]: actor_loss, entropy_loss */) {
  __shared__ double s0[256], s1[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { a += partials[2 * i]; b += partials[2 * i + 1]; }
  s0[threadIdx.x] = a;
  s1[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { s0[threadIdx.x] += s0[threadIdx.x + o]; s1[threadIdx.x] += s1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = (float)(s0[0] * inv_m); out[1] = (float)(s1[0] * inv_m); }
}

// ------------------------------------------------------------------------------------------------ policy head bwd
// dx[m][h] = bf16( sum_a dlogits[m][a] * table[a][h] ) (+ dx_other in bf16 if given) ;
// dtable[a][h] += sum_m dlogits[m][a] * f32(x[m][h]).   Masked logits are constants: dlogits is zero there.
// One warp per token row, lane owns HPL = H/32 feature columns; dtable partials live in registers (AMAX x HPL),
// meet in shared memory per block, one atomic per table entry per block.
template <int AMAX, int HPL>
__global__ void __launch_bounds__(HD_WARPS * 32, (AMAX * HPL <= 32) ? 3 : 1) policy_head_bwd_kernel(
    const float* __restrict__ dlogits, const __nv_bfloat16* __restrict__ x, const float* __restrict__ table,
    const __nv_bfloat16* __restrict__ dx_other, __nv_bfloat16* __restrict__ dx, float* __restrict__ dtable, long long M,
    int A) {
  constexpr int H = HPL * 32;
  extern __shared__ float s_buf[];   // [A][H] table, then [A][H] reduction buffer
  float* s_table = s_buf;
  float* s_red = s_buf + A * H;
  for (int i = threadIdx.x; i < A * H; i += blockDim.x) {
    s_table[i] = table[i];
    s_red[i] = 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float dt[AMAX][HPL];
#pragma unroll
  for (int a = 0; a < AMAX; ++a)
#pragma unroll
    for (int j = 0; j < HPL; ++j) dt[a][j] = 0.f;
  long long row = (long long)blockIdx.x * HD_WARPS + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * HD_WARPS;
  if constexpr (HPL == 4) {
    // H = 128: lane owns the 4 CONSECUTIVE columns [4 lane, 4 lane + 4) - one 8-byte load per operand row, one 8-byte
    // store, the table as one float4 per action.  The operands of the next TWO rows are in flight while the current
    // row is multiplied (the loop is otherwise one row of global-load latency per iteration).
    uint2 xq[2], oq[2];
    float dq[2];
    auto fetch = [&](long long r, int slot) {
      xq[slot] = make_uint2(0u, 0u); oq[slot] = make_uint2(0u, 0u); dq[slot] = 0.f;
      if (r < M) {
        xq[slot] = *reinterpret_cast<const uint2*>(x + r * H + 4 * lane);
        if (dx_other) oq[slot] = *reinterpret_cast<const uint2*>(dx_other + r * H + 4 * lane);
        dq[slot] = lane < A ? dlogits[r * A + lane] : 0.f;
      }
    };
    fetch(row, 0);
    fetch(row + stride, 1);
    for (; row < M; row += 2 * stride) {
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const long long r = row + ph * stride;
        const uint2 xw = xq[ph], ow = oq[ph];
        const float dl_mine = dq[ph];
        fetch(r + 2 * stride, ph);
        if (r < M) {
          const float xv[4] = {bf16_lo(xw.x), bf16_hi(xw.x), bf16_lo(xw.y), bf16_hi(xw.y)};
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int a = 0; a < AMAX; ++a) {
            if (a < A) {
              const float d = __shfl_sync(0xffffffffu, dl_mine, a);
              const float4 tv = *reinterpret_cast<const float4*>(s_table + a * H + 4 * lane);
              acc[0] = fmaf(d, tv.x, acc[0]); acc[1] = fmaf(d, tv.y, acc[1]);
              acc[2] = fmaf(d, tv.z, acc[2]); acc[3] = fmaf(d, tv.w, acc[3]);
#pragma unroll
              for (int j = 0; j < 4; ++j) dt[a][j] = fmaf(d, xv[j], dt[a][j]);
            }
          }
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = bf16_round(acc[j]);
          if (dx_other) {
            o[0] += bf16_lo(ow.x); o[1] += bf16_hi(ow.x); o[2] += bf16_lo(ow.y); o[3] += bf16_hi(ow.y);
          }
          *reinterpret_cast<uint2*>(dx + r * H + 4 * lane) = make_uint2(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]));
        }
      }
    }
#pragma unroll
    for (int a = 0; a < AMAX; ++a)
      if (a < A)
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&s_red[a * H + 4 * lane + j], dt[a][j]);
  } else {
    // the next row's operands are fetched while the current row is multiplied
    float xn[HPL], on[HPL], dn = 0.f;
    auto fetch = [&](long long r) {
      if (r < M) {
#pragma unroll
        for (int j = 0; j < HPL; ++j) {
          xn[j] = __bfloat162float(x[r * H + j * 32 + lane]);
          on[j] = dx_other ? __bfloat162float(dx_other[r * H + j * 32 + lane]) : 0.f;
        }
        dn = lane < A ? dlogits[r * A + lane] : 0.f;
      }
    };
    fetch(row);
    for (; row < M; row += stride) {
      float xv[HPL], ov[HPL], acc[HPL];
#pragma unroll
      for (int j = 0; j < HPL; ++j) {
        xv[j] = xn[j];
        ov[j] = on[j];
        acc[j] = 0.f;
      }
      const float dl_mine = dn;
      fetch(row + stride);
#pragma unroll
      for (int a = 0; a < AMAX; ++a) {
        if (a < A) {
          const float d = __shfl_sync(0xffffffffu, dl_mine, a);
#pragma unroll
          for (int j = 0; j < HPL; ++j) {
            acc[j] = fmaf(d, s_table[a * H + j * 32 + lane], acc[j]);
            dt[a][j] = fmaf(d, xv[j], dt[a][j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < HPL; ++j) {
        float o = bf16_round(acc[j]);
        if (dx_other) o += ov[j];
        dx[row * H + j * 32 + lane] = __float2bfloat16_rn(o);
      }
    }
#pragma unroll
    for (int a = 0; a < AMAX; ++a)
      if (a < A)
#pragma unroll
        for (int j = 0; j < HPL; ++j) atomicAdd(&s_red[a * H + j * 32 + lane], dt[a][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < A * H; i += blockDim.x) atomicAdd(dtable + i, s_red[i]);
}

}  // namespace mpx

using namespace mpx;

static int grid_for(long long n, int per_block, int waves = 16) {
  long long g = (n + per_block - 1) / per_block;
  const long long cap = (long long)num_sms() * waves;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

extern "C" int mpx_embed_combine(const mpx_bf16* obs_emb, const float* reward, const int32_t* last_action,
                                 const int32_t* task_ids, const float* w_reward, const float* b_reward,
                                 const float* action_table, int A, const float* task_table, int n_tasks, mpx_bf16* x,
                                 long long M, int H, mpx_stream_t stream) {
  MPX_REQUIRE(obs_emb && reward && last_action && w_reward && b_reward && action_table && x && M > 0 && H > 0 && A > 0,
              "mpx_embed_combine: bad arguments");
  MPX_REQUIRE(M * H < (1LL << 31), "mpx_embed_combine: too many elements for 32-bit indexing");
  const bool vec8 = (H % 8 == 0) &&
                    ((reinterpret_cast<uintptr_t>(obs_emb) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_reward) |
                      reinterpret_cast<uintptr_t>(b_reward) | reinterpret_cast<uintptr_t>(action_table) |
                      reinterpret_cast<uintptr_t>(task_table)) & 15) == 0;
  if (vec8)
    embed_combine8_kernel<<<grid_for(M * (H / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(obs_emb), reward, last_action, task_ids, w_reward, b_reward, action_table,
        A, task_table, n_tasks, reinterpret_cast<__nv_bfloat16*>(x), M, H);
  else
    embed_combine_kernel<<<grid_for(M * H, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(obs_emb), reward, last_action, task_ids, w_reward, b_reward, action_table,
        A, task_table, n_tasks, reinterpret_cast<__nv_bfloat16*>(x), M, H);
  return check_launch("mpx_embed_combine");
}

extern "C" int mpx_embed_bwd(const mpx_bf16* dx, const float* reward, const int32_t* last_action,
                             const int32_t* task_ids, int A, int n_tasks, float* d_w_reward, float* d_b_reward,
                             float* d_action_table, float* d_task_table, long long M, int H, mpx_stream_t stream) {
  MPX_REQUIRE(dx && reward && last_action && d_w_reward && d_b_reward && d_action_table && M > 0 && H > 0 && A > 0,
              "mpx_embed_bwd: bad arguments");
  MPX_REQUIRE(!d_task_table || task_ids, "mpx_embed_bwd: task table gradient requested without task ids");
  {
    const int ntab = A + (d_task_table ? n_tasks : 0);
    const int ncg = H / 8;
    const size_t smem8 = (H % 8 == 0 && ncg <= EB_THREADS && EB_THREADS % ncg == 0)
                             ? (size_t)(EB_THREADS / ncg) * (ntab + 2) * H * sizeof(float) : 0;
    if (smem8 > 0 && smem8 <= 48 * 1024 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0) {
      int rows8 = (int)((M + (long long)num_sms() * 4 - 1) / ((long long)num_sms() * 4));
      if (rows8 < 128) rows8 = 128;
      embed_bwd8_kernel<<<(int)((M + rows8 - 1) / rows8), EB_THREADS, smem8, (cudaStream_t)stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(dx), reward, last_action, task_ids, A, n_tasks, d_w_reward, d_b_reward,
          d_action_table, d_task_table, M, H, rows8);
      return check_launch("mpx_embed_bwd");
    }
  }
  int rows = (int)((M + (long long)num_sms() * 8 - 1) / ((long long)num_sms() * 8));
  if (rows < 64) rows = 64;
  const int grid = (int)((M + rows - 1) / rows);
  const size_t smem = (size_t)(A + (d_task_table ? n_tasks : 0)) * H * sizeof(float);
  MPX_REQUIRE(smem <= 48 * 1024, "mpx_embed_bwd: tables too large for shared memory (%zu B)", smem);
  embed_bwd_kernel<<<grid, H < 256 ? 128 : 256, smem, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dx), reward, last_action, task_ids, A, n_tasks, d_w_reward, d_b_reward,
      d_action_table, d_task_table, M, H, rows);
  return check_launch("mpx_embed_bwd");
}

extern "C" int mpx_policy_logits(const mpx_bf16* x, const float* table, const uint8_t* action_mask, float* logits,
                                 long long M, int H, int A, mpx_stream_t stream) {
  MPX_REQUIRE(x && table && logits && M > 0 && H > 0, "mpx_policy_logits: bad arguments");
  MPX_REQUIRE(A > 0 && A <= MAX_A, "mpx_policy_logits: A=%d not in [1,%d]", A, MAX_A);
  MPX_REQUIRE(H % 64 == 0, "mpx_policy_logits: H=%d must be a multiple of 64", H);
  const size_t smem = (size_t)A * (H + 32) * sizeof(float);      // 4-float pad per lane slot
  MPX_REQUIRE(smem <= 48 * 1024, "mpx_policy_logits: table too large for shared memory");
  if (A <= 8)
    policy_logits_kernel<8><<<grid_for(M, HD_WARPS * 8, 8), HD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), table, action_mask, logits, M, H, A);
  else
    policy_logits_kernel<MAX_A><<<grid_for(M, HD_WARPS * 8, 8), HD_WARPS * 32, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), table, action_mask, logits, M, H, A);
  return check_launch("mpx_policy_logits");
}

extern "C" int mpx_policy_sample(const float* logits, const float* gumbel, int32_t* action, float* log_prob,
                                 long long M, int A, mpx_stream_t stream) {
  MPX_REQUIRE(logits && gumbel && action && log_prob && M > 0 && A > 0, "mpx_policy_sample: bad arguments");
  policy_sample_kernel<<<(int)((M + 127) / 128), 128, 0, (cudaStream_t)stream>>>(logits, gumbel, action, log_prob, M, A);
  return check_launch("mpx_policy_sample");
}

extern "C" int mpx_gumbel_noise(float* out, long long n, unsigned long long seed, unsigned long long stream_id,
                                const unsigned long long* stream_offset, mpx_stream_t stream) {
  MPX_REQUIRE(out && n > 0, "mpx_gumbel_noise: bad arguments");
  gumbel_noise_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, stream_id,
                                                                                    stream_offset);
  return check_launch("mpx_gumbel_noise");
}

extern "C" size_t mpx_ppo_policy_loss_workspace_bytes(long long M) {
  return (size_t)grid_for(M > 0 ? M : 1, 256, 8) * 2 * sizeof(double);
}

extern "C" int mpx_ppo_policy_loss(const float* logits, const int32_t* actions, const float* old_log_prob,
                                   const float* advantages, float clip_eps, float entropy_coef,
                                   const float* entropy_coef_dev, float* losses, float* dlogits, void* workspace,
                                   long long M, int A, mpx_stream_t stream) {
  MPX_REQUIRE(logits && actions && old_log_prob && advantages && losses && workspace && M > 0,
              "mpx_ppo_policy_loss: bad arguments");
  MPX_REQUIRE(A > 0 && A <= MAX_A, "mpx_ppo_policy_loss: A=%d not in [1,%d]", A, MAX_A);
  const int grid = grid_for(M, 256, 8);
  double* partials = reinterpret_cast<double*>(workspace);
  ppo_policy_loss_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, actions, old_log_prob, advantages, clip_eps,
                                                                 entropy_coef, entropy_coef_dev, 1.0f / (float)M,
                                                                 dlogits, partials, M, A);
  ppo_loss_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, grid, 1.0 / (double)M, losses);
  return check_launch("mpx_ppo_policy_loss");
}

template <int AMAX>
static int launch_phb(const float* dlogits, const __nv_bfloat16* x, const float* table, const __nv_bfloat16* dx_other,
                      __nv_bfloat16* dx, float* dtable, long long M, int H, int A, cudaStream_t s) {
  const size_t smem = (size_t)2 * A * H * sizeof(float);
  if (smem > 48 * 1024) {
    set_error("mpx_policy_head_bwd: table too large for shared memory");
    return MPX_ERR_UNSUPPORTED;
  }
  long long g = (M + HD_WARPS - 1) / HD_WARPS;
  const long long cap = (long long)num_sms() * 6;
  const int grid = (int)(g < cap ? g : cap);
  switch (H / 32) {
    case 4: policy_head_bwd_kernel<AMAX, 4><<<grid, HD_WARPS * 32, smem, s>>>(dlogits, x, table, dx_other, dx, dtable, M, A); break;
    case 8: policy_head_bwd_kernel<AMAX, 8><<<grid, HD_WARPS * 32, smem, s>>>(dlogits, x, table, dx_other, dx, dtable, M, A); break;
    default: set_error("mpx_policy_head_bwd: H=%d unsupported (128 or 256)", H); return MPX_ERR_UNSUPPORTED;
  }
  return check_launch("mpx_policy_head_bwd");
}

extern "C" int mpx_policy_head_bwd(const float* dlogits, const mpx_bf16* x, const float* table, const mpx_bf16* dx_other,
                                   mpx_bf16* dx, float* dtable, long long M, int H, int A, mpx_stream_t stream) {
  MPX_REQUIRE(dlogits && x && table && dx && dtable && M > 0, "mpx_policy_head_bwd: bad arguments");
  MPX_REQUIRE(A > 0 && A <= MAX_A, "mpx_policy_head_bwd: unsupported A=%d", A);
  auto X = reinterpret_cast<const __nv_bfloat16*>(x);
  auto DO = reinterpret_cast<const __nv_bfloat16*>(dx_other);
  auto DX = reinterpret_cast<__nv_bfloat16*>(dx);
  cudaStream_t s = (cudaStream_t)stream;
  if (A <= 8) return launch_phb<8>(dlogits, X, table, DO, DX, dtable, M, H, A, s);
  if (A <= 16) return launch_phb<16>(dlogits, X, table, DO, DX, dtable, M, H, A, s);
  return launch_phb<32>(dlogits, X, table, DO, DX, dtable, M, H, A, s);
}

extern "C" int mpx_policy_head_sample(const mpx_bf16* x, const float* norm_scale, float eps, const float* table,
                                      const uint8_t* action_mask, const float* gumbel_in, unsigned long long seed,
                                      unsigned long long stream_id, const unsigned long long* stream_offset,
                                      float* logits_out, mpx_bf16* xf_out, int32_t* action, float* log_prob, long long M,
                                      int H, int A, mpx_stream_t stream) {
  MPX_REQUIRE(x && table && action && log_prob && M > 0, "mpx_policy_head_sample: bad arguments");
  MPX_REQUIRE(A > 0 && A <= MAX_A, "mpx_policy_head_sample: A=%d not in [1,%d]", A, MAX_A);
  MPX_REQUIRE(H % 64 == 0 && H <= 8 * PH_MAX_PER, "mpx_policy_head_sample: H=%d must be a multiple of 64, <= %d", H,
              8 * PH_MAX_PER);
  const size_t smem = (size_t)A * (H + 32) * sizeof(float);      // 4-float pad per lane slot
  MPX_REQUIRE(smem <= 48 * 1024, "mpx_policy_head_sample: table too large for shared memory");
  cudaError_t le = launch_ex(policy_head_sample_kernel, dim3(grid_for(M, HD_WARPS * 4, 8)), dim3(HD_WARPS * 32), smem,
                             (cudaStream_t)stream, 0, 16, reinterpret_cast<const __nv_bfloat16*>(x), norm_scale, eps, table,
                             action_mask, gumbel_in, seed, stream_id, stream_offset, logits_out,
                             reinterpret_cast<__nv_bfloat16*>(xf_out), action, log_prob, M, H, A);
  if (le != cudaSuccess) {
    set_error("launch policy_head_sample_kernel: %s", cudaGetErrorString(le));
    return MPX_ERR_CUDA;
  }
  return check_launch("mpx_policy_head_sample");
}
