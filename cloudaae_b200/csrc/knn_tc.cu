// knn_tc.cu — k nearest neighbours of every point of a cloud with the Gram matrix on the tensor cores.
//
// Same result as knn.cu (reference utils/tf_util.py:597-632: D = |x_i|^2 - 2 x_i.x_j + |x_j|^2 by a batched
// matmul, then tf.nn.top_k): the k smallest of the fp32 FFMA distances, ascending, ties to the lower index —
// bit-identical indices to knn_kernel, which evaluates all n^2 distances on the FP32 pipe (1.07 GFLOP per
// 64-channel layer = 59 us at B = 128, FMA pipe 40 % busy).  Here the n x n distance matrix is only SCREENED on
// the tensor cores and the exact FFMA arithmetic is spent on a shortlist of ~k+3 candidates per row:
//
//   1. one CTA per cloud (n <= 256, c <= 64) stages the centred rows A = X - mean (knn.cu: the distances are
//      evaluated on centred features) as two split-precision halves H = tf32(A), L = A - H (exact) in the K-major
//      SWIZZLE_128B layout tcgen05 reads;
//   2. one thread issues G~ = H H^T + H L^T + L H^T (tcgen05.mma kind::tf32, M = 128, N = 256, two row tiles,
//      48 instructions) into 512 TMEM columns: |G~ - x_i.x_j| <= 2^-16 |x_i||x_j| (a single TF32 pass would
//      be 2^-9: a shortlist of dozens);
//   3. thread i owns query row i = one TMEM lane: a bound T >= (k-th smallest d~) from the minima of 32 strided
//      column groups, then every j with d~_ij <= T + 2 eps_i joins the row's shortlist, eps_i = 2^-15 (|x_i|^2 + max|x_j|^2)
//      bounding |d~ - d| for both the tensor-core and the fp32-chain rounding — so the shortlist provably
//      contains the exact top k;
//   4. the shortlist is re-evaluated with knn_kernel's arithmetic (fmaf chain over the channels in ascending
//      order, d = (|x_i|^2 + (-2 acc)) + |x_j|^2) and ranked by (distance, index).
// Rows whose shortlist overflows its 48 slots (mass ties: padded duplicate points) sort what they hold and keep
// streaming: every further candidate that passes the screen is evaluated exactly and inserted into the row's sorted
// k-list, so the kernel is complete for every input and no second launch is needed.
#include "tcgen05.cuh"

namespace caae {

constexpr int KT_THREADS = 256;
constexpr int KT_ROWS = 256;             // points per cloud handled by one CTA (rows past n are zero)
constexpr int KT_CAP = 48;               // shortlist capacity per row (real batches: mean 11, p99 16, max 46 — tools/knn_shortlist_stats.py)
constexpr int KT_LD = KT_CAP + 1;        // shortlist row pitch in elements (odd: lane i, entry e -> bank (i + e) % 32 for the keys)
constexpr uint32_t KT_SLAB = KT_ROWS * 128;   // one 32-channel slab: 256 rows x 128 bytes

__device__ __forceinline__ uint32_t kt_sortable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int CPAD>   // channels padded to 32 or 64 (zero columns change neither the Gram matrix nor the fmaf chains)
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc_kernel(int n, int c, int k, const float* __restrict__ x, int ldx, int* __restrict__ idx_out,
              const int* __restrict__ skip, int* __restrict__ dbg_cnt) {
  pdl_wait();
  constexpr int NS = CPAD / 32;                       // slabs
  constexpr uint32_t OPER = NS * KT_SLAB;             // bytes of one operand half (H or L)
  extern __shared__ uint8_t kt_smem_raw[];
  const uint32_t base = (smem_u32(kt_smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = kt_smem_raw + (base - smem_u32(kt_smem_raw));
  const uint32_t sH = base, sL = base + OPER;
  float* nrm = reinterpret_cast<float*>(gbase + 2 * OPER);                    // [256] squared norms (+inf past n)
  // shortlist rows have a pitch of KT_LD = 33 words: lane i reads word e of row i from bank (i + e) % 32 — no
  // conflicts (a pitch of 32 put all 32 lanes of the ranking loop on ONE bank: 80 us per cloud instead of ~20)
  uint32_t* skey = reinterpret_cast<uint32_t*>(gbase + 2 * OPER + 1024);      // [256][KT_LD] exact keys of the shortlist
  uint16_t* sidx = reinterpret_cast<uint16_t*>(skey + KT_ROWS * KT_LD);       // [256][KT_LD] candidate indices (+1 dump slot)
  const uint32_t bar = base + 2 * OPER + 1024 + ((KT_ROWS * KT_LD * 6 + 15) & ~15); // mbarrier (8 bytes), TMEM slot behind it
  const uint32_t tmem_slot = bar + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(gbase + (tmem_slot - base));
  __shared__ float s_red[KT_THREADS / 32];
  __shared__ __align__(16) float mu[CPAD];
  __shared__ float mu_part[4 * CPAD];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cloud = blockIdx.x;
  if (skip != nullptr && skip[cloud] != 0) return;    // heavily padded cloud: routed to the all-pairs kernel (caae_knn_part)
  const float* __restrict__ xc = x + (size_t)cloud * n * ldx;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // ---- stage the raw rows into the L half: 16-byte chunk q of row r sits at slab(q / 8) + r * 128 + ((q % 8) ^ (r % 8)) * 16
  // (all of a thread's global loads are issued before its first shared-memory store: one L2 round trip per batch of 8)
  constexpr int CHUNKS = CPAD / 4;
  const bool vec_ok = ((c & 3) == 0) && ((ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  constexpr int PER_THREAD = KT_ROWS * CHUNKS / KT_THREADS, BATCH = 8;
  static_assert(PER_THREAD % BATCH == 0, "staging batches");
  auto chunk_off = [](int r, int q) {
    return (uint32_t)(q >> 3) * KT_SLAB + (uint32_t)r * 128u + (uint32_t)(((q & 7) ^ (r & 7)) << 4);
  };
#pragma unroll 1
  for (int e0 = 0; e0 < PER_THREAD; e0 += BATCH) {
    float4 v[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int e = tid + (e0 + u) * KT_THREADS;
      const int r = e / CHUNKS, q = e - r * CHUNKS;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < n) {
        const float* src = xc + (size_t)r * ldx + 4 * q;
        if (vec_ok) {
          if (4 * q < c) v[u] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          if (4 * q + 0 < c) v[u].x = __ldg(src + 0);
          if (4 * q + 1 < c) v[u].y = __ldg(src + 1);
          if (4 * q + 2 < c) v[u].z = __ldg(src + 2);
          if (4 * q + 3 < c) v[u].w = __ldg(src + 3);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int e = tid + (e0 + u) * KT_THREADS;
      const int r = e / CHUNKS, q = e - r * CHUNKS;
      *reinterpret_cast<float4*>(gbase + (sL - base) + chunk_off(r, q)) = v[u];
    }
  }
  __syncthreads();
  // ---- cloud mean per channel, knn_kernel's summation order: four interleaved partial sums over ascending rows
  // (lane = channel: the 32 lanes of a warp read 8 different chunks x 4 elements of one row = 32 distinct banks)
  for (int e = tid; e < 4 * CPAD; e += KT_THREADS) {
    const int ch = e % CPAD, part = e / CPAD;
    float sm = 0.f;
    for (int r = part; r < n; r += 4)
      sm = __fadd_rn(sm, *reinterpret_cast<const float*>(gbase + (sL - base) + chunk_off(r, ch >> 2) + (uint32_t)((ch & 3) << 2)));
    mu_part[part * CPAD + ch] = sm;
  }
  __syncthreads();
  if (tid < CPAD)
    mu[tid] = __fdiv_rn(__fadd_rn(__fadd_rn(mu_part[tid], mu_part[CPAD + tid]), __fadd_rn(mu_part[2 * CPAD + tid], mu_part[3 * CPAD + tid])),
                        (float)n);
  __syncthreads();
  // ---- centre, split: a = x - mu (exactly what knn_kernel evaluates), H = tf32(a), L = a - H; rows past n stay zero
#pragma unroll 2
  for (int i = 0; i < PER_THREAD; ++i) {
    const int e = tid + i * KT_THREADS;
    const int r = e / CHUNKS, q = e - r * CHUNKS;
    const uint32_t off = chunk_off(r, q);
    float4 v = *reinterpret_cast<const float4*>(gbase + (sL - base) + off);
    const float4 m4 = *reinterpret_cast<const float4*>(mu + 4 * q);
    if (r < n) { v.x = __fsub_rn(v.x, m4.x); v.y = __fsub_rn(v.y, m4.y); v.z = __fsub_rn(v.z, m4.z); v.w = __fsub_rn(v.w, m4.w); }
    if (4 * q + 0 >= c) v.x = 0.f;
    if (4 * q + 1 >= c) v.y = 0.f;
    if (4 * q + 2 >= c) v.z = 0.f;
    if (4 * q + 3 >= c) v.w = 0.f;
    const float4 h = make_float4(tf32_rne(v.x), tf32_rne(v.y), tf32_rne(v.z), tf32_rne(v.w));
    const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
    st_shared_v4(sH + off, h);
    st_shared_v4(sL + off, l);
  }
  fence_proxy_async_smem();                       // generic-proxy writes -> visible to the tensor core's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  // ---- Gram matrix: rows [128 mt, 128 mt + 128) x all 256 columns into TMEM columns [256 mt, 256 mt + 256)
  if (tid == 0) {
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      uint32_t first = 0;
#pragma unroll
      for (int term = 0; term < 3; ++term) {       // H H^T, H L^T, L H^T
        const uint32_t sa = (term == 2) ? sL : sH, sb = (term == 1) ? sL : sH;
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adesc = make_smem_desc(sa + s * KT_SLAB + mt * (128u * 128u) + kk * 32u, 16u, 1024u, 2u);
            const uint64_t bdesc = make_smem_desc(sb + s * KT_SLAB + kk * 32u, 16u, 1024u, 2u);
            umma_tf32(tmem_base + (uint32_t)(mt * 256), adesc, bdesc, idesc, first);
            first = 1;
          }
      }
    }
    umma_commit(bar);
  }

  // ---- while the tensor core works: this thread's own row (exact values) and its squared norm, knn_kernel's chain
  const int row = tid;
  float xr[CPAD];
#pragma unroll
  for (int q = 0; q < CHUNKS; ++q) {
    const uint32_t off = (uint32_t)(q >> 3) * KT_SLAB + (uint32_t)row * 128u + (uint32_t)(((q & 7) ^ (row & 7)) << 4);
    const float4 h = *reinterpret_cast<const float4*>(gbase + (sH - base) + off);
    const float4 l = *reinterpret_cast<const float4*>(gbase + (sL - base) + off);
    xr[4 * q + 0] = h.x + l.x; xr[4 * q + 1] = h.y + l.y; xr[4 * q + 2] = h.z + l.z; xr[4 * q + 3] = h.w + l.w;
  }
  float sq = 0.f;
#pragma unroll
  for (int ch = 0; ch < CPAD; ++ch) sq = fmaf(xr[ch], xr[ch], sq);
  nrm[row] = (row < n) ? sq : INFINITY;
  float wmax = (row < n) ? sq : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) s_red[warp] = wmax;
  __syncthreads();
  float nmax = s_red[0];
#pragma unroll
  for (int w = 1; w < KT_THREADS / 32; ++w) nmax = fmaxf(nmax, s_red[w]);
  const float margin = 6.103515625e-05f * (sq + nmax);          // 2 eps_i, eps_i = 2^-15 (|x_i|^2 + max_j |x_j|^2)

  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);

  // ---- pass 1: T >= k-th smallest of d~_j = |x_j|^2 - 2 G~_ij (the row constant |x_i|^2 is left out)
  // 32 STRIDED groups, group s = {j : j % 32 == s}: at least k groups have a minimum <= T, hence at least k candidates.
  // (Groups of 8 CONSECUTIVE indices gave a uselessly loose T whenever near-duplicate points sit at consecutive
  // indices — a YCB model with 72 repeats of one point put 73 candidates on 72 shortlists and made that CTA the
  // kernel's tail; strided, a run of near-duplicates lands in every group.)
  float gm[32];
#pragma unroll
  for (int s = 0; s < 32; ++s) gm[s] = INFINITY;
#pragma unroll 1
  for (int ch8 = 0; ch8 < 8; ++ch8) {
    uint32_t r[32];
    tmem_ld32(trow + (uint32_t)(ch8 * 32), r);
    tmem_wait_ld(r);
#pragma unroll
    for (int j4 = 0; j4 < 32; j4 += 4) {
      const float4 n4 = *reinterpret_cast<const float4*>(nrm + ch8 * 32 + j4);       // (same address in every lane: broadcast)
      gm[j4 + 0] = fminf(gm[j4 + 0], fmaf(-2.f, __uint_as_float(r[j4 + 0]), n4.x));
      gm[j4 + 1] = fminf(gm[j4 + 1], fmaf(-2.f, __uint_as_float(r[j4 + 1]), n4.y));
      gm[j4 + 2] = fminf(gm[j4 + 2], fmaf(-2.f, __uint_as_float(r[j4 + 2]), n4.z));
      gm[j4 + 3] = fminf(gm[j4 + 3], fmaf(-2.f, __uint_as_float(r[j4 + 3]), n4.w));
    }
  }
  float T = -INFINITY;
  for (int rr = 0; rr < k; ++rr) {     // k rounds of "smallest group minimum above the previous one" (ties retire together)
    float c4[4] = {INFINITY, INFINITY, INFINITY, INFINITY};     // four independent chains (latency, not issue, bounds this)
#pragma unroll
    for (int s = 0; s < 32; ++s) c4[s & 3] = fminf(c4[s & 3], (gm[s] > T) ? gm[s] : INFINITY);
    T = fminf(fminf(c4[0], c4[1]), fminf(c4[2], c4[3]));
  }
  const float thresh = fminf(T + margin, 3.0e38f);     // finite: masked / padded candidates (d~ = +inf) never pass

  // ---- pass 2: shortlist
  int cnt = 0;
  uint16_t* my_idx = sidx + row * KT_LD;
#pragma unroll 1
  for (int ch8 = 0; ch8 < 8; ++ch8) {
    uint32_t r[32];
    tmem_ld32(trow + (uint32_t)(ch8 * 32), r);
    tmem_wait_ld(r);
#pragma unroll
    for (int j4 = 0; j4 < 32; j4 += 4) {
      const float4 n4 = *reinterpret_cast<const float4*>(nrm + ch8 * 32 + j4);
      const float nv[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d = fmaf(-2.f, __uint_as_float(r[j4 + u]), nv[u]);
        // Six instructions per candidate: the index is stored UNCONDITIONALLY at the list's current end and the end only
        // advances when the candidate passes (rows past n carry |x_j|^2 = +inf and thresh is finite, so they never do);
        // entries past the capacity land in the dump slot KT_CAP of the row.
        my_idx[min(cnt, KT_CAP)] = (uint16_t)(ch8 * 32 + j4 + u);
        cnt += (d <= thresh) ? 1 : 0;
      }
    }
  }

  // The tensor core is done with the operands: the L half becomes the EXACT rows X = H + L (this thread's own row is
  // in registers), so the re-evaluation below costs one LDS.128 per four channels instead of two plus four adds.
#pragma unroll
  for (int q = 0; q < CHUNKS; ++q) {
    const uint32_t off = (uint32_t)(q >> 3) * KT_SLAB + (uint32_t)row * 128u + (uint32_t)(((q & 7) ^ (row & 7)) << 4);
    *reinterpret_cast<float4*>(gbase + (sL - base) + off) = make_float4(xr[4 * q], xr[4 * q + 1], xr[4 * q + 2], xr[4 * q + 3]);
  }
  __syncthreads();

  if (dbg_cnt != nullptr && row < n) dbg_cnt[(size_t)cloud * n + row] = cnt;   // shortlist sizes (tools/knn_shortlist_stats.py)
  // ---- pass 3: exact distances of the shortlist (knn_kernel's arithmetic), rank by (distance, index)
  auto exact_key = [&](int j) -> uint32_t {
    float acc = 0.f;
#pragma unroll
    for (int q = 0; q < CHUNKS; ++q) {
      const float4 xa = *reinterpret_cast<const float4*>(gbase + (sL - base) + (uint32_t)(q >> 3) * KT_SLAB + (uint32_t)j * 128u +
                                                         (uint32_t)(((q & 7) ^ (j & 7)) << 4));
      acc = fmaf(xr[4 * q + 0], xa.x, acc); acc = fmaf(xr[4 * q + 1], xa.y, acc);
      acc = fmaf(xr[4 * q + 2], xa.z, acc); acc = fmaf(xr[4 * q + 3], xa.w, acc);
    }
    return kt_sortable(__fadd_rn(__fadd_rn(sq, __fmul_rn(-2.f, acc)), nrm[j]));
  };
  uint32_t* my_key = skey + row * KT_LD;
  int* out = idx_out + ((size_t)cloud * n + min(row, n - 1)) * k;
  const bool overflow = (row < n) && (cnt > KT_CAP);
  const int m = min(cnt, KT_CAP);
  if (row < n) {
    for (int e = 0; e < m; e += 2) {              // two candidates at a time: two independent fmaf chains
      const int ja = (int)my_idx[e], jb = (int)my_idx[min(e + 1, m - 1)];
      float acca = 0.f, accb = 0.f;
#pragma unroll
      for (int q = 0; q < CHUNKS; ++q) {
        const uint32_t qa = (uint32_t)(q >> 3) * KT_SLAB + (uint32_t)(((q & 7) ^ (ja & 7)) << 4);
        const uint32_t qb = (uint32_t)(q >> 3) * KT_SLAB + (uint32_t)(((q & 7) ^ (jb & 7)) << 4);
        const float4 xa = *reinterpret_cast<const float4*>(gbase + (sL - base) + (uint32_t)ja * 128u + qa);
        const float4 xb = *reinterpret_cast<const float4*>(gbase + (sL - base) + (uint32_t)jb * 128u + qb);
        acca = fmaf(xr[4 * q + 0], xa.x, acca); accb = fmaf(xr[4 * q + 0], xb.x, accb);
        acca = fmaf(xr[4 * q + 1], xa.y, acca); accb = fmaf(xr[4 * q + 1], xb.y, accb);
        acca = fmaf(xr[4 * q + 2], xa.z, acca); accb = fmaf(xr[4 * q + 2], xb.z, accb);
        acca = fmaf(xr[4 * q + 3], xa.w, acca); accb = fmaf(xr[4 * q + 3], xb.w, accb);
      }
      my_key[e] = kt_sortable(__fadd_rn(__fadd_rn(sq, __fmul_rn(-2.f, acca)), nrm[ja]));
      if (e + 1 < m) my_key[e + 1] = kt_sortable(__fadd_rn(__fadd_rn(sq, __fmul_rn(-2.f, accb)), nrm[jb]));
    }
    if (!overflow) {
      // k selection rounds over the m entries (not m^2 rank counting: one row with m = 46 made its warp, and with it
      // the whole CTA, 2.5x slower than the median).  Entries are in ascending index order, so (key, position) is the
      // (distance, index) order; a round takes the smallest pair above the previous one.
      if (m <= 16) {
        // the common case (mean 11.6, p99 16 entries): keys in registers, every pair compared once, no dependent chain
        uint32_t kr[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) kr[e] = e < m ? my_key[e] : 0xffffffffu;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          int rank = 0;
#pragma unroll
          for (int f = 0; f < 16; ++f) {
            if (f < e) rank += (kr[f] <= kr[e]) ? 1 : 0;        // equal keys: the lower position (= lower index) first
            else if (f > e) rank += (kr[f] < kr[e]) ? 1 : 0;
          }
          if (e < m && rank < k) out[rank] = (int)my_idx[e];
        }
      } else {
      unsigned long long prev = 0ull;
      for (int r = 0; r < k; ++r) {
        unsigned long long best = ~0ull;
        for (int e = 0; e < m; ++e) {
          const unsigned long long v = ((unsigned long long)my_key[e] << 8) | (unsigned)e;
          best = (v > prev || r == 0) && v < best ? v : best;
        }
        if (best != ~0ull) out[r] = (int)my_idx[(int)(best & 0xffull)];
        prev = best;
      }
      }
    } else {
      // mass ties (padded duplicate points): sort the 32 collected entries in place — stable insertion sort, so equal
      // keys keep their ascending index order — and keep streaming below
      for (int e = 1; e < KT_CAP; ++e) {
        const uint32_t ke = my_key[e];
        const uint16_t je = my_idx[e];
        int p = e;
        while (p > 0 && my_key[p - 1] > ke) { my_key[p] = my_key[p - 1]; my_idx[p] = my_idx[p - 1]; --p; }
        my_key[p] = ke; my_idx[p] = je;
      }
    }
  }
  // ---- overflow rows: every further candidate that passes the screen (ascending index, from `resume`) is evaluated
  // exactly and inserted into the row's sorted k-list; a later candidate with an equal key loses (higher index).
  // Warp-uniform loop (tcgen05.ld is warp-collective); warps without an overflow row skip it.
  if (__any_sync(0xffffffffu, overflow)) {
    // the k-th exact distance W held so far tightens the screen as the list improves: d_j >= d~_j + |x_i|^2 - eps, so
    // a candidate with d~_j > W - |x_i|^2 + 2 eps cannot enter the list and needs no exact evaluation
    auto unsort = [](uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); };
    float tight = overflow ? fminf(thresh, unsort(my_key[k - 1]) - sq + margin) : thresh;
    int seen = 0;                          // candidates that passed the original screen so far: the first KT_CAP are in the list
#pragma unroll 1
    for (int ch8 = 0; ch8 < 8; ++ch8) {
      uint32_t r[32];
      tmem_ld32(trow + (uint32_t)(ch8 * 32), r);
      tmem_wait_ld(r);
      if (overflow) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int jg = ch8 * 32 + j;
          const float d = fmaf(-2.f, __uint_as_float(r[j]), nrm[jg]);
          const bool fresh = seen >= KT_CAP;
          seen += (d <= thresh) ? 1 : 0;
          if (fresh && d <= tight) {
            const uint32_t key = exact_key(jg);
            if (key < my_key[k - 1]) {
              int p = k - 1;
              while (p > 0 && my_key[p - 1] > key) { my_key[p] = my_key[p - 1]; my_idx[p] = my_idx[p - 1]; --p; }
              my_key[p] = key; my_idx[p] = (uint16_t)jg;
              tight = fminf(thresh, unsort(my_key[k - 1]) - sq + margin);
            }
          }
        }
      }
    }
    if (overflow)
      for (int e = 0; e < k; ++e) out[e] = (int)my_idx[e];
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int CPAD>
static int launch_knn_tc(int b, int n, int c, int k, const float* x, int ldx, int* idx, const int* skip, cudaStream_t s,
                         int* dbg_cnt = nullptr) {
  const size_t smem = 2 * (size_t)(CPAD / 32) * KT_SLAB + 1024 + ((KT_ROWS * KT_LD * 6 + 15) & ~15) + 64 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(knn_tc_kernel<CPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  caae::launch(knn_tc_kernel<CPAD>, b, KT_THREADS, smem, s, n, c, k, x, ldx, idx, skip, dbg_cnt);
  return CAAE_LAUNCH_STATUS();
}

// 1 when the tensor-core screen applies: one cloud per CTA, k within the shortlist.
bool knn_tc_applicable(int n, int c, int k) { return n <= KT_ROWS && c <= 64 && k <= 24 && k <= n; }

int knn_tc_launch(int b, int n, int c, int k, const float* x, int ldx, int* idx, const int* skip, cudaStream_t s) {
  return c <= 32 ? launch_knn_tc<32>(b, n, c, k, x, ldx, idx, skip, s) : launch_knn_tc<64>(b, n, c, k, x, ldx, idx, skip, s);
}

// flags[cloud] = 1 when at least n/8 rows of the cloud repeat an earlier row exactly (the synthesis pads clouds with
// random repeats of their visible points, utils/hidden_point_removal.py:38-40).  Mass ties are the one input the
// tensor-core screen handles badly (every copy passes the screen and needs an exact evaluation); such clouds take the
// all-pairs kernel instead.  A routing decision only: both kernels return the same indices.  One CTA per cloud;
// exact comparison through an open-addressing table keyed by a hash of the row.
constexpr int KC_SLOTS = 1024;
__global__ void __launch_bounds__(256)
knn_classify_kernel(int n, int c, const float* __restrict__ x, int ldx, int* __restrict__ flags) {
  pdl_wait();
  __shared__ int table[KC_SLOTS];
  __shared__ int dups;
  const int cloud = blockIdx.x, tid = threadIdx.x;
  const float* __restrict__ xc = x + (size_t)cloud * n * ldx;
  for (int s = tid; s < KC_SLOTS; s += 256) table[s] = -1;
  if (tid == 0) dups = 0;
  __syncthreads();
  for (int r = tid; r < n; r += 256) {
    const float* row = xc + (size_t)r * ldx;
    uint32_t h = 2166136261u;
    for (int ch = 0; ch < c; ++ch) h = (h ^ __float_as_uint(__ldg(row + ch))) * 16777619u;
    int slot = (int)(h & (KC_SLOTS - 1));
    for (int probe = 0; probe < KC_SLOTS; ++probe) {
      const int prev = atomicCAS(&table[slot], -1, r);
      if (prev == -1) break;                                   // first row with this content (so far)
      const float* other = xc + (size_t)prev * ldx;
      bool same = true;
      for (int ch = 0; ch < c && same; ++ch) same = __float_as_uint(__ldg(other + ch)) == __float_as_uint(__ldg(row + ch));
      if (same) { atomicAdd(&dups, 1); break; }
      slot = (slot + 1) & (KC_SLOTS - 1);
    }
  }
  __syncthreads();
  if (tid == 0) flags[cloud] = (dups * 8 >= n) ? 1 : 0;
}

int knn_tc_debug_counts(int b, int n, int c, int k, const float* x, int ldx, int* idx, int* counts, cudaStream_t s) {
  return c <= 32 ? launch_knn_tc<32>(b, n, c, k, x, ldx, idx, nullptr, s, counts) : launch_knn_tc<64>(b, n, c, k, x, ldx, idx, nullptr, s, counts);
}

int knn_classify_launch(int b, int n, int c, const float* x, int ldx, int* flags, cudaStream_t s) {
  caae::launch(knn_classify_kernel, b, 256, 0, s, n, c, x, ldx, flags);
  return CAAE_LAUNCH_STATUS();
}

}  // namespace caae
