// gemm_tcgen05.cu — TF32 tensor-core GEMM for sm_100a: tcgen05.mma with the accumulator in TMEM,
// operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through an mbarrier pipeline.
//
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N]),  fp32 in / fp32 out, TF32 multiply, fp32 accumulate.
//
// This is the only dense contraction of the model that matters for time: conv2d 320->1024
// `dgcnn_agg` (reference models/pointnet_ycb_23_decoder_4.py:410-413 via utils/tf_util.py:161) with
// M = B*N = 32768 rows — forward, data gradient and weight gradient are 64 GFLOP of the ~70 GFLOP
// step — plus the FC stack and the EdgeConv projections.  All four operand layouts are native:
// an operand stored with K contiguous is consumed as a K-major UMMA operand, one stored with M/N
// contiguous as an MN-major operand (instruction-descriptor major bits + a different TMA box), so
// neither the weight gradient X^T dY (both operands MN-major) nor the data gradient dY W^T needs a
// transposed copy in HBM.
//
// CTA = one 128 x 128 output tile (x one K split).  192 threads, warp-specialised:
//   warp 0      TMA producer (one elected lane): per 32-wide K block, 16 KB of A and 16 KB of B
//   warp 1      TMEM allocator + MMA issuer (one elected lane): 4 x tcgen05.mma 128x128x8 per K block
//   warps 2..5  epilogue: tcgen05.ld 32x32b (each warp its own TMEM lane quadrant) -> bias /
//               accumulate / split-K reduction -> global stores of whole 128-byte lines
// 3 stages x 32 KB of shared memory and 128 TMEM columns per CTA, so two CTAs are co-resident per
// SM and one CTA's epilogue overlaps the other's main loop.
//
// Shared-memory operand layouts (fp32 elements, T = 4 elements per 16 bytes):
//   K-major  (SWIZZLE_128B): row r (an M or N index) at r*128 B holds 32 consecutive K elements;
//            8-row groups every 1024 B (SBO = 1024).  One TMA box {32 (K), 128 (rows)} per stage.
//            A K-step of 8 elements advances the descriptor start address by 32 B.
//   MN-major (SWIZZLE_128B with 32-byte atoms — the only MN-major layout tcgen05 accepts for 32-bit
//            operands; TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, descriptor layout type 1):
//            K-row r at r*128 B holds 32 consecutive M/N elements; 4-row groups every 512 B (SBO);
//            the next 32 M/N elements start LBO = 4096 B later.  Four TMA boxes {32 (M/N), 32 (K rows)}
//            per stage.  A K-step of 8 rows advances the start address by 1024 B.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tcgen05.cuh"

namespace caae {

constexpr int TBM = 128, TBN = 128, TBK = 32;  // CTA tile; TBK fp32 = one 128-byte swizzle row
constexpr int TSTAGES = 3;
constexpr int TTHREADS = 192;
constexpr uint32_t TSTAGE_A = TBM * TBK * 4, TSTAGE_B = TBN * TBK * 4;  // 16 KB each
constexpr uint32_t TSMEM_BYTES = TSTAGES * (TSTAGE_A + TSTAGE_B) + 1024 /*align*/ + 256 /*barriers*/;
constexpr uint32_t TMEM_COLS = 128;

struct GemmParams {
  int M, N, K;
  float* C;
  int ldc;
  const float* bias;
  int accumulate;   // C += (splits == 1)
  int kb_per_split; // K blocks per blockIdx.z
  int atomic;       // split-K: reduce with red.global.add
  int a_mn, b_mn;   // operand majors (0 = K-major, 1 = MN-major)
  uint32_t idesc;
  double* stats_parts;  // persistent kernel: per (row tile, lane quadrant) column sums / sums of squares of C (or NULL)
  int tma_store;    // 128 x 128 kernel: 1 = write C with TMA bulk stores from a swizzled staging box (map_c valid),
                    // 2 = C += tile with TMA bulk reductions (accumulate without split-K)
  // pooled epilogue (inference: batch norm is affine, so bias + BN + ReLU + mean / max over a cloud's points end in the
  // GEMM and the [M, N] activation is never stored): pool_mode 1 = sum, 2 = max of relu(v * pool_scale[c] + pool_shift[c])
  // over each (256-row tile, lane quadrant) -> pool_parts f32[tiles_m * 4][N]; 0 = off
  int pool_mode;
  const float* pool_scale;
  const float* pool_shift;
  float* pool_parts;
  int x3;           // split-precision ("3xTF32") product: C = A*B + A_lo*B + A*B_lo with A_lo = A - tf32(A), B_lo likewise
                    // (map_a_lo / map_b_lo valid) — fp32-grade accuracy at three tensor-core passes
};

__global__ void __launch_bounds__(TTHREADS)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_lo, const GemmParams p) {
  pdl_wait();
  __shared__ __align__(16) float s_bias[TBN];
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-byte alignment
  const uint32_t smem_a = base, smem_b = base + TSTAGES * TSTAGE_A;
  const uint32_t bars = smem_b + TSTAGES * TSTAGE_B;
  const uint32_t full0 = bars, empty0 = bars + 8 * TSTAGES, tmem_full = bars + 16 * TSTAGES;
  const uint32_t tmem_slot = bars + 16 * TSTAGES + 8;
  // generic pointer to the TMEM-address slot
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  const int num_kb_total = (p.K + TBK - 1) / TBK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
  const int num_kb_real = kb1 - kb0;
  // x3: the K loop runs three times over the same K range — (A, B), (A_lo, B), (A, B_lo) — into one accumulator
  const int num_kb = p.x3 ? 3 * num_kb_real : num_kb_real;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % TSTAGES;
        const uint32_t ph = (i / TSTAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, TSTAGE_A + TSTAGE_B);
        const int pass = p.x3 ? i / num_kb_real : 0;
        const int k0 = (kb0 + (i - pass * num_kb_real)) * TBK;
        const CUtensorMap* ma = (pass == 1) ? &map_a_lo : &map_a;
        const CUtensorMap* mb = (pass == 2) ? &map_b_lo : &map_b;
        const uint32_t da = smem_a + s * TSTAGE_A, db = smem_b + s * TSTAGE_B;
        if (p.a_mn) {
#pragma unroll
          for (int c = 0; c < TBM / 32; ++c) tma_load_2d(da + c * 4096, ma, full0 + 8 * s, m0 + 32 * c, k0);
        } else {
          tma_load_2d(da, ma, full0 + 8 * s, k0, m0);
        }
        if (p.b_mn) {
#pragma unroll
          for (int c = 0; c < TBN / 32; ++c) tma_load_2d(db + c * 4096, mb, full0 + 8 * s, n0 + 32 * c, k0);
        } else {
          tma_load_2d(db, mb, full0 + 8 * s, k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t a_lbo = p.a_mn ? 4096u : 16u, b_lbo = p.b_mn ? 4096u : 16u;
      const uint32_t a_kstep = p.a_mn ? 1024u : 32u, b_kstep = p.b_mn ? 1024u : 32u;
      const uint32_t a_sbo = p.a_mn ? 512u : 1024u, b_sbo = p.b_mn ? 512u : 1024u;
      const uint32_t a_lt = p.a_mn ? 1u : 2u, b_lt = p.b_mn ? 1u : 2u;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % TSTAGES;
        const uint32_t ph = (i / TSTAGES) & 1;
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t da = smem_a + s * TSTAGE_A, db = smem_b + s * TSTAGE_B;
#pragma unroll
        for (int k = 0; k < TBK / 8; ++k) {
          const uint64_t adesc = make_smem_desc(da + k * a_kstep, a_lbo, a_sbo, a_lt);
          const uint64_t bdesc = make_smem_desc(db + k * b_kstep, b_lbo, b_sbo, b_lt);
          umma_tf32(tmem_base, adesc, bdesc, p.idesc, (uint32_t)((i | k) != 0));
        }
        umma_commit(empty0 + 8 * s);  // frees the stage once these MMAs have read it
      }
      umma_commit(tmem_full);
    }
  } else {
    // ===== epilogue: warps 2..5 -> TMEM lane quadrants 2,3,0,1 =====
    const int quad = warp & 3;
    const int row = m0 + quad * 32 + lane;
    const bool add_bias = (p.bias != nullptr) && (blockIdx.z == 0);
    // bias row of the tile -> shared memory during the main loop (keeps L2 latency out of the serial epilogue)
    for (int t = (int)threadIdx.x - 64; t < TBN; t += 128) s_bias[t] = (add_bias && n0 + t < p.N) ? p.bias[n0 + t] : 0.f;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int sbuf = 0;
#pragma unroll 1
    for (int c0 = 0; c0 < TBN; c0 += 32) {
      uint32_t r[32];
      if (num_kb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, r);
        tmem_wait_ld(r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (p.tma_store) {
        // every MMA has completed (tmem_full), so the operand stages are free: the A stages become the staging
        // boxes (two 32-row x 128-byte SWIZZLE_128B boxes per epilogue warp) of one TMA store per chunk —
        // full 128-byte lines instead of 16 bytes to 32 different lines per STG.128; rows past M are clipped
        if (n0 + c0 < p.N && m0 + quad * 32 < p.M) {
          const uint32_t box = smem_a + (uint32_t)(warp - 2) * 8192u + (uint32_t)sbuf * 4096u;
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          const uint32_t rowaddr = box + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
            const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            st_shared_v4(rowaddr + (uint32_t)(((j >> 2) ^ (lane & 7)) << 4), v);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (p.tma_store == 2) tma_reduce_add_2d(&map_c, box, n0 + c0, m0 + quad * 32);   // C += tile
            else tma_store_2d(&map_c, box, n0 + c0, m0 + quad * 32);
            tma_store_commit();
          }
          sbuf ^= 1;
        }
        continue;
      }
      if (row < p.M) {
        float* crow = p.C + (size_t)row * p.ldc + n0 + c0;
        const int ncols = min(32, p.N - (n0 + c0));
        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
            {
              const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
              v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            }
            if (p.atomic) {   // split-K partial tile: one red.global.add.v4.f32 per four outputs
              atomicAdd(reinterpret_cast<float4*>(crow + j), v);
              continue;
            }
            if (p.accumulate) {
              const float4 o = *reinterpret_cast<const float4*>(crow + j);
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *reinterpret_cast<float4*>(crow + j) = v;
          }
        } else {
          // (compile-time register indices: a run-time index would put r[] in local memory)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j < ncols) {
              const float v = __uint_as_float(r[j]) + s_bias[c0 + j];
              if (p.atomic) atomicAdd(crow + j, v);
              else crow[j] = p.accumulate ? crow[j] + v : v;
            }
          }
        }
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();   // the staging boxes must outlive the stores reading them
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
}

// ---- large-tile variant ------------------------------------------------------------------------------
// CTA tile (MH x 128) x BN with MH = 2 row halves (two M = 128 accumulators) and BN up to 256 columns.
// The 128 x 128 kernel above moves 32 KB of operands per 2 x 128 x 128 x 32 FLOP; for the dgcnn_agg
// contractions (K = 320 forward, K = 1024 data gradient) that makes it L2 -> SM bandwidth bound (measured
// 6.4 TB/s of operand reads at 209 TFLOP/s).  Larger tiles cut the operand bytes per FLOP: 256 x 160 for the
// N = 320 data gradient (dY is read twice instead of three times).  K-major or MN-major A; B either major.
// One CTA per SM (3-4 stages of 52-64 KB), 512 TMEM columns: accumulator of row block h at column ACC_COLS h.
// Third shape: the weight gradient [320,1024] = X^T dY (both operands MN-major, K = B*N rows, split-K): MH = 3
// row blocks cover all 320 rows (the last block is half padding, zero-filled by TMA), so dY — the large
// operand — is read once instead of three times; partial tiles are combined with 16-byte vector reductions.
template <int BN, int MH_>
struct BigCfg {
  static constexpr int MH = MH_;
  static constexpr int ACC_COLS = (BN <= 128) ? 128 : 256;         // TMEM column pitch of the row blocks
  static constexpr uint32_t STAGE_A = MH * TBM * TBK * 4;          // 16 KB per row block
  static constexpr uint32_t STAGE_B = BN * TBK * 4;                // 32 KB (BN = 256) / 20 KB (160) / 16 KB (128)
  static constexpr int STAGES = (STAGE_A + STAGE_B > 56 * 1024) ? 3 : 4;
  static constexpr uint32_t SMEM = STAGES * (STAGE_A + STAGE_B) + 1024 + 256;
  static_assert(MH * ACC_COLS <= 512, "TMEM has 512 columns");
};

constexpr int BIG_THREADS = 320;   // TMA warp, MMA warp, eight epilogue warps

template <int BN, int MH_, bool A_MN>
__global__ void __launch_bounds__(BIG_THREADS)
gemm_tf32_big_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_c, const GemmParams p) {
  pdl_wait();
  using Cfg = BigCfg<BN, MH_>;
  constexpr int ST = Cfg::STAGES, MH = Cfg::MH;
  __shared__ __align__(16) float s_bias[BN];
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = base, smem_b = base + ST * Cfg::STAGE_A;
  const uint32_t bars = smem_b + ST * Cfg::STAGE_B;
  const uint32_t full0 = bars, empty0 = bars + 8 * ST, tmem_full = bars + 16 * ST, tmem_slot = tmem_full + 8;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * (MH * TBM), n0 = blockIdx.x * BN;
  const int num_kb_total = (p.K + TBK - 1) / TBK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int num_kb = min(num_kb_total, kb0 + p.kb_per_split) - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % ST;
        const uint32_t ph = (i / ST) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, Cfg::STAGE_A + Cfg::STAGE_B);
        const int k0 = (kb0 + i) * TBK;
        const uint32_t da = smem_a + s * Cfg::STAGE_A, db = smem_b + s * Cfg::STAGE_B;
        if (A_MN) {
#pragma unroll
          for (int c = 0; c < MH * TBM / 32; ++c) tma_load_2d(da + c * 4096, &map_a, full0 + 8 * s, m0 + 32 * c, k0);
        } else {
#pragma unroll
          for (int h = 0; h < MH; ++h) tma_load_2d(da + h * TSTAGE_A, &map_a, full0 + 8 * s, k0, m0 + h * TBM);
        }
        if (p.b_mn) {
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) tma_load_2d(db + c * 4096, &map_b, full0 + 8 * s, n0 + 32 * c, k0);
        } else {
          tma_load_2d(db, &map_b, full0 + 8 * s, k0, n0);   // box {32, BN}
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t b_lbo = p.b_mn ? 4096u : 16u, b_kstep = p.b_mn ? 1024u : 32u;
      const uint32_t b_sbo = p.b_mn ? 512u : 1024u, b_lt = p.b_mn ? 1u : 2u;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % ST;
        const uint32_t ph = (i / ST) & 1;
        mbar_wait(full0 + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t da = smem_a + s * Cfg::STAGE_A, db = smem_b + s * Cfg::STAGE_B;
#pragma unroll
        for (int k = 0; k < TBK / 8; ++k) {
          const uint64_t bdesc = make_smem_desc(db + k * b_kstep, b_lbo, b_sbo, b_lt);
#pragma unroll
          for (int h = 0; h < MH; ++h) {
            const uint64_t adesc = A_MN ? make_smem_desc(da + h * TSTAGE_A + k * 1024u, 4096u, 512u, 1u)
                                        : make_smem_desc(da + h * TSTAGE_A + k * 32u, 16u, 1024u, 2u);
            umma_tf32(tmem_base + (uint32_t)(h * Cfg::ACC_COLS), adesc, bdesc, p.idesc, (uint32_t)((i | k) != 0));
          }
        }
        umma_commit(empty0 + 8 * s);
      }
      umma_commit(tmem_full);
    }
  } else {
    // ===== epilogue: EIGHT warps (2..9).  A warp may only read the TMEM lane quadrant warp % 4, so two warps
    // share each quadrant and split the 32-column chunks between them (even / odd).  With one CTA per SM the
    // epilogue is serial after the main loop: twice the warps and a TMEM load issued one chunk ahead of the
    // stores roughly halve it.
    const int quad = warp & 3, part = (warp - 2) >> 2;
    // the tile's bias row goes to shared memory while the main loop runs: a global load per output chunk
    // here would expose the (loaded) L2 latency once per chunk — measured +55 us on the dgcnn_agg forward GEMM
    const bool add_bias = (p.bias != nullptr) && (blockIdx.z == 0);
    for (int t = (int)threadIdx.x - 64; t < BN; t += 256) s_bias[t] = (add_bias && n0 + t < p.N) ? p.bias[n0 + t] : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr int NC = BN / 32;                       // chunks per row block
    constexpr int NWORK = MH * NC;                    // (row block, chunk) items; this warp takes item % 2 == part
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    auto taddr = [&](int w) { return tlane + (uint32_t)((w / NC) * Cfg::ACC_COLS + (w % NC) * 32); };
    int sbuf = 0;
    auto store_item = [&](uint32_t (&r)[32], int w) {
      if (num_kb <= 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      const int h = w / NC, c0 = (w % NC) * 32;
      if (p.tma_store) {
        // the operand stages are free once tmem_full has fired: their first 64 KB become the staging boxes (two
        // 32-row x 128-byte SWIZZLE_128B boxes per epilogue warp) of one TMA store / reduction per chunk
        if (n0 + c0 >= p.N || m0 + h * TBM + quad * 32 >= p.M) return;   // nothing of this box lies inside C
        const uint32_t box = smem_a + (uint32_t)(warp - 2) * 8192u + (uint32_t)sbuf * 4096u;
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const uint32_t rowaddr = box + (uint32_t)lane * 128u;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                 __uint_as_float(r[j + 3]));
          const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          st_shared_v4(rowaddr + (uint32_t)(((j >> 2) ^ (lane & 7)) << 4), v);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (p.tma_store == 2) tma_reduce_add_2d(&map_c, box, n0 + c0, m0 + h * TBM + quad * 32);
          else tma_store_2d(&map_c, box, n0 + c0, m0 + h * TBM + quad * 32);
          tma_store_commit();
        }
        sbuf ^= 1;
        return;
      }
      const int row = m0 + h * TBM + quad * 32 + lane;
      if (row >= p.M || n0 + c0 >= p.N) return;
      float* crow = p.C + (size_t)row * p.ldc + n0 + c0;
      const int ncols = min(32, p.N - (n0 + c0));
      if (ncols == 32 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                 __uint_as_float(r[j + 3]));
          const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + j);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          if (p.atomic) {   // split-K partial tile: red.global.add.v4.f32
            atomicAdd(reinterpret_cast<float4*>(crow + j), v);
            continue;
          }
          if (p.accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(crow + j);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *reinterpret_cast<float4*>(crow + j) = v;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {   // compile-time register indices (a run-time index would put r[] in local memory)
          if (j < ncols) {
            const float v = __uint_as_float(r[j]) + s_bias[c0 + j];
            if (p.atomic) atomicAdd(crow + j, v);
            else crow[j] = p.accumulate ? crow[j] + v : v;
          }
        }
      }
    };
    // two register buffers: the TMEM load of the next item is in flight while the current one is stored
    uint32_t r0[32], r1[32];
    const bool have = num_kb > 0;
    if (part < NWORK && have) tmem_ld32(taddr(part), r0);
#pragma unroll 1
    for (int w = part; w < NWORK; w += 4) {
      tmem_wait_ld(r0);
      if (w + 2 < NWORK && have) tmem_ld32(taddr(w + 2), r1);
      store_item(r0, w);
      if (w + 2 < NWORK) {
        tmem_wait_ld(r1);
        if (w + 4 < NWORK && have) tmem_ld32(taddr(w + 4), r0);
        store_item(r1, w + 2);
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (p.tma_store && lane == 0) tma_store_wait_all();   // the staging boxes must outlive the bulk operations reading them
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---- persistent variant for tall, short-K, wide-N problems (dgcnn_agg / pn_conv5 forward) -------------------
// 256 x 128 tiles (two row blocks share one B stage: 48 KB of operands per 2 x 128 x 128 x 32 MACs, 1.5x
// fewer operand bytes per FLOP than the 128 x 128 kernel), one CTA per SM walking over tiles, and TWO
// accumulator sets in TMEM (2 x 256 columns): the eight epilogue warps drain tile i (128 KB of stores)
// while the TMA and MMA warps run the main loop of tile i+1, so the output stream — the real bound of a
// K = 320 GEMM with a 134 MB result — overlaps the tensor work instead of following it.  Barrier set-up,
// TMEM allocation and descriptor fetch happen once per CTA.  K-major A; B either major; no split-K.
// X3 = split-precision product (GemmParams::x3): a stage holds FOUR operand tiles (A, B, A_lo, B_lo) and every K step
// issues three MMAs per row block — A*B, A_lo*B, A*B_lo — so the operand bytes per useful FLOP only double while the
// result is fp32-grade (the TMA rounds fp32 -> tf32 to nearest-even on the way in: measured, tools/probe_tf32_rounding.py;
// the low parts A - tf32(A) are prepared by the producers of A / by caae_split_tf32).  Two 96 KB stages, one staging box
// per epilogue warp, bias read through the read-only path instead of shared memory (227 KB per CTA is the limit).
constexpr uint32_t PS_STAGE_A = 2 * TSTAGE_A, PS_STAGE_B = TSTAGE_B;
// epilogue staging: per epilogue warp 32-row x 128-byte boxes (SWIZZLE_128B, the layout the C tensor map expects)
constexpr uint32_t PS_STG_BOX = 32 * 128;
constexpr int PS_MAXN = 2048;
template <bool X3>
struct PsCfg {
  static constexpr int STAGES = X3 ? 2 : 3;
  static constexpr uint32_t STAGE = (PS_STAGE_A + PS_STAGE_B) * (X3 ? 2u : 1u);
  static constexpr int BOXES = X3 ? 1 : 2;                       // staging boxes per epilogue warp
  static constexpr uint32_t STG_WARP = BOXES * PS_STG_BOX, STG = 8 * STG_WARP;
  static constexpr uint32_t SMEM = STAGES * STAGE + STG + 1024 + 256;
  static constexpr int BIAS_SMEM = X3 ? 4 : PS_MAXN;
};

template <bool X3>
__global__ void __launch_bounds__(BIG_THREADS)
gemm_tf32_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_a_lo,
                         const __grid_constant__ CUtensorMap map_b_lo, const GemmParams p, const int tiles_m,
                         const int tiles_n) {
  pdl_wait();
  using Cfg = PsCfg<X3>;
  constexpr int ST = Cfg::STAGES;
  __shared__ __align__(16) float s_bias[Cfg::BIAS_SMEM];
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // stage s: [A (2 row blocks, 32 KB) | B (16 KB) | A_lo | B_lo]
  const uint32_t stg = base + ST * Cfg::STAGE;   // 1024-byte aligned: every stage is a multiple of 1 KB
  const uint32_t bars = stg + Cfg::STG;
  const uint32_t full0 = bars, empty0 = bars + 8 * ST;
  const uint32_t tfull0 = empty0 + 8 * ST, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + TBK - 1) / TBK;
  const int tiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (!X3)
    for (int t = threadIdx.x; t < p.N; t += BIG_THREADS) s_bias[t] = (p.bias != nullptr) ? p.bias[t] : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===== TMA producer: tiles id = blockIdx.x, + gridDim.x, ...; id -> (row tile id / tiles_n, column tile id % tiles_n),
    // so the CTAs of one wave share A row tiles in L2 =====
    if (lane == 0) {
      int it = 0;
      for (int id = blockIdx.x; id < tiles; id += gridDim.x) {
        const int m0 = (id / tiles_n) * (2 * TBM), n0 = (id % tiles_n) * TBN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          mbar_expect_tx(full0 + 8 * s, Cfg::STAGE);
          const int k0 = kb * TBK;
          const uint32_t da = base + s * Cfg::STAGE, db = da + PS_STAGE_A;
#pragma unroll
          for (int v = 0; v < (X3 ? 2 : 1); ++v) {
            const CUtensorMap* ma = v ? &map_a_lo : &map_a;
            const CUtensorMap* mb = v ? &map_b_lo : &map_b;
            const uint32_t dav = da + v * (PS_STAGE_A + PS_STAGE_B), dbv = db + v * (PS_STAGE_A + PS_STAGE_B);
            tma_load_2d(dav, ma, full0 + 8 * s, k0, m0);
            tma_load_2d(dav + TSTAGE_A, ma, full0 + 8 * s, k0, m0 + TBM);
            if (p.b_mn) {
#pragma unroll
              for (int c = 0; c < TBN / 32; ++c) tma_load_2d(dbv + c * 4096, mb, full0 + 8 * s, n0 + 32 * c, k0);
            } else {
              tma_load_2d(dbv, mb, full0 + 8 * s, k0, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t b_lbo = p.b_mn ? 4096u : 16u, b_kstep = p.b_mn ? 1024u : 32u;
      const uint32_t b_sbo = p.b_mn ? 512u : 1024u, b_lt = p.b_mn ? 1u : 2u;
      int it = 0, lt = 0;
      for (int id = blockIdx.x; id < tiles; id += gridDim.x, ++lt) {
        const int acc = lt & 1;
        mbar_wait(tempty0 + 8 * acc, (uint32_t)(((lt >> 1) & 1) ^ 1));   // the epilogue has drained this accumulator set
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tacc = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % ST;
          const uint32_t ph = (it / ST) & 1;
          mbar_wait(full0 + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t da = base + s * Cfg::STAGE, db = da + PS_STAGE_A;
          constexpr uint32_t LO = PS_STAGE_A + PS_STAGE_B;
#pragma unroll
          for (int k = 0; k < TBK / 8; ++k) {
            const uint64_t bdesc = make_smem_desc(db + k * b_kstep, b_lbo, b_sbo, b_lt);
            const uint64_t bdesc_lo = make_smem_desc(db + LO + k * b_kstep, b_lbo, b_sbo, b_lt);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint64_t adesc = make_smem_desc(da + h * TSTAGE_A + k * 32u, 16u, 1024u, 2u);
              umma_tf32(tacc + (uint32_t)(h * TBN), adesc, bdesc, p.idesc, (uint32_t)((kb | k) != 0));
              if (X3) {
                const uint64_t adesc_lo = make_smem_desc(da + LO + h * TSTAGE_A + k * 32u, 16u, 1024u, 2u);
                umma_tf32(tacc + (uint32_t)(h * TBN), adesc_lo, bdesc, p.idesc, 1u);
                umma_tf32(tacc + (uint32_t)(h * TBN), adesc, bdesc_lo, p.idesc, 1u);
              }
            }
          }
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull0 + 8 * acc);
      }
    }
  } else {
    // ===== epilogue: eight warps; two per TMEM lane quadrant, splitting the (row block, 32-column chunk) items =====
    const int quad = warp & 3, part = (warp - 2) >> 2;
    constexpr int NC = TBN / 32, NWORK = 2 * NC;
    int lt = 0, sbuf = 0;
    for (int id = blockIdx.x; id < tiles; id += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const int m0 = (id / tiles_n) * (2 * TBM), n0 = (id % tiles_n) * TBN;
      mbar_wait(tfull0 + 8 * acc, (uint32_t)((lt >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 256);
      auto taddr = [&](int w) { return tlane + (uint32_t)((w / NC) * TBN + (w % NC) * 32); };
      double st_sum[2] = {0.0, 0.0}, st_sq[2] = {0.0, 0.0};   // this warp's two column chunks (part, part + 2), both row blocks
      auto bias4 = [&](int col) -> float4 {
        if (X3) return p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        return *reinterpret_cast<const float4*>(s_bias + col);
      };
      float pool_acc[2] = {p.pool_mode == 2 ? -INFINITY : 0.f, p.pool_mode == 2 ? -INFINITY : 0.f};
      auto store_item = [&](uint32_t (&r)[32], int w) {
        const int h = w / NC, c0 = (w % NC) * 32;
        if (p.pool_mode != 0) {
          // inference epilogue: nothing is stored.  relu(bn(v)) goes through the warp's staging box so that lane l can
          // reduce column l over the 32 rows (the statistics path's access pattern: 32 distinct banks per row)
          const uint32_t box = stg + (uint32_t)(warp - 2) * Cfg::STG_WARP;
          const uint32_t rowaddr = box + (uint32_t)lane * 128u;
          __syncwarp();                                     // the previous item's column reads are done
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = bias4(n0 + c0 + j);
            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.pool_scale + n0 + c0 + j));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.pool_shift + n0 + c0 + j));
            float4 v;
            v.x = fmaxf(fmaf(__uint_as_float(r[j]) + bv.x, sc.x, sh.x), 0.f);
            v.y = fmaxf(fmaf(__uint_as_float(r[j + 1]) + bv.y, sc.y, sh.y), 0.f);
            v.z = fmaxf(fmaf(__uint_as_float(r[j + 2]) + bv.z, sc.z, sh.z), 0.f);
            v.w = fmaxf(fmaf(__uint_as_float(r[j + 3]) + bv.w, sc.w, sh.w), 0.f);
            st_shared_v4(rowaddr + (uint32_t)((((j >> 2) ^ (lane & 7))) << 4), v);
          }
          __syncwarp();
          const int rmax = min(32, p.M - (m0 + h * TBM + quad * 32));
          const uint32_t col = box + (uint32_t)((lane & 3) << 2);
          float red = p.pool_mode == 2 ? -INFINITY : 0.f;
          for (int rr = 0; rr < rmax; ++rr) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(col + (uint32_t)rr * 128u + (uint32_t)((((lane >> 2) ^ (rr & 7))) << 4)));
            red = p.pool_mode == 2 ? fmaxf(red, v) : red + v;
          }
          const int slot = ((w - part) >> 1) & 1;
          pool_acc[slot] = p.pool_mode == 2 ? fmaxf(pool_acc[slot], red) : pool_acc[slot] + red;
          return;
        }
        if (!p.accumulate) {
          // registers -> swizzled shared-memory box -> one TMA store of 32 rows x 128 bytes (full lines; rows past M
          // are clipped by the tensor map).  A thread-per-row STG.128 writes 16 bytes to 32 different lines per
          // instruction: that, not the tensor pipe, bounded this kernel (134 MB at 1.9 TB/s).
          const uint32_t box = stg + (uint32_t)(warp - 2) * Cfg::STG_WARP + (uint32_t)sbuf * PS_STG_BOX;
          if (lane == 0) tma_store_wait_read<Cfg::BOXES - 1>();   // the store that last read this box is done with it
          __syncwarp();
          const uint32_t rowaddr = box + (uint32_t)lane * 128u;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                   __uint_as_float(r[j + 3]));
            const float4 bv = bias4(n0 + c0 + j);
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            st_shared_v4(rowaddr + (uint32_t)((((j >> 2) ^ (lane & 7))) << 4), v);   // SWIZZLE_128B: chunk ^= row % 8
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_c, box, n0 + c0, m0 + h * TBM + quad * 32);
            tma_store_commit();
          }
          if (p.stats_parts != nullptr) {
            // batch-norm statistics of the tile while it sits in shared memory: lane l sums column l of the box
            // (swizzled address: 32 distinct banks per row), fp32 over the 32 rows, fp64 across boxes
            const int rmax = min(32, p.M - (m0 + h * TBM + quad * 32));
            float fs = 0.f, fq = 0.f;
            const uint32_t col = box + (uint32_t)((lane & 3) << 2);
            for (int rr = 0; rr < rmax; ++rr) {
              float v;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(col + (uint32_t)rr * 128u + (uint32_t)((((lane >> 2) ^ (rr & 7))) << 4)));
              fs += v;
              fq = fmaf(v, v, fq);
            }
            const int slot = ((w - part) >> 1) & 1;
            st_sum[slot] += (double)fs;
            st_sq[slot] += (double)fq;
          }
          if (Cfg::BOXES == 2) sbuf ^= 1;
          return;
        }
        const int row = m0 + h * TBM + quad * 32 + lane;
        if (row >= p.M) return;
        float* crow = p.C + (size_t)row * p.ldc + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                 __uint_as_float(r[j + 3]));
          const float4 bv = bias4(n0 + c0 + j);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          const float4 o = *reinterpret_cast<const float4*>(crow + j);
          v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          *reinterpret_cast<float4*>(crow + j) = v;
        }
      };
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr(part), r0);
#pragma unroll 1
      for (int w = part; w < NWORK; w += 4) {
        tmem_wait_ld(r0);
        if (w + 2 < NWORK) tmem_ld32(taddr(w + 2), r1);
        store_item(r0, w);
        if (w + 2 < NWORK) {
          tmem_wait_ld(r1);
          if (w + 4 < NWORK) tmem_ld32(taddr(w + 4), r0);
          store_item(r1, w + 2);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (p.pool_mode != 0) {
        float* prow = p.pool_parts + (size_t)((id / tiles_n) * 4 + quad) * p.N;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) prow[n0 + (part + 2 * sl) * 32 + lane] = pool_acc[sl];
      }
      if (p.stats_parts != nullptr && !p.accumulate && p.pool_mode == 0) {
        // one partial row per (row tile, lane quadrant): the layout caae_bn_finalize reduces ([row][2][N])
        double* prow = p.stats_parts + (size_t)((id / tiles_n) * 4 + quad) * 2 * p.N;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
          const int colg = n0 + (part + 2 * sl) * 32 + lane;
          prow[colg] = st_sum[sl];
          prow[p.N + colg] = st_sq[sl];
        }
      }
    }
    if (lane == 0) tma_store_wait_all();   // the staging boxes must outlive the bulk stores that read them
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// lo = x - tf32_rne(x): the low part of a split-precision operand (exact in fp32).  [rows, cols] with row pitches.
__global__ void split_tf32_kernel(long rows, int cols, const float* __restrict__ x, int ldx, float* __restrict__ lo, int ldlo) {
  pdl_wait();
  const long total = rows * cols;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long r = e / cols;
    const int c = (int)(e - r * cols);
    const float v = x[r * ldx + c];
    lo[r * ldlo + c] = v - tf32_rne(v);
  }
}

__global__ void zero_matrix2_kernel(int M, int N, float* __restrict__ C, int ldc) {
  pdl_wait();
  const long total = (long)M * N;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    C[(e / N) * ldc + (e % N)] = 0.f;
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 2-D fp32 tensor with `inner` contiguous elements per row, `outer` rows, row pitch ld elements
static int make_map(CUtensorMap* map, const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                    uint32_t box_outer, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return CAAE_E_UNSUPPORTED;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : CAAE_E_UNSUPPORTED;
}

// Output map for TMA-store epilogues: C[outer = rows, inner = columns] fp32, boxes of 32 columns (128 bytes) x 32 rows
static int make_map_c(CUtensorMap* map, float* ptr, uint64_t cols, uint64_t rows, uint64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return CAAE_E_UNSUPPORTED;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : CAAE_E_UNSUPPORTED;
}

}  // namespace caae

using namespace caae;

// 1 when the TF32 tensor-core kernel accepts this problem (TMA alignment rules), else 0.
extern "C" int caae_gemm_tf32_supported(int transa, int transb, int M, int N, int K, const float* A, int lda,
                                        const float* B, int ldb) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return 0;
  if ((lda & 3) || (ldb & 3)) return 0;
  (void)transa; (void)transb;
  return 1;
}

struct PoolEpilogue { int mode; const float* scale; const float* shift; float* parts; };

static int gemm_tf32_impl(int transa, int transb, int M, int N, int K, const float* A, int lda, const float* B,
                          int ldb, float* C, int ldc, const float* bias, int accumulate, double* stats_parts,
                          caae_stream_t stream, const float* A_lo = nullptr, const float* B_lo = nullptr,
                          const PoolEpilogue* pool = nullptr) {
  CAAE_RETURN_IF(M < 0 || N < 0 || K < 0, CAAE_E_BADSHAPE);
  if (M == 0 || N == 0) return CAAE_OK;
  CAAE_RETURN_IF((!C && !pool) || !A || !B, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(ldc < N || lda < (transa ? M : K) || ldb < (transb ? K : N), CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(!caae_gemm_tf32_supported(transa, transb, M, N, K, A, lda, B, ldb), CAAE_E_UNSUPPORTED);
  const bool x3 = A_lo != nullptr || B_lo != nullptr;
  CAAE_RETURN_IF(x3 && (!A_lo || !B_lo), CAAE_E_NULLPTR);
  CAAE_RETURN_IF(x3 && ((reinterpret_cast<uintptr_t>(A_lo) & 15) || (reinterpret_cast<uintptr_t>(B_lo) & 15)), CAAE_E_UNSUPPORTED);
  cudaStream_t s = as_stream(stream);

  CUtensorMap map_a, map_b, map_a_lo, map_b_lo;
  int rc;
  // A: transa = 0 -> stored [M,K] (K contiguous, K-major); transa = 1 -> stored [K,M] (M contiguous, MN-major)
  rc = transa ? make_map(&map_a, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 32, TBK, true)
              : make_map(&map_a, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TBK, TBM, false);
  if (rc) return rc;
  // B: transb = 1 -> stored [N,K] (K-major); transb = 0 -> stored [K,N] (N contiguous, MN-major)
  rc = transb ? make_map(&map_b, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, TBK, TBN, false)
              : make_map(&map_b, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, TBK, true);
  if (rc) return rc;
  map_a_lo = map_a; map_b_lo = map_b;   // placeholders unless the split-precision product is requested
  if (x3) {   // the low parts share the layout (and leading dimension) of their operand
    rc = transa ? make_map(&map_a_lo, A_lo, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 32, TBK, true)
                : make_map(&map_a_lo, A_lo, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TBK, TBM, false);
    if (rc) return rc;
    rc = transb ? make_map(&map_b_lo, B_lo, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, TBK, TBN, false)
                : make_map(&map_b_lo, B_lo, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, TBK, true);
    if (rc) return rc;
  }

  GemmParams p;
  p.pool_mode = pool ? pool->mode : 0;
  p.pool_scale = pool ? pool->scale : nullptr; p.pool_shift = pool ? pool->shift : nullptr; p.pool_parts = pool ? pool->parts : nullptr;
  p.x3 = x3 ? 1 : 0;
  p.tma_store = 0;
  p.stats_parts = nullptr;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias;
  p.a_mn = transa ? 1 : 0;
  p.b_mn = transb ? 0 : 1;
  // instruction descriptor: c_format F32 (bit 4), a/b format TF32 (=2) at bits 7 / 10, majors at 15 / 16,
  // N>>3 at bit 17, M>>4 at bit 24
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
            ((uint32_t)(TBN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
  const int tiles_m = (M + TBM - 1) / TBM, tiles_n = (N + TBN - 1) / TBN;
  const int num_kb = (K + TBK - 1) / TBK;
  // tall / wide / short-K forward contractions: persistent 256 x 128 tiles with the epilogue overlapped
  static const bool persist_enabled = [] { const char* e = getenv("CAAE_GEMM_PERSIST"); return !(e && e[0] == '0'); }();
  const bool persist_ok = persist_enabled && !transa && N % TBN == 0 && N <= PS_MAXN && N >= 512 &&
                          (M >= 256 * kNumSMs / 2 || pool != nullptr) && num_kb >= 2 && num_kb <= 16 && ldc % 4 == 0 &&
                          (reinterpret_cast<uintptr_t>(C) & 15) == 0 &&
                          (bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 3) == 0);
  if (stats_parts != nullptr && (!persist_ok || accumulate)) return CAAE_E_UNSUPPORTED;
  if (pool != nullptr && (!persist_ok || accumulate || N % 4 != 0)) return CAAE_E_UNSUPPORTED;
  p.stats_parts = stats_parts;
  if (persist_ok) {
    CUtensorMap map_a2, map_b2;
    rc = make_map(&map_a2, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TBK, TBM, false);
    if (rc) return rc;
    rc = transb ? make_map(&map_b2, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, TBK, TBN, false)
                : make_map(&map_b2, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, TBK, true);
    if (rc) return rc;
    static bool psattr = false;
    if (!psattr) {
      cudaError_t e = cudaFuncSetAttribute(gemm_tf32_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)PsCfg<false>::SMEM);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(gemm_tf32_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)PsCfg<true>::SMEM);
      if (e != cudaSuccess) return (int)e;
      psattr = true;
    }
    CUtensorMap map_c2 = map_a2;   // (unused by the pooled epilogue)
    if (pool == nullptr) {
      rc = make_map_c(&map_c2, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc);
      if (rc) return rc;
    }
    const int tm = (M + 2 * TBM - 1) / (2 * TBM), tn = N / TBN;
    p.kb_per_split = num_kb; p.atomic = 0; p.accumulate = accumulate;
    const int ctas = tm * tn < kNumSMs ? tm * tn : kNumSMs;
    if (x3)   // (the *_lo maps built above have the same boxes as map_a2 / map_b2: K-major A, either-major B)
      caae::launch(gemm_tf32_persist_kernel<true>, ctas, BIG_THREADS, PsCfg<true>::SMEM, s, map_a2, map_b2, map_c2, map_a_lo, map_b_lo, p, tm, tn);
    else
      caae::launch(gemm_tf32_persist_kernel<false>, ctas, BIG_THREADS, PsCfg<false>::SMEM, s, map_a2, map_b2, map_c2, map_a2, map_b2, p, tm, tn);
    return CAAE_LAUNCH_STATUS();
  }
  // the three big dgcnn_agg-shaped contractions: large tiles, one CTA per SM (see gemm_tf32_big_kernel)
  static const bool big_enabled = [] { const char* e = getenv("CAAE_GEMM_BIG"); return !(e && e[0] == '0'); }();
  // (N % 256 == 0 shapes such as the dgcnn_agg forward GEMM stay on the 128 x 128 kernel: measured 85 us vs
  // 94 us for a 256 x 256 tile, whose serial epilogue of 256 KB per tile outweighs the saved operand traffic)
  const bool big_fwd = !x3 && !transa && M >= 256 * kNumSMs / 2 && K >= 256 && N >= 160 && N % 160 == 0 && N % 256 != 0;
  const bool big_wgrad = !x3 && transa && !transb && M > 256 && M <= 384 && N % 128 == 0 && N >= 512 && K >= 64 * TBK * 8;
  if (big_enabled && (big_fwd || big_wgrad)) {
    const int bn = big_wgrad ? 128 : 160;
    const int mh = big_wgrad ? 3 : 2;
    CUtensorMap map_a2, map_b2;
    rc = transa ? make_map(&map_a2, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 32, TBK, true)
                : make_map(&map_a2, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, TBK, TBM, false);
    if (rc) return rc;
    rc = transb ? make_map(&map_b2, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, TBK, (uint32_t)bn, false)
                : make_map(&map_b2, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, TBK, true);
    if (rc) return rc;
    p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
              ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
    const int tiles_big = (N / bn) * ((M + mh * TBM - 1) / (mh * TBM));
    int sp = 1;
    if (big_wgrad) {   // split K so that the grid fills the SMs once
      sp = kNumSMs / tiles_big;
      if (sp < 1) sp = 1;
      if (sp > num_kb / 8) sp = num_kb / 8;
    }
    p.kb_per_split = (num_kb + sp - 1) / sp;
    sp = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.atomic = sp > 1;
    p.accumulate = p.atomic ? 0 : accumulate;
    if (p.atomic && !accumulate) {
      const long total = (long)M * N;
      const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
      caae::launch(zero_matrix2_kernel, blocks, 256, 0, s, M, N, C, ldc);
    }
    dim3 grid(N / bn, (M + mh * TBM - 1) / (mh * TBM), sp);
    CAAE_RETURN_IF(grid.y > 65535, CAAE_E_BADSHAPE);
    CUtensorMap map_c2 = map_a2;   // placeholder unless the TMA epilogue applies
    static const bool big_tma = [] { const char* e = getenv("CAAE_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
    if (big_tma && N % 32 == 0 && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
      rc = make_map_c(&map_c2, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc);
      if (rc) return rc;
      p.tma_store = (p.atomic || p.accumulate) ? 2 : 1;   // split-K partial tiles and C += combine by bulk reduction
    }
#define CAAE_LAUNCH_BIG(BN_, MH_, AMN_)                                                                              \
    do {                                                                                                             \
      cudaError_t e_ = cudaFuncSetAttribute(gemm_tf32_big_kernel<BN_, MH_, AMN_>,                                    \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BigCfg<BN_, MH_>::SMEM); \
      if (e_ != cudaSuccess) return (int)e_;                                                                         \
      caae::launch(gemm_tf32_big_kernel<BN_, MH_, AMN_>, grid, BIG_THREADS, BigCfg<BN_, MH_>::SMEM, s, map_a2, map_b2, map_c2, p);        \
    } while (0)
    if (big_wgrad) CAAE_LAUNCH_BIG(128, 3, true);
    else CAAE_LAUNCH_BIG(160, 2, false);
#undef CAAE_LAUNCH_BIG
    return CAAE_LAUNCH_STATUS();
  }
  int splits = 1;
  const int tiles = tiles_m * tiles_n;
  if (tiles < kNumSMs && num_kb >= 8) {   // (an x3 product runs its three passes over each split's own K range)
    splits = (2 * kNumSMs + tiles - 1) / tiles;
    // >= 16 K blocks per split when K is long: with 147 splits of 7 blocks the [64,256] EdgeConv weight gradient
    // spent its time in prologues and in 147-way contended reductions (57 us for 1 GFLOP)
    const int min_kb = num_kb >= 64 ? 16 : 4;
    if (splits > num_kb / min_kb) splits = num_kb / min_kb;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.atomic = splits > 1;
  p.accumulate = p.atomic ? 0 : accumulate;
  if (p.atomic && !accumulate) {
    const long total = (long)M * N;
    const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    caae::launch(zero_matrix2_kernel, blocks, 256, 0, s, M, N, C, ldc);
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TSMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(tiles_n, tiles_m, splits);
  CAAE_RETURN_IF(grid.y > 65535 || grid.z > 65535, CAAE_E_BADSHAPE);
  CUtensorMap map_c = map_a;   // placeholder unless the TMA-store epilogue applies
  static const bool tma_store_enabled = [] { const char* e = getenv("CAAE_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  p.tma_store = 0;
  if (tma_store_enabled && !p.atomic && N % 32 == 0 && ldc % 4 == 0 && M >= 4096 &&
      (reinterpret_cast<uintptr_t>(C) & 15) == 0) {
    rc = make_map_c(&map_c, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc);
    if (rc) return rc;
    p.tma_store = p.accumulate ? 2 : 1;
  }
  caae::launch(gemm_tf32_kernel, grid, TTHREADS, TSMEM_BYTES, s, map_a, map_b, map_c, map_a_lo, map_b_lo, p);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_gemm_tf32(int transa, int transb, int M, int N, int K, const float* A, int lda, const float* B,
                              int ldb, float* C, int ldc, const float* bias, int accumulate, caae_stream_t stream) {
  return gemm_tf32_impl(transa, transb, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, nullptr, stream);
}

// Split-precision variant: C (+)= A*B + A_lo*B + A*B_lo (+ bias) with the low parts x - tf32(x) supplied by the caller
// (caae_split_tf32, or the producing kernel).  parts != NULL: also the batch-norm column statistics (persistent shapes).
extern "C" int caae_gemm_tf32x3(int transa, int transb, int M, int N, int K, const float* A, const float* A_lo, int lda,
                                const float* B, const float* B_lo, int ldb, float* C, int ldc, const float* bias,
                                int accumulate, double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(!A_lo || !B_lo, CAAE_E_NULLPTR);
  return gemm_tf32_impl(transa, transb, M, N, K, A, lda, B, ldb, C, ldc, bias, accumulate, parts, stream, A_lo, B_lo);
}

extern "C" int caae_split_tf32(long rows, int cols, const float* x, int ldx, float* lo, int ldlo, caae_stream_t stream) {
  CAAE_RETURN_IF(rows < 0 || cols <= 0 || ldx < cols || ldlo < cols, CAAE_E_BADSHAPE);
  if (rows == 0) return CAAE_OK;
  CAAE_RETURN_IF(!x || !lo, CAAE_E_NULLPTR);
  const long total = rows * cols;
  long blocks = (total + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  caae::launch(split_tf32_kernel, (int)blocks, 256, 0, as_stream(stream), rows, cols, x, ldx, lo, ldlo);
  return CAAE_LAUNCH_STATUS();
}

// Inference epilogue of the last encoder convolution (models/pointnet_ycb_23_decoder_4.py:410-426 dgcnn_agg -> mean over
// the points; :55-60 pn_conv5 -> max): with moving-average batch norm the layer is affine (utils/tf_util.py:507-510), so
// pooled[g][c] = mean / max over the `group` = 256 rows of cloud g of relu((A B + bias)[r][c] * scale[c] + shift[c]) comes
// straight out of the GEMM epilogue and the [M, N] activation (134 MB at B = 128) is never written.  mode 1 = mean,
// 2 = max.  parts f32[ceil(M / 256) * 4][N] is scratch.  Split-precision product when A_lo / B_lo are given.
__global__ void pool_finalize_kernel(int groups, int N, const float* __restrict__ parts, int mode, float inv_group,
                                     float* __restrict__ pooled) {
  pdl_wait();
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < (long)groups * N; e += (long)gridDim.x * blockDim.x) {
    const long g = e / N;
    const int c = (int)(e - g * N);
    const float a = parts[(g * 4 + 0) * N + c], b = parts[(g * 4 + 1) * N + c], c2 = parts[(g * 4 + 2) * N + c],
                d = parts[(g * 4 + 3) * N + c];
    pooled[e] = mode == 2 ? fmaxf(fmaxf(a, b), fmaxf(c2, d)) : ((a + b) + (c2 + d)) * inv_group;
  }
}

extern "C" int caae_gemm_tf32_pool(int M, int N, int K, const float* A, const float* A_lo, int lda, const float* B,
                                   const float* B_lo, int ldb, const float* bias, const float* scale, const float* shift,
                                   int mode, int group, float* parts, float* pooled, caae_stream_t stream) {
  CAAE_RETURN_IF(mode != 1 && mode != 2, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(group != 2 * TBM || M % group != 0, CAAE_E_UNSUPPORTED);   // one 256-row tile = one cloud
  CAAE_RETURN_IF(!scale || !shift || !parts || !pooled, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) != 0 ||
                 (bias && (reinterpret_cast<uintptr_t>(bias) & 15)), CAAE_E_UNSUPPORTED);
  PoolEpilogue pe{mode, scale, shift, parts};
  // (C is not written: the address only has to satisfy the alignment checks of the shared code path)
  const int rc = gemm_tf32_impl(0, 0, M, N, K, A, lda, B, ldb, parts, N, bias, 0, nullptr, stream, A_lo, B_lo, &pe);
  if (rc) return rc;
  const long total = (long)(M / group) * N;
  caae::launch(pool_finalize_kernel, (int)((total + 255) / 256), 256, 0, as_stream(stream), M / group, N, parts, mode, 1.f / (float)group, pooled);
  return CAAE_LAUNCH_STATUS();
}

// Partial rows caae_gemm_tf32_stats writes for this shape (0: the fused-statistics path does not apply).
extern "C" int caae_gemm_tf32_stats_parts(int M, int N, int K, int ldc) {
  static const bool persist_enabled = [] { const char* e = getenv("CAAE_GEMM_PERSIST"); return !(e && e[0] == '0'); }();
  const int num_kb = (K + TBK - 1) / TBK;
  if (!persist_enabled || M <= 0 || N % TBN != 0 || N > PS_MAXN || N < 512 || M < 256 * kNumSMs / 2 || num_kb < 2 ||
      num_kb > 16 || ldc % 4 != 0)
    return 0;
  return 4 * ((M + 2 * TBM - 1) / (2 * TBM));
}

extern "C" int caae_gemm_tf32_stats(int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                                    int ldc, const float* bias, double* parts, caae_stream_t stream) {
  CAAE_RETURN_IF(parts == nullptr, CAAE_E_NULLPTR);
  return gemm_tf32_impl(0, 0, M, N, K, A, lda, B, ldb, C, ldc, bias, 0, parts, stream);
}
