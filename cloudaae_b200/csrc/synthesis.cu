// synthesis.cu — on-line segment synthesis on the GPU: pose transform, spherical occluders, spherical
// flip and hidden-point-removal visibility, visible-prefix selection; plus the Philox normal/uniform
// generator that feeds it.
//
// Reference (all per SAMPLE on CPU threads, the hull inside a tf.py_func holding the GIL):
//   train_cloudAAE_ycbv.py:79-93          get_rotation_matrix / transform_object_model
//   utils/generate_occluder.py:38-81      get_random_spherical_occluder('ycbv')
//   utils/hidden_point_removal.py:6-73    sphericalFlip[_org], convexHull (scipy Qhull), padding
//
// Hidden point removal without a hull data structure.  The reference flips every point p to
// f = p + 2(R-|p|)p/|p|, appends the viewpoint (origin) and calls Qhull; the visible points are the
// hull vertices.  All flipped points have z > 0, so the projective map (x,y,z) -> (x/z, y/z, -1/z)
// is defined on them, preserves convexity and sends the origin to the point at infinity in -w:
// the vertices of conv(F u {0}) are exactly the vertices of the UPPER hull of the lifted points
// (u, v, w).  Point i is such a vertex iff a plane through it lies above every other lifted point:
//        exists s in R^2 :  s . (u_j - u_i, v_j - v_i)  >=  w_j - w_i      for all j != i.
// After subtracting the paraboloid  -rho/2 |(u,v)|^2  of the reference sphere (an exact change of
// variables, rho = max |f|) s = 0 means "tangent to the sphere", distant points can never violate
// for small s, and the test becomes the LP-type problem "minimum-norm s subject to n half-planes",
// solved per point by Seidel's incremental algorithm: keep the current optimum; a violated
// constraint moves it to the minimum-norm point of that constraint's boundary line clipped by all
// constraints seen before; an empty clip interval proves the point hidden.  Constraints are visited
// neighbourhood-first through a 32x32 grid over (u,v) (so re-solves happen while "seen before" is
// ~100 points), then every remaining point as the global verification.
// fp64 predicates on the fp32 flipped coordinates; measured against Qhull on the reference's own
// fixtures the visible sets are identical (tests/test_gpu_synthesis.py reports the IoU).
//
// One CTA (1024 threads) per cloud; everything a cloud needs lives in shared memory.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "hpr_lp.cuh"

namespace cg = cooperative_groups;

namespace caae {

constexpr int SY_THREADS = 1024;
constexpr int SY_MAXN = 2688;            // points per cloud the HPR kernel supports (82 B of shared memory each)

// ---- Philox4x32-10 -----------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

// out[i] ~ N(0,1) (mode 0) or U(0,1) (mode 1); counter = (i/4, stream, *offset) so a replayed CUDA
// graph draws fresh numbers by bumping the device-side offset.
__global__ void philox_fill_kernel(long n, float* __restrict__ out, uint64_t seed, uint32_t stream_id,
                                   const int* __restrict__ offset, int mode) {
  const uint32_t off = offset ? (uint32_t)*offset : 0u;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream_id, off),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float v[4];
    if (mode == 0) {
      const float a0 = sqrtf(-2.f * logf(u01(r.x))), a1 = sqrtf(-2.f * logf(u01(r.z)));
      float s0, c0, s1, c1;
      sincospif(2.f * u01(r.y), &s0, &c0);
      sincospif(2.f * u01(r.w), &s1, &c1);
      v[0] = a0 * c0; v[1] = a0 * s0; v[2] = a1 * c1; v[3] = a1 * s1;
    } else {
      v[0] = u01(r.x); v[1] = u01(r.y); v[2] = u01(r.z); v[3] = u01(r.w);
    }
    for (int e = 0; e < 4; ++e)
      if (q * 4 + e < n) out[q * 4 + e] = v[e];
  }
}

// ---- pose transform + occluder + spherical flip -------------------------------------------------
__device__ __forceinline__ float block_max(float v, float* s_red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = s_red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmaxf(r, s_red[w]);
  __syncthreads();
  return r;
}

// points [b, nm+no, 3] = (model[class] R^T + t) followed by the occluder; flip_all over all of them,
// flip_org over the nm model points only (its own radius).  One CTA per cloud.
__global__ void __launch_bounds__(SY_THREADS)
synth_points_kernel(int nm, int no, const float* __restrict__ models, const int* __restrict__ class_id,
                    const float* __restrict__ axisangle, const float* __restrict__ translation,
                    const float* __restrict__ z_centers, const float* __restrict__ z_points, float hnear, float wnear,
                    float near_dist, float flip_pow, float* __restrict__ points, float* __restrict__ flip_all,
                    float* __restrict__ flip_org) {
  __shared__ float s_R[9], s_t[3], s_c[6], s_red[32];
  const int cloud = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    // R = f32(expmap_f64(axag))  (losses/angular_distance_taylor.py:30-66)
    double a[3];
    for (int c = 0; c < 3; ++c) a[c] = (double)axisangle[cloud * 3 + c];
    const double tsq = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    double t1, t2;
    if (tsq < 1e-2) {
      const double t4 = tsq * tsq, t6 = t4 * tsq, t8 = t4 * t4;
      t1 = 1 - (tsq / 6) + (t4 / 120) - (t6 / 5040) + (t8 / 362880);
      t2 = 0.5 - (tsq / 24) + (t4 / 720) - (t6 / 40320) + (t8 / 3628800);
    } else {
      const double th = sqrt(tsq);
      t1 = sin(th) / th; t2 = (1 - cos(th)) / tsq;
    }
    const double x = a[0], y = a[1], z = a[2];
    const double R[9] = {1 - t2 * (y * y + z * z), -t1 * z + t2 * x * y, t1 * y + t2 * x * z,
                         t1 * z + t2 * x * y, 1 - t2 * (x * x + z * z), -t1 * x + t2 * y * z,
                         -t1 * y + t2 * x * z, t1 * x + t2 * y * z, 1 - t2 * (x * x + y * y)};
    for (int e = 0; e < 9; ++e) s_R[e] = (float)R[e];
    for (int c = 0; c < 3; ++c) s_t[c] = translation[cloud * 3 + c];
    // occluder centres (generate_occluder.py:63-68)
    if (no > 0) {
      const float tz = s_t[2];
      const float mean_z = (near_dist + tz) / 2.f, std_z = (tz - near_dist) / 6.f;
      for (int o = 0; o < 2; ++o) {
        s_c[o * 3 + 0] = z_centers[(cloud * 2 + o) * 3 + 0] * (wnear / 10.f);
        s_c[o * 3 + 1] = z_centers[(cloud * 2 + o) * 3 + 1] * (hnear / 10.f);
        s_c[o * 3 + 2] = __fadd_rn(__fmul_rn(z_centers[(cloud * 2 + o) * 3 + 2], std_z), mean_z);
      }
    }
  }
  __syncthreads();
  const int n = nm + no;
  const float* __restrict__ mdl = models + (size_t)class_id[cloud] * nm * 3;
  float* __restrict__ pts = points + (size_t)cloud * n * 3;
  float nmax_all = 0.f, nmax_org = 0.f;
  for (int i = tid; i < n; i += SY_THREADS) {
    float p[3];
    if (i < nm) {
      const float m0 = mdl[i * 3 + 0], m1 = mdl[i * 3 + 1], m2 = mdl[i * 3 + 2];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        p[c] = __fadd_rn(fmaf(m2, s_R[c * 3 + 2], fmaf(m1, s_R[c * 3 + 1], __fmul_rn(m0, s_R[c * 3 + 0]))), s_t[c]);
    } else {
      // rows alternate blob 1 / blob 2 (concat of the six columns then reshape(-1,3), :76-79)
      const int r = i - nm, blob = r & 1, k = r >> 1;
      const int half = no / 2;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        p[c] = __fadd_rn(__fmul_rn(z_points[(((size_t)cloud * 2 + blob) * half + k) * 3 + c], 0.01f), s_c[blob * 3 + c]);
    }
    pts[i * 3 + 0] = p[0]; pts[i * 3 + 1] = p[1]; pts[i * 3 + 2] = p[2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p[0], p[0]), __fmul_rn(p[1], p[1])), __fmul_rn(p[2], p[2])));
    nmax_all = fmaxf(nmax_all, nrm);
    if (i < nm) nmax_org = fmaxf(nmax_org, nrm);
  }
  nmax_all = block_max(nmax_all, s_red);
  nmax_org = block_max(nmax_org, s_red);
  // R = max|p| * 10^param ; f = (2(R-|p|) p)/|p| + p   (hidden_point_removal.py:13-17)
  const float R_all = __fmul_rn(nmax_all, flip_pow), R_org = __fmul_rn(nmax_org, flip_pow);
  __syncthreads();  // pts[] written above is re-read below by the same threads only; barrier kept for clarity
  for (int i = tid; i < n; i += SY_THREADS) {
    const float p0 = pts[i * 3 + 0], p1 = pts[i * 3 + 1], p2 = pts[i * 3 + 2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(p0, p0), __fmul_rn(p1, p1)), __fmul_rn(p2, p2)));
    const float ga = __fmul_rn(2.f, __fsub_rn(R_all, nrm));
    float* fa = flip_all + ((size_t)cloud * n + i) * 3;
    fa[0] = __fadd_rn(__fdiv_rn(__fmul_rn(ga, p0), nrm), p0);
    fa[1] = __fadd_rn(__fdiv_rn(__fmul_rn(ga, p1), nrm), p1);
    fa[2] = __fadd_rn(__fdiv_rn(__fmul_rn(ga, p2), nrm), p2);
    if (i < nm) {
      const float go = __fmul_rn(2.f, __fsub_rn(R_org, nrm));
      float* fo = flip_org + ((size_t)cloud * nm + i) * 3;
      fo[0] = __fadd_rn(__fdiv_rn(__fmul_rn(go, p0), nrm), p0);
      fo[1] = __fadd_rn(__fdiv_rn(__fmul_rn(go, p1), nrm), p1);
      fo[2] = __fadd_rn(__fdiv_rn(__fmul_rn(go, p2), nrm), p2);
    }
  }
}

// ---- hidden point removal + visible-prefix selection ---------------------------------------------
//
// Shared-memory layout of one cloud (n input points, all arrays sized by n):
//   Us, Vs, Ws   fp64   lifted coordinates, SORTED by grid cell (row-major cells, ascending index inside a cell)
//   SA, SB       fp64   set-up: (u, v) by original index; afterwards the survivors' optima, by survivor slot
//   F4           f32x4  sorted fp32 copy (u, v, w + rho, -) for the verification filter; the 4th word of
//                       entry `slot` holds the survivor's worst-violator key of the current round
//   id           u16    sorted position -> original index
//   surv         u16    set-up: unordered cell fill; afterwards survivor slot -> sorted position
//   vlist        u16    set-up: original index -> cell; afterwards the slots to verify in this round
//   wlist        u16    the slots to re-solve in this round
//   ext          u16x8  per survivor slot: positions of the violated far constraints added to its LP
//   flag         u8     original index -> visible
//   state        u8     set-up: duplicate marks; afterwards per slot: number of ext entries, or kHidden
constexpr int SY_EXTRA = 8;     // far constraints a survivor's LP can take before the full re-solve
constexpr int SY_BYTES_PER_POINT = 5 * 8 + 16 + 4 * 2 + SY_EXTRA * 2 + 2;
constexpr unsigned char kHidden = 0xFF;
static_assert(SY_MAXN < (1 << hpr::kPosBits), "violation_key stores the position in 12 bits");

// Exclusive scan of one int per thread over the 1024-thread CTA; returns the exclusive prefix, `total` = sum.
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp_tot, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  __syncthreads();  // s_warp_tot may still be read by a previous call
  if (lane == 31) s_warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = s_warp_tot[lane];
    int iw = w;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iw, o); if (lane >= o) iw += t; }
    s_warp_tot[lane] = iw - w;
    if (lane == 31) s_warp_tot[32] = iw;
  }
  __syncthreads();
  total = s_warp_tot[32];
  return s_warp_tot[warp] + inc - v;
}

__device__ __forceinline__ double shfl_d(double v, int src) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_sync(0xffffffffu, lo, src); hi = __shfl_sync(0xffffffffu, hi, src);
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, m); hi = __shfl_xor_sync(0xffffffffu, hi, m);
  return __hiloint2double(hi, lo);
}

// a / b for b > 0 in the normal range: reciprocal seed + two Newton steps + one residual correction
// (<= 1 ulp).  The compiler's IEEE division tests every quotient and sends the whole warp through a
// ~65-instruction slow path whenever one lane's numerator or quotient is zero or tiny — which some lane
// of almost every clip is (measured: 13 % of the kernel's instructions).
__device__ __forceinline__ double div_pos(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = fma(fma(-b, r, 1.0), r, r);
  r = fma(fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// Persistent LP solver: every 8-lane GROUP of the calling warp runs the incremental LP of one point
// (the math of hpr::lp_lane, eight constraints per step), and fetches its next point the moment it
// finishes one, so the four groups of a warp never wait for each other.  One loop body serves the
// groups that are scanning and the groups that are clipping; the group-wide exchanges (broadcast of a
// violated constraint, reduction of a clip interval) run only in the steps where some group needs one.
//   fetch(want, gbase, self, seq, tag, sa, sb, p0): warp-uniform call; groups with want = true receive
//        their next point (sorted position `self`, its constraint sequence, a caller tag, and the state
//        to resume from: optimum (sa, sb) of the first p0 constraints) or self = -1 when the work list
//        is exhausted.
//   finish(fin, gbase, self, tag, status, sa, sb): warp-uniform call; groups with fin = true report.
// Lanes per LP group (a power of two).  Eight is the measured choice on B200; HPR_GROUP=4 for the A/B.
#ifndef HPR_GROUP
#define HPR_GROUP 8
#endif
constexpr int kGW = HPR_GROUP;
constexpr unsigned kGMask = (1u << kGW) - 1u;
static_assert(kGW == 4 || kGW == 8 || kGW == 16, "group width");

template <class Fetch, class Finish>
__device__ __forceinline__ void hpr_solve_groups(const hpr::View& h, Fetch fetch, Finish finish) {
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31, g = lane & (kGW - 1), gbase = lane & ~(kGW - 1);
  hpr::RangesPlusList seq;
  int self = -1, tag = 0, L = 0;
  bool exhausted = false;
  double ui = 0.0, vi = 0.0, wi = 0.0, sa = 0.0, sb = 0.0;
  const double hk = 0.5 * h.kappa;
  int p = 0, q = -1, qend = 0;
  double p0a = 0.0, p0b = 0.0, da = 0.0, db = 0.0;
  double ln = -1.0, ld = 0.0, hn = 1.0, hd = 0.0;
  bool infeas = false;
  while (true) {
    // ---- groups without a point fetch one
    const bool want = self < 0 && !exhausted;
    if (__any_sync(kFull, want)) {
      fetch(want, gbase, self, seq, tag, sa, sb, p);
      __syncwarp();   // a group's lanes read what its leader (or fetch itself) wrote to shared memory
      if (want) {
        if (self < 0) exhausted = true;
        else { L = seq.length(); ui = h.U[self]; vi = h.V[self]; wi = h.W[self]; q = -1; }
      }
    }
    if (__all_sync(kFull, self < 0)) break;
    const bool busy = self >= 0;
    // ---- one step: eight constraints of the scan or of the clip
    const bool clipping = q >= 0;
    const int cur = (clipping ? q : p) + g;
    const int j = (busy && cur < (clipping ? qend : L)) ? seq(cur) : -1;
    const bool valid = j >= 0 && j != self;
    double du = 0.0, dv = 0.0, dw = 0.0, r2 = 0.0, rhs = 0.0;
    if (valid) {
      du = h.U[j] - ui; dv = h.V[j] - vi; dw = h.W[j] - wi;
      r2 = du * du + dv * dv;
      rhs = dw - hk * r2;
    }
    const bool viol = valid && !clipping && r2 != 0.0 && (rhs - (sa * du + sb * dv) > 0.0);
    const bool hid = valid && !clipping && r2 == 0.0 && (dw > 0.0 || (dw == 0.0 && h.id[j] < h.id[self]));
    const unsigned bv = (__ballot_sync(kFull, viol) >> gbase) & kGMask;
    const unsigned bh = (__ballot_sync(kFull, hid) >> gbase) & kGMask;
    if (valid && clipping && r2 != 0.0) {
      const double den = da * du + db * dv, num = rhs - (p0a * du + p0b * dv);
      if (den > 0.0) { if (num * ld > ln * den) { ln = num; ld = den; } }
      else if (den < 0.0) { const double nn = -num, dd = -den; if (nn * hd < hn * dd) { hn = nn; hd = dd; } }
      else if (num > 0.0) infeas = true;
    }
    int status = -1;                 // >= 0: the group's point is decided in this step
    bool start_clip = false, end_clip = false;
    int fv = 0;
    if (busy) {
      if (!clipping) {
        fv = bv ? __ffs(bv) - 1 : kGW;
        const int fh = bh ? __ffs(bh) - 1 : kGW;
        if (fh < fv) status = hpr::kLpHidden;          // a "hidden" verdict earlier than any violation
        else if (fv == kGW) { p += kGW; if (p >= L) status = hpr::kLpVisible; }
        else start_clip = true;
      } else {
        q += kGW;
        end_clip = q >= qend;
      }
    }
    if (__any_sync(kFull, start_clip)) {
      const int src = gbase + (start_clip ? fv : 0);
      const double bdu = shfl_d(du, src), bdv = shfl_d(dv, src), brhs = shfl_d(rhs, src);
      if (start_clip) {   // the new optimum lies on the violated constraint's boundary line p0 + t (da, db)
        const double br2 = bdu * bdu + bdv * bdv;
        const double inv = div_pos(brhs, br2 > 0.0 ? br2 : 1.0);
        p0a = bdu * inv; p0b = bdv * inv; da = -bdv; db = bdu;
        ln = -1.0; ld = 0.0; hn = 1.0; hd = 0.0; infeas = false;
        qend = p + fv; p = p + fv + 1; q = 0;
        end_clip = qend == 0;
      }
    }
    if (__any_sync(kFull, end_clip)) {
      // (absent bounds divide by 1)
      const double lds = ld > 0.0 ? ld : 1.0, hds = hd > 0.0 ? hd : 1.0;
      double lo = ld > 0.0 ? div_pos(ln, lds) : -INFINITY, hi = hd > 0.0 ? div_pos(hn, hds) : INFINITY;
#pragma unroll
      for (int m = 1; m < kGW; m <<= 1) { lo = fmax(lo, shfl_xor_d(lo, m)); hi = fmin(hi, shfl_xor_d(hi, m)); }
      const unsigned binf = (__ballot_sync(kFull, infeas) >> gbase) & kGMask;
      if (end_clip) {
        if (binf || lo > hi) status = hpr::kLpHidden;
        else {
          const double t = fmin(fmax(0.0, lo), hi);
          sa = p0a + t * da; sb = p0b + t * db; q = -1;
          if (p >= L) status = hpr::kLpVisible;
        }
      }
    }
    const bool fin = status >= 0;
    if (__any_sync(kFull, fin)) {
      __syncwarp();   // the leader's writes below follow every lane's reads of the slot's state in fetch
      finish(fin, gbase, self, tag, status, sa, sb);
      if (fin) self = -1;
    }
  }
}

// One launch serves up to two independent HPR problems over the same batch (the occluded cloud and the
// bare object of every sample): 2b CTAs keep all SMs busy where two launches of b <= 148 CTAs would
// each wait for their slowest cloud.
struct HprJob {
  int n, org_stride_pts, take;
  const float* flipped; const float* org; const float* pad_uniform;
  float* out_pts; int* num_vis; unsigned char* flags_out;
};
struct HprJobs { HprJob job[2]; int b; };

// Diagnostics: per CTA (modulo 512) clock64 deltas {set-up, phase 1, first verification, remaining rounds,
// compaction+select}, then {survivors, rounds, re-solved in round 0}; read with caae_debug_hpr_timing.
__device__ long long g_hpr_timing[512 * 8];

__global__ void __launch_bounds__(SY_THREADS)
hpr_select_kernel(const __grid_constant__ HprJobs jobs) {
  // One thread-block CLUSTER per cloud.  Every CTA of the cluster builds the same cell-sorted copy of the
  // cloud in its own shared memory (the set-up is ~5 % of the work) and then solves the LPs of the sorted
  // positions p with p % cluster_size == rank: a cloud's latency drops by the cluster size, which is what
  // balances 2b clouds of very different cost over 148 SMs.  The CTAs only exchange their visible counts
  // (per index window) and, at the end, their flags, through distributed shared memory.
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int cluster_id = blockIdx.x / csize;
  // cluster -> (job, cloud): all clouds of job 0 first, then job 1
  const HprJob& J = jobs.job[cluster_id / jobs.b];
  const int n = J.n, org_stride_pts = J.org_stride_pts, take = J.take;
  const float* __restrict__ flipped = J.flipped;
  const float* __restrict__ org = J.org;
  const float* __restrict__ pad_uniform = J.pad_uniform;
  float* __restrict__ out_pts = J.out_pts;
  int* __restrict__ num_vis = J.num_vis;
  unsigned char* __restrict__ flags_out = J.flags_out;
  static_assert(hpr::G * hpr::G == SY_THREADS, "one grid cell per thread");
  constexpr int G = hpr::G;
  extern __shared__ __align__(16) unsigned char sy_smem[];
  double* Us = reinterpret_cast<double*>(sy_smem);
  double* Vs = Us + n;
  double* Ws = Vs + n;
  double* SA = Ws + n;
  double* SB = SA + n;
  float4* F4 = reinterpret_cast<float4*>(SB + n);
  unsigned short* id = reinterpret_cast<unsigned short*>(F4 + n);
  unsigned short* surv = id + n;
  unsigned short* vlist = surv + n;
  unsigned short* wlist = vlist + n;
  unsigned short* ext = wlist + n;                                     // [n][SY_EXTRA]
  unsigned char* flag = reinterpret_cast<unsigned char*>(ext + (size_t)n * SY_EXTRA);
  unsigned char* state = flag + n;
  int* cell_start = reinterpret_cast<int*>(state + ((n + 3) & ~3));    // [G*G + 1]
  int* cell_fill = cell_start + G * G + 1;                             // [G*G]
  int* ids = reinterpret_cast<int*>(sy_smem);                          // reuses the Us region after the LP phases
  unsigned short* tmp = surv;                                          // set-up aliases
  unsigned short* cell_of_orig = vlist;
  __shared__ float s_redf[32];
  __shared__ int s_warp_tot[33];
  __shared__ float s_box[4], s_fzrow[hpr::G], s_fzmax;
  __shared__ int s_count, s_queue, s_nlist, s_nwork;
  __shared__ int s_xchg[2];   // this CTA's visible count of the current index window (double-buffered)

  const int cloud = cluster_id % jobs.b, tid = threadIdx.x, lane = tid & 31;
  const float* __restrict__ f = flipped + (size_t)cloud * n * 3;

  long long tk[6] = {0, 0, 0, 0, 0, 0};
  int dbg_rounds = 0, dbg_resolved = 0;
  if (tid == 0) tk[0] = clock64();
  // ---- rho = max |f|; (u, v) in fp64 by original index; bounding box of (u, v)
  float nmax = 0.f;
  for (int i = tid; i < n; i += SY_THREADS) {
    const float x = f[i * 3 + 0], y = f[i * 3 + 1], z = f[i * 3 + 2];
    nmax = fmaxf(nmax, sqrtf(x * x + y * y + z * z));
  }
  nmax = block_max(nmax, s_redf);
  const double rho = (double)nmax;
  float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
  for (int i = tid; i < n; i += SY_THREADS) {
    const double z = (double)f[i * 3 + 2];
    const double u = (double)f[i * 3 + 0] / z, v = (double)f[i * 3 + 1] / z;
    SA[i] = u; SB[i] = v;
    umin = fminf(umin, (float)u); umax = fmaxf(umax, (float)u);
    vmin = fminf(vmin, (float)v); vmax = fmaxf(vmax, (float)v);
  }
  umax = block_max(umax, s_redf); vmax = block_max(vmax, s_redf);
  umin = -block_max(-umin, s_redf); vmin = -block_max(-vmin, s_redf);
  if (tid == 0) {
    s_box[0] = umin; s_box[1] = vmin; s_box[2] = fmaxf(umax - umin, 1e-30f); s_box[3] = fmaxf(vmax - vmin, 1e-30f);
    s_count = 0; s_queue = 0; s_nlist = 0;
  }
  cell_fill[tid] = 0;
  __syncthreads();

  // grid cell of a point from its fp32 (u, v): the set-up and every later lookup (F4[p].x/.y hold the
  // same fp32 values) evaluate the same expression, so no per-position cell table or search is needed
  const float cell_gx = (float)G / s_box[2], cell_gy = (float)G / s_box[3];
  auto cell_of = [&](float uf, float vf) {
    int cx = (int)((uf - s_box[0]) * cell_gx), cy = (int)((vf - s_box[1]) * cell_gy);
    cx = min(max(cx, 0), G - 1); cy = min(max(cy, 0), G - 1);
    return cy * G + cx;
  };
  // ---- cells; unordered fill of the cell lists
  for (int i = tid; i < n; i += SY_THREADS) {
    const int c = cell_of((float)SA[i], (float)SB[i]);
    cell_of_orig[i] = (unsigned short)c;
    flag[i] = 0;
    atomicAdd(&cell_fill[c], 1);
  }
  __syncthreads();
  int total;
  {
    const int excl = block_excl_scan(cell_fill[tid], s_warp_tot, total);
    cell_start[tid] = excl;
    if (tid == SY_THREADS - 1) cell_start[G * G] = total;
    cell_fill[tid] = excl;
  }
  __syncthreads();
  for (int i = tid; i < n; i += SY_THREADS) tmp[atomicAdd(&cell_fill[cell_of_orig[i]], 1)] = (unsigned short)i;
  __syncthreads();
  // ---- exact duplicates (the shipped models contain some: class 17 stores 574 copies of one point):
  // only the lowest index of identical points takes part; the copies are hidden by definition.
  for (int i = tid; i < n; i += SY_THREADS) {
    const int c = cell_of_orig[i];
    const float x = f[i * 3 + 0], y = f[i * 3 + 1], z = f[i * 3 + 2];
    bool dup = false;
    for (int k = cell_start[c]; k < cell_start[c + 1] && !dup; ++k) {
      const int j = tmp[k];
      dup = j < i && f[j * 3 + 0] == x && f[j * 3 + 1] == y && f[j * 3 + 2] == z;
    }
    state[i] = dup ? 2 : 0;
  }
  cell_fill[tid] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += SY_THREADS)
    if (state[i] != 2) atomicAdd(&cell_fill[cell_of_orig[i]], 1);
  __syncthreads();
  int n_unique;
  {
    const int excl = block_excl_scan(cell_fill[tid], s_warp_tot, n_unique);
    cell_fill[tid] = excl;  // start of the cell in the final order
  }
  __syncthreads();
  // ---- final order: cell-major, ascending original index inside a cell (deterministic)
  for (int i = tid; i < n; i += SY_THREADS) {
    if (state[i] == 2) continue;
    const int c = cell_of_orig[i];
    int rank = 0;
    for (int k = cell_start[c]; k < cell_start[c + 1]; ++k) {
      const int j = tmp[k];
      rank += (j < i && state[j] != 2) ? 1 : 0;
    }
    id[cell_fill[c] + rank] = (unsigned short)i;
  }
  __syncthreads();
  cell_start[tid] = cell_fill[tid];
  if (tid == 0) cell_start[G * G] = n_unique;
  __syncthreads();
  // ---- lifted coordinates in sorted order
  for (int p = tid; p < n_unique; p += SY_THREADS) {
    const int i = id[p];
    const double u = SA[i], v = SB[i], z = (double)f[i * 3 + 2];
    const double w = -rho * rho / z + 0.5 * rho * (u * u + v * v);  // paraboloid-shifted lift (hpr::lift)
    Us[p] = u; Vs[p] = v; Ws[p] = w;
    F4[p] = make_float4((float)u, (float)v, (float)(w + rho), 0.f);  // w ~ -rho: recentre before rounding to fp32
  }
  __syncthreads();
  if (tid < G) {   // per grid row: maximum of the fp32 (w + rho) — the verification's bound on w_j
    float m = -3.4e38f;
    for (int p = cell_start[tid * G]; p < cell_start[(tid + 1) * G]; ++p) m = fmaxf(m, F4[p].z);
    s_fzrow[tid] = m;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));   // tid < G = one warp
    if (tid == 0) s_fzmax = m;
  }
  __syncthreads();

  hpr::View h;
  h.U = Us; h.V = Vs; h.W = Ws; h.id = id; h.cell_start = cell_start; h.kappa = rho; h.n_unique = n_unique;

  if (tid == 0) tk[1] = clock64();
  // ---- visible-prefix mode.  The caller consumes only the first `take` visible points in index order
  // (the training graph slices vis[:, :num_point], train_cloudAAE_ycbv.py:213-214), so only points up to
  // the index where take + 1 visible ones are known need an LP of their own (the +1: the reference drops
  // the highest visible index, hidden_point_removal.py:36).  Every point still acts as a CONSTRAINT of
  // those LPs.  Indices are handed out in growing windows [lo, hi) until enough are found; the object
  // models are stored in farthest-point order, so ~3 take indices usually suffice and the dense
  // occluder blobs at the end of the cloud are never solved for.
  const bool prefix = flags_out == nullptr && take + 1 < n;
  int lo = 0, hi = prefix ? min(n, 3 * take) : n;
  int slot_base = 0, nsurv = 0, window = 0;
  const float kh = 0.5f * (float)rho;
  while (true) {
  // ---- phase 1: incremental LP of every point of the window over its 3x3 cell neighbourhood (own grid
  // row first), eight lanes per point, points handed out from a work list.  Survivors (~40 %) are
  // recorded with their optimum.
  if (tid == 0) { s_nwork = 0; s_queue = 0; }
  __syncthreads();
  for (int p = tid; p < n_unique; p += SY_THREADS) {
    const int i = id[p];
    if (i >= lo && i < hi && p % csize == crank) wlist[atomicAdd(&s_nwork, 1)] = (unsigned short)p;
  }
  __syncthreads();
  const int nwork1 = s_nwork;
  {
    auto fetch = [&](bool want, int gbase, int& self, hpr::RangesPlusList& seq, int& tag, double& sa, double& sb, int& p0) {
      int t = 0;
      if (want && (lane & (kGW - 1)) == 0) t = atomicAdd(&s_queue, 1);
      t = __shfl_sync(0xffffffffu, t, gbase);
      if (want) {
        sa = 0.0; sb = 0.0; p0 = 0;
        if (t < nwork1) {
          const int pos = wlist[t];
          const int c = cell_of(F4[pos].x, F4[pos].y);
          int A[3], B[3];
          hpr::nbhd_ranges(cell_start, c % G, c / G, hpr::nbhd_halfwidth(cell_start, c), A, B);
          seq = hpr::RangesPlusList(A, B, nullptr, 0);
          self = pos; tag = 0;
        } else self = -1;
      }
    };
    auto finish = [&](bool fin, int gbase, int self, int tag, int status, double sa, double sb) {
      const bool alive = fin && status == hpr::kLpVisible;
      int slot = 0;
      if (alive && (lane & (kGW - 1)) == 0) {
        slot = atomicAdd(&s_count, 1);
        surv[slot] = (unsigned short)self; SA[slot] = sa; SB[slot] = sb; state[slot] = 0;
      }
    };
    hpr_solve_groups(h, fetch, finish);
  }
  __syncthreads();
  // ---- phases 2 / 3, in rounds.  Verify: every listed survivor's optimum against all points, as
  // (survivor, 256-position slice) work items; lanes of a warp share the slice (broadcast LDS.128 of the
  // fp32 copy); a conservative fp32 evaluation dismisses constraints that are clearly slack, the few
  // within `tol` of tight are re-evaluated in fp64 and the worst genuine violation is recorded in the
  // slot's key.  Re-solve: one thread per violated survivor adds that constraint to its LP (neighbourhood
  // + added constraints) and solves it again; it is verified again in the next round.  A survivor is
  // visible once a verification finds nothing.
  if (tid == 0 && lo == 0) tk[2] = clock64();
  nsurv = s_count;
  int nlist = nsurv - slot_base;   // round 0 verifies every new survivor (list = identity)
  bool first = true;
  while (nlist > 0) {
    if (tid == 0) { s_nwork = 0; s_queue = 0; s_nlist = 0; }
    // ---- verify: eight lanes per listed survivor.  A violator of the survivor's optimum s lies inside a
    // disk around (u_i, v_i) - s/kappa (hpr::verify_disk2, evaluated per grid row with the row's maximum
    // of w), so the group walks only the grid rows and columns that disk touches, minus the 3x3
    // neighbourhood phase 1 already enforced — typically a few dozen positions instead of all n.  A
    // conservative fp32 evaluation dismisses constraints that are clearly slack, the rest are re-evaluated
    // in fp64 and the worst genuine violation becomes the slot's key.
    {
      const int g = lane & 7;
      const unsigned gmask = 0xFFu << (lane & ~7);
      const float gx = (float)G / s_box[2], gy = (float)G / s_box[3];
      const float chh = s_box[3] / (float)G;   // cell height
      for (int li = tid >> 3; li < nlist; li += SY_THREADS / 8) {
        const int slot = first ? slot_base + li : (int)vlist[li];
        const int ne = state[slot];
        if (ne == kHidden) continue;
        const int i = surv[slot];
        const float4 fi = F4[i];
        const double sa = SA[slot], sb = SB[slot];
        const float saf = (float)sa, sbf = (float)sb;
        const int c = cell_of(F4[i].x, F4[i].y);
        const int ccx = c % G, ccy = c / G, nk = hpr::nbhd_halfwidth(cell_start, c);
        const int nx0 = max(ccx - nk, 0), nx1 = min(ccx + nk, G - 1);
        const float inv2kh = 1.f / (2.f * kh);
        const float ctru = fi.x - saf * inv2kh, ctrv = fi.y - sbf * inv2kh;   // disk centre
        unsigned key = 0;
        bool hidden = false;
        // rows the disk can touch at all (radius from the cloud-wide maximum of w), then row by row
        const float rg = sqrtf(fmaxf(hpr::verify_disk2(saf, sbf, fi.z, s_fzmax, kh), 0.f)) * 1.001f;
        const int cy0 = max((int)floorf((ctrv - rg - s_box[1]) * gy - 0.01f), 0);
        const int cy1 = min((int)floorf((ctrv + rg - s_box[1]) * gy + 0.01f), G - 1);
        for (int cy = cy0; cy <= cy1; ++cy) {
          const bool nrow = cy >= ccy - nk && cy <= ccy + nk;
          const float r2 = hpr::verify_disk2(saf, sbf, fi.z, s_fzrow[cy], kh);
          const float lo = s_box[1] + (float)cy * chh, hi = lo + chh;
          const float dv = fmaxf(fmaxf(lo - ctrv, ctrv - hi), 0.f) * 0.999f - 0.01f * chh;  // rounded inwards
          const float w2 = r2 - (dv > 0.f ? dv * dv : 0.f);
          int cx0 = nx0, cx1 = nx0 - 1;   // empty
          if (w2 >= 0.f) {
            const float hw = sqrtf(w2) * 1.001f;
            cx0 = max((int)floorf((ctru - hw - s_box[0]) * gx - 0.01f), 0);
            cx1 = min((int)floorf((ctru + hw - s_box[0]) * gx + 0.01f), G - 1);
          }
          if (cx1 < cx0) continue;   // (the neighbourhood's own cells need no visit: phase 1 enforced them)
          // the row's window [ja, jd) minus the neighbourhood's columns [jb, jc) when the row is one of its three
          const int ja = cell_start[cy * G + cx0], jd = cell_start[cy * G + cx1 + 1];
          const int jb = nrow ? max(cell_start[cy * G + nx0], ja) : jd;
          const int jc = nrow ? min(cell_start[cy * G + nx1 + 1], jd) : jd;
          for (int j0 = ja; j0 < jd; j0 += 8) {
            const int j = j0 + g;
            if (j >= jd || (j >= jb && j < jc)) continue;
            const float4 fj = F4[j];
            if (hpr::clearly_slack(fi.x, fi.y, fi.z, fj.x, fj.y, fj.z, saf, sbf, kh)) continue;
            bool known = false;   // already one of the LP's constraints: tight up to rounding, never re-added
            for (int e = 0; e < ne; ++e) known = known || ext[slot * SY_EXTRA + e] == j;
            if (known) continue;
            bool same_dir;
            const double viol = hpr::violation(h, i, j, sa, sb, same_dir);
            if (same_dir) { hidden = hidden || Ws[j] > Ws[i] || (Ws[j] == Ws[i] && id[j] < id[i]); continue; }
            if (viol > 0.0) key = max(key, hpr::violation_key(viol, j));
          }
        }
#pragma unroll
        for (int m = 1; m < 8; m <<= 1) {
          key = max(key, __shfl_xor_sync(gmask, key, m));
          hidden = __shfl_xor_sync(gmask, (int)hidden, m) || hidden;
        }
        if (g == 0) {
          if (hidden) state[slot] = kHidden;
          else if (key) F4[slot].w = __uint_as_float(key);
        }
      }
    }
    __syncthreads();
    // ---- the listed slots whose key is set go to the work list (hidden / verified-clean ones drop out)
    for (int li = tid; li < nlist; li += SY_THREADS) {
      const int slot = first ? slot_base + li : (int)vlist[li];
      const int ne = state[slot];
      if (ne == kHidden || __float_as_uint(F4[slot].w) == 0u) continue;
      if (ne == SY_EXTRA) {   // constraint list full (never seen on the fixtures): the full LP settles it
        const int i = surv[slot];
        const int c = cell_of(F4[i].x, F4[i].y);
        int A[3], B[3], FA[7], FB[7];
        const int nk = hpr::nbhd_halfwidth(cell_start, c);
        hpr::nbhd_ranges(cell_start, c % G, c / G, nk, A, B);
        hpr::full_ranges(A, B, c / G, nk, n_unique, FA, FB);
        double sa, sb;
        if (hpr::lp_lane(h, i, hpr::Ranges<7>(FA, FB), sa, sb) != hpr::kLpVisible) state[slot] = kHidden;
        continue;
      }
      wlist[atomicAdd(&s_nwork, 1)] = (unsigned short)slot;
    }
    __syncthreads();
    const int nwork = s_nwork;
    if (tid == 0 && first && lo == 0) { tk[3] = clock64(); dbg_resolved = nwork; }
    ++dbg_rounds;
    // ---- re-solve: the worst violator joins the slot's constraint list and the incremental LP takes
    // that one more step (clip its boundary line against neighbourhood + list); slots that stay feasible
    // are verified again in the next round
    {
      auto fetch = [&](bool want, int gbase, int& self, hpr::RangesPlusList& seq, int& tag, double& sa, double& sb, int& p0) {
        int t = 0;
        if (want && (lane & (kGW - 1)) == 0) t = atomicAdd(&s_queue, 1);
        t = __shfl_sync(0xffffffffu, t, gbase);
        if (want) {
          if (t < nwork) {
            const int slot = wlist[t];
            const int i = surv[slot];
            const int c = cell_of(F4[i].x, F4[i].y);
            int A[3], B[3];
            hpr::nbhd_ranges(cell_start, c % G, c / G, hpr::nbhd_halfwidth(cell_start, c), A, B);
            const int ne = state[slot];
            const unsigned key = __float_as_uint(F4[slot].w);
            // written by the group's leader; the __syncwarp after fetch orders it before the other lanes' reads
            if ((lane & (kGW - 1)) == 0) ext[slot * SY_EXTRA + ne] = (unsigned short)(key & ((1u << hpr::kPosBits) - 1u));
            seq = hpr::RangesPlusList(A, B, ext + slot * SY_EXTRA, ne + 1);
            self = i; tag = slot;
            // resume: (SA, SB) is the optimum of everything before the new constraint, which is violated
            sa = SA[slot]; sb = SB[slot]; p0 = seq.length() - 1;
          } else self = -1;
        }
      };
      auto finish = [&](bool fin, int gbase, int self, int tag, int status, double sa, double sb) {
        if (fin && (lane & (kGW - 1)) == 0) {
          F4[tag].w = 0.f;
          if (status != hpr::kLpVisible) state[tag] = kHidden;
          else {
            SA[tag] = sa; SB[tag] = sb;
            state[tag] = (unsigned char)(state[tag] + 1);
            vlist[atomicAdd(&s_nlist, 1)] = (unsigned short)tag;
          }
        }
      };
      hpr_solve_groups(h, fetch, finish);
    }
    __syncthreads();
    nlist = s_nlist;
    first = false;
    __syncthreads();
  }
  for (int slot = slot_base + tid; slot < nsurv; slot += SY_THREADS)
    if (state[slot] != kHidden) flag[id[surv[slot]]] = 1;
  __syncthreads();
  if (hi >= n) break;
  {  // enough visible points known?  (sum over the cluster's CTAs; every CTA takes the same decision)
    int c = 0;
    for (int i = tid; i < hi; i += SY_THREADS) c += flag[i];
    int vis_known;
    block_excl_scan(c, s_warp_tot, vis_known);
    if (csize > 1) {
      if (tid == 0) s_xchg[window & 1] = vis_known;
      cluster.sync();
      for (int r = 0; r < csize; ++r)
        if (r != crank) vis_known += *cluster.map_shared_rank(&s_xchg[window & 1], r);
    }
    if (vis_known >= take + 1) break;
  }
  slot_base = nsurv; lo = hi; hi = min(n, hi + 2 * take); ++window;
  }
  if (tid == 0) { tk[4] = clock64(); if (tk[3] == 0) tk[3] = tk[4]; }
  if (csize > 1) {   // rank 0 collects the other CTAs' flags and finishes alone
    cluster.sync();
    if (crank == 0)
      for (int r = 1; r < csize; ++r) {
        const unsigned char* rf = cluster.map_shared_rank(flag, r);
        for (int i = tid; i < n; i += SY_THREADS) flag[i] |= rf[i];
      }
    cluster.sync();   // remote shared memory stays valid until every reader is done
    if (crank != 0) return;
  }

  // ---- ordered compaction by ORIGINAL index (visible ids ascending)
  const int per = (n + SY_THREADS - 1) / SY_THREADS;
  const int i0 = tid * per, i1 = min(n, i0 + per);
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += flag[i];
  int nvis_all;
  int pos = block_excl_scan(cnt, s_warp_tot, nvis_all);   // its barriers also retire the LP phases: Us may be reused
  for (int i = i0; i < i1; ++i) {
    if (flags_out) flags_out[(size_t)cloud * n + i] = flag[i];
    if (flag[i]) ids[pos++] = i;
  }
  __syncthreads();
  // reference quirk: visibleId[:-1] drops the highest-index visible point (hidden_point_removal.py:36)
  const int nv = max(nvis_all - 1, 0);
  if (tid == 0) num_vis[cloud] = nv;
  const float* __restrict__ op = org + (size_t)cloud * org_stride_pts * 3;
  float* __restrict__ dst = out_pts + (size_t)cloud * take * 3;
  for (int r = tid; r < take; r += SY_THREADS) {
    int src;
    if (r < nv) src = ids[r];
    else if (nv > 0) {  // np.random.choice(visibleId, ...) padding (:38-40)
      const float uu = pad_uniform ? pad_uniform[(size_t)cloud * take + r] : ((float)((r - nv) % nv) + 0.5f) / (float)nv;
      src = ids[max(0, min((int)(uu * (float)nv), nv - 1))];   // clamped: a draw outside [0,1) must not index out of range
    } else src = -1;
    dst[r * 3 + 0] = src >= 0 ? op[src * 3 + 0] : 0.f;
    dst[r * 3 + 1] = src >= 0 ? op[src * 3 + 1] : 0.f;
    dst[r * 3 + 2] = src >= 0 ? op[src * 3 + 2] : 0.f;
  }
  if (tid == 0) {
    tk[5] = clock64();
    long long* g = g_hpr_timing + (cluster_id % 512) * 8;
    for (int k = 0; k < 5; ++k) g[k] = tk[k + 1] - tk[k];
    g[5] = nsurv; g[6] = dbg_rounds; g[7] = dbg_resolved;
  }
}

}  // namespace caae

using namespace caae;

extern "C" int caae_philox_fill(long n, float* out, unsigned long long seed, int stream_id, const int* offset,
                                int uniform, caae_stream_t stream) {
  CAAE_RETURN_IF(n < 0, CAAE_E_BADSHAPE);
  if (n == 0) return CAAE_OK;
  CAAE_RETURN_IF(!out, CAAE_E_NULLPTR);
  long blocks = ((n + 3) / 4 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  philox_fill_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(n, out, (uint64_t)seed, (uint32_t)stream_id, offset,
                                                                 uniform);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_synth_points(int b, int nm, int no, const float* models, const int* class_id,
                                 const float* axisangle, const float* translation, const float* z_centers,
                                 const float* z_points, float hnear, float wnear, float near_dist, float flip_pow,
                                 float* points, float* flip_all, float* flip_org, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || nm <= 0 || no < 0 || (no & 1), CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!models || !class_id || !axisangle || !translation || !points || !flip_all || !flip_org, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(no > 0 && (!z_centers || !z_points), CAAE_E_NULLPTR);
  synth_points_kernel<<<b, SY_THREADS, 0, as_stream(stream)>>>(nm, no, models, class_id, axisangle, translation,
                                                              z_centers, z_points, hnear, wnear, near_dist, flip_pow,
                                                              points, flip_all, flip_org);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_debug_hpr_timing(long long* host_buf) {
  CAAE_RETURN_IF(!host_buf, CAAE_E_NULLPTR);
  return (int)cudaMemcpyFromSymbol(host_buf, g_hpr_timing, sizeof(long long) * 512 * 8);
}

static int hpr_launch(int b, int njobs, const HprJob* jobs, caae_stream_t stream) {
  HprJobs J;
  J.b = b;
  int nmax = 0;
  for (int k = 0; k < 2; ++k) {
    J.job[k] = jobs[k < njobs ? k : 0];
    const HprJob& j = J.job[k];
    CAAE_RETURN_IF(j.n <= 0 || j.n > SY_MAXN || j.take <= 0 || j.org_stride_pts < j.n, CAAE_E_BADSHAPE);
    CAAE_RETURN_IF(!j.flipped || !j.org || !j.out_pts || !j.num_vis, CAAE_E_NULLPTR);
    nmax = j.n > nmax ? j.n : nmax;
  }
  size_t smem = (size_t)nmax * SY_BYTES_PER_POINT + 8 + sizeof(int) * (2 * hpr::G * hpr::G + 1) + 16;
  smem = (smem + 15) & ~(size_t)15;
  static size_t smem_set = 0;   // opt in once per size class, not on every call
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(hpr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    smem_set = smem;
  }
  // cluster size = CTAs per cloud (CAAE_HPR_CLUSTER = 1, 2 or 4; read once).  Measured on B200, batch 128:
  // stand-alone the kernel is fastest with 2 (1.37 ms vs 1.48 ms), but next to the train step of the
  // previous batch (CloudAAETrainer.capture_online_pipelined) the extra set-up work of the second CTA costs
  // more than the shorter tail saves (step 3.00 ms vs 2.81 ms), so the default is 1.
  static int cluster = 0;
  if (cluster == 0) {
    const char* e = getenv("CAAE_HPR_CLUSTER");
    const int v = e ? atoi(e) : 1;
    cluster = (v == 1 || v == 2 || v == 4) ? v : 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(b * njobs * cluster));
  cfg.blockDim = dim3(SY_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, hpr_select_kernel, J);
  if (le != cudaSuccess) return (int)le;
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_hpr_select(int b, int n, const float* flipped, const float* org, int org_stride_pts, int take,
                               const float* pad_uniform, float* out_pts, int* num_vis, unsigned char* flags_out,
                               caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  const HprJob job = {n, org_stride_pts, take, flipped, org, pad_uniform, out_pts, num_vis, flags_out};
  return hpr_launch(b, 1, &job, stream);
}

extern "C" int caae_hpr_select_pair(int b, int n_a, const float* flipped_a, int take_a, const float* pad_uniform_a,
                                    float* out_pts_a, int* num_vis_a, int n_b, const float* flipped_b, int take_b,
                                    const float* pad_uniform_b, float* out_pts_b, int* num_vis_b, const float* org,
                                    int org_stride_pts, caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  const HprJob jobs[2] = {{n_a, org_stride_pts, take_a, flipped_a, org, pad_uniform_a, out_pts_a, num_vis_a, nullptr},
                          {n_b, org_stride_pts, take_b, flipped_b, org, pad_uniform_b, out_pts_b, num_vis_b, nullptr}};
  return hpr_launch(b, 2, jobs, stream);
}
