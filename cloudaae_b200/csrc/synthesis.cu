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
#include "common.cuh"

namespace caae {

constexpr int SY_THREADS = 1024;
constexpr int SY_G = 32;                 // grid cells per axis over the (u,v) bounding box (one cell per thread)
constexpr int SY_MAXN = 4096;            // points per cloud the HPR kernel supports
constexpr int SY_LOCAL = 9;              // phase 1 visits the 3x3 cell neighbourhood

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

// Everything the per-point LP needs, in shared memory.
struct HprShared {
  const double *U, *V, *W;          // lifted coordinates (paraboloid-shifted), fp64
  const unsigned short* order;      // point ids sorted by grid cell
  const unsigned char *cellx, *celly;
  const int* cell_start;            // [SY_G*SY_G + 1]
  const unsigned char* dup;         // 2 = exact copy of a lower-index point (not a constraint)
  double kappa;
  int n;
};

// The constraint sequence of point i is: its 3x3 cell neighbourhood (9 cell segments of `order`),
// then every point OUTSIDE that neighbourhood in index order.  Nbhd caches the 9 segments.
struct Nbhd {
  int* start;  // [SY_LOCAL]      per-warp shared-memory scratch
  int* pre;    // [SY_LOCAL + 1]  exclusive prefix of the segment lengths
  int cx, cy;
};
__device__ __forceinline__ void make_nbhd(const HprShared& h, int i, int* scratch, Nbhd& nb) {
  nb.start = scratch; nb.pre = scratch + SY_LOCAL;
  nb.cx = h.cellx[i]; nb.cy = h.celly[i];
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    int run = 0;
    nb.pre[0] = 0;
    for (int o = 0; o < SY_LOCAL; ++o) {
      const int x = nb.cx + (o % 3) - 1, y = nb.cy + (o / 3) - 1;
      const bool valid = x >= 0 && x < SY_G && y >= 0 && y < SY_G;
      const int cell = valid ? y * SY_G + x : 0;
      nb.start[o] = valid ? h.cell_start[cell] : 0;
      run += valid ? h.cell_start[cell + 1] - h.cell_start[cell] : 0;
      nb.pre[o + 1] = run;
    }
  }
  __syncwarp();
}
__device__ __forceinline__ int nbhd_elem(const HprShared& h, const Nbhd& nb, int e) {
  int o = 0;
#pragma unroll
  for (int q = 1; q < SY_LOCAL; ++q) o += (e >= nb.pre[q]) ? 1 : 0;
  return h.order[nb.start[o] + (e - nb.pre[o])];
}
__device__ __forceinline__ bool in_nbhd(const HprShared& h, const Nbhd& nb, int j) {
  return abs((int)h.cellx[j] - nb.cx) <= 1 && abs((int)h.celly[j] - nb.cy) <= 1;
}

// Warp-cooperative clip: minimum-norm point of the boundary line of constraint (du,dv,rhs) of point i,
// subject to the first `e_end` neighbourhood constraints and the non-neighbourhood points of index
// < j_end.  All arguments are warp-uniform; all lanes return the same result.
__device__ bool hpr_clip_warp(const HprShared& h, const Nbhd& nb, int i, double ui, double vi, double wi, double du,
                              double dv, double rhs, int e_end, int j_end, double& sa, double& sb) {
  const int lane = threadIdx.x & 31;
  const double r2 = du * du + dv * dv;
  const double p0a = du * rhs / r2, p0b = dv * rhs / r2, da = -dv, db = du;
  double lo = -INFINITY, hi = INFINITY;
  auto clip = [&](int j) {
    const double eu = h.U[j] - ui, ev = h.V[j] - vi;
    const double s2 = eu * eu + ev * ev;
    if (s2 == 0.0) return;
    const double rk = (h.W[j] - wi) - 0.5 * h.kappa * s2;
    const double den = da * eu + db * ev, num = rk - (p0a * eu + p0b * ev);
    if (den > 0.0) lo = fmax(lo, num / den);
    else if (den < 0.0) hi = fmin(hi, num / den);
    else if (num > 0.0) lo = INFINITY;
  };
  for (int e = lane; e < e_end; e += 32) {
    const int j = nbhd_elem(h, nb, e);
    if (j != i) clip(j);
  }
  for (int j = lane; j < j_end; j += 32)
    if (!in_nbhd(h, nb, j) && h.dup[j] != 2) clip(j);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) { lo = fmax(lo, shfl_xor_d(lo, m)); hi = fmin(hi, shfl_xor_d(hi, m)); }
  if (lo > hi) return false;
  const double t = fmin(fmax(0.0, lo), hi);
  sa = p0a + t * da; sb = p0b + t * db;
  return true;
}

// One warp runs the incremental LP of point i over elements [e_from, m) of its neighbourhood and then
// (when j_to > 0) over the non-neighbourhood points [0, j_to).  32 constraints are tested per step; the
// first violated one (in sequence order) triggers a cooperative clip and the scan resumes right after
// it.  Returns false when the point is proven hidden; (sa, sb) is the running optimum.
__device__ bool hpr_lp_warp(const HprShared& h, const Nbhd& nb, int i, int j_to, double& sa, double& sb) {
  const int lane = threadIdx.x & 31;
  const double ui = h.U[i], vi = h.V[i], wi = h.W[i];
  const int m = nb.pre[SY_LOCAL];
  // pos runs over the concatenated sequence: [0, m) neighbourhood elements, [m, m + j_to) = index j = pos - m
  const int total = m + j_to;
  int pos = 0;
  while (pos < total) {
    const int q = pos + lane;
    int j = -1;
    if (q < m) j = nbhd_elem(h, nb, q);
    else if (q < total) { j = q - m; if (in_nbhd(h, nb, j) || h.dup[j] == 2) j = -1; }
    bool viol = false, hidden = false;
    double du = 0.0, dv = 0.0, rhs = 0.0;
    if (j >= 0 && j != i) {
      du = h.U[j] - ui; dv = h.V[j] - vi;
      const double r2 = du * du + dv * dv;
      const double dw = h.W[j] - wi;
      if (r2 == 0.0) hidden = dw > 0.0 || (dw == 0.0 && j < i);   // same direction: nearer / lower index stays
      else { rhs = dw - 0.5 * h.kappa * r2; viol = rhs - (sa * du + sb * dv) > 0.0; }
    }
    const unsigned mv = __ballot_sync(0xffffffffu, viol), mh = __ballot_sync(0xffffffffu, hidden);
    const int fv = mv ? __ffs(mv) - 1 : 32, fh = mh ? __ffs(mh) - 1 : 32;
    if (fh < fv) return false;                 // a "hidden" verdict earlier in the sequence than any violation
    if (fv == 32) { pos += 32; continue; }
    du = shfl_d(du, fv); dv = shfl_d(dv, fv); rhs = shfl_d(rhs, fv);
    const int qv = pos + fv;                   // sequence position of the violated constraint
    if (!hpr_clip_warp(h, nb, i, ui, vi, wi, du, dv, rhs, min(qv, m), max(qv - m, 0), sa, sb)) return false;
    pos = qv + 1;
  }
  return true;
}

constexpr int SY_SLICE = 256;  // phase-2 verification: points per (survivor, slice) work item

__global__ void __launch_bounds__(SY_THREADS)
hpr_select_kernel(int n, const float* __restrict__ flipped, const float* __restrict__ org, int org_stride_pts, int take,
                  const float* __restrict__ pad_uniform, float* __restrict__ out_pts, int* __restrict__ num_vis,
                  unsigned char* __restrict__ flags_out) {
  extern __shared__ __align__(16) unsigned char sy_smem[];
  double* U = reinterpret_cast<double*>(sy_smem);
  double* V = U + n;
  double* W = V + n;
  double* SA = W + n;                                          // survivors' optima after phase 1
  double* SB = SA + n;
  float4* F4 = reinterpret_cast<float4*>(SB + n);              // fp32 copy (u, v, w - w_ref, 0) for the phase-2 filter
  unsigned short* order = reinterpret_cast<unsigned short*>(F4 + n);
  unsigned short* surv = order + n;
  unsigned char* cellx = reinterpret_cast<unsigned char*>(surv + n);
  unsigned char* celly = cellx + n;
  unsigned char* flag = celly + n;
  unsigned char* dirty = flag + n;
  int* cell_start = reinterpret_cast<int*>(dirty + ((n + 3) & ~3));   // [SY_G*SY_G + 1]
  int* cell_fill = cell_start + SY_G * SY_G + 1;                       // [SY_G*SY_G]
  int* ids = reinterpret_cast<int*>(sy_smem);  // reuses the U region after the LP phases
  __shared__ float s_redf[32];
  __shared__ int s_warp_tot[32];
  __shared__ float s_box[4];
  __shared__ int s_count, s_queue;
  __shared__ int s_nb[SY_THREADS / 32][2 * SY_LOCAL + 2];

  const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* __restrict__ f = flipped + (size_t)cloud * n * 3;

  // ---- rho = max |f|, then lifted coordinates in fp64
  float nmax = 0.f;
  for (int i = tid; i < n; i += SY_THREADS) {
    const float x = f[i * 3 + 0], y = f[i * 3 + 1], z = f[i * 3 + 2];
    nmax = fmaxf(nmax, sqrtf(x * x + y * y + z * z));
  }
  nmax = block_max(nmax, s_redf);
  const double rho = (double)nmax;
  float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
  for (int i = tid; i < n; i += SY_THREADS) {
    const double x = (double)f[i * 3 + 0], y = (double)f[i * 3 + 1], z = (double)f[i * 3 + 2];
    const double u = x / z, v = y / z;
    U[i] = u; V[i] = v;
    const double w = -rho * rho / z + 0.5 * rho * (u * u + v * v);  // paraboloid-shifted lift
    W[i] = w;
    F4[i] = make_float4((float)u, (float)v, (float)(w + rho), 0.f);  // w ~ -rho: recentre before rounding to fp32
    umin = fminf(umin, (float)u); umax = fmaxf(umax, (float)u);
    vmin = fminf(vmin, (float)v); vmax = fmaxf(vmax, (float)v);
  }
  umax = block_max(umax, s_redf); vmax = block_max(vmax, s_redf);
  umin = -block_max(-umin, s_redf); vmin = -block_max(-vmin, s_redf);
  if (tid == 0) {
    s_box[0] = umin; s_box[1] = vmin; s_box[2] = fmaxf(umax - umin, 1e-30f); s_box[3] = fmaxf(vmax - vmin, 1e-30f);
    s_count = 0; s_queue = 0;
  }
  for (int c = tid; c < SY_G * SY_G; c += SY_THREADS) cell_fill[c] = 0;
  __syncthreads();

  // ---- counting sort of the points into grid cells
  for (int i = tid; i < n; i += SY_THREADS) {
    int cx = (int)(((float)U[i] - s_box[0]) / s_box[2] * SY_G), cy = (int)(((float)V[i] - s_box[1]) / s_box[3] * SY_G);
    cx = min(max(cx, 0), SY_G - 1); cy = min(max(cy, 0), SY_G - 1);
    cellx[i] = (unsigned char)cx; celly[i] = (unsigned char)cy;
    flag[i] = 0;
    atomicAdd(&cell_fill[cy * SY_G + cx], 1);
  }
  __syncthreads();
  {  // exclusive scan of SY_G*SY_G (= SY_THREADS) counts, one per thread
    static_assert(SY_G * SY_G == SY_THREADS, "one cell per thread");
    const int v = cell_fill[tid];
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp_tot[lane];
      int iw = w;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iw, o); if (lane >= o) iw += t; }
      s_warp_tot[lane] = iw - w;
    }
    __syncthreads();
    const int excl = s_warp_tot[warp] + inc - v;
    cell_start[tid] = excl;
    if (tid == SY_THREADS - 1) cell_start[SY_G * SY_G] = excl + v;
    cell_fill[tid] = excl;
  }
  __syncthreads();
  for (int i = tid; i < n; i += SY_THREADS) order[atomicAdd(&cell_fill[celly[i] * SY_G + cellx[i]], 1)] = (unsigned short)i;
  __syncthreads();
  // ---- exact duplicates (the shipped models contain some: class 17 stores 574 copies of one point):
  // only the lowest index of identical points takes part; the copies are hidden by definition and are
  // removed from the cell lists so they cost nothing as constraints.
  for (int i = tid; i < n; i += SY_THREADS) {
    const int cell = celly[i] * SY_G + cellx[i];
    const float x = f[i * 3 + 0], y = f[i * 3 + 1], z = f[i * 3 + 2];
    bool dup = false;
    for (int k = cell_start[cell]; k < cell_start[cell + 1] && !dup; ++k) {
      const int j = order[k];
      dup = j < i && f[j * 3 + 0] == x && f[j * 3 + 1] == y && f[j * 3 + 2] == z;
    }
    dirty[i] = dup ? 2 : 0;
  }
  __syncthreads();
  cell_fill[tid] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += SY_THREADS)
    if (dirty[i] != 2) atomicAdd(&cell_fill[celly[i] * SY_G + cellx[i]], 1);
  __syncthreads();
  {
    const int v = cell_fill[tid];
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp_tot[lane];
      int iw = w;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iw, o); if (lane >= o) iw += t; }
      s_warp_tot[lane] = iw - w;
    }
    __syncthreads();
    const int excl = s_warp_tot[warp] + inc - v;
    cell_start[tid] = excl;
    if (tid == SY_THREADS - 1) cell_start[SY_G * SY_G] = excl + v;
    cell_fill[tid] = excl;
  }
  __syncthreads();
  for (int i = tid; i < n; i += SY_THREADS)
    if (dirty[i] != 2) order[atomicAdd(&cell_fill[celly[i] * SY_G + cellx[i]], 1)] = (unsigned short)i;
  __syncthreads();
  const int n_unique = cell_start[SY_G * SY_G];

  HprShared h;
  h.U = U; h.V = V; h.W = W; h.order = order; h.cellx = cellx; h.celly = celly; h.cell_start = cell_start;
  h.kappa = rho; h.n = n; h.dup = dirty;

  // ---- phase 1: one WARP per point (dynamic queue, cell order): incremental LP over the point's 3x3
  // cell neighbourhood, where nearly all re-solves happen.  Survivors (~45 %) are queued with their optimum.
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&s_queue, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= n_unique) break;
    const int i = order[t];
    Nbhd nb;
    make_nbhd(h, i, s_nb[warp], nb);
    double sa = 0.0, sb = 0.0;
    if (hpr_lp_warp(h, nb, i, 0, sa, sb) && lane == 0) {
      const int slot = atomicAdd(&s_count, 1);
      surv[slot] = (unsigned short)i; SA[slot] = sa; SB[slot] = sb;
    }
  }
  __syncthreads();

  // ---- phase 2: verify every survivor's optimum against all points, as (survivor, 256-point slice)
  // work items.  Lanes of a warp share the slice (broadcast LDS.128 of the fp32 copy); a conservative
  // fp32 evaluation dismisses constraints that are clearly slack, the few within `tol` of tight are
  // re-evaluated in fp64; a genuine violation marks the survivor dirty for phase 3.
  const int nsurv = s_count;
  const int spad = (nsurv + 31) & ~31;
  const int nslice = (n + SY_SLICE - 1) / SY_SLICE;
  const float kh = 0.5f * (float)rho;
  for (int q = tid; q < spad * nslice; q += SY_THREADS) {
    const int sl = q / spad, sidx = q - sl * spad;
    if (sidx >= nsurv) continue;
    const int i = surv[sidx];
    if (dirty[i] == 1) continue;
    const float4 fi = F4[i];
    const double sa = SA[sidx], sb = SB[sidx];
    const float saf = (float)sa, sbf = (float)sb;
    const int cx = cellx[i], cy = celly[i];
    const int j1 = min(n, (sl + 1) * SY_SLICE);
    bool bad = false;
    for (int j = sl * SY_SLICE; j < j1; ++j) {
      const float4 fj = F4[j];
      const float duf = fj.x - fi.x, dvf = fj.y - fi.y;
      const float r2f = fmaf(duf, duf, dvf * dvf);
      const float rhsf = (fj.z - fi.z) - kh * r2f;
      const float dotf = fmaf(saf, duf, sbf * dvf);
      const float tol = 1e-3f + 2e-5f * (fabsf(rhsf) + fabsf(dotf) + kh * r2f);
      if (rhsf - dotf <= -tol) continue;                                  // clearly slack
      if (dirty[j] == 2) continue;                                        // copy of a lower-index point
      if (abs((int)cellx[j] - cx) <= 1 && abs((int)celly[j] - cy) <= 1) continue;  // handled in phase 1
      const double du = U[j] - U[i], dv = V[j] - V[i];
      const double r2 = du * du + dv * dv, dw = W[j] - W[i];
      if (r2 == 0.0) { if (dw > 0.0 || (dw == 0.0 && j < i)) bad = true; continue; }
      if ((dw - 0.5 * rho * r2) - (sa * du + sb * dv) > 0.0) bad = true;
    }
    if (bad) dirty[i] = 1;
  }
  __syncthreads();

  // ---- phase 3: the rare dirty survivors redo the LP over the full sequence (neighbourhood, then all
  // other points in index order), one warp each; clean survivors are visible.
  if (tid == 0) s_queue = 0;
  __syncthreads();
  while (true) {
    int t = 0;
    if (lane == 0) t = atomicAdd(&s_queue, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= nsurv) break;
    const int i = surv[t];
    if (dirty[i] != 1) { if (lane == 0) flag[i] = 1; continue; }
    Nbhd nb;
    make_nbhd(h, i, s_nb[warp], nb);
    double sa = 0.0, sb = 0.0;
    const bool vis = hpr_lp_warp(h, nb, i, n, sa, sb);
    if (lane == 0) flag[i] = vis ? 1 : 0;
  }
  __syncthreads();

  // ---- ordered compaction by ORIGINAL index (visible ids ascending)
  const int per = (n + SY_THREADS - 1) / SY_THREADS;
  const int i0 = tid * per, i1 = min(n, i0 + per);
  int cnt = 0;
  for (int i = i0; i < i1; ++i) cnt += flag[i];
  int inc = cnt;
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp_tot[warp] = inc;
  __syncthreads();  // every thread is past the LP phases: the U region may now be reused for ids[]
  if (warp == 0) {
    int v = s_warp_tot[lane], iv = v;
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
    s_warp_tot[lane] = iv - v;
    if (lane == 31) s_count = iv;
  }
  __syncthreads();
  int pos = s_warp_tot[warp] + inc - cnt;
  for (int i = i0; i < i1; ++i) {
    if (flags_out) flags_out[(size_t)cloud * n + i] = flag[i];
    if (flag[i]) ids[pos++] = i;
  }
  __syncthreads();
  // reference quirk: visibleId[:-1] drops the highest-index visible point (hidden_point_removal.py:36)
  const int nv = max(s_count - 1, 0);
  if (tid == 0) num_vis[cloud] = nv;
  const float* __restrict__ op = org + (size_t)cloud * org_stride_pts * 3;
  float* __restrict__ dst = out_pts + (size_t)cloud * take * 3;
  for (int r = tid; r < take; r += SY_THREADS) {
    int src;
    if (r < nv) src = ids[r];
    else if (nv > 0) {  // np.random.choice(visibleId, ...) padding (:38-40)
      const float uu = pad_uniform ? pad_uniform[(size_t)cloud * take + r] : ((float)((r - nv) % nv) + 0.5f) / (float)nv;
      src = ids[min((int)(uu * (float)nv), nv - 1)];
    } else src = -1;
    dst[r * 3 + 0] = src >= 0 ? op[src * 3 + 0] : 0.f;
    dst[r * 3 + 1] = src >= 0 ? op[src * 3 + 1] : 0.f;
    dst[r * 3 + 2] = src >= 0 ? op[src * 3 + 2] : 0.f;
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

extern "C" int caae_hpr_select(int b, int n, const float* flipped, const float* org, int org_stride_pts, int take,
                               const float* pad_uniform, float* out_pts, int* num_vis, unsigned char* flags_out,
                               caae_stream_t stream) {
  CAAE_RETURN_IF(b < 0 || n <= 0 || n > SY_MAXN || take <= 0 || org_stride_pts < n, CAAE_E_BADSHAPE);
  if (b == 0) return CAAE_OK;
  CAAE_RETURN_IF(!flipped || !org || !out_pts || !num_vis, CAAE_E_NULLPTR);
  size_t smem = (size_t)n * (5 * sizeof(double) + sizeof(float4) + 2 * sizeof(unsigned short) + 4) + 8 +
                sizeof(int) * (2 * SY_G * SY_G + 1) + 16;
  smem = (smem + 15) & ~(size_t)15;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(hpr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  hpr_select_kernel<<<b, SY_THREADS, smem, as_stream(stream)>>>(n, flipped, org, org_stride_pts, take, pad_uniform, out_pts, num_vis,
                                                                flags_out);
  return CAAE_LAUNCH_STATUS();
}

