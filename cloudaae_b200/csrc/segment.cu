// segment.cu — the real-segment front end of evaluation for sm_100a.
//
// Replaces, for a batch of (frame, class) segments at once, the CPU tf.data / py_func stages that
// precede the network in the reference's evaluate_cloudAAE_ycbv.py:
//   get_pointcloud                (:164-178)  depth -> camera-frame cloud, fp32, op by op
//   segment_not_empty             (:262-272)  label mask & valid depth
//   segment_mean_distance_filter  (:219-223)  drop points farther than 0.2 m from the segment mean
//   get_outlier_idx               (:250-258)  open3d remove_radius_outlier(nb_points=100, radius=0.02)
//   FPS_random                    (:230-247)  NumPy float64 farthest point sampling, random first index
//
// Design: every segment is independent, so (like FPS / HPR) the unit of parallelism is the segment.
//   * segment_extract_kernel: ONE CTA per segment walks the frame three times (L2-resident: 0.3 MB of
//     labels, 0.6 MB of depth): fp64 sums -> mean, then an ordered stream compaction of both point
//     lists (label-masked, distance-filtered) with one packed block scan per 4096-pixel chunk.
//   * radius_count_kernel: (query tile x segment) CTAs, candidates staged through shared memory; the
//     pair test runs in fp32 and only pairs within 1e-5 of r^2 are re-evaluated in fp64, so the flag
//     equals the fp64 count exactly; CTAs leave as soon as all their queries have enough neighbours.
//   * radius_compact_kernel: ordered compaction of the inlier flags + the reference's "< 512 inliers ->
//     keep everything" rule.
//   * fps_seeded_f64_kernel: one CTA per segment, float64 distances (NumPy promotes float32 - float64),
//     np.argmax semantics (first maximum), caller-chosen first index.
#include "common.cuh"

namespace caae {

constexpr int kSegThreads = 1024;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegPix = 4;                           // consecutive pixels per thread per chunk
constexpr int kSegChunk = kSegThreads * kSegPix;     // 4096 pixels

// get_pointcloud, one pixel, each TF op rounded to fp32 on its own (evaluate…:164-178).
__device__ __forceinline__ void pixel_to_xyz(int p, int w, unsigned short dep, float fx, float fy, float cx,
                                             float cy, float factor, float& x, float& y, float& z) {
  z = __fdiv_rn((float)dep, factor);
  const float X = (float)(p % w), Y = (float)(p / w);
  x = __fdiv_rn(__fmul_rn(__fsub_rn(X, cx), z), fx);
  y = __fdiv_rn(__fmul_rn(__fsub_rn(Y, cy), z), fy);
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kSegThreads)
segment_extract_kernel(int h, int w, const int* __restrict__ frame_of_seg, const int* __restrict__ class_of_seg,
                       const unsigned short* __restrict__ depth, const unsigned char* __restrict__ label,
                       const float* __restrict__ intrinsics, const float* __restrict__ threshold_per_class, int cap,
                       float* __restrict__ xyz_org, int* __restrict__ n_org, float* __restrict__ xyz_filt,
                       int* __restrict__ pix_filt, int* __restrict__ n_filt, float* __restrict__ seg_mean) {
  __shared__ double s_part[4][kSegWarps];
  __shared__ float s_mean[3];
  __shared__ unsigned int s_scan[2][kSegWarps + 1];

  const int seg = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int frame = frame_of_seg[seg], cls = class_of_seg[seg];
  const int npix = h * w;
  const unsigned short* __restrict__ dep = depth + (size_t)frame * npix;
  const unsigned char* __restrict__ lab = label + (size_t)frame * npix;
  const float fx = intrinsics[frame * 5 + 0], fy = intrinsics[frame * 5 + 1], cx = intrinsics[frame * 5 + 2],
              cy = intrinsics[frame * 5 + 3], factor = intrinsics[frame * 5 + 4];
  const float thr = threshold_per_class[cls];
  const int want = cls + 1;  // labels are one-based (evaluate…:263)

  // ---- pass 1: mean of the label-masked points (fp64 accumulation in a fixed order) ----
  double sx = 0.0, sy = 0.0, sz = 0.0, sc = 0.0;
  for (int p = t; p < npix; p += kSegThreads) {
    const unsigned short d = __ldg(dep + p);
    if ((int)__ldg(lab + p) == want && d != 0) {
      float x, y, z;
      pixel_to_xyz(p, w, d, fx, fy, cx, cy, factor, x, y, z);
      sx += (double)x; sy += (double)y; sz += (double)z; sc += 1.0;
    }
  }
  sx = warp_sum_f64(sx); sy = warp_sum_f64(sy); sz = warp_sum_f64(sz); sc = warp_sum_f64(sc);
  if (lane == 0) { s_part[0][warp] = sx; s_part[1][warp] = sy; s_part[2][warp] = sz; s_part[3][warp] = sc; }
  __syncthreads();
  if (warp == 0) {
    double a = warp_sum_f64(s_part[0][lane]), b = warp_sum_f64(s_part[1][lane]), c = warp_sum_f64(s_part[2][lane]),
           n = warp_sum_f64(s_part[3][lane]);
    if (lane == 0) {
      // empty segment: 0/0 = NaN like tf.reduce_mean of an empty tensor; every comparison below is false
      s_mean[0] = (float)(a / n); s_mean[1] = (float)(b / n); s_mean[2] = (float)(c / n);
      if (seg_mean != nullptr) {
        seg_mean[seg * 3 + 0] = s_mean[0]; seg_mean[seg * 3 + 1] = s_mean[1]; seg_mean[seg * 3 + 2] = s_mean[2];
      }
    }
  }
  __syncthreads();
  const float mx = s_mean[0], my = s_mean[1], mz = s_mean[2];

  // ---- pass 2: ordered compaction of both lists, one packed scan per chunk ----
  float* __restrict__ o_org = xyz_org != nullptr ? xyz_org + (size_t)seg * cap * 3 : nullptr;
  float* __restrict__ o_flt = xyz_filt + (size_t)seg * cap * 3;
  int* __restrict__ o_pix = pix_filt != nullptr ? pix_filt + (size_t)seg * cap : nullptr;
  unsigned int tot_org = 0, tot_flt = 0;
  int buf = 0;
  for (int base = 0; base < npix; base += kSegChunk) {
    const int p0 = base + t * kSegPix;
    float x[kSegPix], y[kSegPix], z[kSegPix];
    unsigned int m_org = 0, m_flt = 0;
#pragma unroll
    for (int i = 0; i < kSegPix; ++i) {
      const int p = p0 + i;
      if (p < npix) {
        const unsigned short d = __ldg(dep + p);
        if ((int)__ldg(lab + p) == want && d != 0) {
          pixel_to_xyz(p, w, d, fx, fy, cx, cy, factor, x[i], y[i], z[i]);
          m_org |= 1u << i;
          // tf.norm(xyz - mean): sqrt(reduce_sum(square(diff))), fp32 (evaluate…:222)
          const float dx = __fsub_rn(x[i], mx), dy = __fsub_rn(y[i], my), dz = __fsub_rn(z[i], mz);
          const float dist = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
          if (dist <= thr) m_flt |= 1u << i;
        }
      }
    }
    if (__syncthreads_or((int)m_org) == 0) continue;  // chunk without a segment pixel (uniform branch)
    const unsigned int mine = (unsigned int)__popc(m_org) | ((unsigned int)__popc(m_flt) << 16);
    unsigned int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int up = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += up;
    }
    if (lane == 31) s_scan[buf][warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const unsigned int v = s_scan[buf][lane];
      unsigned int winc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += up;
      }
      s_scan[buf][lane] = winc - v;  // exclusive offset of each warp
      if (lane == 31) s_scan[buf][kSegWarps] = winc;
    }
    __syncthreads();
    const unsigned int excl = s_scan[buf][warp] + inc - mine;
    const unsigned int total = s_scan[buf][kSegWarps];
    unsigned int pos_org = tot_org + (excl & 0xffffu), pos_flt = tot_flt + (excl >> 16);
#pragma unroll
    for (int i = 0; i < kSegPix; ++i) {
      if (m_org >> i & 1u) {
        if (o_org != nullptr && pos_org < (unsigned int)cap) {
          o_org[pos_org * 3 + 0] = x[i]; o_org[pos_org * 3 + 1] = y[i]; o_org[pos_org * 3 + 2] = z[i];
        }
        ++pos_org;
      }
      if (m_flt >> i & 1u) {
        if (pos_flt < (unsigned int)cap) {
          o_flt[pos_flt * 3 + 0] = x[i]; o_flt[pos_flt * 3 + 1] = y[i]; o_flt[pos_flt * 3 + 2] = z[i];
          if (o_pix != nullptr) o_pix[pos_flt] = p0 + i;
        }
        ++pos_flt;
      }
    }
    tot_org += total & 0xffffu;
    tot_flt += total >> 16;
    buf ^= 1;  // the next chunk scans in the other buffer: no third barrier needed
  }
  if (t == 0) {
    if (n_org != nullptr) n_org[seg] = (int)tot_org;
    n_filt[seg] = (int)tot_flt;
  }
}

// ---- radius outlier removal ---------------------------------------------------------------------
constexpr int kRadThreads = 256;
constexpr int kRadTile = 1024;

__global__ void __launch_bounds__(kRadThreads)
radius_count_kernel(int cap, const float* __restrict__ xyz, const int* __restrict__ n_pts, int nb_points, double r2,
                    unsigned char* __restrict__ flag) {
  __shared__ float s_x[kRadTile], s_y[kRadTile], s_z[kRadTile];
  const int seg = blockIdx.y, t = threadIdx.x;
  const int n = min(n_pts[seg], cap);
  const int q0 = blockIdx.x * kRadThreads;
  if (q0 >= n) return;
  const float* __restrict__ pts = xyz + (size_t)seg * cap * 3;
  const int q = q0 + t;
  const bool live = q < n;
  const float qx = live ? pts[q * 3 + 0] : 0.f, qy = live ? pts[q * 3 + 1] : 0.f, qz = live ? pts[q * 3 + 2] : 0.f;
  const float lo = (float)(r2 * (1.0 - 1e-5)), hi = (float)(r2 * (1.0 + 1e-5));
  int cnt = 0;
  for (int base = 0; base < n; base += kRadTile) {
    const int len = min(kRadTile, n - base);
    __syncthreads();
    for (int i = t; i < len; i += kRadThreads) {
      s_x[i] = pts[(base + i) * 3 + 0]; s_y[i] = pts[(base + i) * 3 + 1]; s_z[i] = pts[(base + i) * 3 + 2];
    }
    __syncthreads();
    if (live && cnt <= nb_points) {
      for (int j = 0; j < len; ++j) {
        const float dx = s_x[j] - qx, dy = s_y[j] - qy, dz = s_z[j] - qz;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d < lo) {
          ++cnt;
        } else if (d <= hi) {  // within 1e-5 of the radius: decide in fp64 like the KD-tree does
          const double ex = (double)s_x[j] - (double)qx, ey = (double)s_y[j] - (double)qy,
                       ez = (double)s_z[j] - (double)qz;
          const double e = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
          if (e < r2) ++cnt;
        }
      }
    }
    if (__syncthreads_and(!live || cnt > nb_points)) break;  // every query of this CTA is already an inlier
  }
  if (live) flag[(size_t)seg * cap + q] = cnt > nb_points ? 1 : 0;
}

__global__ void __launch_bounds__(kSegThreads)
radius_compact_kernel(int cap, const int* __restrict__ n_pts, const unsigned char* __restrict__ flag, int min_keep,
                      int* __restrict__ inlier_idx, int* __restrict__ n_inlier) {
  __shared__ unsigned int s_scan[2][kSegWarps + 1];
  const int seg = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = min(n_pts[seg], cap);
  const unsigned char* __restrict__ f = flag + (size_t)seg * cap;
  int* __restrict__ out = inlier_idx + (size_t)seg * cap;
  unsigned int total = 0;
  int buf = 0;
  for (int base = 0; base < n; base += kSegThreads) {
    const int i = base + t;
    const bool in = i < n && f[i] != 0;
    const unsigned int bal = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_scan[buf][warp] = (unsigned int)__popc(bal);
    __syncthreads();
    if (warp == 0) {
      const unsigned int v = s_scan[buf][lane];
      unsigned int winc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += up;
      }
      s_scan[buf][lane] = winc - v;
      if (lane == 31) s_scan[buf][kSegWarps] = winc;
    }
    __syncthreads();
    if (in) out[total + s_scan[buf][warp] + __popc(bal & ((1u << lane) - 1u))] = i;
    total += s_scan[buf][kSegWarps];
    buf ^= 1;
  }
  __syncthreads();
  // `if len(idx) < 512: idx = np.arange(xyz.shape[0])` (evaluate…:256-257)
  const bool keep_all = (int)total < min_keep;
  const int kept = keep_all ? n : (int)total;
  if (keep_all)
    for (int i = t; i < n; i += kSegThreads) out[i] = i;
  for (int i = kept + t; i < cap; i += kSegThreads) out[i] = 0;  // tail: a valid index, so gathers stay in range
  if (t == 0) n_inlier[seg] = kept;
}

// ---- FPS_random: float64, first index given, np.argmax tie rule ---------------------------------
__device__ __forceinline__ void argmax_pair(double& v, int& i, double ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__global__ void __launch_bounds__(kSegThreads)
fps_seeded_f64_kernel(int cap, int k, const float* __restrict__ xyz, const int* __restrict__ n_pts,
                      const int* __restrict__ first_idx, double* __restrict__ temp, int* __restrict__ out_idx,
                      float* __restrict__ out_xyz) {
  __shared__ double s_v[2][kSegWarps];
  __shared__ int s_i[2][kSegWarps];
  const int seg = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int n = min(n_pts[seg], cap);
  const float* __restrict__ pts = xyz + (size_t)seg * cap * 3;
  double* __restrict__ td = temp + (size_t)seg * cap;
  int* __restrict__ o = out_idx + (size_t)seg * k;
  float* __restrict__ ox = out_xyz != nullptr ? out_xyz + (size_t)seg * k * 3 : nullptr;
  if (n <= 0) {  // the reference would raise (randint(0, -1)); defined here: all-zero output
    for (int i = t; i < k; i += kSegThreads) {
      o[i] = 0;
      if (ox != nullptr) { ox[i * 3 + 0] = 0.f; ox[i * 3 + 1] = 0.f; ox[i * 3 + 2] = 0.f; }
    }
    return;
  }
  int cur = min(max(first_idx[seg], 0), n - 1);
  for (int i = t; i < n; i += kSegThreads) td[i] = __longlong_as_double(0x7ff0000000000000LL);  // +inf
  for (int r = 0; r < k; ++r) {
    const float cxf = pts[cur * 3 + 0], cyf = pts[cur * 3 + 1], czf = pts[cur * 3 + 2];
    if (t == 0) {
      o[r] = cur;
      if (ox != nullptr) { ox[r * 3 + 0] = cxf; ox[r * 3 + 1] = cyf; ox[r * 3 + 2] = czf; }
    }
    if (r + 1 == k) break;
    const double x1 = (double)cxf, y1 = (double)cyf, z1 = (double)czf;
    double best = -1.0;
    int besti = 0x7fffffff;
    for (int i = t; i < n; i += kSegThreads) {
      // ((p0 - points)**2).sum(axis=1): float64, square then left-to-right add, no contraction
      const double dx = x1 - (double)pts[i * 3 + 0], dy = y1 - (double)pts[i * 3 + 1], dz = z1 - (double)pts[i * 3 + 2];
      const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      const double m = fmin(td[i], d);
      td[i] = m;
      if (m > best) { best = m; besti = i; }  // ascending i per thread: strict > keeps the first maximum
    }
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, of);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, of);
      argmax_pair(best, besti, ov, oi);
    }
    const int b = r & 1;
    if (lane == 0) { s_v[b][warp] = best; s_i[b][warp] = besti; }
    __syncthreads();
    best = s_v[b][lane];
    besti = s_i[b][lane];
#pragma unroll
    for (int of = 16; of > 0; of >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best, of);
      const int oi = __shfl_xor_sync(0xffffffffu, besti, of);
      argmax_pair(best, besti, ov, oi);
    }
    cur = besti;
  }
}

}  // namespace caae

using namespace caae;

extern "C" int caae_segment_extract(int nseg, int nframes, int h, int w, const int* frame_of_seg,
                                    const int* class_of_seg, const unsigned short* depth, const unsigned char* label,
                                    const float* intrinsics, const float* threshold_per_class, int cap,
                                    float* xyz_org, int* n_org, float* xyz_filt, int* pix_filt, int* n_filt,
                                    float* seg_mean, caae_stream_t stream) {
  CAAE_RETURN_IF(nseg < 0 || nframes < 0 || h < 0 || w < 0 || cap < 0, CAAE_E_BADSHAPE);
  CAAE_RETURN_IF((long long)h * w > (1ll << 30) || cap > 65535 * 16, CAAE_E_BADSHAPE);
  if (nseg == 0) return CAAE_OK;
  CAAE_RETURN_IF(frame_of_seg == nullptr || class_of_seg == nullptr || intrinsics == nullptr ||
                     threshold_per_class == nullptr || n_filt == nullptr,
                 CAAE_E_NULLPTR);
  CAAE_RETURN_IF(h * w > 0 && (depth == nullptr || label == nullptr), CAAE_E_NULLPTR);
  CAAE_RETURN_IF(cap > 0 && xyz_filt == nullptr, CAAE_E_NULLPTR);
  segment_extract_kernel<<<nseg, kSegThreads, 0, as_stream(stream)>>>(
      h, w, frame_of_seg, class_of_seg, depth, label, intrinsics, threshold_per_class, cap, xyz_org, n_org, xyz_filt,
      pix_filt, n_filt, seg_mean);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_radius_outlier(int nseg, int cap, const float* xyz, const int* n_pts, int nb_points,
                                   double radius, int min_keep, unsigned char* flag, int* inlier_idx, int* n_inlier,
                                   caae_stream_t stream) {
  CAAE_RETURN_IF(nseg < 0 || cap < 0 || nb_points < 0 || !(radius >= 0.0), CAAE_E_BADSHAPE);
  CAAE_RETURN_IF(nseg > 65535, CAAE_E_BADSHAPE);
  if (nseg == 0) return CAAE_OK;
  CAAE_RETURN_IF(n_pts == nullptr || n_inlier == nullptr, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(cap > 0 && (xyz == nullptr || inlier_idx == nullptr), CAAE_E_NULLPTR);
  CAAE_RETURN_IF(cap > 0 && flag == nullptr, CAAE_E_SCRATCH);
  if (cap > 0) {
    dim3 grid((cap + kRadThreads - 1) / kRadThreads, nseg);
    radius_count_kernel<<<grid, kRadThreads, 0, as_stream(stream)>>>(cap, xyz, n_pts, nb_points, radius * radius, flag);
    int st = CAAE_LAUNCH_STATUS();
    if (st != 0) return st;
  }
  radius_compact_kernel<<<nseg, kSegThreads, 0, as_stream(stream)>>>(cap, n_pts, flag, min_keep, inlier_idx, n_inlier);
  return CAAE_LAUNCH_STATUS();
}

extern "C" int caae_fps_seeded_f64(int nseg, int cap, int k, const float* xyz, const int* n_pts, const int* first_idx,
                                   double* temp, int* out_idx, float* out_xyz, caae_stream_t stream) {
  CAAE_RETURN_IF(nseg < 0 || cap < 0 || k < 0, CAAE_E_BADSHAPE);
  if (nseg == 0 || k == 0) return CAAE_OK;
  CAAE_RETURN_IF(n_pts == nullptr || first_idx == nullptr || out_idx == nullptr, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(cap > 0 && xyz == nullptr, CAAE_E_NULLPTR);
  CAAE_RETURN_IF(cap > 0 && temp == nullptr, CAAE_E_SCRATCH);
  fps_seeded_f64_kernel<<<nseg, kSegThreads, 0, as_stream(stream)>>>(cap, k, xyz, n_pts, first_idx, temp, out_idx,
                                                                     out_xyz);
  return CAAE_LAUNCH_STATUS();
}
