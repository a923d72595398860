// hpr_lp.cuh — the per-point LP of the hidden-point-removal kernel (synthesis.cu), written so that the
// same source compiles for the device and for the host (tests/hpr_host_harness.cpp runs it on the CPU
// against scipy's Qhull; that harness is test infrastructure, the product only uses the device build).
//
// Problem (see synthesis.cu for the derivation).  Lifted points (u, v, w) sit in shared memory sorted
// by grid cell.  Point i is visible iff
//        exists s in R^2 :  s . (u_j - u_i, v_j - v_i)  >=  (w_j - w_i) - kappa/2 |(u_j-u_i, v_j-v_i)|^2   for all j.
// Seidel's incremental LP with the minimum-norm objective: keep the optimum s of the constraints seen
// so far; a violated constraint moves s to the minimum-norm point of that constraint's boundary line
// clipped by every earlier constraint; an empty clip interval proves the point hidden.
//
// lp_lane: ONE THREAD runs the LP of one point.  "Scan the next constraint" and "clip against an
// earlier constraint" are two modes of ONE loop whose body loads one neighbour and forms its
// constraint, so the lanes of a warp — each at its own position, some scanning, some clipping — keep
// executing the same instructions.  The clip interval is kept as two fractions compared by
// cross-multiplication: no fp64 division per constraint, two per clip.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HPR_HD __host__ __device__ __forceinline__
#else
#define HPR_HD inline
#endif

namespace caae {
namespace hpr {

constexpr int G = 32;  // grid cells per axis over the (u, v) bounding box

// Cell-sorted view of one cloud: position p in [0, n_unique) -> lifted coordinates and original index.
struct View {
  const double* U;
  const double* V;
  const double* W;
  const unsigned short* id;
  const int* cell_start;  // [G*G + 1] positions of the first point of each cell (row-major cells)
  double kappa;
  int n_unique;
};

// Lift of a flipped point (x, y, z), z > 0, with reference radius rho: (x/z, y/z, -rho^2/z + rho/2 (u^2+v^2)).
HPR_HD void lift(double x, double y, double z, double rho, double& u, double& v, double& w) {
  u = x / z; v = y / z;
  w = -rho * rho / z + 0.5 * rho * (u * u + v * v);
}

// Neighbourhood half-width of a cell: 0 (the cell alone) when it is dense enough to hold a point's
// nearest neighbours by itself, else 1 (the 3x3 block).  Verification covers whatever lies outside.
#ifndef HPR_DENSE_CELL
#define HPR_DENSE_CELL 100000  /* measured on B200: shrinking dense cells to k = 0 costs more in verification than it saves */
#endif
HPR_HD int nbhd_halfwidth(const int* cell_start, int c) {
  return (cell_start[c + 1] - cell_start[c] >= HPR_DENSE_CELL) ? 0 : 1;
}

// The (2k+1)x(2k+1) cell neighbourhood of cell (cx, cy), k in {0, 1}, as three position ranges (cells of
// one grid row are consecutive in the sorted order): the point's own row first, then the row above, then
// the row below.
HPR_HD void nbhd_ranges(const int* cell_start, int cx, int cy, int k, int (&A)[3], int (&B)[3]) {
  const int x0 = cx - k > 0 ? cx - k : 0, x1 = cx + k < G - 1 ? cx + k : G - 1;
  A[0] = cell_start[cy * G + x0]; B[0] = cell_start[cy * G + x1 + 1];
  if (k > 0 && cy > 0) { A[1] = cell_start[(cy - 1) * G + x0]; B[1] = cell_start[(cy - 1) * G + x1 + 1]; }
  else { A[1] = 0; B[1] = 0; }
  if (k > 0 && cy < G - 1) { A[2] = cell_start[(cy + 1) * G + x0]; B[2] = cell_start[(cy + 1) * G + x1 + 1]; }
  else { A[2] = 0; B[2] = 0; }
}

// Everything outside the neighbourhood, as up to four more ranges (used by the full re-solve).
HPR_HD void full_ranges(const int (&A)[3], const int (&B)[3], int cy, int k, int n_unique, int (&FA)[7], int (&FB)[7]) {
  for (int r = 0; r < 3; ++r) { FA[r] = A[r]; FB[r] = B[r]; }
  const bool up = k > 0 && cy > 0, down = k > 0 && cy < G - 1;
  const int at = up ? A[1] : 0, bt = up ? B[1] : 0;
  const int ab = down ? A[2] : n_unique, bb = down ? B[2] : n_unique;
  FA[3] = 0; FB[3] = at;            // before the upper row's cells
  FA[4] = bt; FB[4] = A[0];         // between the upper row's cells and the own row's
  FA[5] = B[0]; FB[5] = ab;         // between the own row's and the lower row's
  FA[6] = bb; FB[6] = n_unique;     // after the lower row's cells
}

// Position sequences the LP runs over.  Ranges<NR>: the concatenation of NR position ranges.
template <int NR>
struct Ranges {
  int pre[NR + 1], off[NR];
  HPR_HD Ranges() { for (int r = 0; r < NR; ++r) { pre[r] = 0; off[r] = 0; } pre[NR] = 0; }
  HPR_HD Ranges(const int (&A)[NR], const int (&B)[NR]) {
    pre[0] = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) { pre[r + 1] = pre[r] + (B[r] - A[r]); off[r] = A[r] - pre[r]; }
  }
  HPR_HD int length() const { return pre[NR]; }
  HPR_HD int operator()(int cur) const {
    int o = off[0];
#pragma unroll
    for (int r = 1; r < NR; ++r) o = (cur >= pre[r]) ? off[r] : o;
    return cur + o;
  }
};

// The three neighbourhood ranges followed by `count` individually listed positions (constraints that
// the verification pass found violated).
struct RangesPlusList {
  Ranges<3> nb;
  const unsigned short* list;
  int count;
  HPR_HD RangesPlusList() : nb(), list(nullptr), count(0) {}
  HPR_HD RangesPlusList(const int (&A)[3], const int (&B)[3], const unsigned short* l, int c) : nb(A, B), list(l), count(c) {}
  HPR_HD int length() const { return nb.length() + count; }
  HPR_HD int operator()(int cur) const { return cur < nb.length() ? nb(cur) : (int)list[cur - nb.length()]; }
};

// Visits the positions of `Seq` in bit-reversed (van der Corput) order: every prefix of the visit
// order is spread evenly over the whole neighbourhood, which is what keeps Seidel's algorithm near
// its expected cost (a constraint late in the order is rarely violated, so the long clips are rare).
// The order runs over the next power of two; counters that map past the end yield -1 (skipped).
template <class Seq>
struct Scattered {
  Seq seq;
  int len, shift;
  HPR_HD explicit Scattered(const Seq& s) : seq(s) {
    len = s.length();
    int bits = 0;
    while ((1 << bits) < len) ++bits;
    shift = 32 - bits;
  }
  HPR_HD int length() const { return shift == 32 ? len : (int)(1u << (32 - shift)); }
  HPR_HD int operator()(int cur) const {
    if (shift == 32) return cur < len ? seq(cur) : -1;   // len <= 1
    unsigned v = (unsigned)cur;
#if defined(__CUDA_ARCH__)
    v = __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    v = (v >> 16) | (v << 16);
#endif
    const int idx = (int)(v >> shift);
    return idx < len ? seq(idx) : -1;
  }
};

enum LpStatus { kLpHidden = 0, kLpVisible = 1, kLpUnfinished = 2 };

// Incremental LP of the point at sorted position `self` over the position sequence `seq` (entries
// < 0 are skipped).  kLpHidden: proven hidden; kLpVisible: (sa, sb) is the minimum-norm feasible s;
// kLpUnfinished: `budget` loop iterations were not enough (budget <= 0: unlimited).
// `iters` (optional) counts loop iterations (= constraint evaluations).
template <class Seq>
HPR_HD int lp_lane(const View& h, int self, const Seq& seq, double& sa_out, double& sb_out, int budget = 0,
                   int* iters = nullptr, double s0a = 0.0, double s0b = 0.0, int* clips = nullptr) {
  const int L = seq.length();
  const double ui = h.U[self], vi = h.V[self], wi = h.W[self];
  const double hk = 0.5 * h.kappa;
  // objective: minimum |s - s0| (any objective decides feasibility; s0 = a neighbour's optimum is nearly feasible)
  double sa = s0a, sb = s0b;
  int nclips = 0;
  int p = 0;             // scan position in the sequence
  int q = -1, qend = 0;  // clip position / end (q < 0: scanning)
  double p0a = 0.0, p0b = 0.0, da = 0.0, db = 0.0;
  double ln = -1.0, ld = 0.0, hn = 1.0, hd = 0.0;  // clip interval [ln/ld, hn/hd]; ld == 0: -inf, hd == 0: +inf
  int it = 0, status = kLpVisible;
  while (true) {
    const bool clipping = q >= 0;
    if (!clipping && p >= L) break;
    if (budget > 0 && it >= budget) { status = kLpUnfinished; break; }
    const int j = seq(clipping ? q : p);
    ++it;
    if (clipping) ++q; else ++p;
    if (j >= 0 && j != self) {
      const double du = h.U[j] - ui, dv = h.V[j] - vi, dw = h.W[j] - wi;
      const double r2 = du * du + dv * dv;
      const double rhs = dw - hk * r2;
      if (!clipping) {
        if (r2 == 0.0) {  // same viewing direction: the nearer point (then the lower index) hides the other
          if (dw > 0.0 || (dw == 0.0 && h.id[j] < h.id[self])) { status = kLpHidden; break; }
        } else if (rhs - (sa * du + sb * dv) > 0.0) {
          // violated: the new optimum lies on this constraint's boundary line  p0 + t (da, db)
          const double inv = (rhs - (s0a * du + s0b * dv)) / r2;   // projection of s0 onto the line
          p0a = s0a + du * inv; p0b = s0b + dv * inv; da = -dv; db = du;
          ln = -1.0; ld = 0.0; hn = 1.0; hd = 0.0;
          q = 0; qend = p - 1; ++nclips;
        }
      } else if (r2 != 0.0) {
        const double den = da * du + db * dv, num = rhs - (p0a * du + p0b * dv);
        if (den > 0.0) { if (num * ld > ln * den) { ln = num; ld = den; } }
        else if (den < 0.0) { const double nn = -num, dd = -den; if (nn * hd < hn * dd) { hn = nn; hd = dd; } }
        else if (num > 0.0) { status = kLpHidden; break; }
      }
    }
    if (q >= 0 && q == qend) {  // the clip has seen every earlier constraint
      if (ld > 0.0 && hd > 0.0 && ln * hd > hn * ld) { status = kLpHidden; break; }
      double t = 0.0;
      if (ld > 0.0) t = fmax(t, ln / ld);
      if (hd > 0.0) t = fmin(t, hn / hd);
      sa = p0a + t * da; sb = p0b + t * db;
      q = -1;
    }
  }
  if (iters) *iters = it;
  if (clips) *clips = nclips;
  sa_out = sa; sb_out = sb;
  return status;
}

// Verification window.  Completing the square, constraint j of point i at optimum s reads
//   c_j = (w_j - w_i) - kappa/2 |a_j|^2 - s.a_j = (w_j - w_i) + |s|^2/(2 kappa) - kappa/2 |a_j + s/kappa|^2,
// so a violator (c_j > 0) lies within distance sqrt(2 (dW + |s|^2/(2 kappa)) / kappa) of the centre
// (u_i, v_i) - s/kappa, where dW bounds w_j - w_i over the candidates considered (a grid row).
// verify_disk2 returns that squared radius (fp32, rounded outwards; negative: no candidate can violate).
// fz = (float)(w + rho) of point i, fzmax = maximum of it over the candidates, kh = kappa/2.
HPR_HD float verify_disk2(float saf, float sbf, float fz, float fzmax, float kh) {
  const float s2 = (saf * saf + sbf * sbf) * 1.0002f;
  const float dw = (fzmax - fz) + 2e-6f * (fabsf(fzmax) + fabsf(fz)) + 1e-12f;
  const float ikh = 1.0f / kh;   // loop-invariant for the caller; the 1.0001 / 1.001 factors absorb the extra rounding
  return (dw + s2 * (0.25f * ikh) * 1.0001f) * ikh * 1.001f;
}

// Cell of sorted position p: the largest c with cell_start[c] <= p.
HPR_HD int find_cell(const int* cell_start, int p) {
  int lo = 0, hi = G * G;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cell_start[mid] <= p) lo = mid; else hi = mid;
  }
  return lo;
}

// Genuine (fp64) violation of constraint j by the optimum (sa, sb) of point i; > 0 means violated.
// r2 == 0 (same viewing direction) is reported through `same_dir`.
HPR_HD double violation(const View& h, int i, int j, double sa, double sb, bool& same_dir) {
  const double du = h.U[j] - h.U[i], dv = h.V[j] - h.V[i], r2 = du * du + dv * dv, dw = h.W[j] - h.W[i];
  same_dir = r2 == 0.0;
  return (dw - 0.5 * h.kappa * r2) - (sa * du + sb * dv);
}

// Key of the worst violator in one 32-bit word: the violation's float bits (positive, so they order
// like unsigned integers) with the low 12 mantissa bits replaced by the position.
constexpr int kPosBits = 12;
HPR_HD unsigned violation_key(double viol, int j) {
  const float fv = (float)viol;
  unsigned bits;
#if defined(__CUDA_ARCH__)
  bits = __float_as_uint(fv);
#else
  union { float f; unsigned u; } cv; cv.f = fv; bits = cv.u;
#endif
  bits &= ~((1u << kPosBits) - 1u);
  if (bits < (1u << kPosBits)) bits = 1u << kPosBits;  // a violation that underflows fp32 still yields a non-zero key
  return bits | (unsigned)j;
}

// Phase-2 filter: conservative fp32 evaluation of constraint j for the optimum (saf, sbf) of point i.
// fi / fj = (u, v, w + rho) rounded to fp32.  true = clearly slack (no fp64 re-evaluation needed).
HPR_HD bool clearly_slack(float fix, float fiy, float fiz, float fjx, float fjy, float fjz, float saf, float sbf,
                          float kh) {
  const float duf = fjx - fix, dvf = fjy - fiy;
  const float r2f = fmaf(duf, duf, dvf * dvf);
  const float rhsf = (fjz - fiz) - kh * r2f;
  const float dotf = fmaf(saf, duf, sbf * dvf);
  const float tol = 1e-3f + 2e-5f * (fabsf(rhsf) + fabsf(dotf) + kh * r2f);
  return rhsf - dotf <= -tol;
}

}  // namespace hpr
}  // namespace caae
