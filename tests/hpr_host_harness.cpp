// hpr_host_harness.cpp — TEST INFRASTRUCTURE.  Runs the hidden-point-removal LP of
// cloudaae_b200/csrc/hpr_lp.cuh (the very source the sm_100a kernel compiles) on the CPU, with the
// kernel's set-up (lift, 32x32 grid, cell sort by index, duplicate removal) restated sequentially, so
// the algorithm can be checked against scipy's Qhull without a GPU (tests/test_hpr_host.py) and its
// work (constraint evaluations per point) can be counted.
//
//   g++ -O2 -std=c++17 -shared -fPIC -I cloudaae_b200/csrc tests/hpr_host_harness.cpp -o <out>.so
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdlib>

#include "hpr_lp.cuh"

using namespace caae::hpr;

// flipped f32[b,n,3] -> flags u8[b,n] (1 = hull vertex / visible).
// stats i64[b,8]: {n_unique, phase1 iterations, survivors, dirty survivors, phase3 iterations,
//                  max re-solve rounds of one point, slow-path (fp64) evaluations in phase 2, full-LP fallbacks}
// iters_out (optional) i32[b,n]: phase-1 loop iterations of the point at each SORTED position (-1 past n_unique).
// ids_out (optional) i32[b,n]: original index of the point at each SORTED position (-1 past n_unique).
extern "C" int hpr_host_iters_ids(const float* flipped, int b, int n, unsigned char* flags, long long* stats, int* iters_out,
                                  int* ids_out);
extern "C" int hpr_host_iters(const float* flipped, int b, int n, unsigned char* flags, long long* stats, int* iters_out) {
  return hpr_host_iters_ids(flipped, b, n, flags, stats, iters_out, nullptr);
}
extern "C" int hpr_host(const float* flipped, int b, int n, unsigned char* flags, long long* stats) {
  return hpr_host_iters_ids(flipped, b, n, flags, stats, nullptr, nullptr);
}
extern "C" int hpr_host_iters_ids(const float* flipped, int b, int n, unsigned char* flags, long long* stats, int* iters_out,
                                  int* ids_out) {
  for (int cloud = 0; cloud < b; ++cloud) {
    const float* f = flipped + (size_t)cloud * n * 3;
    unsigned char* flag = flags + (size_t)cloud * n;
    long long* st = stats + (size_t)cloud * 8;
    std::memset(flag, 0, n);
    std::memset(st, 0, 8 * sizeof(long long));
    float nmax = 0.f;
    for (int i = 0; i < n; ++i) nmax = std::max(nmax, std::sqrt(f[i * 3] * f[i * 3] + f[i * 3 + 1] * f[i * 3 + 1] + f[i * 3 + 2] * f[i * 3 + 2]));
    const double rho = (double)nmax;
    std::vector<double> u(n), v(n), w(n);
    float umin = 3.4e38f, umax = -3.4e38f, vmin = 3.4e38f, vmax = -3.4e38f;
    for (int i = 0; i < n; ++i) {
      lift((double)f[i * 3], (double)f[i * 3 + 1], (double)f[i * 3 + 2], rho, u[i], v[i], w[i]);
      umin = std::min(umin, (float)u[i]); umax = std::max(umax, (float)u[i]);
      vmin = std::min(vmin, (float)v[i]); vmax = std::max(vmax, (float)v[i]);
    }
    const float bw = std::max(umax - umin, 1e-30f), bh = std::max(vmax - vmin, 1e-30f);
    std::vector<int> cell(n);
    std::vector<std::vector<int>> members(G * G);
    for (int i = 0; i < n; ++i) {
      int cx = (int)(((float)u[i] - umin) / bw * G), cy = (int)(((float)v[i] - vmin) / bh * G);
      cx = std::min(std::max(cx, 0), G - 1); cy = std::min(std::max(cy, 0), G - 1);
      cell[i] = cy * G + cx;
      members[cell[i]].push_back(i);  // ascending index
    }
    // (experiment, HPR_ORDER_W=1: inside a cell the points nearest to the viewer — largest lifted w — first, so that the
    //  constraints most likely to hide a point are met early; the kernel's order is ascending index)
    static const int order_w = std::getenv("HPR_ORDER_W") ? std::atoi(std::getenv("HPR_ORDER_W")) : 0;
    if (order_w)
      for (int c = 0; c < G * G; ++c)
        std::stable_sort(members[c].begin(), members[c].end(), [&](int a, int b2) { return order_w > 0 ? w[a] > w[b2] : w[a] < w[b2]; });
    // exact duplicates: only the lowest index of identical points takes part
    std::vector<int> cell_start(G * G + 1, 0);
    std::vector<unsigned short> id;
    std::vector<unsigned short> cell_of;
    for (int c = 0; c < G * G; ++c) {
      cell_start[c] = (int)id.size();
      for (size_t a = 0; a < members[c].size(); ++a) {
        const int i = members[c][a];
        bool dup = false;
        for (size_t k = 0; k < a && !dup; ++k) {
          const int j = members[c][k];
          dup = f[j * 3] == f[i * 3] && f[j * 3 + 1] == f[i * 3 + 1] && f[j * 3 + 2] == f[i * 3 + 2];
        }
        if (!dup) { id.push_back((unsigned short)i); cell_of.push_back((unsigned short)c); }
      }
    }
    const int nu = (int)id.size();
    cell_start[G * G] = nu;
    std::vector<double> U(nu), V(nu), W(nu);
    std::vector<float> F(4 * (size_t)nu);
    for (int p = 0; p < nu; ++p) {
      const int i = id[p];
      U[p] = u[i]; V[p] = v[i]; W[p] = w[i];
      F[4 * p] = (float)u[i]; F[4 * p + 1] = (float)v[i]; F[4 * p + 2] = (float)(w[i] + rho);
    }
    View h{U.data(), V.data(), W.data(), id.data(), cell_start.data(), rho, nu};
    st[0] = nu;
    if (iters_out) for (int p = 0; p < n; ++p) iters_out[(size_t)cloud * n + p] = -1;
    if (ids_out) for (int p = 0; p < n; ++p) ids_out[(size_t)cloud * n + p] = p < nu ? (int)id[p] : -1;
    const float kh = 0.5f * (float)rho;
    std::vector<float> fzrow(G, -3.4e38f);
    for (int p = 0; p < nu; ++p) fzrow[cell_of[p] / G] = std::max(fzrow[cell_of[p] / G], F[4 * p + 2]);
    long long outside_radius = 0;
    for (int p = 0; p < nu; ++p) {
      const int cx = cell_of[p] % G, cy = cell_of[p] / G;
      int A[3], B[3];
      const int kk = nbhd_halfwidth(cell_start.data(), cell_of[p]);
      nbhd_ranges(cell_start.data(), cx, cy, kk, A, B);
      double sa, sb;
      int it = 0;
      static const int scatter = std::getenv("HPR_SCATTER") ? std::atoi(std::getenv("HPR_SCATTER")) : 0;
      static const int warm = std::getenv("HPR_WARM") ? std::atoi(std::getenv("HPR_WARM")) : 0;
      static double wa = 0.0, wb = 0.0;
      if (p == 0 || !warm) { wa = 0.0; wb = 0.0; }
      int ncl = 0;
      const bool alive = scatter ? lp_lane(h, p, Scattered<Ranges<3>>(Ranges<3>(A, B)), sa, sb, 0, &it, wa, wb, &ncl) == kLpVisible
                                 : lp_lane(h, p, Ranges<3>(A, B), sa, sb, 0, &it, wa, wb, &ncl) == kLpVisible;
      if (alive) { wa = sa; wb = sb; }
      st[7] += ncl;  // (debug) clips in phase 1
      st[1] += it;
      if (iters_out) iters_out[(size_t)cloud * n + p] = it;
      if (!alive) continue;
      st[2] += 1;
      // verification / re-solve rounds, as the kernel runs them: check the optimum against every point
      // outside the neighbourhood (fp32 filter, fp64 where it is not clearly slack); the worst violator
      // joins the LP's constraint list and the LP is re-solved, until clean, hidden or the list is full.
      constexpr int kExtra = 8;
      unsigned short ext[kExtra];
      int nextra = 0, rounds = 0;
      bool vis = true;
      while (true) {
        const float saf = (float)sa, sbf = (float)sb;
        unsigned key = 0;
        for (int j = 0; j < nu && vis; ++j) {
          if (clearly_slack(F[4 * p], F[4 * p + 1], F[4 * p + 2], F[4 * j], F[4 * j + 1], F[4 * j + 2], saf, sbf, kh)) continue;
          if ((j >= A[0] && j < B[0]) || (j >= A[1] && j < B[1]) || (j >= A[2] && j < B[2])) continue;
          bool known = false;  // already a constraint of the LP: tight up to rounding, never re-added
          for (int e = 0; e < nextra; ++e) known = known || ext[e] == j;
          if (known) continue;
          st[6] += 1;
          bool same_dir;
          const double viol = violation(h, p, j, sa, sb, same_dir);
          if (same_dir) { if (W[j] > W[p] || (W[j] == W[p] && id[j] < id[p])) vis = false; continue; }
          if (viol > 0.0) {
            key = std::max(key, violation_key(viol, j));
            // the kernel only looks inside the verification radius: a violator outside it would be a bug
            // (bound checked with the per-row maximum, as the kernel applies it)
            const float row_max = fzrow[cell_of[j] / G];
            const double ddu = U[j] - (U[p] - (double)saf / (2.0 * kh)), ddv = V[j] - (V[p] - (double)sbf / (2.0 * kh));
            if (ddu * ddu + ddv * ddv > (double)verify_disk2(saf, sbf, F[4 * p + 2], row_max, kh)) outside_radius += 1;
          }
        }
        if (!vis || key == 0) break;
        if (rounds == 0) st[3] += 1;
        ++rounds;
        if (nextra == kExtra) {  // rare: the full LP (neighbourhood first, then everything else)
          int FA[7], FB[7];
          full_ranges(A, B, cy, kk, nu, FA, FB);
          vis = lp_lane(h, p, Ranges<7>(FA, FB), sa, sb, 0, &it) == kLpVisible;
          st[4] += it;
          break;
        }
        ext[nextra++] = (unsigned short)(key & ((1u << kPosBits) - 1u));
        vis = (scatter ? lp_lane(h, p, Scattered<RangesPlusList>(RangesPlusList(A, B, ext, nextra)), sa, sb, 0, &it)
                       : lp_lane(h, p, RangesPlusList(A, B, ext, nextra), sa, sb, 0, &it)) == kLpVisible;
        st[4] += it;
        if (!vis) break;
      }
      st[5] = std::max<long long>(st[5], rounds);
      if (vis) flag[id[p]] = 1;
    }
    if (outside_radius) return 100;  // verify_radius() bound violated
  }
  return 0;
}
