/* caae_oracle.c — CPU restatement of the reference's custom-op kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cloudaae_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU baseline.
 *
 * Parity status: the reference ships NO golden vectors for these ops
 * (SURVEY.md §4, §8c).  This restatement is instead pinned against the
 * reference itself: oracle/build_ref.sh compiles the reference's own sources
 * from /root/reference into oracle/_ref/ (CPU NnDistance ops through a small
 * TensorFlow header shim, and both .cu files rebuilt for sm_100a), and
 * tests/test_oracle_vs_ref.py + tests/test_gpu_ref_kernels.py require
 * bit-identical outputs.
 *
 * Two arithmetic modes exist in the reference and both are restated:
 *   mode 0 "gpu": d = fmaf(dz,dz, fmaf(dx,dx, dy*dy))   — what nvcc emits for
 *          tf_nndistance_g.cu:33 / tf_sampling_g.cu:141 (mul.f32, fma.rn, fma.rn
 *          in the shipped PTX and in nvcc 12.9's), canonical for index parity.
 *   mode 1 "cpu": d = (dx*dx + dy*dy) + dz*dz, every op rounded to float —
 *          what g++ -O2 emits for nnsearch (tf_nndistance.cpp:29-33).
 * Compile with -ffp-contract=off so the compiler cannot fuse mode 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define FPS_THREADS 512 /* blockDim.x of farthestpointsamplingLauncher, tf_sampling_g.cu:204 */
#define NND_TILE 512    /* `batch` in NmDistanceKernel, tf_nndistance_g.cu:6 */

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static inline float sqdist_gpu(float cx, float cy, float cz, float qx, float qy, float qz) {
    float dx = cx - qx, dy = cy - qy, dz = cz - qz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

static inline float sqdist_cpu(float cx, float cy, float cz, float qx, float qy, float qz) {
    float dx = cx - qx, dy = cy - qy, dz = cz - qz;
    float a = dx * dx, b = dy * dy, c = dz * dz;
    float s = a + b;
    return s + c;
}

/* ---- K1: farthestpointsamplingKernel, tf_sampling_g.cu:105-170 -------------------------------
 * Seed index 0 (:114-116); running min distance `temp` initialised to 1e38f (:118);
 * thread t owns k = t, t+512, ... ascending, strict `>` from best=-1, besti=0 (:125-149);
 * 512-slot tree keeps the LOWER slot on ties (:153-163).
 * Net tie rule: max d, then lowest (k mod 512), then lowest k.
 * min() is CUDA's fminf (returns the non-NaN operand). */
static void fps_one(int n, int m, const float *pts, int *out, float *temp, float *tb, int *tbi) {
    if (m <= 0) return;
    int old = 0;
    out[0] = 0;
    for (int k = 0; k < n; k++) temp[k] = 1e38f;
    for (int j = 1; j < m; j++) {
        float x1 = pts[old * 3 + 0], y1 = pts[old * 3 + 1], z1 = pts[old * 3 + 2];
        for (int t = 0; t < FPS_THREADS; t++) {
            float best = -1.0f;
            int besti = 0;
            for (int k = t; k < n; k += FPS_THREADS) {
                float td = temp[k];
                float d = sqdist_gpu(pts[k * 3 + 0], pts[k * 3 + 1], pts[k * 3 + 2], x1, y1, z1);
                float d2 = fminf(d, td);
                if (d2 != td) temp[k] = d2;
                if (d2 > best) { best = d2; besti = k; }
            }
            tb[t] = best;
            tbi[t] = besti;
        }
        /* the shared-memory tree, literally */
        for (int u = 0; (1 << u) < FPS_THREADS; u++) {
            for (int t = 0; t < (FPS_THREADS >> (u + 1)); t++) {
                int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                if (tb[i1] < tb[i2]) { tb[i1] = tb[i2]; tbi[i1] = tbi[i2]; }
            }
        }
        old = tbi[0];
        out[j] = old;
    }
}

void oracle_fps(int b, int n, int m, const float *inp, int *out, int threads) {
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
#endif
    {
        float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
        float tb[FPS_THREADS];
        int tbi[FPS_THREADS];
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int i = 0; i < b; i++)
            fps_one(n, m, inp + (size_t)i * n * 3, out + (size_t)i * m, temp, tb, tbi);
        free(temp);
    }
    (void)threads;
}

/* ---- K2: gatherpointKernel, tf_sampling_g.cu:172-181 ---------------------------------------- */
void oracle_gather(int b, int n, int m, const float *inp, const int *idx, float *out) {
    for (int i = 0; i < b; i++)
        for (int j = 0; j < m; j++) {
            int a = idx[i * m + j];
            for (int c = 0; c < 3; c++) out[((size_t)i * m + j) * 3 + c] = inp[((size_t)i * n + a) * 3 + c];
        }
}

/* ---- K3: GatherPointGradGpuOp zero-fill (tf_sampling.cpp:174) + scatteraddpointKernel
 * (tf_sampling_g.cu:183-192).  Serial j-ascending order; the GPU's atomic order is undefined, so
 * results agree bit-exactly only when idx has no duplicates inside a cloud (always true for FPS
 * output unless the tail repeats index 0). */
void oracle_gather_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    memset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);
    for (int i = 0; i < b; i++)
        for (int j = 0; j < m; j++) {
            int a = idx[i * m + j];
            for (int c = 0; c < 3; c++) inp_g[((size_t)i * n + a) * 3 + c] += out_g[((size_t)i * m + j) * 3 + c];
        }
}

/* ---- K4 / K5: NmDistanceKernel (tf_nndistance_g.cu:5-127) and nnsearch (tf_nndistance.cpp:21-43)
 * One direction: for every query j of cloud A, min + FIRST argmin over cloud B.
 *   mode 0: GPU tiling made explicit — strict `<` inside a 512-candidate tile with k==0
 *           initialising (:29-114), strict `>` across tiles so the earlier tile wins (:119).
 *   mode 1: nnsearch — float d widened to double, strict `<`, k==0 initialises. */
static void nn_one_dir(int n, int m, const float *a, const float *bq, float *dist, int *idx, int mode) {
    for (int j = 0; j < n; j++) {
        float x1 = a[j * 3 + 0], y1 = a[j * 3 + 1], z1 = a[j * 3 + 2];
        if (mode == 0) {
            float res = 0.0f;
            int res_i = 0;
            for (int k2 = 0; k2 < m; k2 += NND_TILE) {
                int end_k = (m < k2 + NND_TILE ? m : k2 + NND_TILE) - k2;
                float best = 0.0f;
                int best_i = 0;
                for (int k = 0; k < end_k; k++) {
                    const float *c = bq + (size_t)(k2 + k) * 3;
                    float d = sqdist_gpu(c[0], c[1], c[2], x1, y1, z1);
                    if (k == 0 || d < best) { best = d; best_i = k + k2; }
                }
                if (k2 == 0 || res > best) { res = best; res_i = best_i; }
            }
            if (m > 0) { dist[j] = res; idx[j] = res_i; }
        } else {
            double best = 0;
            int besti = 0;
            for (int k = 0; k < m; k++) {
                const float *c = bq + (size_t)k * 3;
                double d = (double)sqdist_cpu(c[0], c[1], c[2], x1, y1, z1);
                if (k == 0 || d < best) { best = d; besti = k; }
            }
            dist[j] = (float)best;
            idx[j] = besti;
        }
    }
}

void oracle_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2, float *dist1, int *idx1,
                        float *dist2, int *idx2, int mode, int threads) {
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3, *p2 = xyz2 + (size_t)i * m * 3;
        nn_one_dir(n, m, p1, p2, dist1 + (size_t)i * n, idx1 + (size_t)i * n, mode);
        nn_one_dir(m, n, p2, p1, dist2 + (size_t)i * m, idx2 + (size_t)i * m, mode);
    }
    (void)threads;
}

/* ---- K6: NnDistanceGradOp CPU loop (tf_nndistance.cpp:120-163); the GPU kernel
 * (tf_nndistance_g.cu:132-157) evaluates the same products (g = gd+gd; v = (p1-p2)*g; the cross
 * term is the exact negation) and differs only in the order the fp32 atomics land. */
void oracle_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1,
                             const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1,
                             float *grad_xyz2) {
    memset(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3);
    memset(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3);
    for (int i = 0; i < b; i++) {
        const float *p1 = xyz1 + (size_t)i * n * 3, *p2 = xyz2 + (size_t)i * m * 3;
        float *g1 = grad_xyz1 + (size_t)i * n * 3, *g2 = grad_xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; j++) {
            int j2 = idx1[(size_t)i * n + j];
            float g = grad_dist1[(size_t)i * n + j] * 2;
            for (int c = 0; c < 3; c++) {
                float v = g * (p1[j * 3 + c] - p2[j2 * 3 + c]);
                g1[j * 3 + c] += v;
                g2[j2 * 3 + c] -= v;
            }
        }
        for (int j = 0; j < m; j++) {
            int j2 = idx2[(size_t)i * m + j];
            float g = grad_dist2[(size_t)i * m + j] * 2;
            for (int c = 0; c < 3; c++) {
                float v = g * (p2[j * 3 + c] - p1[j2 * 3 + c]);
                g2[j * 3 + c] += v;
                g1[j2 * 3 + c] -= v;
            }
        }
    }
}

/* ---- prob_sample: cumsumKernel + binarysearchKernel (tf_sampling_g.cu:7-104) — "next" row §8(f).
 * The GPU cumsum is a blocked scan (4-element serial prefix, padded up/down-sweep over groups,
 * Kahan-style running sum across 8192-element chunks).  For n <= 8192 there is one chunk and
 * runningsum is 0, so the result equals this blocked evaluation order. */
static void cumsum_row(int n, const float *inp, float *out) {
    const int BS4 = 2048 * 4;
    float runningsum = 0, runningsum2 = 0;
    float *buffer4 = (float *)malloc(sizeof(float) * BS4);
    float *buffer = (float *)malloc(sizeof(float) * 2048);
    for (int j = 0; j < n; j += BS4) {
        int n24_i = (n - j < BS4) ? n - j : BS4;
        int n24 = (n24_i + 3) & ~3;
        int n2 = n24 >> 2;
        for (int k = 0; k < n24_i; k += 4) {
            if (k + 3 < n24_i) {
                float v1 = inp[j + k], v2 = inp[j + k + 1];
                v2 += v1;
                float v3 = inp[j + k + 2], v4 = inp[j + k + 3];
                v4 += v3; v3 += v2; v4 += v2;
                buffer4[k] = v1; buffer4[k + 1] = v2; buffer4[k + 2] = v3; buffer4[k + 3] = v4;
                buffer[k >> 2] = v4;
            } else {
                float v = 0;
                for (int k2 = k; k2 < n24_i; k2++) { v += inp[j + k2]; buffer4[k2] = v; }
                for (int k2 = n24_i; k2 < n24; k2++) buffer4[k2] = v;
                buffer[k >> 2] = v;
            }
        }
        int u = 0;
        for (; (2 << u) <= n2; u++)
            for (int k = 0; k < (n2 >> (u + 1)); k++) {
                int i1 = (((k << 1) + 2) << u) - 1, i2 = (((k << 1) + 1) << u) - 1;
                buffer[i1] += buffer[i2];
            }
        u--;
        for (; u >= 0; u--)
            for (int k = 0; k < ((n2 - (1 << u)) >> (u + 1)); k++) {
                int i1 = (((k << 1) + 3) << u) - 1, i2 = (((k << 1) + 2) << u) - 1;
                buffer[i1] += buffer[i2];
            }
        for (int k = 4; k < n24; k += 4) {
            float add = buffer[(k >> 2) - 1];
            buffer4[k] += add; buffer4[k + 1] += add; buffer4[k + 2] += add; buffer4[k + 3] += add;
        }
        for (int k = 0; k < n24_i; k++) out[j + k] = buffer4[k] + runningsum;
        float t = buffer[n2 - 1] + runningsum2;
        float r2 = runningsum + t;
        runningsum2 = t - (r2 - runningsum);
        runningsum = r2;
    }
    free(buffer4);
    free(buffer);
}

void oracle_cumsum(int b, int n, const float *inp, float *out) {
    for (int i = 0; i < b; i++) cumsum_row(n, inp + (size_t)i * n, out + (size_t)i * n);
}

void oracle_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out) {
    oracle_cumsum(b, n, inp_p, temp);
    int base = 1;
    while (base < n) base <<= 1;
    for (int i = 0; i < b; i++)
        for (int j = 0; j < m; j++) {
            const float *ds = temp + (size_t)i * n;
            float q = inp_r[(size_t)i * m + j] * ds[n - 1];
            int r = n - 1;
            for (int k = base; k >= 1; k >>= 1)
                if (r >= k && ds[r - k] >= q) r -= k;
            out[(size_t)i * m + j] = r;
        }
}
