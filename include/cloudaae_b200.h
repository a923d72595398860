/* cloudaae_b200.h — C ABI of libcloudaae_b200.so (hand-written sm_100a kernels).
 *
 * Every entry point replaces one launcher of the reference's tf_ops and keeps its parameter order,
 * adding a trailing CUDA stream and an int status return (0 = ok, >0 = cudaError_t of the launch,
 * <0 = CAAE_E_* argument error).  All pointers are DEVICE pointers unless the name ends in _host.
 * No global state, no allocation, no host synchronisation: every call is CUDA-graph capturable.
 * The library zero-fills what it accumulates into.  Plain C types only (no torch, no C++).
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   caae_fps               farthestpointsamplingLauncher  tf_ops/sampling/tf_sampling.cpp:94,  tf_sampling_g.cu:203-205
 *   caae_gather            gatherpointLauncher            tf_ops/sampling/tf_sampling.cpp:125, tf_sampling_g.cu:206-208
 *   caae_gather_grad       scatteraddpointLauncher (+ the cudaMemset of GatherPointGradGpuOp)
 *                                                          tf_ops/sampling/tf_sampling.cpp:150,174, tf_sampling_g.cu:209-211
 *   caae_prob_sample       probsampleLauncher             tf_ops/sampling/tf_sampling.cpp:65,  tf_sampling_g.cu:198-201
 *   caae_nn_distance       NmDistanceKernelLauncher       tf_ops/nn_distance/tf_nndistance.cpp:168, tf_nndistance_g.cu:128-131
 *   caae_nn_distance_grad  NmDistanceGradKernelLauncher   tf_ops/nn_distance/tf_nndistance.cpp:208, tf_nndistance_g.cu:152-157
 */
#ifndef CLOUDAAE_B200_H_
#define CLOUDAAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAAE_ABI_VERSION 1

/* argument errors (negative); positive return values are cudaError_t */
#define CAAE_OK 0
#define CAAE_E_BADSHAPE (-1)   /* negative size, or a size the kernel cannot address */
#define CAAE_E_NULLPTR (-2)    /* a required pointer is NULL while the tensor is non-empty */
#define CAAE_E_SCRATCH (-3)    /* scratch required for this size but not provided */
#define CAAE_E_UNSUPPORTED (-4)

typedef void* caae_stream_t; /* cudaStream_t */

int caae_abi_version(void);
/* Human-readable text for a status returned by any entry point (static storage). */
const char* caae_status_string(int status);

/* ---- sampling ops -------------------------------------------------------------------------- */

/* Bytes of scratch caae_fps needs for clouds of n points (0 when n fits the register-resident
 * kernel; otherwise b*n floats, the role of the reference's temp[32,n]). */
size_t caae_fps_scratch_bytes(int b, int n);

/* Farthest point sampling. inp f32[b,n,3] -> out i32[b,m].  Seed index 0; ties resolved exactly as
 * the reference kernel does (max distance, then lowest (k mod 512), then lowest k).
 * temp may be NULL when caae_fps_scratch_bytes(b,n)==0. */
int caae_fps(int b, int n, int m, const float* inp, float* temp, int* out, caae_stream_t stream);

/* Same, and additionally writes the sampled coordinates out_xyz f32[b,m,3] (fuses gather_point). */
int caae_fps_gather(int b, int n, int m, const float* inp, float* temp, int* out, float* out_xyz,
                    caae_stream_t stream);

/* out[b,j,:] = inp[b,idx[b,j],:].  inp f32[b,n,3], idx i32[b,m] -> out f32[b,m,3]. */
int caae_gather(int b, int n, int m, const float* inp, const int* idx, float* out, caae_stream_t stream);

/* inp_g[b,idx[b,j],:] += out_g[b,j,:] after zero-filling inp_g f32[b,n,3]. */
int caae_gather_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g,
                     caae_stream_t stream);

/* inp_p f32[b,n] (weights), inp_r f32[b,m] (uniform draws), temp f32[b,n] -> out i32[b,m]. */
int caae_prob_sample(int b, int n, int m, const float* inp_p, const float* inp_r, float* temp, int* out,
                     caae_stream_t stream);

/* ---- chamfer nearest-neighbour distance ------------------------------------------------------ */

/* xyz f32[b,n,3], xyz2 f32[b,m,3] -> result f32[b,n], result_i i32[b,n], result2 f32[b,m],
 * result2_i i32[b,m].  Squared distances, d = fma(dz,dz, fma(dx,dx, dy*dy)); FIRST (lowest-index)
 * argmin.  An empty opposite cloud yields dist 0 / idx 0 (the reference CPU op's behaviour). */
int caae_nn_distance(int b, int n, const float* xyz, int m, const float* xyz2, float* result, int* result_i,
                     float* result2, int* result2_i, caae_stream_t stream);

/* grad_xyz1 f32[b,n,3], grad_xyz2 f32[b,m,3] (zero-filled by the library). */
int caae_nn_distance_grad(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                          const int* idx1, const float* grad_dist2, const int* idx2, float* grad_xyz1,
                          float* grad_xyz2, caae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDAAE_B200_H_ */
