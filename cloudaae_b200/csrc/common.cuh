// common.cuh — shared helpers for the sm_100a kernels of libcloudaae_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cloudaae_b200.h"

#define CAAE_RETURN_IF(cond, code) \
  do { if (cond) return (code); } while (0)

// Launch epilogue: report the launch error (if any) as the positive cudaError_t.
#define CAAE_LAUNCH_STATUS() ((int)cudaPeekAtLastError())

namespace caae {

constexpr int kNumSMs = 148;  // B200

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }

// The reference's squared distance in the exact operation order nvcc emits for it
// (tf_nndistance_g.cu:30-33, tf_sampling_g.cu:141): mul.f32 y; fma.rn x; fma.rn z.
// Spelled with intrinsics so no compiler version can re-associate or re-contract it.
__device__ __forceinline__ float sqdist_ref(float cx, float cy, float cz, float qx, float qy, float qz) {
  const float dx = __fsub_rn(cx, qx), dy = __fsub_rn(cy, qy), dz = __fsub_rn(cz, qz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// fp32 -> tf32 (10 explicit mantissa bits), round to nearest even — what the TMA does to a TFLOAT32 tensor map
// on the way into shared memory (measured: tools/probe_tf32_rounding.py).  x - tf32_rne(x) is exact in fp32.
__device__ __forceinline__ float tf32_rne(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x0FFFu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xFFFFE000u);
}

inline cudaStream_t as_stream(caae_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch ------------------------------------------------------------------------------
// A train step is ~125 short kernels, ~70 of them on one dependent chain: what separates them is launch latency, not
// work.  Every kernel starts with pdl_wait() (griddepcontrol.wait: block until the preceding kernel of the stream has
// completed and flushed — a no-op for a normal launch) and is launched with the programmatic-stream-serialization
// attribute, so its CTAs are scheduled and run their prologue while the predecessor drains.  Because the wait is the
// FIRST statement, ordering semantics are exactly those of a normal launch.  CAAE_PDL=0 launches normally.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 900
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

bool pdl_enabled();   // capi.cu

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // (a launch error is picked up by CAAE_LAUNCH_STATUS)
}

}  // namespace caae
