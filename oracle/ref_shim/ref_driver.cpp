// ref_driver.cpp — C entry points onto the reference's own compiled code.  TEST INFRASTRUCTURE ONLY.
//
// Linked (by oracle/build_ref.sh) with objects compiled directly from
//   /root/reference/tf_ops/nn_distance/tf_nndistance.cpp      (through ref_shim/tensorflow/...)
//   /root/reference/tf_ops/nn_distance/tf_nndistance_g.cu     (nvcc, sm_100a)
//   /root/reference/tf_ops/sampling/tf_sampling_g.cu          (nvcc, sm_100a)
// * ref_cpu_*  run the reference's CPU OpKernels (NnDistanceOp / NnDistanceGradOp::Compute,
//   tf_nndistance.cpp:45-165) on host buffers through the mock OpKernelContext.
// * ref_gpu_*  call the reference's CUDA launchers on caller-provided DEVICE pointers, using the
//   prototypes the reference's own .cpp glue declares (tf_nndistance.cpp:168,208,
//   tf_sampling.cpp:65,94,125,150).  They launch on the legacy default stream, as the reference does.
#include <cstring>
#include <string>
#include "tensorflow/core/framework/op_kernel.h"

using namespace tensorflow;

void NmDistanceKernelLauncher(int b, int n, const float* xyz, int m, const float* xyz2, float* result,
                              int* result_i, float* result2, int* result2_i);
void NmDistanceGradKernelLauncher(int b, int n, const float* xyz1, int m, const float* xyz2,
                                  const float* grad_dist1, const int* idx1, const float* grad_dist2,
                                  const int* idx2, float* grad_xyz1, float* grad_xyz2);
void farthestpointsamplingLauncher(int b, int n, int m, const float* inp, float* temp, int* out);
void gatherpointLauncher(int b, int n, int m, const float* inp, const int* idx, float* out);
void scatteraddpointLauncher(int b, int n, int m, const float* out_g, const int* idx, float* inp_g);
void probsampleLauncher(int b, int n, int m, const float* inp_p, const float* inp_r, float* temp, int* out);

static std::string g_last_error;

static OpKernel* make_kernel(const char* op, const char* device) {
  auto it = KernelRegistry().find(std::make_pair(std::string(op), std::string(device)));
  if (it == KernelRegistry().end()) { g_last_error = std::string("no kernel registered: ") + op; return nullptr; }
  OpKernelConstruction c;
  return it->second(&c);
}

extern "C" {

const char* ref_last_error() { return g_last_error.c_str(); }

// Shapes are passed explicitly (rank + dims) so tests can also drive the reference's shape checks.
int ref_cpu_nn_distance(const float* xyz1, int r1, const long long* s1, const float* xyz2, int r2,
                        const long long* s2, float* dist1, int* idx1, float* dist2, int* idx2) {
  auto mk = [](int r, const long long* s) {
    if (r == 2) return TensorShape{s[0], s[1]};
    if (r == 3) return TensorShape{s[0], s[1], s[2]};
    if (r == 4) return TensorShape{s[0], s[1], s[2], s[3]};
    return TensorShape{s[0]};
  };
  Tensor t1(const_cast<float*>(xyz1), mk(r1, s1)), t2(const_cast<float*>(xyz2), mk(r2, s2));
  OpKernelContext ctx;
  ctx.inputs = {&t1, &t2};
  std::unique_ptr<OpKernel> k(make_kernel("NnDistance", DEVICE_CPU));
  if (!k) return 2;
  k->Compute(&ctx);
  if (!ctx.status.ok()) { g_last_error = ctx.status.error_message(); return 1; }
  long long bn = ctx.outputs[0]->shape().num_elements(), bm = ctx.outputs[2]->shape().num_elements();
  std::memcpy(dist1, ctx.outputs[0]->raw(), 4 * bn);
  std::memcpy(idx1, ctx.outputs[1]->raw(), 4 * bn);
  std::memcpy(dist2, ctx.outputs[2]->raw(), 4 * bm);
  std::memcpy(idx2, ctx.outputs[3]->raw(), 4 * bm);
  return 0;
}

int ref_cpu_nn_distance_grad(int b, int n, const float* xyz1, int m, const float* xyz2, const float* gd1,
                             const int* idx1, const float* gd2, const int* idx2, float* gx1, float* gx2) {
  Tensor t1(const_cast<float*>(xyz1), TensorShape{b, n, 3}), t2(const_cast<float*>(xyz2), TensorShape{b, m, 3});
  Tensor t3(const_cast<float*>(gd1), TensorShape{b, n}), t4(const_cast<int*>(idx1), TensorShape{b, n});
  Tensor t5(const_cast<float*>(gd2), TensorShape{b, m}), t6(const_cast<int*>(idx2), TensorShape{b, m});
  OpKernelContext ctx;
  ctx.inputs = {&t1, &t2, &t3, &t4, &t5, &t6};
  std::unique_ptr<OpKernel> k(make_kernel("NnDistanceGrad", DEVICE_CPU));
  if (!k) return 2;
  k->Compute(&ctx);
  if (!ctx.status.ok()) { g_last_error = ctx.status.error_message(); return 1; }
  std::memcpy(gx1, ctx.outputs[0]->raw(), sizeof(float) * (size_t)b * n * 3);
  std::memcpy(gx2, ctx.outputs[1]->raw(), sizeof(float) * (size_t)b * m * 3);
  return 0;
}

void ref_gpu_nn_distance(int b, int n, const float* xyz1, int m, const float* xyz2, float* d1, int* i1,
                         float* d2, int* i2) {
  NmDistanceKernelLauncher(b, n, xyz1, m, xyz2, d1, i1, d2, i2);
}
void ref_gpu_nn_distance_grad(int b, int n, const float* xyz1, int m, const float* xyz2, const float* gd1,
                              const int* idx1, const float* gd2, const int* idx2, float* gx1, float* gx2) {
  NmDistanceGradKernelLauncher(b, n, xyz1, m, xyz2, gd1, idx1, gd2, idx2, gx1, gx2);
}
// temp must hold 32*n floats (tf_sampling.cpp:115)
void ref_gpu_fps(int b, int n, int m, const float* inp, float* temp, int* out) {
  farthestpointsamplingLauncher(b, n, m, inp, temp, out);
}
void ref_gpu_gather(int b, int n, int m, const float* inp, const int* idx, float* out) {
  gatherpointLauncher(b, n, m, inp, idx, out);
}
// caller zero-fills inp_g first, as GatherPointGradGpuOp does (tf_sampling.cpp:174)
void ref_gpu_gather_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g) {
  scatteraddpointLauncher(b, n, m, out_g, idx, inp_g);
}
// temp must hold b*n floats (tf_sampling.cpp:84)
void ref_gpu_prob_sample(int b, int n, int m, const float* inp_p, const float* inp_r, float* temp, int* out) {
  probsampleLauncher(b, n, m, inp_p, inp_r, temp, out);
}

}  // extern "C"
