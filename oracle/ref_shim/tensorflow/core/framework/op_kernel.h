// Minimal stand-in for the TensorFlow 1.x op-kernel headers — TEST INFRASTRUCTURE ONLY.
//
// Purpose: let oracle/build_ref.sh compile the reference's own, unmodified
// tf_ops/nn_distance/tf_nndistance.cpp (read in place from /root/reference) without TensorFlow,
// so that its CPU NnDistance / NnDistanceGrad OpKernels can be executed here and used to pin the
// oracle.  Only the API surface that file touches exists: Tensor/TensorShape views over caller
// memory, OpKernel{Construction,Context}, Status, errors::InvalidArgument, OP_REQUIRES[_OK],
// REGISTER_OP (ignored) and REGISTER_KERNEL_BUILDER (records a factory by op name + device).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace tensorflow {

typedef long long int64;

class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }
 private:
  bool ok_;
  std::string msg_;
};

namespace errors {
inline Status InvalidArgument(const std::string& msg) { return Status(msg); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<int64> d) : dims_(d) {}
  int dims() const { return (int)dims_.size(); }
  int64 dim_size(int i) const { return dims_[i]; }
  int64 num_elements() const { int64 n = 1; for (auto d : dims_) n *= d; return n; }
  bool operator==(const TensorShape& o) const { return dims_ == o.dims_; }
 private:
  std::vector<int64> dims_;
};

enum DataType { DT_FLOAT = 1, DT_INT32 = 3 };
template <class T> struct DataTypeToEnum;
template <> struct DataTypeToEnum<float> { static const DataType value = DT_FLOAT; };
template <> struct DataTypeToEnum<int> { static const DataType value = DT_INT32; };

template <class T> struct FlatView {
  T* p;
  T& operator()(int64 i) const { return p[i]; }
};

class Tensor {
 public:
  Tensor() : data_(nullptr), owned_(false) {}
  Tensor(void* data, const TensorShape& s) : data_(data), shape_(s), owned_(false) {}
  Tensor(const Tensor&) = delete;
  Tensor& operator=(const Tensor&) = delete;
  ~Tensor() { if (owned_) std::free(data_); }
  void Allocate(const TensorShape& s, size_t elem) {
    shape_ = s; data_ = std::calloc((size_t)(s.num_elements() > 0 ? s.num_elements() : 1), elem); owned_ = true;
  }
  int dims() const { return shape_.dims(); }
  const TensorShape& shape() const { return shape_; }
  template <class T> FlatView<T> flat() { return FlatView<T>{static_cast<T*>(data_)}; }
  template <class T> FlatView<const T> flat() const { return FlatView<const T>{static_cast<const T*>(data_)}; }
  void* raw() const { return data_; }
 private:
  void* data_;
  TensorShape shape_;
  bool owned_;
};

class OpKernelConstruction {
 public:
  std::map<std::string, int> int_attrs;
  Status GetAttr(const std::string& name, int* v) const {
    auto it = int_attrs.find(name);
    if (it == int_attrs.end()) return Status("missing attr " + name);
    *v = it->second; return Status::OK();
  }
  void CtxFailure(const Status& s) { status = s; }
  Status status;
};

class OpKernelContext {
 public:
  std::vector<Tensor*> inputs;                       // borrowed
  std::map<int, std::unique_ptr<Tensor>> outputs;    // owned
  std::vector<std::unique_ptr<Tensor>> temps;
  Status status;
  const Tensor& input(int i) const { return *inputs[i]; }
  Status allocate_output(int i, const TensorShape& s, Tensor** out) {
    std::unique_ptr<Tensor> t(new Tensor()); t->Allocate(s, 4); *out = t.get(); outputs[i] = std::move(t);
    return Status::OK();
  }
  Status allocate_temp(DataType, const TensorShape& s, Tensor* out) { out->Allocate(s, 4); return Status::OK(); }
  void CtxFailure(const Status& s) { status = s; }
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

static const char* const DEVICE_CPU = "CPU";
static const char* const DEVICE_GPU = "GPU";

struct KernelKey {
  std::string op, device;
  KernelKey& Device(const char* d) { device = d; return *this; }
};
inline KernelKey Name(const char* op) { KernelKey k; k.op = op; return k; }

typedef std::function<OpKernel*(OpKernelConstruction*)> KernelFactory;
inline std::map<std::pair<std::string, std::string>, KernelFactory>& KernelRegistry() {
  static std::map<std::pair<std::string, std::string>, KernelFactory> r;
  return r;
}
struct KernelRegistrar {
  KernelRegistrar(const KernelKey& k, KernelFactory f) { KernelRegistry()[std::make_pair(k.op, k.device)] = f; }
};

struct OpDefBuilderStub {
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Attr(const char*) { return *this; }
  template <class F> OpDefBuilderStub& SetShapeFn(F) { return *this; }
};

}  // namespace tensorflow

#define TF_SHIM_CAT2(a, b) a##b
#define TF_SHIM_CAT(a, b) TF_SHIM_CAT2(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderStub TF_SHIM_CAT(tf_shim_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub()
#define REGISTER_KERNEL_BUILDER(key, cls)                                                  \
  static ::tensorflow::KernelRegistrar TF_SHIM_CAT(tf_shim_kernel_, __COUNTER__)(          \
      key, [](::tensorflow::OpKernelConstruction* c) -> ::tensorflow::OpKernel* { return new cls(c); })
#define OP_REQUIRES(ctx, cond, status) \
  do { if (!(cond)) { (ctx)->CtxFailure(status); return; } } while (0)
#define OP_REQUIRES_OK(ctx, expr) \
  do { ::tensorflow::Status _s = (expr); if (!_s.ok()) { (ctx)->CtxFailure(_s); return; } } while (0)
