// Minimal stand-in for the slice of the Caffe2 (pytorch v1.3.0) operator API that the
// reference's CPU operators on the hot path use.  TEST INFRASTRUCTURE ONLY: it exists so
// that oracle/build_ref.sh can compile the UNMODIFIED reference sources
//   detectron/ops/{roi_feature_boost_op,cross_entropy_wsl_op,acm_weightdecay_momentum_sgd_op}.cc
// where they lie under /root/reference and run their RunOnDevice() bodies on host buffers.
// Nothing here restates reference arithmetic; it only provides containers and registration.
#pragma once
#include <cstdint>
#include <cstdio>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <algorithm>
#include <cmath>

namespace caffe2 {
using std::string;
using std::vector;

enum DeviceType { CPU = 0 };
namespace at_shim { struct dtype_tag {}; }

class Tensor {
 public:
  int dim() const { return (int)sizes_.size(); }
  int dim32(int i) const { return (int)sizes_.at(i); }
  int64_t numel() const { int64_t n = 1; for (auto s : sizes_) n *= s; return n; }
  int64_t size_from_dim(int k) const { int64_t n = 1; for (size_t i = k; i < sizes_.size(); ++i) n *= sizes_[i]; return n; }
  const vector<int64_t>& sizes() const { return sizes_; }
  void Resize(const vector<int64_t>& s) { sizes_ = s; buf_.resize((size_t)numel() * 8); }
  template <typename... Ts> void Resize(Ts... ds) { Resize(vector<int64_t>{(int64_t)ds...}); }
  void ResizeLike(const Tensor& o) { if (&o != this) Resize(o.sizes_); }
  template <typename T> const T* data() const { return reinterpret_cast<const T*>(buf_.data()); }
  template <typename T> T* mutable_data() { return reinterpret_cast<T*>(buf_.data()); }
  void set(const vector<int64_t>& s, const void* src, size_t bytes) {
    Resize(s); std::copy((const char*)src, (const char*)src + bytes, buf_.begin());
  }
 private:
  vector<int64_t> sizes_;
  vector<char> buf_;   // 8 bytes per element: wide enough for any scalar type used
};

struct Argument { string name; double f = 0; };
struct OperatorDef {
  string type_;
  std::map<string, double> args;
  const string& type() const { return type_; }
};
struct TensorShape {};
class Workspace {};
template <typename T> Argument MakeArgument(const string& n, T v) { Argument a; a.name = n; a.f = (double)v; return a; }
struct ArgumentHelper {
  explicit ArgumentHelper(const OperatorDef& d) : d_(d) {}
  template <typename T> T GetSingleArgument(const string& n, T dflt) const {
    auto it = d_.args.find(n); return it == d_.args.end() ? dflt : (T)it->second; }
  const OperatorDef& d_;
};

class CPUContext {
 public:
  static constexpr DeviceType GetDeviceType() { return CPU; }
};

class OperatorBase {
 public:
  OperatorBase(const OperatorDef& def, Workspace*) : def_(def) {}
  virtual ~OperatorBase() {}
  virtual bool RunOnDevice() = 0;
  template <typename T> T GetSingleArgument(const string& n, T dflt) const {
    auto it = def_.args.find(n); return it == def_.args.end() ? dflt : (T)it->second; }
  bool InputIsTensorType(int, DeviceType) const { return true; }
  int InputSize() const { return (int)inputs_.size(); }
  int OutputSize() const { return (int)outputs_.size(); }
  // harness side
  vector<Tensor*> inputs_;
  vector<Tensor*> outputs_;
 protected:
  OperatorDef def_;
};

template <class Context>
class Operator : public OperatorBase {
 public:
  Operator(const OperatorDef& def, Workspace* ws) : OperatorBase(def, ws) {}
  const Tensor& Input(int i) { return *inputs_.at(i); }
  Tensor* Output(int i) { return outputs_.at(i); }
  template <typename... A> Tensor* Output(int i, const vector<int64_t>& s, A...) { outputs_.at(i)->Resize(s); return outputs_.at(i); }
 protected:
  Context context_;
};

#define USE_OPERATOR_CONTEXT_FUNCTIONS                      \
  using Operator<Context>::context_;                        \
  using Operator<Context>::Input;                           \
  using Operator<Context>::Output;                          \
  using OperatorBase::InputSize;                            \
  using OperatorBase::OutputSize
#define INPUT_TAGS(...) enum _InputTags { __VA_ARGS__ }
#define OUTPUT_TAGS(...) enum _OutputTags { __VA_ARGS__ }

// ---- registry -------------------------------------------------------------------
using OpFactory = std::function<OperatorBase*(const OperatorDef&, Workspace*)>;
inline std::map<string, OpFactory>& CPUOperatorRegistry() { static std::map<string, OpFactory> r; return r; }
struct OpRegistrar { OpRegistrar(const char* n, OpFactory f) { CPUOperatorRegistry()[n] = std::move(f); } };
#define C2SHIM_CAT_(a, b) a##b
#define C2SHIM_CAT(a, b) C2SHIM_CAT_(a, b)
#define REGISTER_CPU_OPERATOR(name, ...)                                                   \
  static ::caffe2::OpRegistrar C2SHIM_CAT(c2shim_reg_##name##_, __COUNTER__)(              \
      #name, [](const ::caffe2::OperatorDef& d, ::caffe2::Workspace* w) -> ::caffe2::OperatorBase* { \
        return new __VA_ARGS__(d, w); })

// ---- schema: every builder call is accepted and ignored -----------------------------
struct OpSchema {
  template <typename... A> OpSchema& NumInputs(A...) { return *this; }
  template <typename... A> OpSchema& NumOutputs(A...) { return *this; }
  OpSchema& AllowInplace(std::initializer_list<std::pair<int, int>>) { return *this; }
  template <typename... A> OpSchema& IdenticalTypeAndShapeOfInputDim(A...) { return *this; }
  template <typename... A> OpSchema& IdenticalTypeAndShapeOfInput(A...) { return *this; }
  template <typename... A> OpSchema& IdenticalTypeAndShape(A...) { return *this; }
  template <typename... A> OpSchema& SetDoc(A...) { return *this; }
  template <typename... A> OpSchema& Input(A...) { return *this; }
  template <typename... A> OpSchema& Output(A...) { return *this; }
  template <typename... A> OpSchema& Arg(A...) { return *this; }
  template <typename F> OpSchema& TensorInferenceFunction(F) { return *this; }
};
#define OPERATOR_SCHEMA(name) static ::caffe2::OpSchema C2SHIM_CAT(c2shim_schema_##name##_, __COUNTER__) = ::caffe2::OpSchema()

// ---- gradient makers: compiled, never run ---------------------------------------------
class GradientMakerBase {
 public:
  GradientMakerBase() {}
  virtual ~GradientMakerBase() {}
  virtual vector<OperatorDef> GetGradientDefs() { return {}; }
 protected:
  string I(int i) { return "I" + std::to_string(i); }
  string O(int i) { return "O" + std::to_string(i); }
  string GI(int i) { return "GI" + std::to_string(i); }
  string GO(int i) { return "GO" + std::to_string(i); }
  template <typename... A> vector<OperatorDef> SingleGradientDef(const string& type, A...) {
    OperatorDef d; d.type_ = type; return {d}; }
  OperatorDef def_;
};
#define REGISTER_GRADIENT(name, ...) static_assert(sizeof(__VA_ARGS__) > 0, "gradient maker")
#define SHOULD_NOT_DO_GRADIENT(name) static_assert(true, "no gradient")

}  // namespace caffe2
