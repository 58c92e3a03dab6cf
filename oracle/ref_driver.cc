// Host-side harness (TEST INFRASTRUCTURE ONLY) that instantiates the reference's own CPU
// operators -- compiled unmodified from /root/reference/detectron/ops/*.cc against
// oracle/c2shim -- and runs them on caller buffers.  Exposed as a tiny C ABI for ctypes:
//   nawsod_ref_create(type, n_args, names, vals) -> handle (operator instance; keeps the
//       op's hidden state such as ACMWeightDecayMomentumSGDUpdate's iter_count_)
//   nawsod_ref_run(handle, ...)  -> 0 ok / 1 enforce failure / 2 RunOnDevice()==false
//   nawsod_ref_destroy(handle)
#include <cstring>
#include <memory>
#include "caffe2/core/operator.h"
#include "caffe2/core/logging.h"

using namespace caffe2;

namespace {
struct Handle { std::unique_ptr<OperatorBase> op; std::string err; };
std::string g_err;
}

extern "C" {

const char* nawsod_ref_last_error() { return g_err.c_str(); }

int nawsod_ref_has_op(const char* type) { return CPUOperatorRegistry().count(type) ? 1 : 0; }

void* nawsod_ref_create(const char* type, int n_args, const char** names, const double* vals) {
  auto it = CPUOperatorRegistry().find(type);
  if (it == CPUOperatorRegistry().end()) { g_err = std::string("unknown op ") + type; return nullptr; }
  OperatorDef def; def.type_ = type;
  for (int i = 0; i < n_args; ++i) def.args[names[i]] = vals[i];
  auto* h = new Handle();
  try { h->op.reset(it->second(def, nullptr)); }
  catch (const std::exception& e) { g_err = e.what(); delete h; return nullptr; }
  return h;
}

void nawsod_ref_destroy(void* hv) { delete static_cast<Handle*>(hv); }

// in_dims: concatenated dims of every input.  out_alias[k] >= 0 makes output k the same
// tensor object as input out_alias[k] (Caffe2 in-place).  Outputs are copied into
// out_ptrs[k] (caller-sized, capacity out_cap[k] floats); their shapes are returned in
// out_ndims / out_dims (8 slots per output).
int nawsod_ref_run(void* hv, int n_in, const float** in_ptrs, const int* in_ndims, const int64_t* in_dims,
                   int n_out, const int* out_alias, float** out_ptrs, const int64_t* out_cap,
                   int* out_ndims, int64_t* out_dims) {
  auto* h = static_cast<Handle*>(hv);
  std::vector<std::unique_ptr<Tensor>> ins(n_in), outs(n_out);
  const int64_t* d = in_dims;
  h->op->inputs_.clear(); h->op->outputs_.clear();
  for (int i = 0; i < n_in; ++i) {
    std::vector<int64_t> s(d, d + in_ndims[i]); d += in_ndims[i];
    ins[i].reset(new Tensor());
    int64_t n = 1; for (auto v : s) n *= v;
    ins[i]->set(s, in_ptrs[i], (size_t)n * sizeof(float));
    h->op->inputs_.push_back(ins[i].get());
  }
  for (int k = 0; k < n_out; ++k) {
    if (out_alias[k] >= 0) h->op->outputs_.push_back(ins[out_alias[k]].get());
    else { outs[k].reset(new Tensor()); h->op->outputs_.push_back(outs[k].get()); }
  }
  try {
    if (!h->op->RunOnDevice()) { g_err = "RunOnDevice returned false"; return 2; }
  } catch (const std::exception& e) { g_err = e.what(); return 1; }
  for (int k = 0; k < n_out; ++k) {
    Tensor* t = h->op->outputs_[k];
    if (t->numel() > out_cap[k]) { g_err = "output buffer too small"; return 1; }
    out_ndims[k] = t->dim();
    for (int j = 0; j < t->dim(); ++j) out_dims[k * 8 + j] = t->sizes()[j];
    std::memcpy(out_ptrs[k], t->data<float>(), (size_t)t->numel() * sizeof(float));
  }
  return 0;
}

}  // extern "C"
