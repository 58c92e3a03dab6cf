// CAFFE_ENFORCE* -> std::runtime_error (Caffe2 throws EnforceNotMet, surfaced to Python as RuntimeError).
#pragma once
#include <sstream>
#include <stdexcept>
#define C2SHIM_FAIL(msg) do { std::ostringstream os_; os_ << __FILE__ << ":" << __LINE__ << " enforce failed: " << msg; throw std::runtime_error(os_.str()); } while (0)
#define CAFFE_ENFORCE(cond, ...) do { if (!(cond)) C2SHIM_FAIL(#cond); } while (0)
#define CAFFE_ENFORCE_EQ(a, b, ...) do { if (!((a) == (b))) C2SHIM_FAIL(#a " == " #b); } while (0)
#define CAFFE_ENFORCE_GT(a, b, ...) do { if (!((a) > (b))) C2SHIM_FAIL(#a " > " #b); } while (0)
#define CAFFE_ENFORCE_GE(a, b, ...) do { if (!((a) >= (b))) C2SHIM_FAIL(#a " >= " #b); } while (0)
#define CAFFE_ENFORCE_LT(a, b, ...) do { if (!((a) < (b))) C2SHIM_FAIL(#a " < " #b); } while (0)
#define CAFFE_ENFORCE_LE(a, b, ...) do { if (!((a) <= (b))) C2SHIM_FAIL(#a " <= " #b); } while (0)
