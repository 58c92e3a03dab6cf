// ACMWeightDecayMomentumSGDUpdate fused into one pass (row a10 of SURVEY.md section 8).
//
// Replaces detectron/ops/acm_weightdecay_momentum_sgd_op.h:48-112 (five math:: passes + the
// MomentumSGDMultKernel of acm_weightdecay_momentum_sgd_op_gpu.cu:7-33) as wired per parameter
// blob by detectron/modeling/optimizer_wsl.py:96-137.  HBM-bound: with iter_size == 1 it reads
// g, m, p and writes m, p (+ an optional bf16 shadow of p for the tensor-core GEMMs): 20-22 B/param.
#include <algorithm>
#include "common.cuh"

namespace nawsod {
namespace {

constexpr int kMaxGradSources = 16;

struct SgdArgs {
  const float* extra[kMaxGradSources - 1];   // further gradient contributions, summed onto g in order (data-parallel owner)
  int n_extra;
  const float* g; float* m; const float* lr; float* p; float* acc; __nv_bfloat16* p_bf16; float* p_tf32;
  int64_t n;
  float momentum, weight_decay, lr_mult, inv_norm;
  int first_call;   // iter_count == 0: m and acc start from zero whatever the buffers hold (.h:62-69)
  int do_update;    // (iter_count + 1) % iter_size == 0
  int use_acc;      // iter_size > 1 or acc buffer given
  const uint32_t* abort_flag;   // optional: a non-zero word (the exchange's watchdog status) turns the launch into a no-op
};

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Per-element arithmetic in the reference's order (no FMA contraction across its separate passes):
//   acc = g + acc                        (.h:72-75)
//   acc = acc * (1 / (iter_size*gpu_num)) (.h:79-84)
//   acc = acc + wd * p                   (.h:88-90, Axpy)
//   v   = LR * acc + momentum * m        (.h:19)
//   m = v; p = p - v; acc = 0            (.h:20-21, 30, 106-109)
__device__ __forceinline__ void sgd_elem(float g, float& m, float& p, float& acc, const SgdArgs& a, float LR) {
  float ac = __fadd_rn(g, acc);
  if (a.do_update) {
    ac = __fmul_rn(ac, a.inv_norm);
    ac = __fadd_rn(ac, __fmul_rn(a.weight_decay, p));
    const float v = __fadd_rn(__fmul_rn(LR, ac), __fmul_rn(a.momentum, m));
    m = v;
    p = __fsub_rn(p, v);
    ac = 0.f;
  }
  acc = ac;
}

// Measured and removed (profiles/r2y_pytest_pull.log, r2z_bench_n4_tmapull.json vs r2z_bench_n4_default.json): the same reduction
// with the peers' slices pulled by TMA bulk copies (cp.async.bulk, 1 KB per source and stage, a <= 16 KB shared-memory ring so that
// the CTA stays co-resident with the GEMM's 210 KB one) -- bit-identical, but 0.37 instead of 0.29 ms per fc6 panel at 4 GPUs
// (step 4.71 vs 4.32 ms): one 14 KB ring per SM holds far fewer bytes in flight than the load-based kernel's registers do.
// kPre: number of extra gradient sources whose loads are issued TOGETHER before the first add (0: the plain update; 7 / 15: the
// data-parallel owner's reduction over up to 8 / 16 ranks).  The contributions may live in PEER memory (the pull exchange reads
// the ranks' gradient slices in place over NVLink, ~1-2 us per load): a load -> add -> load chain would serialise that latency
// once per rank; with all of a vector's loads in flight at once it is paid once.  The adds still run in rank order.
template <int kPre>
__global__ void __launch_bounds__(256) sgd_kernel(const SgdArgs a) {
  // a peer's contribution never arrived (p2p_wait's watchdog fired): leave m / p / shadow untouched rather than
  // update from an incomplete sum; the host sees the same word and raises (dp.P2PExchange.check)
  if (a.abort_flag && *reinterpret_cast<const volatile uint32_t*>(a.abort_flag) != 0u) return;
  const float LR = __fmul_rn(a.lr[0], a.lr_mult);
  const int64_t n4 = a.n / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = t0; i < n4; i += stride) {
    float4 g = __ldcs(reinterpret_cast<const float4*>(a.g) + i);
    if (kPre > 0) {
      float4 x[kPre > 0 ? kPre : 1];
#pragma unroll
      for (int e = 0; e < kPre; ++e)
        if (e < a.n_extra) x[e] = __ldcv(reinterpret_cast<const float4*>(a.extra[e]) + i);   // never from a stale L1 line
#pragma unroll
      for (int e = 0; e < kPre; ++e)             // fixed (rank) order: the sum is deterministic
        if (e < a.n_extra) {
          g.x = __fadd_rn(g.x, x[e].x); g.y = __fadd_rn(g.y, x[e].y); g.z = __fadd_rn(g.z, x[e].z); g.w = __fadd_rn(g.w, x[e].w);
        }
    }
    float4 m = a.first_call ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcs(reinterpret_cast<const float4*>(a.m) + i);
    float4 p = __ldcs(reinterpret_cast<const float4*>(a.p) + i);
    float4 c = (a.first_call || !a.use_acc) ? make_float4(0.f, 0.f, 0.f, 0.f)
                                            : reinterpret_cast<const float4*>(a.acc)[i];
    sgd_elem(g.x, m.x, p.x, c.x, a, LR);
    sgd_elem(g.y, m.y, p.y, c.y, a, LR);
    sgd_elem(g.z, m.z, p.z, c.z, a, LR);
    sgd_elem(g.w, m.w, p.w, c.w, a, LR);
    if (a.do_update || a.first_call) __stcs(reinterpret_cast<float4*>(a.m) + i, m);
    if (a.do_update) {
      __stcs(reinterpret_cast<float4*>(a.p) + i, p);
      if (a.p_bf16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
        reinterpret_cast<uint2*>(a.p_bf16)[i] =
            make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      }
      if (a.p_tf32) reinterpret_cast<float4*>(a.p_tf32)[i] = make_float4(rna_tf32(p.x), rna_tf32(p.y), rna_tf32(p.z), rna_tf32(p.w));
    }
    if (a.use_acc) reinterpret_cast<float4*>(a.acc)[i] = c;
  }
  // tail (n % 4)
  for (int64_t i = n4 * 4 + t0; i < a.n; i += stride) {
    float m = a.first_call ? 0.f : a.m[i];
    float p = a.p[i];
    float c = (a.first_call || !a.use_acc) ? 0.f : a.acc[i];
    float g = a.g[i];
    for (int e = 0; e < a.n_extra; ++e) g = __fadd_rn(g, a.extra[e][i]);
    sgd_elem(g, m, p, c, a, LR);
    if (a.do_update || a.first_call) a.m[i] = m;
    if (a.do_update) {
      a.p[i] = p;
      if (a.p_bf16) a.p_bf16[i] = __float2bfloat16_rn(p);
      if (a.p_tf32) a.p_tf32[i] = rna_tf32(p);
    }
    if (a.use_acc) a.acc[i] = c;
  }
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

static int sgd_launch(const float* const* grads, int n_grads, float* m, const float* lr, float* p, float* acc, int64_t n,
                      float momentum, float weight_decay, float lr_mult, int iter_size, int gpu_num,
                      int64_t iter_count, void* p_shadow, int shadow_dtype, const void* abort_flag, void* stream) {
  const float* g = grads[0];
  NAWSOD_REQUIRE(n >= 0, NAWSOD_ERR_SHAPE, "sgd_update: negative n");
  NAWSOD_REQUIRE(iter_size >= 1 && gpu_num >= 1 && iter_count >= 0, NAWSOD_ERR_ARG,
                 "sgd_update: need iter_size >= 1, gpu_num >= 1, iter_count >= 0");
  if (n == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(g && m && lr && p, NAWSOD_ERR_ARG, "sgd_update: null pointer");
  NAWSOD_REQUIRE(acc || iter_size == 1, NAWSOD_ERR_ARG, "sgd_update: acc buffer required when iter_size > 1");
  NAWSOD_REQUIRE(aligned16(g) && aligned16(m) && aligned16(p) && (!acc || aligned16(acc)) &&
                     (!p_shadow || aligned16(p_shadow)),
                 NAWSOD_ERR_ALIGN, "sgd_update: buffers must be 16-byte aligned");
  SgdArgs a;
  NAWSOD_REQUIRE(!p_shadow || shadow_dtype == NAWSOD_BF16 || shadow_dtype == NAWSOD_F32, NAWSOD_ERR_ARG,
                 "sgd_update: shadow_dtype must be NAWSOD_BF16 or NAWSOD_F32 (TF32-rounded)");
  a.g = g; a.m = m; a.lr = lr; a.p = p; a.acc = acc;
  a.n_extra = n_grads - 1;
  for (int e = 1; e < n_grads; ++e) {
    NAWSOD_REQUIRE(grads[e] && aligned16(grads[e]), NAWSOD_ERR_ALIGN, "sgd_update: gradient source %d is null or not 16-byte aligned", e);
    a.extra[e - 1] = grads[e];
  }
  a.p_bf16 = (p_shadow && shadow_dtype == NAWSOD_BF16) ? static_cast<__nv_bfloat16*>(p_shadow) : nullptr;
  a.p_tf32 = (p_shadow && shadow_dtype == NAWSOD_F32) ? static_cast<float*>(p_shadow) : nullptr;
  a.n = n; a.momentum = momentum; a.weight_decay = weight_decay; a.lr_mult = lr_mult;
  a.inv_norm = static_cast<float>(1.0 / (static_cast<double>(iter_size) * gpu_num));   // T(1.0 / (iter_size_ * gpu_num_))
  a.first_call = iter_count == 0;
  a.do_update = ((iter_count + 1) % iter_size) == 0;
  a.use_acc = acc != nullptr;
  a.abort_flag = static_cast<const uint32_t*>(abort_flag);
  const int64_t work = std::max<int64_t>(n / 4, 1);
  // sgd_max_ctas: when the update runs beside the tensor-core GEMMs it only has to keep up with them;
  // a narrower grid leaves the memory system's queues to the GEMM epilogues
  const int64_t cap = get_tuning("sgd_max_ctas", 0);
  const int blocks = (int)std::min<int64_t>((work + 255) / 256, cap > 0 ? cap : (int64_t)sm_count() * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a.n_extra == 0) sgd_kernel<0><<<blocks, 256, 0, st>>>(a);
  else if (a.n_extra <= 7) sgd_kernel<7><<<blocks, 256, 0, st>>>(a);
  else sgd_kernel<15><<<blocks, 256, 0, st>>>(a);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_sgd_update(const float* g, float* m, const float* lr, float* p, float* acc, int64_t n,
                                 float momentum, float weight_decay, float lr_mult, int iter_size, int gpu_num,
                                 int64_t iter_count, void* p_shadow, int shadow_dtype, void* stream) {
  const float* grads[1] = {g};
  return sgd_launch(grads, 1, m, lr, p, acc, n, momentum, weight_decay, lr_mult, iter_size, gpu_num, iter_count,
                    p_shadow, shadow_dtype, nullptr, stream);
}

extern "C" int nawsod_sgd_update_reduce(const float* const* grads, int n_grads, float* m, const float* lr, float* p,
                                        int64_t n, float momentum, float weight_decay, float lr_mult, int gpu_num,
                                        int64_t iter_count, void* p_shadow, int shadow_dtype, const void* abort_flag,
                                        void* stream) {
  NAWSOD_REQUIRE(grads && n_grads >= 1 && n_grads <= kMaxGradSources, NAWSOD_ERR_ARG,
                 "sgd_update_reduce: need 1..%d gradient sources", kMaxGradSources);
  NAWSOD_REQUIRE(n == 0 || grads[0], NAWSOD_ERR_ARG, "sgd_update_reduce: null gradient source");
  return sgd_launch(grads, n_grads, m, lr, p, nullptr, n, momentum, weight_decay, lr_mult, 1, gpu_num, iter_count,
                    p_shadow, shadow_dtype, abort_flag, stream);
}
