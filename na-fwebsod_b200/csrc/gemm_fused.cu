// fc6 weight-gradient GEMM with the step that FOLLOWS it fused into the epilogue (sm_100a; experimental, opt-in).
//
// The weight gradient dW[N,K] = dY[M,N]^T . A[M,K] of fc6 is 86 % of the head's parameter bytes.  In the reference it is
// written by FCGradient (cuBLAS), all-reduced per blob (detectron/modeling/optimizer_wsl.py:52-72) and re-read by
// ACMWeightDecayMomentumSGDUpdate (detectron/ops/acm_weightdecay_momentum_sgd_op.h:48-112).  The stand-alone path of this
// library keeps those as separate kernels (gemm.cu -> p2p.cu / NCCL -> sgd.cu); here the consumer is folded into the
// producer so that the 822 MB gradient never makes the round trip through HBM:
//
//   nawsod_fc_bwd_w_sgd      (one GPU)   the epilogue applies the SGD update to momentum / master parameter / GEMM-operand
//                                        shadow of the tile it just accumulated; the gradient itself is optional output;
//   nawsod_fc_bwd_w_scatter  (N GPUs)    the epilogue stores each tile straight into the staging buffer of the rank that
//                                        OWNS those rows (peer-mapped memory over NVLink / NVSwitch, posted 16-byte
//                                        stores), i.e. GEMM + the reduce-scatter's send leg in one kernel, tile by tile.
//
// Main loop: identical to gemm_tcgen05_kernel<256, MN, MN> (TMA producer warp, one-lane tcgen05.mma issuer, double-buffered
// TMEM accumulator).  Epilogue (warps 2..5), SGD mode: tcgen05.ld hands every lane one ROW of 32 columns, which is the wrong
// shape for a read-modify-write of three more arrays (a first version that let every thread walk its own 128-byte row
// segments of m / p ran 14x slower than GEMM + stand-alone update: 32 distinct lines per warp instruction thrash the small
// L1 left beside 200 KB of pipeline stages).  So each warp transposes its 32 x 32 chunk through a padded shared-memory tile
// and then owns whole 128-byte row segments per instruction: lane = column, all m / p loads of the chunk are issued before
// the first use (64 independent loads in flight per lane), and the tile's m / p lines are prefetched into L2 while the
// tensor cores are still working on it.  The arithmetic is sgd.cu's, operation for operation (bit-exact with the
// stand-alone update on the same gradient).
#include "gemm_tc.cuh"

namespace nawsod {
namespace {

constexpr int kFusedBN = 256;
constexpr int kStagePitch = 33;                                  // floats per staged row: conflict-free both ways
constexpr int kStageBytes = 4 * 32 * kStagePitch * 4;            // one 32 x 32 chunk per epilogue warp
constexpr int kMaxOwners = 16;

enum { kModeSgd = 0, kModeScatter = 1 };

struct FusedParams {
  int M, N, K;                 // GEMM dims: out [M, N] (M = rows of W, N = columns of W), reduction over K (RoIs)
  long long ldo;               // row pitch (elements) of dW / m / p / shadow
  int accumulate;              // dW += (SGD mode with a gradient buffer, scatter mode: never)
  // ---- SGD mode
  float* g;                    // gradient out, or null
  float* m; float* p;
  void* shadow;                // GEMM-operand copy of p in the operands' type (bf16, or float rounded to TF32)
  const float* lr;
  float momentum, weight_decay, lr_mult, inv_norm;
  int first_call;
  // ---- scatter mode: rows [k * rows_per_owner, (k + 1) * rows_per_owner) go to owner_out[k] (row 0 of that range first)
  float* owner_out[kMaxOwners];
  int rows_per_owner;
};

__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int ES, int MODE, bool WRITE_G>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_dw_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FusedParams fp_) {
  constexpr int BN = kFusedBN;
  using C = Cfg<BN, ES>;
  const FusedParams& fp = fp_;
  constexpr int BK = C::BK;
  constexpr int ATOM = 128 / ES;                 // MN elements per 128-byte panel
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::STAGES * C::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));
  float* const stage_all = reinterpret_cast<float*>(smem_raw + (bar_base + 256u - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (fp.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (fp.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (fp.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ================= TMA producer (both operands MN-major: 128-byte column panels) =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile % num_m) * BLOCK_M, n0 = (tile / num_m) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          const int k0 = kb * BK;
#pragma unroll
          for (int a = 0; a < BLOCK_M / ATOM; ++a) tma_load_3d(sa + a * BK * 128, &tmA, full_bar(stage), m0 + a * ATOM, k0, 0);
#pragma unroll
          for (int a = 0; a < BN / ATOM; ++a) tma_load_3d(sb + a * BK * 128, &tmB, full_bar(stage), n0 + a * ATOM, k0, 0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(ES, true, true, BLOCK_M, BN);
      constexpr uint32_t lbo = BK * 128;
      constexpr uint32_t lay = (ES == 4) ? 1 : 2;
      constexpr uint32_t sbo = (ES == 4) ? 512 : 1024;
      constexpr uint32_t kstep = (C::UMMA_K * 128) >> 4;        // descriptor start-address step per UMMA_K
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          const uint64_t adesc = make_smem_desc(sa, lbo, sbo, lay), bdesc = make_smem_desc(sb, lbo, sbo, lay);
#pragma unroll
          for (int k = 0; k < BK / C::UMMA_K; ++k)
            tc_mma<ES>(tmem_d, adesc + (uint64_t)(k * kstep), bdesc + (uint64_t)(k * kstep), idesc, (kb | k) != 0);
          tc_commit(empty_bar(stage));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    float* const st = stage_all + q * (32 * kStagePitch);
    int acc = 0; uint32_t acc_phase = 0;
    const float LR = (MODE == kModeSgd) ? __fmul_rn(__ldg(fp.lr), fp.lr_mult) : 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * BLOCK_M, n0 = (tile / num_m) * BN;
      const int mq = m0 + q * 32;                 // first row of this warp's quarter
      if (MODE == kModeSgd) {
        // while the tensor cores work on this tile: pull its m / p lines into L2 (8 lines per row and array)
        const int prow = mq + (lane >> 3) * 8, pcol = n0 + (lane & 7) * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = prow + i;
          if (r < fp.M && pcol < fp.N) {
            const size_t off = (size_t)r * fp.ldo + pcol;
            prefetch_l2(fp.p + off);
            if (!fp.first_call) prefetch_l2(fp.m + off);
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= fp.N) break;                // warp-uniform
        uint32_t r[32];
        __syncwarp();                             // tcgen05.ld is .sync.aligned: the warp must be converged
        tc_ld32(tmem_base + acc * BN + c + (static_cast<uint32_t>(q * 32) << 16), r);
        tc_wait_ld();
        const int n = n0 + c;
        if (MODE == kModeScatter) {
          // lane = row: 128-byte row segments straight to the owner of these rows (local or peer-mapped memory)
          const int mrow = mq + lane;
          const int ncols = min(32, fp.N - n);
          if (mrow < fp.M) {
            const int owner = mrow / fp.rows_per_owner;
            float* o = fp.owner_out[owner] + (size_t)(mrow - owner * fp.rows_per_owner) * fp.ldo + n;
            if (ncols == 32) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(o + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = __uint_as_float(r[i]);
            }
          }
        } else {
          // transpose the chunk: lane = row  ->  lane = column
#pragma unroll
          for (int i = 0; i < 32; ++i) st[lane * kStagePitch + i] = __uint_as_float(r[i]);
          __syncwarp();
          const int col = n + lane;
          const bool col_ok = col < fp.N;
          const int nrows = min(32, fp.M - mq);   // warp-uniform; <= 0 when the whole quarter is padding
          const bool first = fp.first_call != 0;
          float mm[32], pp[32];
          size_t off = (size_t)mq * fp.ldo + col;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr, off += fp.ldo) {
            mm[rr] = 0.f; pp[rr] = 0.f;
            if (rr < nrows && col_ok) {
              if (!first) mm[rr] = __ldcs(fp.m + off);
              pp[rr] = __ldcs(fp.p + off);
            }
          }
          off = (size_t)mq * fp.ldo + col;
          const float* srow = st + lane;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr, off += fp.ldo) {
            if (rr < nrows && col_ok) {
              float g = srow[rr * kStagePitch];
              if (WRITE_G) {
                if (fp.accumulate) g += fp.g[off];
                fp.g[off] = g;
              }
              // sgd.cu sgd_elem with iter_size == 1 (accumulator 0): acm_weightdecay_momentum_sgd_op.h:72-109
              float ac = __fadd_rn(g, 0.f);
              ac = __fmul_rn(ac, fp.inv_norm);
              ac = __fadd_rn(ac, __fmul_rn(fp.weight_decay, pp[rr]));
              const float v = __fadd_rn(__fmul_rn(LR, ac), __fmul_rn(fp.momentum, mm[rr]));
              const float pn = __fsub_rn(pp[rr], v);
              __stcs(fp.m + off, v);
              __stcs(fp.p + off, pn);
              // the GEMM-operand shadow has the operands' type: bf16, or float rounded to the nearest TF32
              if (ES == 2) static_cast<__nv_bfloat16*>(fp.shadow)[off] = __float2bfloat16_rn(pn);
              else static_cast<float*>(fp.shadow)[off] = rna_tf32(pn);
            }
          }
          __syncwarp();                           // the staging tile is rewritten by the next chunk
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

template <int ES, int MODE, bool WRITE_G>
int launch_fused(const void* dY, long long lddy, const void* A, long long lda, const FusedParams& fp, cudaStream_t st) {
  using C = Cfg<kFusedBN, ES>;
  constexpr int ATOM = 128 / ES;
  constexpr int kSmem = C::SMEM_BYTES + kStageBytes;
  static_assert(kSmem <= 227 * 1024, "pipeline stages + epilogue staging exceed the shared memory of one CTA");
  CUtensorMap tmA, tmB;
  // both operands MN-major: stored [K, MN]; box [BK rows, ATOM]
  if (int rc = make_tmap(&tmA, dY, ES, fp.K, fp.M, lddy, C::BK, ATOM, ES == 4, 1, 0)) return rc;
  if (int rc = make_tmap(&tmB, A, ES, fp.K, fp.N, lda, C::BK, ATOM, ES == 4, 1, 0)) return rc;
  auto kern = gemm_dw_fused_kernel<ES, MODE, WRITE_G>;
  static bool attr_set = false;
  if (!attr_set) {
    NAWSOD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  const int num_tiles = ((fp.M + BLOCK_M - 1) / BLOCK_M) * ((fp.N + kFusedBN - 1) / kFusedBN);
  const int cap = (int)get_tuning("gemm_max_ctas", 0);
  const int grid = std::min(num_tiles, cap > 0 ? std::min(cap, sm_count()) : sm_count());
  kern<<<grid, kNumThreads, kSmem, st>>>(tmA, tmB, fp);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

int check_fused(const char* who, const void* dY, const void* A, int M, int N, int K, int ab_dtype, int64_t ldw, int flags) {
  NAWSOD_REQUIRE(M > 0 && N > 0 && K > 0, NAWSOD_ERR_SHAPE, "%s: need M, N, K > 0 (got %d, %d, %d)", who, M, N, K);
  NAWSOD_REQUIRE(ab_dtype == NAWSOD_BF16 || ab_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "%s: bad ab_dtype", who);
  NAWSOD_REQUIRE(dY && A, NAWSOD_ERR_ARG, "%s: null operand", who);
  NAWSOD_REQUIRE(ldw >= K, NAWSOD_ERR_SHAPE, "%s: ldw smaller than K", who);
  NAWSOD_REQUIRE(!(flags & ~NAWSOD_FC_ACCUMULATE), NAWSOD_ERR_ARG, "%s: only ACCUMULATE is a valid flag", who);
  return NAWSOD_OK;
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_fc_bwd_w_sgd(const void* dY, int64_t lddy, const void* A, int64_t lda, int M, int N, int K, int ab_dtype,
                                   float* dW, int64_t ldw, float* db, int flags, float* m, float* p, void* p_shadow,
                                   int shadow_dtype, const float* lr, float momentum, float weight_decay, float lr_mult,
                                   int gpu_num, int64_t iter_count, void* stream) {
  if (int rc = check_fused("fc_bwd_w_sgd", dY, A, M, N, K, ab_dtype, ldw, flags)) return rc;
  NAWSOD_REQUIRE(m && p && lr, NAWSOD_ERR_ARG, "fc_bwd_w_sgd: null momentum / parameter / lr");
  NAWSOD_REQUIRE(dW || !(flags & NAWSOD_FC_ACCUMULATE), NAWSOD_ERR_ARG, "fc_bwd_w_sgd: ACCUMULATE needs the gradient buffer dW");
  NAWSOD_REQUIRE(gpu_num >= 1 && iter_count >= 0, NAWSOD_ERR_ARG, "fc_bwd_w_sgd: need gpu_num >= 1 and iter_count >= 0");
  NAWSOD_REQUIRE(p_shadow && shadow_dtype == ab_dtype, NAWSOD_ERR_UNSUPPORTED,
                 "fc_bwd_w_sgd: needs the GEMM-operand shadow of p, in the operands' type (bf16, or float = TF32-rounded)");
  NAWSOD_REQUIRE((ldw * 4) % 16 == 0, NAWSOD_ERR_ALIGN, "fc_bwd_w_sgd: ldw must be a multiple of 4 floats");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FusedParams fp{};
  // GEMM view: out [N, K] = dY^T [N, M] . A [M, K] -> reduction over M
  fp.M = N; fp.N = K; fp.K = M; fp.ldo = ldw; fp.accumulate = (flags & NAWSOD_FC_ACCUMULATE) ? 1 : 0;
  fp.g = dW; fp.m = m; fp.p = p; fp.shadow = p_shadow; fp.lr = lr;
  fp.momentum = momentum; fp.weight_decay = weight_decay; fp.lr_mult = lr_mult;
  fp.inv_norm = static_cast<float>(1.0 / static_cast<double>(gpu_num));      // T(1.0 / (iter_size_ * gpu_num_)), iter_size 1
  fp.first_call = iter_count == 0;
  int rc;
  if (ab_dtype == NAWSOD_BF16) rc = dW ? launch_fused<2, kModeSgd, true>(dY, lddy, A, lda, fp, st) : launch_fused<2, kModeSgd, false>(dY, lddy, A, lda, fp, st);
  else rc = dW ? launch_fused<4, kModeSgd, true>(dY, lddy, A, lda, fp, st) : launch_fused<4, kModeSgd, false>(dY, lddy, A, lda, fp, st);
  if (rc) return rc;
  return fc_bias_grad(dY, lddy, 0, 1, M, N, ab_dtype, db, 0, flags, st);
}

extern "C" int nawsod_fc_bwd_w_scatter(const void* dY, int64_t lddy, const void* A, int64_t lda, int M, int N, int K,
                                       int ab_dtype, float* const* owner_dW, int n_owners, int rows_per_owner, int64_t ldw,
                                       float* db, void* stream) {
  if (int rc = check_fused("fc_bwd_w_scatter", dY, A, M, N, K, ab_dtype, ldw, 0)) return rc;
  NAWSOD_REQUIRE(owner_dW && n_owners >= 1 && n_owners <= kMaxOwners, NAWSOD_ERR_ARG, "fc_bwd_w_scatter: need 1..%d owners", kMaxOwners);
  NAWSOD_REQUIRE(rows_per_owner > 0 && rows_per_owner % BLOCK_M == 0 && (int64_t)rows_per_owner * n_owners >= N, NAWSOD_ERR_SHAPE,
                 "fc_bwd_w_scatter: rows_per_owner must be a positive multiple of %d covering the %d rows", BLOCK_M, N);
  NAWSOD_REQUIRE((ldw * 4) % 16 == 0, NAWSOD_ERR_ALIGN, "fc_bwd_w_scatter: ldw must be a multiple of 4 floats");
  FusedParams fp{};
  fp.M = N; fp.N = K; fp.K = M; fp.ldo = ldw; fp.rows_per_owner = rows_per_owner;
  for (int k = 0; k < n_owners; ++k) {
    NAWSOD_REQUIRE(owner_dW[k] && aligned16(owner_dW[k]), NAWSOD_ERR_ALIGN, "fc_bwd_w_scatter: owner buffer %d is null or not 16-byte aligned", k);
    fp.owner_out[k] = owner_dW[k];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rc = ab_dtype == NAWSOD_BF16 ? launch_fused<2, kModeScatter, true>(dY, lddy, A, lda, fp, st)
                                         : launch_fused<4, kModeScatter, true>(dY, lddy, A, lda, fp, st);
  if (rc) return rc;
  return fc_bias_grad(dY, lddy, 0, 1, M, N, ab_dtype, db, 0, 0, st);
}
