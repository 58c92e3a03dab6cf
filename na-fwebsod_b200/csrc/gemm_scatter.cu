// fc6 weight-gradient GEMM with the reduce-scatter's SEND LEG fused into the epilogue (sm_100a).
//
// The weight gradient dW[N,K] = dY[M,N]^T . A[M,K] of fc6 is 86 % of the head's parameter bytes.  In the reference it is
// written by FCGradient (cuBLAS) and all-reduced per blob (detectron/modeling/optimizer_wsl.py:52-72).  In the
// data-parallel step of this library (dp.py, sync = "p2p") rank k owns row slice k of every exchange bucket; the
// stand-alone path writes the gradient locally (gemm.cu) and a scatter kernel (p2p.cu) re-reads it and stores the
// peers' slices into their staging areas.  Here the epilogue stores each tile straight into the staging buffer of the rank
// that OWNS those rows (peer-mapped memory over NVLink / NVSwitch, posted 16-byte stores; the local gradient buffer for
// the rank's own slice): GEMM + send leg in ONE kernel, tile by tile, so the transfer overlaps the math and the gradient
// is neither re-read nor written twice.
//
// Main loop: identical to gemm_tcgen05_kernel<256, MN, MN> (TMA producer warp, one-lane tcgen05.mma issuer, double-buffered
// TMEM accumulator).  Epilogue (warps 2..5): tcgen05.ld hands every lane one ROW of 32 columns = one 128-byte segment.
//
// (A second epilogue that applied the SGD update in place -- momentum / master / shadow read-modify-write through a
// transposing shared-memory tile -- was measured in round 2 and REMOVED: 4.44 ms per step against 4.13 ms for GEMM +
// side-stream update, profiles/r2a_bench_n1_fused_sgd.json.)
#include "gemm_tc.cuh"

namespace nawsod {
namespace {

constexpr int kFusedBN = 256;
constexpr int kMaxOwners = 16;

struct FusedParams {
  int M, N, K;                 // GEMM dims: out [M, N] (M = rows of W, N = columns of W), reduction over K (RoIs)
  long long ldo;               // row pitch (elements) of every owner's buffer
  // rows [k * rows_per_owner, (k + 1) * rows_per_owner) go to owner_out[k] (row 0 of that range first)
  float* owner_out[kMaxOwners];
  int rows_per_owner;
};

template <int ES>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_dw_scatter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FusedParams fp_) {
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
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile % num_m) * BLOCK_M, n0 = (tile / num_m) * BN;
      const int mrow = m0 + q * 32 + lane;        // lane = row of this warp's TMEM lane quarter
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      float* o = nullptr;
      if (mrow < fp.M) {
        const int owner = mrow / fp.rows_per_owner;
        o = fp.owner_out[owner] + (size_t)(mrow - owner * fp.rows_per_owner) * fp.ldo;
      }
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= fp.N) break;                // warp-uniform
        uint32_t r[32];
        __syncwarp();                             // tcgen05.ld is .sync.aligned: the warp must be converged
        tc_ld32(tmem_base + acc * BN + c + (static_cast<uint32_t>(q * 32) << 16), r);
        tc_wait_ld();
        const int n = n0 + c;
        const int ncols = min(32, fp.N - n);
        if (o) {                                  // 128-byte row segment straight to the owner (local or peer-mapped)
          if (ncols == 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(o + n + i) = make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                  __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) if (i < ncols) o[n + i] = __uint_as_float(r[i]);
          }
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

template <int ES>
int launch_scatter(const void* dY, long long lddy, const void* A, long long lda, const FusedParams& fp, cudaStream_t st) {
  using C = Cfg<kFusedBN, ES>;
  constexpr int ATOM = 128 / ES;
  constexpr int kSmem = C::SMEM_BYTES;
  CUtensorMap tmA, tmB;
  // both operands MN-major: stored [K, MN]; box [BK rows, ATOM]
  if (int rc = make_tmap(&tmA, dY, ES, fp.K, fp.M, lddy, C::BK, ATOM, ES == 4, 1, 0)) return rc;
  if (int rc = make_tmap(&tmB, A, ES, fp.K, fp.N, lda, C::BK, ATOM, ES == 4, 1, 0)) return rc;
  auto kern = gemm_dw_scatter_kernel<ES>;
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

int check_scatter(const char* who, const void* dY, const void* A, int M, int N, int K, int ab_dtype, int64_t ldw) {
  NAWSOD_REQUIRE(M > 0 && N > 0 && K > 0, NAWSOD_ERR_SHAPE, "%s: need M, N, K > 0 (got %d, %d, %d)", who, M, N, K);
  NAWSOD_REQUIRE(ab_dtype == NAWSOD_BF16 || ab_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "%s: bad ab_dtype", who);
  NAWSOD_REQUIRE(dY && A, NAWSOD_ERR_ARG, "%s: null operand", who);
  NAWSOD_REQUIRE(ldw >= K, NAWSOD_ERR_SHAPE, "%s: ldw smaller than K", who);
  return NAWSOD_OK;
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_fc_bwd_w_scatter(const void* dY, int64_t lddy, const void* A, int64_t lda, int M, int N, int K,
                                       int ab_dtype, float* const* owner_dW, int n_owners, int rows_per_owner, int64_t ldw,
                                       float* db, void* stream) {
  if (int rc = check_scatter("fc_bwd_w_scatter", dY, A, M, N, K, ab_dtype, ldw)) return rc;
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
  const int rc = ab_dtype == NAWSOD_BF16 ? launch_scatter<2>(dY, lddy, A, lda, fp, st)
                                         : launch_scatter<4>(dY, lddy, A, lda, fp, st);
  if (rc) return rc;
  return fc_bias_grad(dY, lddy, 0, 1, M, N, ab_dtype, db, 0, 0, st);
}
