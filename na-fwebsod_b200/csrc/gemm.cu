// FC-stack GEMMs on the 5th-generation tensor cores (tcgen05 + TMEM) fed by TMA, sm_100a.
//
// Replaces Caffe2 FC / FCGradient (cuBLAS sgemm in the reference) + Relu + Dropout and their
// gradient ops as wired by detectron/modeling/wsl_heads.py:674-679 and
// detectron/modeling/webly_heads.py:490-498.  One persistent, warp-specialised kernel serves
// all three GEMM shapes of a layer:
//     fwd    Y [M,N]  = A[M,K]  . W[N,K]^T     A: K-major,  B = W : K-major
//     bwd_x  dA[M,K]  = dY[M,N] . W[N,K]       A: K-major,  B = W : MN-major (no transposed copy)
//     bwd_w  dW[N,K]  = dY[M,N]^T . A[M,K]     A = dY: MN-major, B = A: MN-major
// so no operand is ever transposed in memory: MN-major operands are described to the tensor
// core through the UMMA shared-memory descriptor (a_major / b_major bits) and loaded by TMA as
// 128-byte-wide column panels.
//
// Kernel anatomy (192 threads, one CTA per SM, persistent over output tiles; CTA PAIRS -- (2,1,1) clusters, tcgen05
// cta_group::2 -- for 256-wide tiles, see the PAIR template parameter):
//   warp 0      TMA producer: cp.async.bulk.tensor.3d into a 128B-swizzled smem ring (4 stages of 48 KB; pairs: 6 stages of
//               32 KB, each CTA staging its 128 rows of A and half of the B tile); optionally gated on the peer exchange's
//               "operands have landed" flags (nawsod_fc_fwd_gated)
//   warp 1      MMA issuer: one lane issues tcgen05.mma (M=128 or, for a pair, the even CTA issues M=256; N=BN, K=32 bytes)
//               into a double-buffered TMEM accumulator (2 x BN fp32 columns); tcgen05.commit (multicast to both CTAs of a
//               pair) releases smem slots and publishes finished accumulators
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns) -> [prior output] / bias / ReLU / dropout / ReLU-gradient ->
//               bf16 or fp32 stores; plain fp32 outputs (the weight gradients) are staged in swizzled smem and stored by the
//               TMA (bulk tensor store, or reduce-add when accumulating); overlaps the next tile's main loop
#include "gemm_tc.cuh"

namespace nawsod {
namespace {

// default of the gemm_pair tuning knob (0: one CTA per tile, 1: CTA pairs for BN = 256 tiles; measured r2t: 4.43 -> 4.16 ms per step)
constexpr long long kGemmPairDefault = 1;
// default of the gemm_tma_store tuning knob (1: plain fp32 outputs -- the weight gradients -- are stored by the TMA)
// (measured r2w: step 4.17 -> 4.08 ms, fc6 dW in the step 1.56 -> 1.48 ms)
constexpr long long kGemmTmaStoreDefault = 1;

struct EpiParams {
  void* out; long long ldo; int out_dtype;
  const float* bias;
  const uint8_t* mask; long long ldmask;
  const void* act; long long ldact; int act_dtype;
  int flags;
  unsigned long long seed;   // counter-based dropout (fwd) when mask == nullptr and seed != 0
  int M, N, K;          // GEMM dims: out is [M, N], reduction over K
  // independent problems of one shape in one launch (the head's clean / noisy stacks): problem b uses
  // out + b*so, bias + b*sbias, mask + b*smask, act + b*sact (elements) and seed + b; the operand strides
  // live in the 3-D tensor maps
  int nbatch;
  long long so, sbias, smask, sact;
  // Optional weight gate (data-parallel peer exchange): rows [g * gate_rows, (g + 1) * gate_rows) of the B operand (W) may only
  // be read once the gate_nflags words gate_flags[g * gate_nflags ...] have reached gate_seq -- the "operands of exchange
  // bucket g have landed" flags the owner ranks publish (p2p.cu).  The TMA producer checks them when its tiles enter a new
  // row group, so the GEMM starts on the weights that are there and meets the rest as they arrive.
  const uint32_t* gate_flags; int gate_nflags, gate_rows; uint32_t gate_seq;
  unsigned long long gate_timeout_ns; uint32_t* gate_status;
  // 1: plain fp32 output (no bias / activation / mask; optionally ACCUMULATE) leaves through the TMA: each epilogue warp stages its
  // 32 x 32 chunk in swizzled shared memory and one lane issues a bulk tensor store (reduce-add when accumulating) -- full
  // 128-byte lines instead of 32 scattered 16-byte stores per instruction, and nothing queued in the LSU next to a concurrent
  // update kernel's long-latency loads
  int tma_store;
};

constexpr int kStagingBytes = 4 * 32 * 128;      // one 32-row x 128-byte box per epilogue warp

// spin (one thread) until all n flags have reached `value` (sequence numbers wrap); a watchdog turns a lost peer into a status
// word instead of a hung GPU
__device__ __forceinline__ void gate_wait(const uint32_t* flags, int n, uint32_t value, unsigned long long timeout_ns, uint32_t* status) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < n; ++i) {
    unsigned ns = 32;
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      if (static_cast<int32_t>(v - value) >= 0) break;
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) { if (status) atomicMax(status, 1u); break; }
      __nanosleep(ns);
      if (ns < 512) ns <<= 1;
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");     // the TMA loads that follow read what the flags guard
}

// PAIR: the CTA-pair form (cta_group::2, launched as (2,1,1) clusters): a pair owns a 256 x BN tile, CTA r of the pair stages rows
// [r * 128, r * 128 + 128) of A and rows [r * BN/2, (r + 1) * BN/2) of B, the even CTA issues M = 256 MMAs for both, each CTA's TMEM
// holds its own 128 rows of the accumulator and its epilogue warps store them.  Half the B bytes per FLOP through L2 and shared
// memory, 32 KB instead of 48 KB per stage -> 6 stages instead of 4 of prefetch distance.
template <int BN, bool A_MN, bool B_MN, int ES, bool PAIR = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const EpiParams ep_) {
  using C = typename std::conditional<PAIR, CfgPair<BN, ES>, Cfg<BN, ES>>::type;
  const EpiParams& ep = ep_;
  constexpr int BK = C::BK;
  constexpr int ATOM = 128 / ES;                 // MN elements per 128-byte panel
  constexpr int TILE_M = PAIR ? 2 * BLOCK_M : BLOCK_M;
  constexpr int BNL = PAIR ? BN / 2 : BN;        // B rows this CTA stages
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int cta = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int ncta = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging_base = smem_base + C::STAGES * C::STAGE_BYTES;     // 1024-byte aligned: the swizzle phase of a row is row & 7
  const uint32_t bar_base = staging_base + kStagingBytes;
  // barrier layout (8 bytes each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM pointer
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_ptr_generic =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (ep.M + TILE_M - 1) / TILE_M;
  const int num_n = (ep.N + BN - 1) / BN;
  const int tiles_pb = num_m * num_n;
  const int num_tiles = tiles_pb * ep.nbatch;
  const int num_kb = (ep.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (ep.tma_store) tma_prefetch_desc(&tmO);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    // the accumulator is released by the 4 epilogue warps of every CTA that holds a part of it
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                  // the peer's barriers and TMEM exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_generic;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int gate_open = -1;                        // row groups [0, gate_open] of W have been waited for
      auto load = [&](uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
        if (PAIR) tma_load_3d_pair(dst, map, bar, c0, c1, c2);     // completes on the leader's barrier
        else tma_load_3d(dst, map, bar, c0, c1, c2);
      };
      for (int tile = cta; tile < num_tiles; tile += ncta) {
        const int bi = tile / tiles_pb, tb = tile - bi * tiles_pb;
        const int m0 = (tb % num_m) * TILE_M + static_cast<int>(crank) * BLOCK_M, n0 = (tb / num_m) * BN;
        const int n0l = n0 + static_cast<int>(crank) * BNL;        // this CTA's share of the B tile
        if (ep.gate_flags) {                     // tiles arrive in ascending n0: groups open in order
          const int need = (min(n0 + BN, ep.N) - 1) / ep.gate_rows;
          for (; gate_open < need; ++gate_open)
            gate_wait(ep.gate_flags + (size_t)(gate_open + 1) * ep.gate_nflags, ep.gate_nflags, ep.gate_seq, ep.gate_timeout_ns, ep.gate_status);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          // pair: the leader's barrier counts both CTAs' bytes (a peer's complete_tx may precede the leader's expect_tx: the
          // phase cannot complete before the leader's arrival)
          if (!PAIR) mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          else if (crank == 0) mbar_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
          const int k0 = kb * BK;
          if (!A_MN) load(sa, &tmA, full_bar(stage), k0, m0, bi);
          else {
#pragma unroll
            for (int a = 0; a < BLOCK_M / ATOM; ++a) load(sa + a * BK * 128, &tmA, full_bar(stage), m0 + a * ATOM, k0, bi);
          }
          if (!B_MN) load(sb, &tmB, full_bar(stage), k0, n0l, bi);
          else {
#pragma unroll
            for (int a = 0; a < BNL / ATOM; ++a) load(sb + a * BK * 128, &tmB, full_bar(stage), n0l + a * ATOM, k0, bi);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && crank == 0) {               // pair: the even CTA issues for both
      constexpr uint32_t idesc = make_idesc(ES, A_MN, B_MN, TILE_M, BN);
      // K-major: 8-row groups 1024 B apart (SBO), LBO unused; MN-major: 128-byte panels BK*128 B apart (LBO),
      // 8-row k-groups 1024 B apart (SBO)
      constexpr uint32_t a_lbo = A_MN ? BK * 128 : 16, b_lbo = B_MN ? BK * 128 : 16;
      // 4-byte MN-major operands: 32-byte swizzle atoms, k-groups of 4 rows (512 B apart)
      constexpr uint32_t a_lay = (A_MN && ES == 4) ? 1 : 2, b_lay = (B_MN && ES == 4) ? 1 : 2;
      constexpr uint32_t a_sbo = (A_MN && ES == 4) ? 512 : 1024, b_sbo = (B_MN && ES == 4) ? 512 : 1024;
      constexpr uint32_t a_kstep = A_MN ? (C::UMMA_K * 128) >> 4 : 32 >> 4;   // descriptor start-address step per UMMA_K
      constexpr uint32_t b_kstep = B_MN ? (C::UMMA_K * 128) >> 4 : 32 >> 4;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = cta; tile < num_tiles; tile += ncta) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          const uint64_t adesc = make_smem_desc(sa, a_lbo, a_sbo, a_lay), bdesc = make_smem_desc(sb, b_lbo, b_sbo, b_lay);
#pragma unroll
          for (int k = 0; k < BK / C::UMMA_K; ++k) {
            if (PAIR) tc_mma_pair<ES>(tmem_d, adesc + (uint64_t)(k * a_kstep), bdesc + (uint64_t)(k * b_kstep), idesc, (kb | k) != 0);
            else tc_mma<ES>(tmem_d, adesc + (uint64_t)(k * a_kstep), bdesc + (uint64_t)(k * b_kstep), idesc, (kb | k) != 0);
          }
          if (PAIR) tc_commit_pair(empty_bar(stage)); else tc_commit(empty_bar(stage));     // pair: frees the slot in both CTAs
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) tc_commit_pair(tfull_bar(acc)); else tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5) =================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    const bool vec_ok = ((ep.ldo * (ep.out_dtype == NAWSOD_F32 ? 4 : 2)) % 16 == 0) &&
                        ((reinterpret_cast<uintptr_t>(ep.out) & 15u) == 0);
    const bool bias_vec = ep.bias && (reinterpret_cast<uintptr_t>(ep.bias) & 15u) == 0;
    const bool act_vec = ep.act && (reinterpret_cast<uintptr_t>(ep.act) & 15u) == 0 && (ep.ldact * 2) % 16 == 0;
    const bool mask_vec = ep.mask && (reinterpret_cast<uintptr_t>(ep.mask) & 15u) == 0 && ep.ldmask % 16 == 0;
    for (int tile = cta; tile < num_tiles; tile += ncta) {
      const int bi = tile / tiles_pb, tb = tile - bi * tiles_pb;
      const int m0 = (tb % num_m) * TILE_M + static_cast<int>(crank) * BLOCK_M, n0 = (tb / num_m) * BN;
      // this problem's view of the epilogue operands, as scalars (a per-tile mutable copy of the whole parameter
      // struct proved fragile: one build kept reading the unshifted kernel parameters for stacks > 0)
      const size_t oes = ep.out_dtype == NAWSOD_F32 ? 4 : 2, aes = ep.act_dtype == NAWSOD_BF16 ? 2 : 4;
      char* const out_b = static_cast<char*>(ep.out) + (size_t)bi * ep.so * oes;
      const float* const bias_b = ep.bias ? ep.bias + (size_t)bi * ep.sbias : nullptr;
      const uint8_t* const mask_b = ep.mask ? ep.mask + (size_t)bi * ep.smask : nullptr;
      const char* const act_b = ep.act ? static_cast<const char*>(ep.act) + (size_t)bi * ep.sact * aes : nullptr;
      const unsigned long long seed_b = ep.seed ? ep.seed + bi : 0;   // seed 0 means "no counter-based dropout" for every stack
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < ep.M;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= ep.N) break;                // warp-uniform
        uint32_t r[32];
        __syncwarp();                             // tcgen05.ld is .sync.aligned: the warp must be converged
        tc_ld32(tmem_base + acc * BN + c + (static_cast<uint32_t>(q * 32) << 16), r);
        tc_wait_ld();
        const int n = n0 + c;
        const int ncols = min(32, ep.N - n);
        if (ep.tma_store) {                       // warp-uniform
          // row `lane` of the warp's 32 x 32 box: eight 16-byte chunks, chunk j at (j ^ (row & 7)) -- the 128B swizzle the
          // tensor map undoes; a quarter-warp's eight rows then hit eight different bank groups
          const uint32_t stg = staging_base + static_cast<uint32_t>(q) * (32 * 128);
          if (lane == 0) tma_store_wait_read();   // the previous box of this warp has left shared memory
          __syncwarp();
          const uint32_t rowp = stg + static_cast<uint32_t>(lane) * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + ((j ^ (lane & 7)) << 4)),
                         "r"(r[4 * j]), "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3]) : "memory");
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (ep.flags & NAWSOD_FC_ACCUMULATE) tma_reduce_add_3d(&tmO, stg, n, m0 + q * 32, bi);
            else tma_store_3d(&tmO, stg, n, m0 + q * 32, bi);
            tma_store_commit();
          }
          continue;
        }
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          const bool full = ncols == 32;
          if ((ep.flags & NAWSOD_FC_ACCUMULATE) && ep.out_dtype == NAWSOD_F32) {
            // add what the output holds BEFORE bias / activation: the earlier passes of a split-operand (3 x TF32) product left
            // their raw partial sums there, the weight-gradient GEMMs their running gradient
            const float* o = reinterpret_cast<const float*>(out_b) + (size_t)m * ep.ldo + n;
            if (vec_ok && full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 o4 = *reinterpret_cast<const float4*>(o + i);
                v[i] += o4.x; v[i + 1] += o4.y; v[i + 2] += o4.z; v[i + 3] += o4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += o[i];
            }
          }
          if (bias_b) {
            if (full && bias_vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias_b + n + i));
                v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) v[i] += __ldg(bias_b + n + i);
            }
          }
          if (ep.flags & NAWSOD_FC_RELU) {
            if (act_b) {
              if (full && act_vec && ep.act_dtype == NAWSOD_BF16) {
                const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(act_b) + (size_t)m * ep.ldact + n);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint4 a4 = ap[i];
                  const uint32_t w[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                    if (!((w[j] & 0x8000u) == 0 && (w[j] & 0x7FFFu) != 0)) v[i * 8 + j * 2] = 0.f;
                    if (!((w[j] & 0x80000000u) == 0 && (w[j] & 0x7FFF0000u) != 0)) v[i * 8 + j * 2 + 1] = 0.f;
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  if (i < ncols) v[i] = ld_act(act_b, ep.act_dtype, (size_t)m * ep.ldact + n + i) > 0.f ? v[i] : 0.f;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
          }
          if (ep.flags & NAWSOD_FC_DROPOUT) {
            if (mask_b) {
              if (full && mask_vec) {
                const uint4* mp = reinterpret_cast<const uint4*>(mask_b + (size_t)m * ep.ldmask + n);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const uint4 m4 = mp[i];
                  const uint32_t w[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    v[i * 16 + j] *= 2.0f * static_cast<float>((w[j >> 2] >> (8 * (j & 3))) & 0xFFu);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) if (i < ncols) v[i] *= 2.0f * static_cast<float>(mask_b[(size_t)m * ep.ldmask + n + i]);
              }
            } else if (seed_b != 0) {
              // counter-based keep bits: one 64-bit mix per (row, 32-column chunk), one bit per column
              unsigned long long h = seed_b ^ (0x9E3779B97F4A7C15ull * (unsigned long long)(m + 1)) ^
                                     (0xC2B2AE3D27D4EB4Full * (unsigned long long)(n / 32 + 1));
              h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull; h ^= h >> 33;
              const uint32_t bits = static_cast<uint32_t>(h);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = ((bits >> i) & 1u) ? v[i] * 2.0f : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= 2.0f;
            }
          }
          if (ep.out_dtype == NAWSOD_F32) {
            float* o = reinterpret_cast<float*>(out_b) + (size_t)m * ep.ldo + n;
            if (ep.flags & NAWSOD_FC_ROUND_TF32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = round_tf32(v[i]);
            }
            if (vec_ok && full) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = v[i];
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_b) + (size_t)m * ep.ldo + n;
            if (vec_ok && full) {
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i], v[i + 1]), h1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]), h3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
                *reinterpret_cast<uint4*>(o + i) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                              *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (i < ncols) o[i] = __float2bfloat16_rn(v[i]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (ep.tma_store && warp >= 2 && lane == 0) tma_store_wait_all();          // this thread's bulk stores have been written
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                  // the leader's MMAs read the peer's shared memory until the last tile is done
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// small helpers: column sums (bias gradient) and float -> bf16 conversion
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ X, long long ld, int M, int N, int rows_per_block,
                                                    float* __restrict__ out) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int r = r0 + ty; r < r1; r += 8) {
      const T x = X[(size_t)r * ld + n];
      if constexpr (sizeof(T) == 2) s += __uint_as_float(static_cast<uint32_t>(x) << 16);
      else s += x;
    }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][tx];
    atomicAdd(out + n, t);
  }
}

__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ src, long long ld_src, long long rows,
                                                        long long cols, float* __restrict__ dst, long long ld_dst) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = round_tf32(src[r * ld_src + c]);
  }
}

// x = hi + lo with hi = nearest TF32 of x and lo = nearest TF32 of the (exact) remainder: the operand pair of a split-operand
// (3 x TF32) product, x.y ~ hi.hi' + lo.hi' + hi.lo' to ~2^-21 relative
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, long long ld_src, long long rows,
                                                        long long cols, float* __restrict__ hi, long long ld_hi,
                                                        float* __restrict__ lo, long long ld_lo) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    const float x = src[r * ld_src + c];
    // the 13 bits below the TF32 mantissa are masked explicitly: the remainder must be taken against exactly the value the
    // tensor core will read (it ignores those bits)
    const float h = __uint_as_float(__float_as_uint(round_tf32(x)) & 0xFFFFE000u);
    hi[r * ld_hi + c] = h;
    lo[r * ld_lo + c] = __uint_as_float(__float_as_uint(round_tf32(__fsub_rn(x, h))) & 0xFFFFE000u);
  }
}

__global__ void __launch_bounds__(256) cvt_bf16_kernel(const float* __restrict__ src, long long ld_src, long long rows, long long cols,
                                                      __nv_bfloat16* __restrict__ dst, long long ld_dst) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = __float2bfloat16_rn(src[r * ld_src + c]);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Operands { const void* A; long long lda, sA; const void* B; long long ldb, sB; };

template <int BN, bool A_MN, bool B_MN, int ES, bool PAIR = false>
int launch_gemm(const Operands& o, const EpiParams& ep_in, cudaStream_t st) {
  using C = typename std::conditional<PAIR, CfgPair<BN, ES>, Cfg<BN, ES>>::type;
  constexpr int ATOM = 128 / ES;
  constexpr int TILE_M = PAIR ? 2 * BLOCK_M : BLOCK_M;
  CUtensorMap tmA, tmB;
  int rc;
  EpiParams ep = ep_in;
  // K-major operand: stored [MN, K]; box [BLOCK rows, BK].  MN-major operand: stored [K, MN]; box [BK rows, ATOM].
  // (pair: a CTA stages BN / 2 rows of B per k-block)
  if (!A_MN) rc = make_tmap(&tmA, o.A, ES, ep.M, ep.K, o.lda, BLOCK_M, C::BK, false, ep.nbatch, o.sA);
  else rc = make_tmap(&tmA, o.A, ES, ep.K, ep.M, o.lda, C::BK, ATOM, ES == 4, ep.nbatch, o.sA);
  if (rc) return rc;
  if (!B_MN) rc = make_tmap(&tmB, o.B, ES, ep.N, ep.K, o.ldb, PAIR ? BN / 2 : BN, C::BK, false, ep.nbatch, o.sB);
  else rc = make_tmap(&tmB, o.B, ES, ep.K, ep.N, o.ldb, C::BK, ATOM, ES == 4, ep.nbatch, o.sB);
  if (rc) return rc;
  CUtensorMap tmO = tmA;                         // placeholder unless the output leaves through the TMA
  ep.tma_store = 0;
  if (ep.out_dtype == NAWSOD_F32 && !ep.bias && !ep.mask && !ep.act && !(ep.flags & ~NAWSOD_FC_ACCUMULATE) && aligned16(ep.out) &&
      (ep.ldo * 4) % 16 == 0 && (ep.nbatch <= 1 || (ep.so * 4) % 16 == 0) && get_tuning("gemm_tma_store", kGemmTmaStoreDefault) != 0) {
    // out [M, N] float, 32 x 32 boxes (128 bytes wide, 128B swizzle)
    if (int rc2 = make_tmap(&tmO, ep.out, 4, ep.M, ep.N, ep.ldo, 32, 32, false, ep.nbatch, ep.so)) return rc2;
    ep.tma_store = 1;
  }
  constexpr int kSmem = C::SMEM_BYTES + kStagingBytes;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, ES, PAIR>;
  static bool attr_set = false;
  if (!attr_set) {
    NAWSOD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  const int num_tiles = ((ep.M + TILE_M - 1) / TILE_M) * ((ep.N + BN - 1) / BN) * ep.nbatch;
  // gemm_max_ctas: leave SMs to concurrently running collectives (a persistent one-CTA-per-SM grid
  // would otherwise wait behind them, or they behind it)
  const int cap = (int)get_tuning("gemm_max_ctas", 0);
  const int sms = cap > 0 ? std::min(cap, sm_count()) : sm_count();
  if (PAIR) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * std::min(num_tiles, std::max(sms / 2, 1)));    // whole pairs: (2,1,1) clusters on the SMs of one TPC
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    NAWSOD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmO, ep));
    return NAWSOD_OK;
  }
  const int grid = std::min(num_tiles, sms);
  kern<<<grid, kNumThreads, kSmem, st>>>(tmA, tmB, tmO, ep);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

template <bool A_MN, bool B_MN>
int dispatch_gemm(const Operands& o, EpiParams ep, int ab_dtype, cudaStream_t st) {
  if (ep.nbatch < 1) ep.nbatch = 1;
  if (ep.nbatch > 1) {
    // the vectorised epilogue paths are chosen once per launch: every stack must keep problem 0's alignment
    const long long oes = ep.out_dtype == NAWSOD_F32 ? 4 : 2, aes = ep.act_dtype == NAWSOD_BF16 ? 2 : 4;
    NAWSOD_REQUIRE(ep.so > 0 && (ep.so * oes) % 16 == 0 && (!ep.bias || (ep.sbias * 4) % 16 == 0) &&
                       (!ep.mask || ep.smask % 16 == 0) && (!ep.act || (ep.sact * aes) % 16 == 0),
                   NAWSOD_ERR_ALIGN, "fc: stack strides of the output / bias / mask / activation must be multiples of 16 bytes");
  }
  // BN = 256 for wide outputs; narrower tiles only when N itself is narrow (fc8: N = 2C)
  const int bn = ep.N > 128 ? 256 : (ep.N > 64 ? 128 : 64);
  // gemm_pair: wide outputs on CTA pairs (cta_group::2) when the output has at least two 128-row blocks
  const bool pair = bn == 256 && ep.M > BLOCK_M && get_tuning("gemm_pair", kGemmPairDefault) != 0;
  if (pair) {
    if (ab_dtype == NAWSOD_BF16) return launch_gemm<256, A_MN, B_MN, 2, true>(o, ep, st);
    return launch_gemm<256, A_MN, B_MN, 4, true>(o, ep, st);
  }
  if (ab_dtype == NAWSOD_BF16) {
    if (bn == 256) return launch_gemm<256, A_MN, B_MN, 2>(o, ep, st);
    if (bn == 128) return launch_gemm<128, A_MN, B_MN, 2>(o, ep, st);
    return launch_gemm<64, A_MN, B_MN, 2>(o, ep, st);
  }
  if (bn == 256) return launch_gemm<256, A_MN, B_MN, 4>(o, ep, st);
  if (bn == 128) return launch_gemm<128, A_MN, B_MN, 4>(o, ep, st);
  return launch_gemm<64, A_MN, B_MN, 4>(o, ep, st);
}

int check_common(const char* who, int M, int N, int K, int ab_dtype) {
  NAWSOD_REQUIRE(M > 0 && N > 0 && K > 0, NAWSOD_ERR_SHAPE, "%s: need M, N, K > 0 (got %d, %d, %d)", who, M, N, K);
  NAWSOD_REQUIRE(ab_dtype == NAWSOD_BF16 || ab_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "%s: bad ab_dtype", who);
  return NAWSOD_OK;
}

}  // namespace
}  // namespace nawsod

namespace nawsod {
// Bias gradient db[s] = column sums of dY[s] (FCGradient's db; shared by the plain and the fused weight-gradient entry points).
int fc_bias_grad(const void* dY, int64_t lddy, int64_t sdY, int S, int M, int N, int ab_dtype, float* db, int64_t sdb, int flags,
                 cudaStream_t st) {
  if (db) {
    const int es = ab_dtype == NAWSOD_BF16 ? 2 : 4;
    for (int s = 0; s < S; ++s) {
      float* dbs = db + (size_t)s * sdb;
      const char* dys = static_cast<const char*>(dY) + (size_t)s * sdY * es;
      if (!(flags & NAWSOD_FC_ACCUMULATE)) NAWSOD_CUDA_OK(cudaMemsetAsync(dbs, 0, (size_t)N * sizeof(float), st));
      const int row_blocks = std::max(1, std::min(64, M / 64));
      const int rpb = (M + row_blocks - 1) / row_blocks;
      dim3 grid((N + 31) / 32, row_blocks);
      if (ab_dtype == NAWSOD_BF16)
        colsum_kernel<uint16_t><<<grid, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(dys), lddy, M, N, rpb, dbs);
      else
        colsum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(dys), lddy, M, N, rpb, dbs);
      NAWSOD_LAUNCH_OK();
    }
  }
  return NAWSOD_OK;
}
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_fc_fwd_stacks(const void* A, int64_t lda, int64_t sA, const void* W, int64_t ldw, int64_t sW,
                                    const float* bias, int64_t sbias, const uint8_t* mask, int64_t ldmask, int64_t smask,
                                    uint64_t dropout_seed, int S, int M, int N, int K, int ab_dtype, void* Y, int64_t ldy,
                                    int64_t sY, int y_dtype, int flags, void* stream) {
  if (int rc = check_common("fc_fwd", M, N, K, ab_dtype)) return rc;
  NAWSOD_REQUIRE(S >= 1, NAWSOD_ERR_SHAPE, "fc_fwd: need at least one stack");
  NAWSOD_REQUIRE(A && W && Y, NAWSOD_ERR_ARG, "fc_fwd: null pointer");
  NAWSOD_REQUIRE(y_dtype == NAWSOD_F32 || y_dtype == NAWSOD_BF16, NAWSOD_ERR_ARG, "fc_fwd: bad y_dtype");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_ACCUMULATE) || y_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "fc_fwd: ACCUMULATE needs a float output");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_ROUND_TF32) || y_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "fc_fwd: ROUND_TF32 needs a float output");
  NAWSOD_REQUIRE(ldy >= N && (!mask || ldmask >= N), NAWSOD_ERR_SHAPE, "fc_fwd: ldy / ldmask smaller than N");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_DROPOUT) || mask || dropout_seed != 0, NAWSOD_ERR_ARG,
                 "fc_fwd: DROPOUT needs a mask or a non-zero dropout_seed (it would only scale the activations by 2)");
  EpiParams ep{};
  ep.out = Y; ep.ldo = ldy; ep.out_dtype = y_dtype; ep.bias = bias; ep.mask = (flags & NAWSOD_FC_DROPOUT) ? mask : nullptr;
  ep.ldmask = ldmask; ep.act = nullptr; ep.flags = flags; ep.seed = dropout_seed; ep.M = M; ep.N = N; ep.K = K;
  ep.nbatch = S; ep.so = sY; ep.sbias = sbias; ep.smask = smask; ep.sact = 0;
  const Operands o{A, lda, sA, W, ldw, sW};
  return dispatch_gemm<false, false>(o, ep, ab_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int nawsod_fc_fwd_gated(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const uint8_t* mask,
                                   int64_t ldmask, uint64_t dropout_seed, int M, int N, int K, int ab_dtype, void* Y, int64_t ldy,
                                   int y_dtype, int flags, const void* gate_flags, int gate_groups, int gate_nflags, int gate_rows,
                                   uint32_t gate_seq, int64_t gate_timeout_ms, void* gate_status, void* stream) {
  if (int rc = check_common("fc_fwd_gated", M, N, K, ab_dtype)) return rc;
  NAWSOD_REQUIRE(A && W && Y, NAWSOD_ERR_ARG, "fc_fwd_gated: null pointer");
  NAWSOD_REQUIRE(y_dtype == NAWSOD_F32 || y_dtype == NAWSOD_BF16, NAWSOD_ERR_ARG, "fc_fwd_gated: bad y_dtype");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_ACCUMULATE), NAWSOD_ERR_ARG, "fc_fwd_gated: ACCUMULATE is a bwd_w flag");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_ROUND_TF32) || y_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "fc_fwd_gated: ROUND_TF32 needs a float output");
  NAWSOD_REQUIRE(ldy >= N && (!mask || ldmask >= N), NAWSOD_ERR_SHAPE, "fc_fwd_gated: ldy / ldmask smaller than N");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_DROPOUT) || mask || dropout_seed != 0, NAWSOD_ERR_ARG,
                 "fc_fwd_gated: DROPOUT needs a mask or a non-zero dropout_seed");
  NAWSOD_REQUIRE(gate_flags && gate_groups >= 1 && gate_nflags >= 1 && gate_rows > 0 && gate_rows % 256 == 0 &&
                     (int64_t)gate_groups * gate_rows >= N && gate_timeout_ms > 0,
                 NAWSOD_ERR_ARG, "fc_fwd_gated: need gate flags, groups * rows covering N = %d, rows a multiple of 256, a time-out", N);
  EpiParams ep{};
  ep.out = Y; ep.ldo = ldy; ep.out_dtype = y_dtype; ep.bias = bias; ep.mask = (flags & NAWSOD_FC_DROPOUT) ? mask : nullptr;
  ep.ldmask = ldmask; ep.act = nullptr; ep.flags = flags; ep.seed = dropout_seed; ep.M = M; ep.N = N; ep.K = K;
  ep.nbatch = 1;
  ep.gate_flags = static_cast<const uint32_t*>(gate_flags); ep.gate_nflags = gate_nflags; ep.gate_rows = gate_rows;
  ep.gate_seq = gate_seq; ep.gate_timeout_ns = (unsigned long long)gate_timeout_ms * 1000000ull;
  ep.gate_status = static_cast<uint32_t*>(gate_status);
  const Operands o{A, lda, 0, W, ldw, 0};
  return dispatch_gemm<false, false>(o, ep, ab_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int nawsod_fc_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const uint8_t* mask,
                             int64_t ldmask, uint64_t dropout_seed, int M, int N, int K, int ab_dtype, void* Y, int64_t ldy,
                             int y_dtype, int flags, void* stream) {
  return nawsod_fc_fwd_stacks(A, lda, 0, W, ldw, 0, bias, 0, mask, ldmask, 0, dropout_seed, 1, M, N, K, ab_dtype, Y, ldy, 0,
                              y_dtype, flags, stream);
}

extern "C" int nawsod_fc_bwd_x_stacks(const void* dY, int64_t lddy, int64_t sdY, const void* W, int64_t ldw, int64_t sW,
                                      const void* act_below, int64_t ldact, int64_t sact, int act_dtype,
                                      const uint8_t* mask_below, int64_t ldmask, int64_t smask, int S, int M, int N, int K,
                                      int ab_dtype, void* dA, int64_t ldda, int64_t sdA, int da_dtype, int flags, void* stream) {
  if (int rc = check_common("fc_bwd_x", M, N, K, ab_dtype)) return rc;
  NAWSOD_REQUIRE(S >= 1, NAWSOD_ERR_SHAPE, "fc_bwd_x: need at least one stack");
  NAWSOD_REQUIRE(dY && W && dA, NAWSOD_ERR_ARG, "fc_bwd_x: null pointer");
  NAWSOD_REQUIRE(da_dtype == NAWSOD_F32 || da_dtype == NAWSOD_BF16, NAWSOD_ERR_ARG, "fc_bwd_x: bad da_dtype");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_RELU) || act_below, NAWSOD_ERR_ARG, "fc_bwd_x: RELU needs act_below");
  NAWSOD_REQUIRE(!(flags & NAWSOD_FC_ACCUMULATE) || da_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "fc_bwd_x: ACCUMULATE needs a float output");
  NAWSOD_REQUIRE(ldda >= K, NAWSOD_ERR_SHAPE, "fc_bwd_x: ldda smaller than K");
  EpiParams ep{};
  // GEMM view: out [M, K] = dY [M, N] . W [N, K]  -> reduction over N, "N" of the GEMM is K
  ep.out = dA; ep.ldo = ldda; ep.out_dtype = da_dtype; ep.bias = nullptr;
  ep.mask = (flags & NAWSOD_FC_DROPOUT) ? mask_below : nullptr; ep.ldmask = ldmask;
  ep.act = (flags & NAWSOD_FC_RELU) ? act_below : nullptr; ep.ldact = ldact; ep.act_dtype = act_dtype;
  ep.flags = flags; ep.M = M; ep.N = K; ep.K = N;
  ep.nbatch = S; ep.so = sdA; ep.sbias = 0; ep.smask = smask; ep.sact = sact;
  const Operands o{dY, lddy, sdY, W, ldw, sW};
  return dispatch_gemm<false, true>(o, ep, ab_dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int nawsod_fc_bwd_x(const void* dY, int64_t lddy, const void* W, int64_t ldw, const void* act_below, int64_t ldact,
                               int act_dtype, const uint8_t* mask_below, int64_t ldmask, int M, int N, int K, int ab_dtype,
                               void* dA, int64_t ldda, int da_dtype, int flags, void* stream) {
  return nawsod_fc_bwd_x_stacks(dY, lddy, 0, W, ldw, 0, act_below, ldact, 0, act_dtype, mask_below, ldmask, 0, 1, M, N, K,
                                ab_dtype, dA, ldda, 0, da_dtype, flags, stream);
}

extern "C" int nawsod_fc_bwd_w_stacks(const void* dY, int64_t lddy, int64_t sdY, const void* A, int64_t lda, int64_t sA, int S,
                                      int M, int N, int K, int ab_dtype, float* dW, int64_t lddw, int64_t sdW, float* db,
                                      int64_t sdb, int flags, void* stream) {
  if (int rc = check_common("fc_bwd_w", M, N, K, ab_dtype)) return rc;
  NAWSOD_REQUIRE(S >= 1, NAWSOD_ERR_SHAPE, "fc_bwd_w: need at least one stack");
  NAWSOD_REQUIRE(dY && A && dW, NAWSOD_ERR_ARG, "fc_bwd_w: null pointer");
  NAWSOD_REQUIRE(lddw >= K, NAWSOD_ERR_SHAPE, "fc_bwd_w: lddw smaller than K");
  NAWSOD_REQUIRE(!(flags & (NAWSOD_FC_RELU | NAWSOD_FC_DROPOUT | NAWSOD_FC_ROUND_TF32)), NAWSOD_ERR_ARG,
                 "fc_bwd_w: only ACCUMULATE is valid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EpiParams ep{};
  // GEMM view: out [N, K] = dY^T [N, M] . A [M, K] -> reduction over M
  ep.out = dW; ep.ldo = lddw; ep.out_dtype = NAWSOD_F32; ep.flags = flags; ep.M = N; ep.N = K; ep.K = M;
  ep.nbatch = S; ep.so = sdW;
  const Operands o{dY, lddy, sdY, A, lda, sA};
  if (int rc = dispatch_gemm<true, true>(o, ep, ab_dtype, st)) return rc;
  return fc_bias_grad(dY, lddy, sdY, S, M, N, ab_dtype, db, sdb, flags, st);
}

extern "C" int nawsod_fc_bias_grad(const void* dY, int64_t lddy, int64_t sdY, int S, int M, int N, int ab_dtype, float* db,
                                   int64_t sdb, int flags, void* stream) {
  NAWSOD_REQUIRE(S >= 1 && M > 0 && N > 0 && lddy >= N, NAWSOD_ERR_SHAPE, "fc_bias_grad: need S >= 1, M, N > 0 and lddy >= N");
  NAWSOD_REQUIRE(ab_dtype == NAWSOD_BF16 || ab_dtype == NAWSOD_F32, NAWSOD_ERR_ARG, "fc_bias_grad: bad ab_dtype");
  NAWSOD_REQUIRE(dY && db, NAWSOD_ERR_ARG, "fc_bias_grad: null pointer");
  NAWSOD_REQUIRE(!(flags & ~NAWSOD_FC_ACCUMULATE), NAWSOD_ERR_ARG, "fc_bias_grad: only ACCUMULATE is valid");
  return fc_bias_grad(dY, lddy, sdY, S, M, N, ab_dtype, db, sdb, flags, static_cast<cudaStream_t>(stream));
}

extern "C" int nawsod_fc_bwd_w(const void* dY, int64_t lddy, const void* A, int64_t lda, int M, int N, int K, int ab_dtype,
                               float* dW, int64_t lddw, float* db, int flags, void* stream) {
  return nawsod_fc_bwd_w_stacks(dY, lddy, 0, A, lda, 0, 1, M, N, K, ab_dtype, dW, lddw, 0, db, 0, flags, stream);
}

extern "C" int nawsod_convert_f32_to_bf16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* dst, int64_t ld_dst,
                                          void* stream) {
  NAWSOD_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= cols, NAWSOD_ERR_SHAPE, "convert: bad shape");
  if (rows == 0 || cols == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(src && dst, NAWSOD_ERR_ARG, "convert: null pointer");
  const long long total = rows * cols;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
  cvt_bf16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ld_src, rows, cols, static_cast<__nv_bfloat16*>(dst), ld_dst);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_round_to_tf32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* dst, int64_t ld_dst,
                                    void* stream) {
  NAWSOD_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld_dst >= cols, NAWSOD_ERR_SHAPE, "round_to_tf32: bad shape");
  if (rows == 0 || cols == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(src && dst, NAWSOD_ERR_ARG, "round_to_tf32: null pointer");
  const long long total = rows * cols;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
  round_tf32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ld_src, rows, cols, dst, ld_dst);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_split_tf32(const float* src, int64_t ld_src, int64_t rows, int64_t cols, float* hi, int64_t ld_hi, float* lo,
                                 int64_t ld_lo, void* stream) {
  NAWSOD_REQUIRE(rows >= 0 && cols >= 0 && ld_src >= cols && ld_hi >= cols && ld_lo >= cols, NAWSOD_ERR_SHAPE, "split_tf32: bad shape");
  if (rows == 0 || cols == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(src && hi && lo && lo != src && lo != hi, NAWSOD_ERR_ARG, "split_tf32: null pointer, or lo aliases src / hi");
  const long long total = rows * cols;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 8);
  split_tf32_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ld_src, rows, cols, hi, ld_hi, lo, ld_lo);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}
