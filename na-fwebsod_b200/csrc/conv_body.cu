// Frozen VGG16 conv body in channels-last bf16 (row N4 of SURVEY.md section 8f).
//
// The reference builds conv1_1 .. conv5_3 with Caffe2 `Conv` (cuDNN, NCHW fp32) + `Relu` + `MaxPool`
// (detectron/modeling/VGG16.py:9-58, the flickr configs' MODEL.CONV_BODY) and freezes all of it
// (TRAIN.FREEZE_CONV_BODY, yaml:31), so the body is a forward-only producer of the head's conv5 map.
//
//   nawsod_conv3x3_relu  Conv(3x3, stride 1, pad = dilation) + bias + Relu as an IMPLICIT GEMM on the tcgen05 tensor cores:
//                      no patch matrix.  An output tile is 8 x 16 pixels (128 GEMM rows) x BN output channels; the
//                      reduction runs over 9 taps x Cin / 64 channel blocks.  For tap (kh, kw) the A operand of a k-block is
//                      the input map's [8, 16, 64-channel] box shifted by ((kh - 1) d, (kw - 1) d): ONE 4-D tiled TMA
//                      load per k-block, whose out-of-bounds zero fill is the convolution's zero padding and whose shared-
//                      memory image (128 rows of 128 swizzled bytes) is exactly the K-major operand tile tcgen05.mma reads.
//                      The weights [Cout, (kh, kw, c)] are the K-major B operand as they lie.  Same warp roles, smem ring,
//                      double-buffered TMEM accumulator and epilogue structure as gemm.cu.  Cin a multiple of 64 (all of
//                      VGG16 but conv1_1, which keeps the patch-matrix path: K = 9 x 8 padded planes).
//   nawsod_im2col3x3   X [N,H,W,C] bf16  ->  cols [N*H*W, 9*C] bf16, K-order (kh, kw, c), stride 1, pad = dilation
//                      (conv1_1, and the reference form the implicit GEMM is tested against).
//   nawsod_maxpool2x2  MaxPool(kernel=2, pad=0, stride=1|2) on a channels-last bf16 map (Caffe2's floor output size:
//                      (H - 2) / stride + 1); packed bf16x2 maxima, 16 bytes per thread.
//
// Measured and removed (profiles/r2w_microbench_convbody_pair0.log / _pair1.log): the CTA-pair form of this kernel (cta_group::2,
// 16 x 16-pixel tiles, half of the weight tile per CTA; bit-identical outputs) -- 0.582 vs 0.566 ms for the 480 x 640 body, 0.873 vs
// 0.869 ms for 688 x 912 (only conv3_3 on the large image gained, 53 -> 49 us): these layers are bound by tile count, short
// reductions and the epilogue, not by the weight tile's bytes, unlike the FC GEMMs where pairs took 6 % off the step.
// Measured (profiles/r2a_microbench_convbody.log -> r2*_microbench_convbody.log): the patch-matrix body ran 1.0 ms for a
// 480 x 640 image (235 TFLOP/s, 14 % of the tensor peak: the patch matrix moves ~10x the layers' algorithmic bytes).
#include <cuda.h>
#include <algorithm>
#include "gemm_tc.cuh"


namespace nawsod {
namespace {

constexpr int kThreads = 256;

// one thread = one 16-byte vector (8 channels) of one tap of one output pixel
__global__ void __launch_bounds__(kThreads) im2col3x3_kernel(const uint4* __restrict__ X, int N, int H, int W, int C8, int dil,
                                                              uint4* __restrict__ out, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % C8);
    int64_t t = i / C8;
    const int tap = (int)(t % 9);
    t /= 9;                                   // output pixel index n*H*W + h*W + w
    const int w = (int)(t % W);
    const int64_t t2 = t / W;
    const int h = (int)(t2 % H);
    const int64_t n = t2 / H;
    const int hh = h + (tap / 3 - 1) * dil, ww = w + (tap % 3 - 1) * dil;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(X + ((n * H + hh) * W + ww) * C8 + v);
    out[i] = val;
  }
}

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162 m = __hmax2(x, y);
  return *reinterpret_cast<uint32_t*>(&m);
}

__device__ __forceinline__ uint4 max4(uint4 a, uint4 b) {
  return make_uint4(max_bf16x2(a.x, b.x), max_bf16x2(a.y, b.y), max_bf16x2(a.z, b.z), max_bf16x2(a.w, b.w));
}

__global__ void __launch_bounds__(kThreads) maxpool2x2_kernel(const uint4* __restrict__ X, int H, int W, int C8, int s,
                                                               int Ho, int Wo, uint4* __restrict__ out, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = (int)(i % C8);
    int64_t t = i / C8;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int64_t n = t / Ho;
    const uint4* p = X + ((n * H + (int64_t)ho * s) * W + (int64_t)wo * s) * C8 + v;
    const uint4 a = __ldg(p), b = __ldg(p + C8), c = __ldg(p + (int64_t)W * C8), d = __ldg(p + (int64_t)W * C8 + C8);
    out[i] = max4(max4(a, b), max4(c, d));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// implicit-GEMM 3x3 convolution
// ------------------------------------------------------------------------------------------------------------------
constexpr int kTileH = 8, kTileW = 16;           // 128 output pixels per tile = the GEMM's BLOCK_M rows, row = h_local * 16 + w_local

struct ConvParams {
  int N, H, W, Cin, Cout, dil, relu;
  const float* bias;
  __nv_bfloat16* Y;                              // [N, H, W, Cout]
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
conv3x3_igemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  using C = Cfg<BN, 2>;
  constexpr int BK = C::BK;                      // 64 channels = one 128-byte swizzle atom
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
  const int tiles_w = (p.W + kTileW - 1) / kTileW, tiles_h = (p.H + kTileH - 1) / kTileH;
  const int num_m = p.N * tiles_h * tiles_w;
  const int num_n = (p.Cout + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int cblocks = p.Cin / BK;
  const int num_kb = 9 * cblocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
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

  // tile -> (image, first output row, first output column, first output channel); pixel tiles fastest, so the CTAs that run
  // together share the weight tile through L2
  auto tile_origin = [&](int tile, int& n, int& h0, int& w0, int& n0) {
    const int tm = tile % num_m;
    n0 = (tile / num_m) * BN;
    n = tm / (tiles_h * tiles_w);
    const int r = tm - n * tiles_h * tiles_w;
    h0 = (r / tiles_w) * kTileH;
    w0 = (r % tiles_w) * kTileW;
  };

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int n, h0, w0, n0;
        tile_origin(tile, n, h0, w0, n0);
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / cblocks, c0 = (kb - tap * cblocks) * BK;
          const int dh = (tap / 3 - 1) * p.dil, dw = (tap % 3 - 1) * p.dil;
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES, sb = sa + C::A_BYTES;
          mbar_expect_tx(full_bar(stage), C::STAGE_BYTES);
          // the shifted [8, 16, 64] box of the input map: rows / columns outside the image arrive as zeros (the padding)
          tma_load_4d(sa, &tmX, full_bar(stage), c0, w0 + dw, h0 + dh, n);
          tma_load_3d(sb, &tmW, full_bar(stage), kb * BK, n0, 0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(2, false, false, BLOCK_M, BN);
      constexpr uint32_t kstep = 32 >> 4;
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
          const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2), bdesc = make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
          for (int k = 0; k < BK / C::UMMA_K; ++k)
            tc_mma<2>(tmem_d, adesc + (uint64_t)(k * kstep), bdesc + (uint64_t)(k * kstep), idesc, (kb | k) != 0);
          tc_commit(empty_bar(stage));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue (warps 2..5): bias + Relu -> bf16, 16-byte stores =================
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int n, h0, w0, n0;
      tile_origin(tile, n, h0, w0, n0);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int ml = q * 32 + lane;
      const int h = h0 + ml / kTileW, w = w0 + ml % kTileW;
      const bool pix_ok = h < p.H && w < p.W;
      __nv_bfloat16* const orow = p.Y + (((size_t)n * p.H + h) * p.W + w) * p.Cout;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (n0 + c >= p.Cout) break;              // warp-uniform
        uint32_t r[32];
        __syncwarp();
        tc_ld32(tmem_base + acc * BN + c + (static_cast<uint32_t>(q * 32) << 16), r);
        tc_wait_ld();
        if (pix_ok) {
          const int co = n0 + c;                  // Cout is a multiple of 32 (host check): whole 32-column chunks
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co + i));
            v[i] = __uint_as_float(r[i]) + b4.x; v[i + 1] = __uint_as_float(r[i + 1]) + b4.y;
            v[i + 2] = __uint_as_float(r[i + 2]) + b4.z; v[i + 3] = __uint_as_float(r[i + 3]) + b4.w;
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            __nv_bfloat162 a0 = __floats2bfloat162_rn(v[i], v[i + 1]), a1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
            __nv_bfloat162 a2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]), a3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
            *reinterpret_cast<uint4*>(orow + co + i) = make_uint4(*reinterpret_cast<uint32_t*>(&a0), *reinterpret_cast<uint32_t*>(&a1),
                                                                  *reinterpret_cast<uint32_t*>(&a2), *reinterpret_cast<uint32_t*>(&a3));
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

// 4-D tiled tensor map over the channels-last map X [N, H, W, C] (bf16): box [1, 8, 16, 64 channels], 128-byte swizzle,
// out-of-bounds elements read as zero
int make_tmap_nhwc(CUtensorMap* map, const void* X, int N, int H, int W, int Cch) {
  EncodeTiledFn fn = get_encode_fn();
  NAWSOD_REQUIRE(fn != nullptr, NAWSOD_ERR_CUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t gdim[4] = {(cuuint64_t)Cch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstride[3] = {(cuuint64_t)Cch * 2, (cuuint64_t)W * Cch * 2, (cuuint64_t)H * W * Cch * 2};
  cuuint32_t box[4] = {64, kTileW, kTileH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(X), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NAWSOD_REQUIRE(r == CUDA_SUCCESS, NAWSOD_ERR_CUDA, "cuTensorMapEncodeTiled (4-d) failed with %d", (int)r);
  return NAWSOD_OK;
}

template <int BN>
int launch_conv(const void* X, const void* Wm, const ConvParams& p, cudaStream_t st) {
  using C = Cfg<BN, 2>;
  CUtensorMap tmX, tmW;
  if (int rc = make_tmap_nhwc(&tmX, X, p.N, p.H, p.W, p.Cin)) return rc;
  if (int rc = make_tmap(&tmW, Wm, 2, p.Cout, 9LL * p.Cin, 9LL * p.Cin, BN, C::BK, false, 1, 0)) return rc;
  auto kern = conv3x3_igemm_kernel<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    NAWSOD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int num_tiles = p.N * ((p.H + kTileH - 1) / kTileH) * ((p.W + kTileW - 1) / kTileW) * ((p.Cout + BN - 1) / BN);
  kern<<<std::min(num_tiles, sm_count()), kNumThreads, C::SMEM_BYTES, st>>>(tmX, tmW, p);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

int grid_for(int64_t total) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((total + kThreads - 1) / kThreads, 16LL * sm_count()));
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_conv3x3_relu(const void* X, int N, int H, int W, int Cin, const void* Wmat, const float* bias, int Cout,
                                   int dilation, int relu, void* Y, void* stream) {
  NAWSOD_REQUIRE(N >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, NAWSOD_ERR_SHAPE, "conv3x3_relu: bad shape N=%d H=%d W=%d Cin=%d Cout=%d", N, H, W, Cin, Cout);
  NAWSOD_REQUIRE(Cin % 64 == 0, NAWSOD_ERR_UNSUPPORTED, "conv3x3_relu: Cin=%d must be a multiple of 64 (one 128-byte channel block per k-step; "
                 "use nawsod_im2col3x3 + nawsod_fc_fwd otherwise)", Cin);
  NAWSOD_REQUIRE(Cout % 32 == 0, NAWSOD_ERR_UNSUPPORTED, "conv3x3_relu: Cout=%d must be a multiple of 32", Cout);
  NAWSOD_REQUIRE(dilation == 1 || dilation == 2, NAWSOD_ERR_UNSUPPORTED, "conv3x3_relu: dilation must be 1 or 2 (pad = dilation)");
  if (N == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(X && Wmat && bias && Y, NAWSOD_ERR_ARG, "conv3x3_relu: null pointer");
  NAWSOD_REQUIRE(aligned16(X) && aligned16(Wmat) && aligned16(bias) && aligned16(Y), NAWSOD_ERR_ALIGN, "conv3x3_relu: buffers must be 16-byte aligned");
  ConvParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.dil = dilation; p.relu = relu ? 1 : 0;
  p.bias = bias; p.Y = static_cast<__nv_bfloat16*>(Y);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // BN = 256 only when that still gives every SM a tile; small maps (conv5 at 1/8) take narrower tiles
  const long long mt = (long long)N * ((H + kTileH - 1) / kTileH) * ((W + kTileW - 1) / kTileW);
  if (Cout % 256 == 0 && mt * (Cout / 256) >= 2LL * sm_count()) return launch_conv<256>(X, Wmat, p, st);
  if (Cout % 128 == 0 && mt * (Cout / 128) >= sm_count()) return launch_conv<128>(X, Wmat, p, st);
  if (Cout % 64 == 0) return launch_conv<64>(X, Wmat, p, st);
  return launch_conv<128>(X, Wmat, p, st);        // Cout = 32 * odd: BN covers it with one guarded tile column
}

extern "C" int nawsod_im2col3x3(const void* X, int N, int H, int W, int C, int dilation, void* cols, void* stream) {
  NAWSOD_REQUIRE(N >= 0 && H > 0 && W > 0 && C > 0, NAWSOD_ERR_SHAPE, "im2col3x3: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  NAWSOD_REQUIRE(C % 8 == 0, NAWSOD_ERR_SHAPE, "im2col3x3: C=%d must be a multiple of 8 (16-byte bf16 vectors; pad the 3 image planes to 8)", C);
  NAWSOD_REQUIRE(dilation == 1 || dilation == 2, NAWSOD_ERR_UNSUPPORTED, "im2col3x3: dilation must be 1 or 2 (pad = dilation)");
  if (N == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(X && cols, NAWSOD_ERR_ARG, "im2col3x3: null pointer");
  NAWSOD_REQUIRE(aligned16(X) && aligned16(cols), NAWSOD_ERR_ALIGN, "im2col3x3: buffers must be 16-byte aligned");
  const int C8 = C / 8;
  const int64_t total = (int64_t)N * H * W * 9 * C8;
  im2col3x3_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(X), N, H, W, C8, dilation, static_cast<uint4*>(cols), total);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_maxpool2x2(const void* X, int N, int H, int W, int C, int stride, void* Y, void* stream) {
  NAWSOD_REQUIRE(N >= 0 && H >= 2 && W >= 2 && C > 0, NAWSOD_ERR_SHAPE, "maxpool2x2: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  NAWSOD_REQUIRE(C % 8 == 0, NAWSOD_ERR_SHAPE, "maxpool2x2: C=%d must be a multiple of 8", C);
  NAWSOD_REQUIRE(stride == 1 || stride == 2, NAWSOD_ERR_UNSUPPORTED, "maxpool2x2: stride must be 1 or 2");
  if (N == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(X && Y, NAWSOD_ERR_ARG, "maxpool2x2: null pointer");
  NAWSOD_REQUIRE(aligned16(X) && aligned16(Y), NAWSOD_ERR_ALIGN, "maxpool2x2: buffers must be 16-byte aligned");
  const int Ho = (H - 2) / stride + 1, Wo = (W - 2) / stride + 1;
  const int C8 = C / 8;
  const int64_t total = (int64_t)N * Ho * Wo * C8;
  maxpool2x2_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(X), H, W, C8, stride, Ho, Wo, static_cast<uint4*>(Y), total);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}
