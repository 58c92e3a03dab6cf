// Frozen VGG16 conv body in channels-last bf16 (row N4 of SURVEY.md section 8f; EXPERIMENTAL: compiled, not yet run).
//
// The reference builds conv1_1 .. conv5_3 with Caffe2 `Conv` (cuDNN, NCHW fp32) + `Relu` + `MaxPool`
// (detectron/modeling/VGG16.py:9-58, the flickr configs' MODEL.CONV_BODY) and freezes all of it
// (TRAIN.FREEZE_CONV_BODY, yaml:31), so the body is a forward-only producer of the head's conv5 map.  Here every 3x3
// convolution is the library's tcgen05 FC GEMM (gemm.cu: bias + ReLU fused in its epilogue, bf16 channels-last output =
// the next layer's input) over a patch matrix this file builds:
//
//   nawsod_im2col3x3   X [N,H,W,C] bf16  ->  cols [N*H*W, 9*C] bf16, K-order (kh, kw, c), stride 1,
//                      pad = dilation (the only combinations VGG16.py uses: pad 1 / dilation 1, pad 2 / dilation 2);
//                      out-of-image taps are zeros.  Pure 16-byte copies: HBM-bound, 2 * 9 * C bytes written per pixel.
//   nawsod_maxpool2x2  MaxPool(kernel=2, pad=0, stride=1|2) on a channels-last bf16 map (Caffe2's floor output size:
//                      (H - 2) / stride + 1); packed bf16x2 maxima, 16 bytes per thread.
//
// The weights [Cout, Cin, 3, 3] are permuted once on the host to [Cout, (kh, kw, c)] (conv_body.py).  A TMA-im2col
// implicit GEMM would remove the patch matrix's round trip through HBM (it is ~3x the layers' algorithmic bytes); this
// first version reuses the verified GEMM unchanged.
#include <algorithm>
#include "common.cuh"

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

int grid_for(int64_t total) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((total + kThreads - 1) / kThreads, 16LL * sm_count()));
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

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
