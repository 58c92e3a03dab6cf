// RoIPoolF (+ fused RoIFeatureBoost) forward and backward for sm_100a.
//
// Replaces Caffe2 RoIPoolF / RoIPoolFGradient as wired by detectron/modeling/detector.py:321-329
// and RoIFeatureBoost (detectron/ops/roi_feature_boost_op.cc:8-64).  The bin arithmetic follows
// detectron/ops/roi_loop_pool_op.cu:41-100 with RoIPoolF's deltas (stride-5 rois, no inner
// rectangle, maxval = empty ? 0 : -FLT_MAX) and is bit-exact in fp32.
//
// Forward design (HBM-write bound: 2 x 100 KB written per RoI, the map is L2 resident):
//   the map is channels-last; a CTA owns one (image, channel slab) and stages the whole
//   H*W x SC slab in shared memory once (16-byte cp.async), then walks a chunk of RoIs.  A
//   work item is (RoI, 16-byte channel vector): it scans each bin's window out of shared
//   memory with 16-byte loads in the reference's (h, w) order with strict '>' so the argmax
//   tie-break is identical, and writes 16-byte vectors of Y / argmax in pooled-NHWC order
//   ([R, PH, PW, C]).  Maps too large for shared memory take the same code path reading the
//   (L2-resident) map directly.
#include <algorithm>
#include <cfloat>
#include <type_traits>
#include "common.cuh"

namespace nawsod {
namespace {

// default of the pool_rows2 tuning knob: 0 = roi_pool_fwd_rows_kernel, 1 = roi_pool_fwd_rows2_kernel for maps staged in shared
// memory, 2 = also for maps read from L2 directly
constexpr long long kPoolRows2Default = 1;

struct PoolParams {
  const void* X;        // channels-last map [N, H, W, C]
  const float* rois;    // [R, 5]
  const float* boost;   // [R] or null
  int N, C, H, W, R, PH, PW;
  float scale;
  int SC;               // channels per slab (multiple of VEC, divides C)
  int rois_per_chunk;
  void* Y;              // [R, PH, PW, C]
  int32_t* argmax;      // same layout or null
};

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void unpack(const uint4& q, float (&v)[4]) {
    v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y);
    v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void unpack(const uint4& q, float (&v)[8]) {
    v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
    v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
    v[4] = __uint_as_float(q.z << 16); v[5] = __uint_as_float(q.z & 0xffff0000u);
    v[6] = __uint_as_float(q.w << 16); v[7] = __uint_as_float(q.w & 0xffff0000u);
  }
};

template <int VEC>
__device__ __forceinline__ void store_vals(float* dst, const float (&v)[VEC]) {
#pragma unroll
  for (int k = 0; k < VEC; k += 4)
    *reinterpret_cast<float4*>(dst + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}
template <int VEC>
__device__ __forceinline__ void store_vals(__nv_bfloat16* dst, const float (&v)[VEC]) {
  uint32_t w[VEC / 2];
#pragma unroll
  for (int k = 0; k < VEC; k += 2) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[k], v[k + 1]);
    w[k / 2] = *reinterpret_cast<uint32_t*>(&h);
  }
  if (VEC == 4) *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
  else *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[VEC / 2 - 2], w[VEC / 2 - 1]);
}
template <int VEC>
__device__ __forceinline__ void store_idx(int32_t* dst, const int (&v)[VEC]) {
#pragma unroll
  for (int k = 0; k < VEC; k += 4)
    *reinterpret_cast<int4*>(dst + k) = make_int4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}

// Per-bin scanners.  A lane owns VEC channels of one bin; the window is walked in the reference's
// (h, w) row-major order with strict '>' so ties resolve to the first cell (roi_loop_pool_op.cu:77-95).
template <bool kArgmax>
struct ScanF32 {   // fp32 map, VEC = 4
  float m[4]; int i[4];
  __device__ __forceinline__ void init(bool empty) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = empty ? 0.f : -FLT_MAX; i[k] = -1; }
  }
  // start of a bin: the cached row's scan when the bin shares that row, else the empty scan
  __device__ __forceinline__ void select(bool share, const ScanF32& c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = share ? c.m[k] : -FLT_MAX; i[k] = share ? c.i[k] : -1; }
  }
  __device__ __forceinline__ void visit(const uint4& q, int idx) {
    const float x[4] = {__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (kArgmax) { if (x[k] > m[k]) { m[k] = x[k]; i[k] = idx; } }
      else m[k] = fmaxf(m[k], x[k]);
    }
  }
  // fold a later scan segment in: it wins only where strictly greater, so ties keep the earlier cell
  __device__ __forceinline__ void merge(const ScanF32& o) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (kArgmax) { if (o.m[k] > m[k]) { m[k] = o.m[k]; i[k] = o.i[k]; } }
      else m[k] = fmaxf(m[k], o.m[k]);
    }
  }
  __device__ __forceinline__ void result(bool, float (&v)[4], int (&a)[4]) const {
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = m[k]; a[k] = i[k]; }
  }
};

// bf16 map, VEC = 8: packed bf16x2 compare/select (3 instructions per 2 elements with argmax,
// 1 per 2 without).  Cell indices are carried as packed u16 pairs (host guarantees H*W <= 65535),
// 0xFFFF = "never selected".  The running maximum starts at -inf: every finite bf16 beats it,
// exactly like -FLT_MAX in the fp32 reference.
template <bool kArgmax>
struct ScanBF16 {
  uint32_t m[4]; uint32_t i[4];
  __device__ __forceinline__ void init(bool empty) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = empty ? 0u : 0xFF80FF80u; i[k] = 0xFFFFFFFFu; }
  }
  __device__ __forceinline__ void select(bool share, const ScanBF16& c) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { m[k] = share ? c.m[k] : 0xFF80FF80u; i[k] = share ? c.i[k] : 0xFFFFFFFFu; }
  }
  __device__ __forceinline__ void merge(const ScanBF16& o) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(&o.m[k]);
      const __nv_bfloat162 mv = *reinterpret_cast<const __nv_bfloat162*>(&m[k]);
      if (kArgmax) {
        const uint32_t gt = __hgt2_mask(xv, mv);
        m[k] = (o.m[k] & gt) | (m[k] & ~gt);
        i[k] = (o.i[k] & gt) | (i[k] & ~gt);
      } else {
        const __nv_bfloat162 r = __hmax2(xv, mv);
        m[k] = *reinterpret_cast<const uint32_t*>(&r);
      }
    }
  }
  __device__ __forceinline__ void visit(const uint4& q, int idx) {
    const uint32_t x[4] = {q.x, q.y, q.z, q.w};
    const uint32_t idx2 = static_cast<uint32_t>(idx) * 0x00010001u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(&x[k]);
      const __nv_bfloat162 mv = *reinterpret_cast<const __nv_bfloat162*>(&m[k]);
      if (kArgmax) {
        const uint32_t gt = __hgt2_mask(xv, mv);                  // 0xFFFF per half where x > m (strict)
        m[k] = (x[k] & gt) | (m[k] & ~gt);
        i[k] = (idx2 & gt) | (i[k] & ~gt);
      } else {
        const __nv_bfloat162 r = __hmax2(xv, mv);
        m[k] = *reinterpret_cast<const uint32_t*>(&r);
      }
    }
  }
  __device__ __forceinline__ void result(bool, float (&v)[8], int (&a)[8]) const {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = __uint_as_float(m[k] << 16);            // an empty bin was initialised to +0
      v[2 * k + 1] = __uint_as_float(m[k] & 0xffff0000u);
      const uint32_t lo = i[k] & 0xFFFFu, hi = i[k] >> 16;
      a[2 * k] = (lo == 0xFFFFu) ? -1 : static_cast<int>(lo);
      a[2 * k + 1] = (hi == 0xFFFFu) ? -1 : static_cast<int>(hi);
    }
  }
};

// Forward.  grid = (channel slabs, RoI chunks, images).  One WARP owns one RoI at a time (fetched
// dynamically from the CTA's chunk); its 32 lanes are (bin slot, 16-byte channel vector) pairs, so
// all lanes share the RoI geometry and loop trip counts differ by at most one cell.
template <typename TIn, typename TOut, bool kSmem, bool kArgmax, bool k7x7>
__global__ void __launch_bounds__(1024, 1) roi_pool_fwd_kernel(const PoolParams p) {
  constexpr int VEC = Vec<TIn>::N;
  using Scan = typename std::conditional<sizeof(TIn) == 4, ScanF32<kArgmax>, ScanBF16<kArgmax>>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int next_roi;
  const int slab = blockIdx.x, chunk = blockIdx.y, n = blockIdx.z;
  const int HW = p.H * p.W;
  const int vpr = p.SC / VEC;          // lanes per bin slot (power of two <= 32)
  const int slots = 32 / vpr;          // bins handled concurrently by one warp
  const TIn* gbase = static_cast<const TIn*>(p.X) + (size_t)n * HW * p.C + (size_t)slab * p.SC;
  const int r0 = chunk * p.rois_per_chunk;
  const int r1 = min(p.R, r0 + p.rois_per_chunk);
  if (threadIdx.x == 0) next_roi = r0;

  const TIn* src;
  int row_stride;   // elements between consecutive cells
  if (kSmem) {
    TIn* s = reinterpret_cast<TIn*>(smem_raw);
    const int total = HW * vpr;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int cell = i / vpr, v = i - cell * vpr;
      cp_async16(s + (size_t)cell * p.SC + v * VEC, gbase + (size_t)cell * p.C + v * VEC);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    src = s;
    row_stride = p.SC;
  } else {
    src = gbase;
    row_stride = p.C;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int v = lane % vpr, slot = lane / vpr;
  const int W = p.W, H = p.H;
  const int PH = k7x7 ? 7 : p.PH, PW = k7x7 ? 7 : p.PW;
  const int bins = PH * PW;
  const TIn* vsrc = src + v * VEC;
  const int cell_bytes = row_stride * (int)sizeof(TIn);
  // Slot grid: slotsW slots across bin columns (a lane keeps ONE column pw when slotsW >= PW, so its
  // w-bounds are computed once per RoI), slotsH bin rows per round (the h-range is warp-uniform when 1).
  int slotsW = 1;
  while (slotsW < PW && slotsW < slots) slotsW <<= 1;
  const int slotsH = slots / slotsW;
  const int sw = slot % slotsW, sh = slot / slotsW;
  const float fPH = static_cast<float>(PH), fPW = static_cast<float>(PW);

  while (true) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&next_roi, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= r1) break;
    const float* roi = p.rois + (size_t)r * 5;
    if (static_cast<int>(__ldg(roi)) != n) continue;
    // detectron/ops/roi_loop_pool_op.cu:42-57
    const int roi_start_w = static_cast<int>(roundf(__ldg(roi + 1) * p.scale));
    const int roi_start_h = static_cast<int>(roundf(__ldg(roi + 2) * p.scale));
    const int roi_end_w = static_cast<int>(roundf(__ldg(roi + 3) * p.scale));
    const int roi_end_h = static_cast<int>(roundf(__ldg(roi + 4) * p.scale));
    const int roi_width = max(roi_end_w - roi_start_w + 1, 1);
    const int roi_height = max(roi_end_h - roi_start_h + 1, 1);
    const float bin_size_h = __fdiv_rn(static_cast<float>(roi_height), fPH);
    const float bin_size_w = __fdiv_rn(static_cast<float>(roi_width), fPW);
    const float s = p.boost ? __ldg(p.boost + r) : 1.0f;
    TOut* yrow = static_cast<TOut*>(p.Y) + ((size_t)r * bins * p.C + (size_t)slab * p.SC + v * VEC);
    int32_t* arow = kArgmax ? p.argmax + ((size_t)r * bins * p.C + (size_t)slab * p.SC + v * VEC) : nullptr;

#pragma unroll 1
    for (int ph = sh; ph < PH; ph += slotsH) {
      int hstart = static_cast<int>(floorf(__fmul_rn(static_cast<float>(ph), bin_size_h)));
      int hend = static_cast<int>(ceilf(__fmul_rn(static_cast<float>(ph + 1), bin_size_h)));
      hstart = min(max(hstart + roi_start_h, 0), H);
      hend = min(max(hend + roi_start_h, 0), H);
#pragma unroll 1
      for (int pw = sw; pw < PW; pw += slotsW) {
        int wstart = static_cast<int>(floorf(__fmul_rn(static_cast<float>(pw), bin_size_w)));
        int wend = static_cast<int>(ceilf(__fmul_rn(static_cast<float>(pw + 1), bin_size_w)));
        wstart = min(max(wstart + roi_start_w, 0), W);
        wend = min(max(wend + roi_start_w, 0), W);
        const bool is_empty = (hend <= hstart) || (wend <= wstart);
        Scan sc;
        sc.init(is_empty);
        if (!is_empty) {
          int idx0 = hstart * W + wstart;
          const unsigned char* rowp = reinterpret_cast<const unsigned char*>(vsrc) + (size_t)idx0 * cell_bytes;
          const int row_bytes = W * cell_bytes;
#pragma unroll 1
          for (int h = hstart; h < hend; ++h, idx0 += W, rowp += row_bytes) {
            const unsigned char* cp = rowp;
            int idx = idx0;
#pragma unroll 1
            for (int w = wstart; w < wend; ++w, ++idx, cp += cell_bytes)
              sc.visit(*reinterpret_cast<const uint4*>(cp), idx);
          }
        }
        float maxv[VEC];
        int maxi[VEC];
        sc.result(is_empty, maxv, maxi);
        if (p.boost) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) maxv[k] = __fmul_rn(maxv[k], s);
        }
        const int o = (ph * PW + pw) * p.C;
        store_vals<VEC>(yrow + o, maxv);
        if (kArgmax) store_idx<VEC>(arow + o, maxi);
      }
    }
  }
}

// Forward, bin-row variant (the default whenever one warp can hold a whole row of bins, PW <= 32 / vpr).
// Same CTA / slab / dynamic-RoI structure as above, but the per-bin index arithmetic is hoisted out of
// the bin loops, which the first kernel spent ~2/3 of its issue slots on (ncu: 79-83 % issue-active):
//   * RoI geometry is computed lane-parallel ONCE per RoI: lane l rounds coordinate l of the RoI, lanes
//     0..PH-1 evaluate the h-bounds of bin row l and lanes 8.. the w-bounds of bin column l-8 (one fp32
//     division each), and every lane then pulls what it needs with shuffles;
//   * a lane keeps ONE bin column for the whole RoI (w-range, shared-memory column offset and output
//     pointer are per-RoI constants); the warp walks the PH bin rows together, so the h-loop is uniform;
//   * the w-loop is unrolled by two with a clamped second load (a duplicate visit never wins a strict '>').
// The scan order per bin is still (h, w) row-major with strict '>', so values and argmax stay bit-exact.
//   * (kRowCache) consecutive bin rows overlap by one map row whenever (ph+1)*bin_size_h is not an integer
//     (hend(ph) = ceil, hstart(ph+1) = floor), and bins of RoIs shorter than PH cells repeat the same row: the
//     lane keeps the scan result of the last map row it walked (its own w-range of that row: first maximum and
//     its cell index) and folds it into the next bin instead of re-reading the row.  Folding whole rows in
//     increasing h with strict '>' selects the same (first) cell as the reference's flat (h, w) scan, so values
//     and argmax are unchanged; shared-memory reads and issue slots drop by the overlap factor (~1.7x for the
//     10 x 20-cell windows of the 38 x 50 map).
template <typename TIn, typename TOut, bool kSmem, bool kArgmax, bool kRowCache>
__global__ void __launch_bounds__(1024, 1) roi_pool_fwd_rows_kernel(const PoolParams p) {
  constexpr int VEC = Vec<TIn>::N;
  using Scan = typename std::conditional<sizeof(TIn) == 4, ScanF32<kArgmax>, ScanBF16<kArgmax>>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int next_roi;
  const int slab = blockIdx.x, chunk = blockIdx.y, n = blockIdx.z;
  const int HW = p.H * p.W;
  const int vpr = p.SC / VEC;          // lanes per bin column (power of two, 32 / vpr >= PW)
  const int vshift = __ffs(vpr) - 1;
  const TIn* gbase = static_cast<const TIn*>(p.X) + (size_t)n * HW * p.C + (size_t)slab * p.SC;
  const int r0 = chunk * p.rois_per_chunk;
  const int r1 = min(p.R, r0 + p.rois_per_chunk);
  if (threadIdx.x == 0) next_roi = r0;

  const unsigned char* src;
  int cell_bytes;
  if (kSmem) {
    const int total = HW * vpr;                 // 16-byte vectors in the slab
    const unsigned char* g = reinterpret_cast<const unsigned char*>(gbase);
    const size_t gcell = (size_t)p.C * sizeof(TIn);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int cell = i >> vshift, v = i & (vpr - 1);
      cp_async16(smem_raw + (size_t)i * 16, g + (size_t)cell * gcell + v * 16);
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    src = smem_raw;
    cell_bytes = p.SC * (int)sizeof(TIn);
  } else {
    src = reinterpret_cast<const unsigned char*>(gbase);
    cell_bytes = p.C * (int)sizeof(TIn);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int v = lane & (vpr - 1), sw = lane >> vshift;   // sw = this lane's bin column (idle when >= PW)
  const int W = p.W, H = p.H, PH = p.PH, PW = p.PW;
  const bool col_ok = sw < PW;
  const int row_bytes = W * cell_bytes;
  const unsigned char* lane_src = src + v * 16;
  // lane-parallel bound evaluation: lanes [0, 8) -> bin row `lane` (h), lanes [8, 32) -> bin column `lane - 8` (w)
  const bool is_h = lane < 8;
  const int pidx = is_h ? lane : lane - 8;
  const float fdiv = static_cast<float>(is_h ? PH : PW);
  const int lim = is_h ? H : W;
  const size_t bin_stride = (size_t)PW * p.C;    // elements between consecutive bin rows of Y

  while (true) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&next_roi, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= r1) break;
    // lane l (1..4) owns coordinate l of the RoI; detectron/ops/roi_loop_pool_op.cu:42-45
    const float coord = __ldg(p.rois + (size_t)r * 5 + min(lane, 4));
    if (static_cast<int>(__shfl_sync(0xffffffffu, coord, 0)) != n) continue;
    const int rounded = static_cast<int>(roundf(coord * p.scale));
    const int roi_start_w = __shfl_sync(0xffffffffu, rounded, 1);
    const int roi_start_h = __shfl_sync(0xffffffffu, rounded, 2);
    const int roi_end_w = __shfl_sync(0xffffffffu, rounded, 3);
    const int roi_end_h = __shfl_sync(0xffffffffu, rounded, 4);
    // roi_loop_pool_op.cu:54-68, one (row or column) bound pair per lane
    const int extent = is_h ? max(roi_end_h - roi_start_h + 1, 1) : max(roi_end_w - roi_start_w + 1, 1);
    const int offs = is_h ? roi_start_h : roi_start_w;
    const float bin_size = __fdiv_rn(static_cast<float>(extent), fdiv);
    int bstart = static_cast<int>(floorf(__fmul_rn(static_cast<float>(pidx), bin_size)));
    int bend = static_cast<int>(ceilf(__fmul_rn(static_cast<float>(pidx + 1), bin_size)));
    bstart = min(max(bstart + offs, 0), lim);
    bend = min(max(bend + offs, 0), lim);
    const int wstart = __shfl_sync(0xffffffffu, bstart, 8 + min(sw, 23));
    const int wend = __shfl_sync(0xffffffffu, bend, 8 + min(sw, 23));
    const int ncols = wend - wstart;               // cells per row of this lane's bins (<= 0: empty column)
    const float s = p.boost ? __ldg(p.boost + r) : 1.0f;
    const unsigned char* col_src = lane_src + (size_t)wstart * cell_bytes;
    const size_t out_off = ((size_t)r * PH * PW + sw) * p.C + (size_t)slab * p.SC + v * VEC;
    TOut* yout = static_cast<TOut*>(p.Y) + out_off;
    int32_t* aout = kArgmax ? p.argmax + out_off : nullptr;

    int cached_h = -1;                             // map row whose scan `cached` holds (kRowCache)
    Scan cached;
    cached.init(false);
#pragma unroll 1
    for (int ph = 0; ph < PH; ++ph, yout += bin_stride, aout += (kArgmax ? bin_stride : 0)) {
      const int hstart = __shfl_sync(0xffffffffu, bstart, ph);
      const int hend = __shfl_sync(0xffffffffu, bend, ph);
      if (!col_ok) continue;                       // idle lanes only take part in the shuffles
      const bool is_empty = (hend <= hstart) || (ncols <= 0);
      Scan sc;
      sc.init(is_empty);
      if (!is_empty) {
        const unsigned char* rowp = col_src + (size_t)hstart * row_bytes;
        int idx0 = hstart * W + wstart;
#pragma unroll 1
        for (int h = hstart; h < hend; ++h, rowp += row_bytes, idx0 += W) {
          if (kRowCache) {
            if (h != cached_h) {                   // uniform over the lanes that scan (h-ranges are per warp)
              cached.init(false);
              const unsigned char* cp = rowp;
              int idx = idx0;
#pragma unroll 1
              for (int j = 0; j < ncols; j += 2, cp += 2 * cell_bytes, idx += 2) {
                cached.visit(*reinterpret_cast<const uint4*>(cp), idx);
                const int more = (j + 1 < ncols) ? 1 : 0;      // clamped second cell: re-visiting a cell never wins '>'
                cached.visit(*reinterpret_cast<const uint4*>(cp + (more ? cell_bytes : 0)), idx + more);
              }
              cached_h = h;
            }
            sc.merge(cached);
          } else {
            const unsigned char* cp = rowp;
            int idx = idx0;
#pragma unroll 1
            for (int j = 0; j < ncols; j += 2, cp += 2 * cell_bytes, idx += 2) {
              sc.visit(*reinterpret_cast<const uint4*>(cp), idx);
              const int more = (j + 1 < ncols) ? 1 : 0;
              sc.visit(*reinterpret_cast<const uint4*>(cp + (more ? cell_bytes : 0)), idx + more);
            }
          }
        }
      }
      float maxv[VEC];
      int maxi[VEC];
      sc.result(is_empty, maxv, maxi);
      if (p.boost) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) maxv[k] = __fmul_rn(maxv[k], s);
      }
      store_vals<VEC>(yout, maxv);
      if (kArgmax) store_idx<VEC>(aout, maxi);
    }
  }
}

// Forward, bin-row variant 2 ("rows2"): the same mapping and scan order as roi_pool_fwd_rows_kernel (a warp owns one
// RoI, a lane one bin column x one 16-byte channel vector, bin rows walked together, last map row cached), with the
// control overhead that kernel spent ~85 % of its issue slots on (ncu: 81 % issue-active, 16 K warp instructions per
// RoI of which ~2 K are loads and compares) removed:
//   * four lanes per bin column, always (SC = 4 vectors), so lane -> (column, vector) is two bit operations;
//   * the w-scan has NO per-lane loop: its trip count is the warp-uniform maximum NC of the column widths, the cells
//     beyond a lane's own width are re-reads of its last cell (a re-visited cell never wins the strict '>' and
//     never changes a maximum) at offsets computed once per RoI, and the whole bin loop is instantiated for
//     NC = 1..4 (straight-line loads, then compares) with one generic instance for wider bins;
//   * the map-row pointer runs on across bin rows: consecutive bins either share exactly one row (the cached one:
//     hend(ph) = ceil, hstart(ph+1) = floor of the same product) or continue with the next row, so a bin costs no
//     address arithmetic and the cached row is folded in by a register move before the row loop;
//   * the idle lanes (columns >= PW) shadow the last column instead of branching around the scan, only their stores
//     write the same values to the same addresses; RoIs with an empty bin (clipped by the map border) take a
//     separate instance of the bin loop, chosen by one vote per RoI, so the common instance has no emptiness tests.
// Scan order per bin is still (h, w) row-major with strict '>': values and argmax are bit-identical.
template <typename Scan, int NC>
__device__ __forceinline__ void rows2_scan_row(Scan& c, const unsigned char* rowp, int idx0, int o1, int o2, int o3,
                                               int ncmax, int last, int cell_bytes) {
  const uint4 q0 = *reinterpret_cast<const uint4*>(rowp);
  uint4 q1 = q0, q2 = q0, q3 = q0;
  if (NC == 0 || NC >= 2) q1 = *reinterpret_cast<const uint4*>(rowp + o1);
  if (NC == 0 || NC >= 3) q2 = *reinterpret_cast<const uint4*>(rowp + o2);
  if (NC == 0 || NC >= 4) q3 = *reinterpret_cast<const uint4*>(rowp + o3);
  c.visit(q0, idx0);
  if (NC == 0 || NC >= 2) c.visit(q1, idx0 + 1);
  if (NC == 0 || NC >= 3) c.visit(q2, idx0 + 2);
  if (NC == 0 || NC >= 4) c.visit(q3, idx0 + 3);
  if (NC == 0) {
#pragma unroll 1
    for (int j = 4; j < ncmax; ++j)
      c.visit(*reinterpret_cast<const uint4*>(rowp + min(j, last) * cell_bytes), idx0 + j);
  }
}

struct Rows2Roi {            // per-RoI, per-lane constants of the bin loop
  const unsigned char* col_src;   // first cell of the lane's w-range in map row 0 (clamped into the row)
  int wstart, ncols, ncmax, last, o1, o2, o3, cell_bytes, row_bytes, W, PH;
  int bstart, bend;               // lane l < 8: h-bounds of bin row l
  unsigned share_mask;            // bit ph: bin row ph starts on the last map row of bin row ph - 1
  float s;                        // boost factor (1 when there is none: x * 1 is exact)
  size_t bin_stride;
};

// kEmpty = false: every bin of the RoI is non-empty (the common case, decided by one vote per RoI).  Then bin row
// ph + 1 starts either on the last row of bin row ph (share_mask) or right below it, so the row pointer simply runs on.
// kEmpty = true: RoIs clipped by the map border or degenerate; bins are tested one by one.
template <typename TIn, typename TOut, bool kArgmax, int NC, bool kEmpty>
__device__ __forceinline__ void rows2_bins(const Rows2Roi& g, TOut* yout, int32_t* aout) {
  constexpr int VEC = Vec<TIn>::N;
  using Scan = typename std::conditional<sizeof(TIn) == 4, ScanF32<kArgmax>, ScanBF16<kArgmax>>::type;
  int next_h = kEmpty ? -8 : __shfl_sync(0xffffffffu, g.bstart, 0);   // map row `rowp` points at; `cached` = row next_h - 1
  const unsigned char* rowp = g.col_src + (kEmpty ? (size_t)0 : (size_t)next_h * g.row_bytes);
  int idx0 = kEmpty ? 0 : next_h * g.W + g.wstart;
  Scan cached;
  cached.init(false);
#pragma unroll 1
  for (int ph = 0; ph < g.PH; ++ph, yout += g.bin_stride, aout += (kArgmax ? g.bin_stride : 0)) {
    const int hend = __shfl_sync(0xffffffffu, g.bend, ph);
    Scan sc;
    bool is_empty = false, row_empty = false;   // row_empty: warp-uniform (no map row in this bin row)
    if (!kEmpty) {
      sc.select((g.share_mask >> ph) & 1u, cached);
    } else {
      const int hstart = __shfl_sync(0xffffffffu, g.bstart, ph);
      row_empty = hend <= hstart;
      is_empty = row_empty || (g.ncols <= 0);
      if (!row_empty) {                         // warp-uniform
        sc.select(hstart == next_h - 1, cached);
        if (hstart != next_h - 1 && hstart != next_h) {          // first non-empty bin of the RoI
          next_h = hstart;
          rowp = g.col_src + (size_t)hstart * g.row_bytes;
          idx0 = hstart * g.W + g.wstart;
        }
      } else {
        sc.init(true);
      }
    }
    const int hstop = (kEmpty && row_empty) ? next_h : hend;
#pragma unroll 1
    for (; next_h < hstop; ++next_h, rowp += g.row_bytes, idx0 += g.W) {
      cached.init(false);
      rows2_scan_row<Scan, NC>(cached, rowp, idx0, g.o1, g.o2, g.o3, g.ncmax, g.last, g.cell_bytes);
      sc.merge(cached);
    }
    if (kEmpty && is_empty) sc.init(true);
    float maxv[VEC];
    int maxi[VEC];
    sc.result(is_empty, maxv, maxi);
#pragma unroll
    for (int k = 0; k < VEC; ++k) maxv[k] = __fmul_rn(maxv[k], g.s);
    // lanes of the idle columns (>= PW) shadow the last column: they store the same values to the same addresses
    store_vals<VEC>(yout, maxv);
    if (kArgmax) store_idx<VEC>(aout, maxi);
  }
}

// kSkipIdle (staged maps; tuning knob pool_skip_idle, default 1): RoIs arrive grouped by image, so of the N CTAs that share a
// (slab, chunk) usually one finds work; with kSkipIdle the others return BEFORE staging 120 KB of map (at N = 2 a third of
// the launch's CTA time; measured 79.9 -> 73.7 us bf16, 139.3 -> 131.1 us bf16 + argmax, profiles/r2a_microbench_pool.log).
//
// Measured and REMOVED in round 2 (profiles/r2a_microbench_pool.log, r2e / r2g_microbench_pool_*.log, r2g_ncu_pool_rows4_summary.txt):
// a one-RoI-ahead prefetch of the RoI coordinates (no change), the seven bin rows unrolled (slower: instruction cache), and a
// rebuilt bf16 / no-argmax kernel ("rows4": 32-bit shared addresses, three-input packed maxima, no -inf fill or selects,
// packed-bf16 FMA boost, then a ping-pong prefetch of the next map row, then 4- / 5-cell specialisations).  Its three builds
// ran 42.8-46.1 M warp instructions with very different schedules and all took 73-78 us -- the same as this kernel.  The bound
// they share is the LSU / shared-memory pipe: 11.7 M shared wavefronts (2.1 M of them bank conflicts between the two bin columns
// of a quarter-warp, whose cell parities are data dependent) + the output stores keep l1tex at 72 % of peak over the whole launch
// and ~85 % while the SMs are active; reading each RoI's window out of a 64-byte-per-cell slab costs ~6 bytes of shared-memory
// traffic per byte written (bin columns overlap by a cell, windows are clamped to the widest column of the warp).
template <typename TIn, typename TOut, bool kSmem, bool kArgmax, bool kSkipIdle = false>
__global__ void __launch_bounds__(1024, 1) roi_pool_fwd_rows2_kernel(const PoolParams p) {
  constexpr int VEC = Vec<TIn>::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int next_roi;
  const int slab = blockIdx.x, chunk = blockIdx.y, n = blockIdx.z;
  const int HW = p.H * p.W;
  const TIn* gbase = static_cast<const TIn*>(p.X) + (size_t)n * HW * p.C + (size_t)slab * (4 * VEC);
  const int r0 = chunk * p.rois_per_chunk;
  const int r1 = min(p.R, r0 + p.rois_per_chunk);
  if (threadIdx.x == 0) next_roi = r0;
  if (kSkipIdle) {
    int mine = 0;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) mine |= (static_cast<int>(__ldg(p.rois + (size_t)r * 5)) == n) ? 1 : 0;
    if (!__syncthreads_or(mine)) return;
  }

  Rows2Roi g;
  const unsigned char* src;
  if (kSmem) {
    const int total = HW * 4;                   // 16-byte vectors in the slab (64 bytes per cell)
    const unsigned char* gsrc = reinterpret_cast<const unsigned char*>(gbase);
    const size_t gcell = (size_t)p.C * sizeof(TIn);
    for (int i = threadIdx.x; i < total; i += blockDim.x)
      cp_async16(smem_raw + (size_t)i * 16, gsrc + (size_t)(i >> 2) * gcell + (i & 3) * 16);
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    src = smem_raw;
    g.cell_bytes = 64;
  } else {
    src = reinterpret_cast<const unsigned char*>(gbase);
    g.cell_bytes = p.C * (int)sizeof(TIn);
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int v = lane & 3;
  const int W = p.W, H = p.H, PH = p.PH, PW = p.PW;
  const int sw = min(lane >> 2, PW - 1);         // idle lanes shadow the last bin column
  g.W = W; g.PH = PH;
  g.row_bytes = W * g.cell_bytes;
  g.bin_stride = (size_t)PW * p.C;               // elements between consecutive bin rows of Y
  const unsigned char* lane_src = src + v * 16;
  // lane-parallel bound evaluation: lanes [0, 8) -> bin row `lane` (h), lanes [8, 32) -> bin column `lane - 8` (w)
  const bool is_h = lane < 8;
  const int pidx = is_h ? lane : lane - 8;
  const float fdiv = static_cast<float>(is_h ? PH : PW);
  const int lim = is_h ? H : W;

  while (true) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&next_roi, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= r1) break;
    // lane l (1..4) owns coordinate l of the RoI; detectron/ops/roi_loop_pool_op.cu:42-45
    const float coord = __ldg(p.rois + (size_t)r * 5 + min(lane, 4));
    if (static_cast<int>(__shfl_sync(0xffffffffu, coord, 0)) != n) continue;
    const int rounded = static_cast<int>(roundf(coord * p.scale));
    const int roi_start_w = __shfl_sync(0xffffffffu, rounded, 1);
    const int roi_start_h = __shfl_sync(0xffffffffu, rounded, 2);
    const int roi_end_w = __shfl_sync(0xffffffffu, rounded, 3);
    const int roi_end_h = __shfl_sync(0xffffffffu, rounded, 4);
    // roi_loop_pool_op.cu:54-68, one (row or column) bound pair per lane
    const int extent = is_h ? max(roi_end_h - roi_start_h + 1, 1) : max(roi_end_w - roi_start_w + 1, 1);
    const int offs = is_h ? roi_start_h : roi_start_w;
    const float bin_size = __fdiv_rn(static_cast<float>(extent), fdiv);
    const int bstart = static_cast<int>(floorf(__fmul_rn(static_cast<float>(pidx), bin_size)));
    const int bend = static_cast<int>(ceilf(__fmul_rn(static_cast<float>(pidx + 1), bin_size)));
    g.bstart = min(max(bstart + offs, 0), lim);
    g.bend = min(max(bend + offs, 0), lim);
    g.wstart = __shfl_sync(0xffffffffu, g.bstart, 8 + sw);
    const int wend = __shfl_sync(0xffffffffu, g.bend, 8 + sw);
    g.ncols = wend - g.wstart;                     // cells per row of this lane's bins (<= 0: empty column)
    g.ncmax = __reduce_max_sync(0xffffffffu, g.ncols);     // warp-uniform trip count of the w-scan
    g.last = max(g.ncols - 1, 0);
    // byte offsets of cells 1..3 of the lane's w-range, clamped to its last cell
    g.o1 = min(1, g.last) * g.cell_bytes; g.o2 = min(2, g.last) * g.cell_bytes; g.o3 = min(3, g.last) * g.cell_bytes;
    g.s = p.boost ? __ldg(p.boost + r) : 1.0f;
    // an empty column may start at W: keep its (discarded) reads inside the row
    g.col_src = lane_src + (size_t)min(g.wstart, W - 1) * g.cell_bytes;
    const size_t out_off = ((size_t)r * PH * PW + sw) * p.C + (size_t)slab * (4 * VEC) + v * VEC;
    TOut* yout = static_cast<TOut*>(p.Y) + out_off;
    int32_t* aout = kArgmax ? p.argmax + out_off : nullptr;
    // bin row ph starts on the last map row of bin row ph - 1 (lanes < 8 hold the h-bounds)
    const int prev_end = __shfl_up_sync(0xffffffffu, g.bend, 1);
    g.share_mask = __ballot_sync(0xffffffffu, lane > 0 && lane < PH && g.bstart == prev_end - 1);
    const bool bound_empty = (is_h ? lane < PH : lane - 8 < PW) && g.bend <= g.bstart;
    const bool any_empty = __any_sync(0xffffffffu, bound_empty);
#define NAWSOD_ROWS2(NC_)                                                        \
  do {                                                                           \
    if (any_empty) rows2_bins<TIn, TOut, kArgmax, NC_, true>(g, yout, aout);     \
    else rows2_bins<TIn, TOut, kArgmax, NC_, false>(g, yout, aout);              \
  } while (0)
    switch (g.ncmax) {                             // warp-uniform
      case 2: NAWSOD_ROWS2(2); break;
      case 3: NAWSOD_ROWS2(3); break;
      case 4: NAWSOD_ROWS2(4); break;
      default:
        if (g.ncmax <= 1) NAWSOD_ROWS2(1);         // <= 0: every column empty, reads stay in bounds
        else NAWSOD_ROWS2(0);
    }
#undef NAWSOD_ROWS2
  }
}

// ---------------------------------------------------------------------------------------------
// Backward.  NHWC: dX[b, argmax, c] += dY[r, bin, c] (lanes = channels: coalesced reads, one
// red per element).  NCHW: lanes = consecutive (c, bin): neighbouring bins often share their
// argmax cell, so lanes with the same target address are combined with __match_any_sync and
// one lane issues the red ("warp-aggregated atomics keyed on the argmax").
// ---------------------------------------------------------------------------------------------
template <typename TDy>
__device__ __forceinline__ float load_dy(const TDy* p, size_t i);
template <> __device__ __forceinline__ float load_dy<float>(const float* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ float load_dy<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) {
  return __bfloat162float(p[i]);
}

template <typename TDy>
__global__ void __launch_bounds__(256) roi_pool_bwd_nhwc_kernel(const TDy* __restrict__ dY,
                                                               const int32_t* __restrict__ argmax,
                                                               const float* __restrict__ rois,
                                                               const float* __restrict__ boost, int C, int HW,
                                                               int bins, size_t total, float* __restrict__ dX) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int a = argmax[i];
    if (a < 0) continue;
    const int c = static_cast<int>(i % C);
    const size_t r = i / ((size_t)C * bins);
    const int b = static_cast<int>(rois[r * 5]);
    float g = load_dy<TDy>(dY, i);
    if (boost) g *= boost[r];
    atomicAdd(dX + ((size_t)b * HW + a) * C + c, g);
  }
}

template <typename TDy>
__global__ void __launch_bounds__(256) roi_pool_bwd_nchw_kernel(const TDy* __restrict__ dY,
                                                               const int32_t* __restrict__ argmax,
                                                               const float* __restrict__ rois,
                                                               const float* __restrict__ boost, int C, int HW,
                                                               int bins, size_t total, float* __restrict__ dX) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // total is padded by the caller's loop bound so that whole warps stay converged
  const size_t padded = (total + 31) / 32 * 32;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += stride) {
    long long key = -1;
    float g = 0.f;
    if (i < total) {
      const int a = argmax[i];
      if (a >= 0) {
        const size_t rc = i / bins;               // r * C + c
        const size_t r = rc / C;
        const int c = static_cast<int>(rc - r * C);
        const int b = static_cast<int>(rois[r * 5]);
        key = ((long long)b * C + c) * HW + a;
        g = load_dy<TDy>(dY, i);
        if (boost) g *= boost[r];
      }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    if (key >= 0) {
      const int lane = threadIdx.x & 31;
      const int leader = __ffs(peers) - 1;
      float sum = 0.f;
      // fixed lane order -> deterministic partial sums inside the warp
      for (unsigned m = peers; m; m &= m - 1) {
        const int src = __ffs(m) - 1;
        sum += __shfl_sync(peers, g, src);
      }
      if (lane == leader) atomicAdd(dX + key, sum);
    }
  }
}

__global__ void boost_kernel(const float* __restrict__ X, const float* __restrict__ S, int64_t F4, int64_t total4,
                             float* __restrict__ Y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = S[i / F4];
    float4 v = reinterpret_cast<const float4*>(X)[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    reinterpret_cast<float4*>(Y)[i] = v;
  }
}
__global__ void boost_kernel_scalar(const float* __restrict__ X, const float* __restrict__ S, int64_t F, int64_t total,
                                    float* __restrict__ Y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    Y[i] = X[i] * S[i / F];
}

// Batched transpose of 4-byte elements: in [B, rows, cols] -> out [B, cols, rows].
template <typename TOut>
__global__ void transpose_kernel(const uint32_t* __restrict__ in, int64_t rows, int64_t cols, TOut* __restrict__ out) {
  __shared__ uint32_t tile[32][33];
  const int64_t b = blockIdx.z;
  const uint32_t* src = in + b * rows * cols;
  TOut* dst = out + b * rows * cols;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int64_t c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) {
      const uint32_t w = tile[threadIdx.x][j];
      if (sizeof(TOut) == 4) reinterpret_cast<uint32_t*>(dst)[c * rows + r] = w;
      else reinterpret_cast<__nv_bfloat16*>(dst)[c * rows + r] = __float2bfloat16_rn(__uint_as_float(w));
    }
  }
}

template <typename TIn, typename TOut, bool kSmem, bool kArgmax, bool k7x7>
int launch_pool_fwd3(const PoolParams& p, size_t smem_bytes, dim3 grid, int threads, cudaStream_t st) {
  auto k = roi_pool_fwd_kernel<TIn, TOut, kSmem, kArgmax, k7x7>;
  if (kSmem) NAWSOD_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  k<<<grid, threads, smem_bytes, st>>>(p);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

template <typename TIn, typename TOut, bool kSmem, bool kArgmax>
int launch_pool_fwd2(const PoolParams& p, size_t smem_bytes, dim3 grid, int threads, cudaStream_t st) {
  // bin-row kernel: one warp holds a whole row of bins (h-bounds on lanes 0..7, w-bounds on lanes 8..31)
  const int slots = 32 / (p.SC / Vec<TIn>::N);
  if (p.PH <= 8 && p.PW <= 8 && slots == 8 && get_tuning("pool_generic", 0) == 0 && get_tuning("pool_rows2", kPoolRows2Default) != 0) {
    if constexpr (kSmem) {
      if (get_tuning("pool_skip_idle", 1) != 0) {
        auto k = roi_pool_fwd_rows2_kernel<TIn, TOut, true, kArgmax, true>;
        NAWSOD_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        k<<<grid, threads, smem_bytes, st>>>(p);
        NAWSOD_LAUNCH_OK();
        return NAWSOD_OK;
      }
    }
    auto k = roi_pool_fwd_rows2_kernel<TIn, TOut, kSmem, kArgmax>;
    if (kSmem) NAWSOD_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    k<<<grid, threads, smem_bytes, st>>>(p);
    NAWSOD_LAUNCH_OK();
    return NAWSOD_OK;
  }
  if (p.PH <= 8 && p.PW <= slots && slots <= 8 && get_tuning("pool_generic", 0) == 0) {
    if (get_tuning("pool_rowcache", 1) != 0) {
      auto k = roi_pool_fwd_rows_kernel<TIn, TOut, kSmem, kArgmax, true>;
      if (kSmem) NAWSOD_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      k<<<grid, threads, smem_bytes, st>>>(p);
    } else {
      auto k = roi_pool_fwd_rows_kernel<TIn, TOut, kSmem, kArgmax, false>;
      if (kSmem) NAWSOD_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      k<<<grid, threads, smem_bytes, st>>>(p);
    }
    NAWSOD_LAUNCH_OK();
    return NAWSOD_OK;
  }
  if (p.PH == 7 && p.PW == 7) return launch_pool_fwd3<TIn, TOut, kSmem, kArgmax, true>(p, smem_bytes, grid, threads, st);
  return launch_pool_fwd3<TIn, TOut, kSmem, kArgmax, false>(p, smem_bytes, grid, threads, st);
}

template <typename TIn, typename TOut>
int launch_pool_fwd(const PoolParams& p, bool use_smem, size_t smem_bytes, dim3 grid, int threads, cudaStream_t st) {
  const bool am = p.argmax != nullptr;
  if (use_smem) return am ? launch_pool_fwd2<TIn, TOut, true, true>(p, smem_bytes, grid, threads, st)
                          : launch_pool_fwd2<TIn, TOut, true, false>(p, smem_bytes, grid, threads, st);
  return am ? launch_pool_fwd2<TIn, TOut, false, true>(p, 0, grid, threads, st)
            : launch_pool_fwd2<TIn, TOut, false, false>(p, 0, grid, threads, st);
}

}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_transpose_batched(const void* in, int64_t B, int64_t rows, int64_t cols, void* out,
                                        int out_dtype, void* stream) {
  NAWSOD_REQUIRE(B >= 0 && rows >= 0 && cols >= 0, NAWSOD_ERR_SHAPE, "transpose: negative size");
  NAWSOD_REQUIRE(out_dtype == NAWSOD_F32 || out_dtype == NAWSOD_BF16, NAWSOD_ERR_ARG, "transpose: bad out_dtype");
  if (B == 0 || rows == 0 || cols == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(in && out, NAWSOD_ERR_ARG, "transpose: null pointer");
  NAWSOD_REQUIRE(B <= 65535, NAWSOD_ERR_SHAPE, "transpose: batch %lld > 65535", (long long)B);
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)B), block(32, 8);
  NAWSOD_REQUIRE(grid.y <= 65535, NAWSOD_ERR_SHAPE, "transpose: too many rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == NAWSOD_F32)
    transpose_kernel<uint32_t><<<grid, block, 0, st>>>(static_cast<const uint32_t*>(in), rows, cols,
                                                       static_cast<uint32_t*>(out));
  else
    transpose_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(static_cast<const uint32_t*>(in), rows, cols,
                                                            static_cast<__nv_bfloat16*>(out));
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

// Forward on a channels-last map into pooled-NHWC output (the native fast path).
static int pool_fwd_nhwc(const void* X, int x_dtype, const float* rois, const float* boost, int N, int C, int H, int W,
                         int R, float scale, int PH, int PW, void* Y, int y_dtype, int32_t* argmax, cudaStream_t st) {
  const int esize = x_dtype == NAWSOD_F32 ? 4 : 2;
  const int VEC = 16 / esize;
  NAWSOD_REQUIRE(C % VEC == 0, NAWSOD_ERR_SHAPE, "roi_pool_f: C=%d must be a multiple of %d", C, VEC);
  NAWSOD_REQUIRE(aligned16(X) && aligned16(Y) && (!argmax || aligned16(argmax)), NAWSOD_ERR_ALIGN,
                 "roi_pool_f: X, Y and argmax must be 16-byte aligned");
  NAWSOD_REQUIRE(N <= 65535, NAWSOD_ERR_SHAPE, "roi_pool_f: N too large");
  NAWSOD_REQUIRE(x_dtype == NAWSOD_F32 || (int64_t)H * W <= 65535, NAWSOD_ERR_UNSUPPORTED,
                 "roi_pool_f: a bf16 map is limited to H*W <= 65535 cells (packed 16-bit argmax)");
  // Channel slab SC = VEC * 2^k lanes-per-bin (<= 32 lanes), dividing C.  Prefer the largest slab whose
  // H*W*SC tile fits the shared-memory budget; otherwise read the (L2-resident) map directly with a
  // whole warp of channel vectors per bin.
  const int64_t budget = std::min<int64_t>(get_tuning("pool_slab_bytes", 200 * 1024), 200 * 1024);
  const bool force_global = get_tuning("pool_force_global", 0) != 0;
  // the rows2 kernel maps exactly four lanes to a bin column: prefer the 4-vector slab wherever it applies, also for
  // small maps whose wider slabs would fit (those fall back to the slower generic kernel)
  // (pool_rows2 = 2 also takes it for maps that are read from L2 directly)
  const int rows2 = (get_tuning("pool_generic", 0) == 0 && PH <= 8 && PW <= 8) ? (int)get_tuning("pool_rows2", kPoolRows2Default) : 0;
  int SC = 0;
  if (!force_global)
    for (int lanes = rows2 ? 4 : 32; lanes >= 1; lanes >>= 1) {
      const int cand = lanes * VEC;
      if (cand > C || C % cand) continue;
      if ((int64_t)H * W * cand * esize <= budget) { SC = cand; break; }
    }
  const bool use_smem = SC > 0;
  if (!use_smem)
    for (int lanes = rows2 == 2 ? 4 : 32; lanes >= 1; lanes >>= 1)
      if (lanes * VEC <= C && C % (lanes * VEC) == 0) { SC = lanes * VEC; break; }
  NAWSOD_REQUIRE(SC > 0, NAWSOD_ERR_SHAPE, "roi_pool_f: C=%d has no power-of-two multiple of %d as a divisor", C, VEC);
  const size_t smem_bytes = use_smem ? (size_t)H * W * SC * esize : 0;
  const int slabs = C / SC;
  int threads = (int)get_tuning("pool_threads", 0);
  if (threads <= 0) threads = use_smem ? 1024 : 256;
  threads = std::max(32, std::min(1024, threads / 32 * 32));
  const int warps = threads / 32;
  int chunks = (int)get_tuning("pool_chunks", 0);
  if (chunks <= 0) {
    // ~2 CTAs per SM slot overall; never fewer than one RoI per warp
    // fill whole waves: choose the chunk count (2..8 waves' worth) whose CTA total wastes the least
    // of its last wave; never fewer than ~4 RoIs per warp
    // RoIs arrive grouped by image (roi_data/wsl.py:59-85 concatenates per-image blobs), so a chunk holds the RoIs of
    // ONE image: of the N CTAs that share a (slab, chunk) only one finds work, the others return after staging.  The
    // waves that matter are therefore counted over slabs x chunks (measured, 2 x 2000 RoIs bf16: 18 chunks 80.8 us,
    // 9 chunks 96.3 us; profiles/r1j_microbench_pool.log).
    const int64_t base = (int64_t)slabs;
    const int64_t slots = (int64_t)sm_count() * (use_smem ? 1 : (2048 / threads));
    const int64_t max_chunks = std::max<int64_t>(1, (int64_t)R / ((use_smem ? 4 : 1) * warps));
    double best_eff = -1.0;
    chunks = 1;
    for (int64_t c = 1; c <= std::min<int64_t>(max_chunks, ((use_smem ? 4 : 8) * slots) / base + 1); ++c) {
      const int64_t ctas = c * base;
      const int64_t waves = (ctas + slots - 1) / slots;
      double eff = (double)ctas / (double)(waves * slots);
      if (waves < 2) eff *= 0.9;      // a single wave has no slack for uneven RoI costs
      if (eff > best_eff + 1e-9) { best_eff = eff; chunks = (int)c; }
    }
  }
  chunks = std::min(chunks, 65535);
  PoolParams p;
  p.X = X; p.rois = rois; p.boost = boost;
  p.N = N; p.C = C; p.H = H; p.W = W; p.R = R; p.PH = PH; p.PW = PW;
  p.scale = scale; p.SC = SC; p.rois_per_chunk = (R + chunks - 1) / chunks;
  p.Y = Y; p.argmax = argmax;
  dim3 grid(slabs, chunks, N);
  if (x_dtype == NAWSOD_F32 && y_dtype == NAWSOD_F32)
    return launch_pool_fwd<float, float>(p, use_smem, smem_bytes, grid, threads, st);
  if (x_dtype == NAWSOD_F32 && y_dtype == NAWSOD_BF16)
    return launch_pool_fwd<float, __nv_bfloat16>(p, use_smem, smem_bytes, grid, threads, st);
  if (x_dtype == NAWSOD_BF16 && y_dtype == NAWSOD_BF16)
    return launch_pool_fwd<__nv_bfloat16, __nv_bfloat16>(p, use_smem, smem_bytes, grid, threads, st);
  if (x_dtype == NAWSOD_BF16 && y_dtype == NAWSOD_F32)
    return launch_pool_fwd<__nv_bfloat16, float>(p, use_smem, smem_bytes, grid, threads, st);
  set_error("roi_pool_f: unsupported dtype combination");
  return NAWSOD_ERR_UNSUPPORTED;
}

extern "C" int nawsod_roi_pool_f_fwd(const void* X, int x_dtype, int x_layout, const float* rois, const float* boost,
                                     int N, int C, int H, int W, int R, float spatial_scale, int pooled_h,
                                     int pooled_w, void* Y, int y_dtype, int y_layout, int32_t* argmax,
                                     void* stream) {
  NAWSOD_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && pooled_h > 0 && pooled_w > 0, NAWSOD_ERR_SHAPE,
                 "roi_pool_f: bad shape N=%d C=%d H=%d W=%d R=%d pooled=%dx%d", N, C, H, W, R, pooled_h, pooled_w);
  NAWSOD_REQUIRE((x_dtype == NAWSOD_F32 || x_dtype == NAWSOD_BF16) && (y_dtype == NAWSOD_F32 || y_dtype == NAWSOD_BF16),
                 NAWSOD_ERR_ARG, "roi_pool_f: bad dtype");
  NAWSOD_REQUIRE(x_layout == NAWSOD_NHWC, NAWSOD_ERR_UNSUPPORTED,
                 "roi_pool_f: the kernel consumes a channels-last map; convert NCHW with nawsod_transpose_batched");
  NAWSOD_REQUIRE(y_layout == NAWSOD_NHWC, NAWSOD_ERR_UNSUPPORTED,
                 "roi_pool_f: the kernel emits pooled-NHWC; convert with nawsod_transpose_batched for NCHW");
  if (R == 0 || N == 0) return NAWSOD_OK;   // empty rois: nothing to write (roi_loop_pool_op.cu:149-158)
  NAWSOD_REQUIRE(X && rois && Y, NAWSOD_ERR_ARG, "roi_pool_f: null pointer");
  NAWSOD_REQUIRE((int64_t)H * W < (1ll << 31) / 4, NAWSOD_ERR_SHAPE, "roi_pool_f: map too large");
  return pool_fwd_nhwc(X, x_dtype, rois, boost, N, C, H, W, R, spatial_scale, pooled_h, pooled_w, Y, y_dtype, argmax,
                       static_cast<cudaStream_t>(stream));
}

extern "C" int nawsod_roi_pool_f_bwd(const void* dY, int dy_dtype, int y_layout, const int32_t* argmax,
                                     const float* rois, const float* boost, int N, int C, int H, int W, int R,
                                     int pooled_h, int pooled_w, float* dX, int dx_layout, void* stream) {
  NAWSOD_REQUIRE(N >= 0 && C > 0 && H > 0 && W > 0 && R >= 0 && pooled_h > 0 && pooled_w > 0, NAWSOD_ERR_SHAPE,
                 "roi_pool_f_grad: bad shape");
  NAWSOD_REQUIRE(dy_dtype == NAWSOD_F32 || dy_dtype == NAWSOD_BF16, NAWSOD_ERR_ARG, "roi_pool_f_grad: bad dtype");
  NAWSOD_REQUIRE(y_layout == dx_layout && (y_layout == NAWSOD_NCHW || y_layout == NAWSOD_NHWC), NAWSOD_ERR_UNSUPPORTED,
                 "roi_pool_f_grad: dY/argmax and dX must share one layout (NCHW or NHWC)");
  NAWSOD_REQUIRE(dX || N == 0, NAWSOD_ERR_ARG, "roi_pool_f_grad: null dX");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N > 0) NAWSOD_CUDA_OK(cudaMemsetAsync(dX, 0, (size_t)N * C * H * W * sizeof(float), st));
  if (R == 0 || N == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(dY && argmax && rois, NAWSOD_ERR_ARG, "roi_pool_f_grad: null pointer");
  const int bins = pooled_h * pooled_w;
  const size_t total = (size_t)R * C * bins;
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, (size_t)sm_count() * 32);
#define NAWSOD_BWD(KERNEL, T)                                                                               \
  KERNEL<T><<<blocks, threads, 0, st>>>(static_cast<const T*>(dY), argmax, rois, boost, C, H * W, bins, total, dX)
  if (y_layout == NAWSOD_NHWC) {
    if (dy_dtype == NAWSOD_F32) NAWSOD_BWD(roi_pool_bwd_nhwc_kernel, float);
    else NAWSOD_BWD(roi_pool_bwd_nhwc_kernel, __nv_bfloat16);
  } else {
    if (dy_dtype == NAWSOD_F32) NAWSOD_BWD(roi_pool_bwd_nchw_kernel, float);
    else NAWSOD_BWD(roi_pool_bwd_nchw_kernel, __nv_bfloat16);
  }
#undef NAWSOD_BWD
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_roi_feature_boost(const float* X, const float* S, int R, int64_t feature_size, float* Y,
                                        void* stream) {
  NAWSOD_REQUIRE(R >= 0 && feature_size >= 0, NAWSOD_ERR_SHAPE, "roi_feature_boost: negative size");
  if (R == 0 || feature_size == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(X && S && Y, NAWSOD_ERR_ARG, "roi_feature_boost: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)R * feature_size;
  if (feature_size % 4 == 0 && aligned16(X) && aligned16(Y)) {
    const int64_t total4 = total / 4;
    const int blocks = (int)std::min<int64_t>((total4 + 255) / 256, (int64_t)sm_count() * 16);
    boost_kernel<<<blocks, 256, 0, st>>>(X, S, feature_size / 4, total4, Y);
  } else {
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    boost_kernel_scalar<<<blocks, 256, 0, st>>>(X, S, feature_size, total, Y);
  }
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}
