// CUDA-core kernels of the hot path: everything that is HBM-bound rather than tensor-bound.
//   * first_conv_kernel   - e11 (Cin = in_channels, K = 9*Cin, src/unet/model/unet.py:82,141): reads the image
//                           (uint8 pixels -> x/255 as src/unet/evaluate.py:45, or float [0,1]) and writes the
//                           split-bf16 NHWC activation with its reflect halo.
//   * filter_ws_kernel    - KB / AVG / AVG9 / identity predictor (src/filters/evaluate.py:29-50,136-141) fused with
//                           the WS estimator (src/ws/estimate.py:83-128): LSB flip, residual, local-variance
//                           weights, per-image partial sums. One pass over the uint8 image: 1 B/pixel read.
//   * ws_from_pred_kernel - WS reduction against a caller-supplied x_hat (second pass of correct_bias, or any
//                           external predictor output).
//   * finalize_kernel     - deterministic fixed-order reduction of the partial sums to beta_hat / l1 per image.
//   * pack / unpack       - NCHW fp32 <-> split-bf16 NHWC, used by the per-layer parity tests.
#include <algorithm>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "stencil.h"
#include "ws_math.cuh"

namespace wsu {

namespace {

// ------------------------------------------------------------------------------------------------ e11
// One CTA = kE11Rows rows of one 128-pixel row segment of one image. The input rows (reflect-resolved, already scaled to
// [0,1]) are staged in shared memory; thread (pixel lane pl = tid/8, channel group cg = tid%8) keeps its 8 x 9 weights in
// registers as fp32 pairs and walks 4 pixels per row. 8 threads write one pixel's 64 channels = one 128-byte line per plane.
// The kernel is instruction-issue bound (ncu, round 2: 199 instructions per pixel and thread at 0.78 IPC per scheduler, 16 % of
// the time in the staging loop's dependent global loads), so everything here is about instruction count: the output format is a
// template parameter, row pointers are formed once per row and the four pixels use immediate offsets, the e4m3 operands come
// from packed f32x2 arithmetic, uint8 pixels are loaded four at a time and scaled through a 256-entry table of x / 255.
constexpr int kE11Seg = 128;   // pixels per row segment
constexpr int kE11Rows = 16;   // rows per CTA (weights are loaded into registers once per CTA)
constexpr int kE11Pitch = kE11Seg + 4;

// kIn: 0 = uint8 pixels (x / 255), 1 = float32 in [0,1], 2 = uint8 pixels, LSB DIFFERENCE image (x_bar - x) / 255 = +-1/255:
// what the reference feeds the predictor for bias correction, pixel_estimator(x_bar - x) (src/ws/estimate.py:126-127)
// FMT: ACT_SPLIT or ACT_F16F8 (the two formats a full-resolution map can have)
template <int kIn, int FMT>
__global__ void __launch_bounds__(256, 2) first_conv_kernel(const void* __restrict__ img, const float* __restrict__ w,
                                                            const float* __restrict__ bias, Act out, int segs_per_row,
                                                            int row_groups) {
  __shared__ __align__(16) float rows[kE11Rows + 2][kE11Pitch];   // column c holds pixel x0 + c - 1
  __shared__ float lut[256];
  const int H = out.H, W = out.W;
  const int seg = blockIdx.x % segs_per_row;
  const int y0 = ((blockIdx.x / segs_per_row) % row_groups) * kE11Rows;
  const int b = blockIdx.x / (segs_per_row * row_groups);
  const int x0 = seg * kE11Seg;
  const int tid = threadIdx.x;
  auto reflect_row = [&](int r) {
    int yy = min(y0 + r - 1, 2 * H - 2);                    // rows past a ragged last group: keep the index valid
    return yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);  // reflect (unet.py:73)
  };
  auto reflect_col = [&](int col) {
    int xx = x0 + col - 1;
    xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
    return min(max(xx, 0), W - 1);                          // columns past a ragged last segment: any valid pixel
  };
  // x / 255. of src/unet/evaluate.py:45 as an IEEE division; kIn == 2: (x ^ 1) - x = +-1
  auto scale_u8 = [&](uint32_t px) -> float {
    if constexpr (kIn == 2) return (px & 1u) ? -lut[1] : lut[1];
    else return lut[px];
  };
  if constexpr (kIn != 1) {
    lut[tid] = __fdiv_rn(float(tid), 255.f);
    __syncthreads();
  }
  const uint8_t* im8 = static_cast<const uint8_t*>(img) + size_t(b) * H * W;
  const float* imf = static_cast<const float*>(img) + size_t(b) * H * W;
  const bool vec = kIn != 1 && (W % 4 == 0) && (reinterpret_cast<uintptr_t>(img) % 4 == 0);
  if (vec) {
    // interior columns as uchar4 groups (a group lies entirely inside or entirely outside the image because W % 4 == 0)
    uint32_t v[3];
    const int g = tid & 31, r0 = tid >> 5;
    const bool in = x0 + 4 * g < W;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int r = r0 + 8 * k;
      if (r < kE11Rows + 2 && in) v[k] = *reinterpret_cast<const uint32_t*>(im8 + size_t(reflect_row(r)) * W + x0 + 4 * g);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int r = r0 + 8 * k;
      if (r < kE11Rows + 2) {
        if (in) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rows[r][1 + 4 * g + j] = scale_u8((v[k] >> (8 * j)) & 0xffu);
        } else {
          for (int j = 0; j < 4; ++j) rows[r][1 + 4 * g + j] = scale_u8(im8[size_t(reflect_row(r)) * W + reflect_col(1 + 4 * g + j)]);
        }
      }
    }
    if (tid < 2 * (kE11Rows + 2)) {   // the two halo columns
      const int r = tid >> 1, col = (tid & 1) ? kE11Seg + 1 : 0;
      rows[r][col] = scale_u8(im8[size_t(reflect_row(r)) * W + reflect_col(col)]);
    }
  } else {
    for (int i = tid; i < (kE11Rows + 2) * (kE11Seg + 2); i += 256) {
      const int col = i % (kE11Seg + 2), r = i / (kE11Seg + 2);
      const size_t o = size_t(reflect_row(r)) * W + reflect_col(col);
      float v;
      if constexpr (kIn == 1) v = imf[o];
      else v = scale_u8(im8[o]);
      rows[r][col] = v;
    }
  }
  const int cg = tid & 7, pl = tid >> 3;
  // weights / accumulators as fp32 pairs: one FFMA2 advances two output channels
  uint64_t wr[4][9], bs[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bs[i] = pack_f32x2(bias[cg * 8 + 2 * i], bias[cg * 8 + 2 * i + 1]);
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[i][t] = pack_f32x2(w[(cg * 8 + 2 * i) * 9 + t], w[(cg * 8 + 2 * i + 1) * 9 + t]);
  }
  __syncthreads();
  // byte offset of this thread's piece inside a pixel's plane-1 row. ACT_F16F8: a pixel's plane 1 = 4 groups of [16 a2s bytes |
  // 16 a1q bytes]; this thread owns 8 channels = bytes [8 (cg & 1), +8) of both halves of group cg / 2 (two 8-byte stores).
  const int p1_off = FMT == ACT_F16F8 ? (cg >> 1) * 32 + (cg & 1) * 8 : cg * 16;
  uint32_t xborder = 0;   // bit s: pixel s of this thread's row also owns reflect-halo columns
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int x = x0 + s * 32 + pl;
    xborder |= uint32_t(x == 1 || x == W - 2) << s;
  }
#pragma unroll 1
  for (int ry = 0; ry < kE11Rows; ++ry) {
    const int y = y0 + ry;
    if (y >= H) break;
    const bool yborder = (y == 1 || y == H - 2);
    const size_t rowpix = (size_t(b) * (H + 2) + (y + 1)) * (W + 2) + (x0 + 1 + pl);
    __nv_bfloat16* q0 = out.base + rowpix * 64 + cg * 8;
    uint8_t* q1 = reinterpret_cast<uint8_t*>(out.base + out.plane + rowpix * 64) + p1_off;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int lx = s * 32 + pl;
      if (x0 + lx >= W) break;
      uint64_t acc[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = bs[i];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float v = rows[ry + dy][lx + dx];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fma2_bcast(v, wr[i][dy * 3 + dx], acc[i]);
        }
      uint4 vh, v1;
      uint32_t h[4], l[4];
      if constexpr (FMT == ACT_F16F8) {
        // plane 0: 8 fp16 values; plane 1: l[0..1] = the 8 a2s bytes e4m3((v - fp16 v) 2^14), l[2..3] = the 8 a1q bytes e4m3(8 v).
        // (v - h) 2^14 = fma(h, -2^14, v 2^14) exactly: v - h is representable and the scale is a power of two.
        static_assert(kF8ScaleA2 == 16384.f, "kNegA2h below is -2^14 as an fp16 bit pattern");
        const uint64_t kA2 = pack_f32x2(kF8ScaleA2, kF8ScaleA2);
        const uint64_t kA1 = pack_f32x2(kF8ScaleA1, kF8ScaleA1);
        const uint32_t kNegA2h = 0xF400F400u;
        uint32_t a2[4], a1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v0, v1f;
          unpack_f32x2(acc[i], v0, v1f);
          const uint64_t v = pack_f32x2(fmaxf(v0, 0.f), fmaxf(v1f, 0.f));
          unpack_f32x2(v, v0, v1f);
          h[i] = cvt_f16x2(v0, v1f);
          float e0, e1, s0, s1;
          unpack_f32x2(mul2(v, kA2), e0, e1);
          const float r0 = fhfma_lo(h[i], kNegA2h, e0), r1 = fhfma_hi(h[i], kNegA2h, e1);   // one FHFMA per value, halves read in place
          unpack_f32x2(mul2(v, kA1), s0, s1);
          a2[i] = cvt_e4m3x2(r0, r1);
          a1[i] = cvt_e4m3x2(s0, s1);
        }
        l[0] = a2[0] | (a2[1] << 16); l[1] = a2[2] | (a2[3] << 16);
        l[2] = a1[0] | (a1[1] << 16); l[3] = a1[2] | (a1[3] << 16);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a0, a1;
          unpack_f32x2(acc[i], a0, a1);
          split_pack2(fmaxf(a0, 0.f), fmaxf(a1, 0.f), h[i], l[i]);
        }
      }
      vh = make_uint4(h[0], h[1], h[2], h[3]);
      v1 = make_uint4(l[0], l[1], l[2], l[3]);
      // d0 / d1: this thread's plane-0 / plane-1 address for some pixel
      auto store_planes = [&](__nv_bfloat16* d0, uint8_t* d1) {
        *reinterpret_cast<uint4*>(d0) = vh;
        if constexpr (FMT == ACT_F16F8) {
          *reinterpret_cast<uint2*>(d1) = make_uint2(v1.x, v1.y);
          *reinterpret_cast<uint2*>(d1 + 16) = make_uint2(v1.z, v1.w);
        } else {
          *reinterpret_cast<uint4*>(d1) = v1;
        }
      };
      store_planes(q0 + s * 32 * 64, q1 + s * 32 * 128);
      if (yborder || ((xborder >> s) & 1)) {   // reflect-halo duplicates (border pixels only)
        const int x = x0 + lx;
        int ys[3], xs[3];
        const int ny = halo_targets(y, H, ys), nx = halo_targets(x, W, xs);
        for (int iy = 0; iy < ny; ++iy)
          for (int ix = 0; ix < nx; ++ix) {
            if (iy == 0 && ix == 0) continue;
            const size_t pix = ((size_t(b) * (H + 2) + ys[iy]) * (W + 2) + xs[ix]) * 64;
            store_planes(out.base + pix + cg * 8, reinterpret_cast<uint8_t*>(out.base + out.plane + pix) + p1_off);
          }
      }
    }
  }
}

// generic fallback for in_channels > kE11MaxCin: 8 threads per pixel, weights in shared memory
template <bool kFloatIn>
__global__ void __launch_bounds__(256) first_conv_generic_kernel(const void* __restrict__ img, int cin, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, Act out) {
  extern __shared__ float sw[];  // [64][cin*9] + [64] bias
  const int kk = cin * 9;
  for (int i = threadIdx.x; i < 64 * kk; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sw[64 * kk + i] = bias[i];
  __syncthreads();
  const int H = out.H, W = out.W;
  const size_t npix = size_t(out.B) * H * W;
  const size_t gid = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t pix = gid >> 3;
  if (pix >= npix) return;
  const int cg = int(gid & 7);
  const int x = int(pix % W);
  const int y = int((pix / W) % H);
  const int b = int(pix / (size_t(W) * H));
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = sw[64 * kk + cg * 8 + i];
  for (int ci = 0; ci < cin; ++ci) {
    float in[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      int yy = y + dy - 1;
      yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        int xx = x + dx - 1;
        xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
        const size_t o = ((size_t(b) * cin + ci) * H + yy) * W + xx;
        if constexpr (kFloatIn) in[dy * 3 + dx] = static_cast<const float*>(img)[o];
        else in[dy * 3 + dx] = __fdiv_rn(float(static_cast<const uint8_t*>(img)[o]), 255.f);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* wr = &sw[(cg * 8 + i) * kk + ci * 9];
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[i] = fmaf(in[t], wr[t], acc[i]);
    }
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pack2(fmaxf(acc[2 * i], 0.f), fmaxf(acc[2 * i + 1], 0.f), h[i], l[i]);
  int ys[3], xs[3];
  const int ny = halo_targets(y, H, ys), nx = halo_targets(x, W, xs);
  for (int iy = 0; iy < ny; ++iy)
    for (int ix = 0; ix < nx; ++ix) {
      const size_t off = ((size_t(b) * (H + 2) + ys[iy]) * (W + 2) + xs[ix]) * out.C + cg * 8;
      *reinterpret_cast<uint4*>(out.base + off) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(out.base + out.plane + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ------------------------------------------------------------------------------------------------ filter WS
// One CTA = one (image, 8-row strip). 128 threads x 4 pixels = one 512-pixel row segment per row; wider images loop.
// The strip and its +-1 rows are staged in shared memory with 16-byte loads where alignment allows.
constexpr int kStripRows = 8;
constexpr int kFThreads = 256;

__device__ __forceinline__ float predict_linear(int kind, const float (&n)[9]) {
  // n = 3x3 neighbourhood row-major, n[4] = centre. All sums of <= 9 uint8 values are exact in fp32.
  const float cross = (n[1] + n[3]) + (n[5] + n[7]);
  const float diag = (n[0] + n[2]) + (n[6] + n[8]);
  switch (kind) {
    case PRED_KB: return (2.f * cross - diag) * 0.25f;           // [[-1,2,-1],[2,0,2],[-1,2,-1]]/4
    case PRED_AVG: return (cross + diag) * 0.125f;               // 8 neighbours / 8
    case PRED_AVG9: return __fdiv_rn((cross + diag) + n[4], 9.f);  // 3x3 box / 9
    default: return n[4];                                        // '1'
  }
}

template <bool kFloatIn>
__global__ void __launch_bounds__(kFThreads) filter_ws_kernel(const void* __restrict__ img, int B, int H, int W, int kind,
                                                              int weighted, int want_bias, float* __restrict__ xhat_out,
                                                              float* __restrict__ partials, int strips) {
  extern __shared__ float tile[];  // (kStripRows + 2) x (W) pixel values as float
  const int b = blockIdx.x / strips;
  const int strip = blockIdx.x - b * strips;
  const int y0 = 1 + strip * kStripRows;               // first interior row of this strip
  const int rows = min(kStripRows, (H - 1) - y0);      // interior rows y0 .. y0+rows-1
  const int lrows = rows + 2;
  // stage rows y0-1 .. y0+rows
  if constexpr (kFloatIn) {
    const float* src = static_cast<const float*>(img) + (size_t(b) * H + (y0 - 1)) * W;
    for (int i = threadIdx.x; i < lrows * W; i += kFThreads) tile[i] = src[i] * 255.f;
  } else {
    const uint8_t* src = static_cast<const uint8_t*>(img) + (size_t(b) * H + (y0 - 1)) * W;
    const int total = lrows * W;
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const int nvec = total >> 4;
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      for (int i = threadIdx.x; i < nvec; i += kFThreads) {
        const uint4 v = __ldg(s4 + i);
        const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int k = 0; k < 4; ++k) tile[i * 16 + q * 4 + k] = float((wds[q] >> (8 * k)) & 0xffu);
      }
      for (int i = (nvec << 4) + threadIdx.x; i < total; i += kFThreads) tile[i] = float(src[i]);
    } else {
      for (int i = threadIdx.x; i < total; i += kFThreads) tile[i] = float(src[i]);
    }
  }
  __syncthreads();

  WsAcc acc;
  const int wi = W - 2;  // interior width
  for (int i = threadIdx.x; i < rows * wi; i += kFThreads) {
    const int r = i / wi;
    const int x = 1 + (i - r * wi);
    const float* c = tile + (r + 1) * W + x;
    float n[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) n[dy * 3 + dx] = c[(dy - 1) * W + (dx - 1)];
    const float xv = n[4];
    const float xbar = float(__float2int_rn(xv) ^ 1);
    const float xhat = predict_linear(kind, n);
    float s1 = 0.f, s2 = 0.f;
    if (weighted != WS_UNWEIGHTED) {
#pragma unroll
      for (int k = 0; k < 9; ++k)
        if (k != 4) {
          s1 += n[k];
          s2 = fmaf(n[k], n[k], s2);
        }
    }
    const float wgt = ws_weight(weighted, s1, s2);
    ws_accumulate(acc, xv, xbar, xhat, wgt);
    if (want_bias) {
      // predictor applied to the difference image x_bar - x (estimate.py:127); linear => apply to parities
      float dn[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) dn[k] = float(__float2int_rn(n[k]) ^ 1) - n[k];
      const float xb = predict_linear(kind, dn);
      acc.wb = fmaf(wgt * (xv - xbar), xb, acc.wb);
    }
    if (xhat_out) xhat_out[(size_t(b) * (H - 2) + (y0 - 1 + r)) * wi + (x - 1)] = xhat;
  }
  // block reduce in fixed order: warp shuffles, then warp 0 sums the 8 warp results sequentially
  __shared__ float red[kFThreads / 32][kPartialSlots];
  const float v0 = warp_sum(acc.wr), v1 = warp_sum(acc.w), v2 = warp_sum(acc.l1), v3 = warp_sum(acc.wb);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = v0;
    red[warp][1] = v1;
    red[warp][2] = v2;
    red[warp][3] = v3;
  }
  __syncthreads();
  if (threadIdx.x < kPartialSlots) {
    float s = 0.f;
    for (int wv = 0; wv < kFThreads / 32; ++wv) s += red[wv][threadIdx.x];
    partials[(size_t(b) * strips + strip) * kPartialSlots + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------ fast filter WS
// Register sliding-window version for the common case (uint8 image, KB or AVG predictor, no bias term, no x_hat
// output, W % 4 == 0): a thread owns 4 adjacent columns and walks down kFastRows rows, keeping the 3x6 neighbourhood
// in registers; one 32-bit load per thread and row (6 rows in flight), neighbour columns through warp shuffles.
// The stencil is evaluated in exact integer arithmetic (4*(x - x_hat) for KB, 8*(x - x_hat) for AVG); the unweighted
// sums stay integers until the per-CTA partial, which is emitted as two exactly representable floats.
constexpr int kFastRows = 30;      // interior rows per CTA (5 chunks of 6)
constexpr int kFastThreads = 128;  // x 4 pixels = 512 columns per CTA

template <int KIND, int WEIGHTED>
__global__ void __launch_bounds__(kFastThreads) filter_ws_fast_kernel(const uint8_t* __restrict__ img, int H, int W,
                                                                      float* __restrict__ partials, int strips, int xtiles) {
  const int xt = blockIdx.x % xtiles;
  const int strip = (blockIdx.x / xtiles) % strips;
  const int b = blockIdx.x / (xtiles * strips);
  const int lane = threadIdx.x & 31;
  const int x = xt * (kFastThreads * 4) + threadIdx.x * 4;
  const bool active = x < W;
  const int y0 = 1 + strip * kFastRows;
  const int yend = min(y0 + kFastRows, H - 1);
  const uint8_t* base = img + size_t(b) * H * W;
  bool inside[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) inside[c] = active && (x + c >= 1) && (x + c <= W - 2);

  int top[6], mid[6], tsq[6], msq[6];
  // own aligned word plus the words left and right of it (immediate offsets from one row pointer, L1 hits)
  const bool has_l = active && x > 0, has_r = active && (x + 4 < W);
  const uint8_t* rp = base + size_t(y0 - 1) * W + (active ? x : 0);
  int yrow = y0 - 1;
  auto fetch = [&](uint32_t& w, uint32_t& lw, uint32_t& rw) {
    w = active ? __ldg(reinterpret_cast<const uint32_t*>(rp)) : 0u;
    lw = has_l ? __ldg(reinterpret_cast<const uint32_t*>(rp - 4)) : 0u;
    rw = has_r ? __ldg(reinterpret_cast<const uint32_t*>(rp + 4)) : 0u;
    if (yrow < H - 1) rp += W;
    ++yrow;
  };
  auto unpack = [&](uint32_t w, uint32_t lw, uint32_t rw, int (&v)[6]) {
    v[0] = int(lw >> 24);
    v[1] = int(w & 0xffu);
    v[2] = int((w >> 8) & 0xffu);
    v[3] = int((w >> 16) & 0xffu);
    v[4] = int(w >> 24);
    v[5] = int(rw & 0xffu);
  };
  {
    uint32_t w0, l0, r0, w1, l1, r1;
    fetch(w0, l0, r0);
    fetch(w1, l1, r1);
    unpack(w0, l0, r0, top);
    unpack(w1, l1, r1, mid);
    if (WEIGHTED) {
#pragma unroll
      for (int c = 0; c < 6; ++c) { tsq[c] = top[c] * top[c]; msq[c] = mid[c] * mid[c]; }
    }
  }
  int acc_r = 0, acc_l1 = 0, acc_n = 0;
  float facc_r = 0.f, facc_w = 0.f;

  for (int yc = y0; yc < yend; yc += 6) {
    uint32_t w[6], le[6], re[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) fetch(w[r], le[r], re[r]);
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      int bot[6], bsq[6];
      unpack(w[r], le[r], re[r], bot);
      if (yc + r < yend) {
        int vs[6], vq[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) vs[c] = top[c] + bot[c];
        if (WEIGHTED) {
#pragma unroll
          for (int c = 0; c < 6; ++c) { bsq[c] = bot[c] * bot[c]; vq[c] = tsq[c] + bsq[c]; }
        }
#pragma unroll
        for (int c = 1; c <= 4; ++c) {
          const int xc = mid[c];
          const int s8 = vs[c - 1] + vs[c] + vs[c + 1] + mid[c - 1] + mid[c + 1];   // 8-neighbour sum
          int e;                                                                     // scaled residual x - x_hat
          if (KIND == PRED_KB) e = 4 * xc - (2 * (vs[c] + mid[c - 1] + mid[c + 1]) - (vs[c - 1] + vs[c + 1]));
          else e = 8 * xc - s8;
          const int de = (xc & 1) ? e : -e;                                          // (x - x_bar) * (x - x_hat), scaled
          if (inside[c - 1]) {
            acc_l1 += abs(e);
            if (WEIGHTED) {
              const int q8 = vq[c - 1] + vq[c] + vq[c + 1] + msq[c - 1] + msq[c + 1];
              // 5 + var = (320 + 8*S2 - S1^2) / 64 exactly (integers < 2^23); the common factor 64 cancels in
              // sum(w r) / sum(w), so w = 1/D (weighted) or D (anti-weighted). Integers are converted with the
              // 2^23 magic-number trick (no I2F), the reciprocal is MUFU.RCP + one Newton step (<= 1 ulp).
              const float dd = __int_as_float(0x4B000000 | (320 + 8 * q8 - s8 * s8)) - 8388608.f;
              float wgt = dd;
              if (WEIGHTED == WS_WEIGHTED) {
                float r;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(dd));
                wgt = r * fmaf(-dd, r, 2.f);
              }
              const float fde = __int_as_float(0x4B000000 | (de + 4096)) - 8392704.f;
              facc_r = fmaf(wgt, fde, facc_r);
              facc_w += wgt;
            } else {
              acc_r += de;
              acc_n += 1;
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 6; ++c) { top[c] = mid[c]; mid[c] = bot[c]; }
      if (WEIGHTED) {
#pragma unroll
        for (int c = 0; c < 6; ++c) { tsq[c] = msq[c]; msq[c] = bsq[c]; }
      }
    }
  }
  constexpr float kScale = (KIND == PRED_KB) ? 0.25f : 0.125f;
  __shared__ int ired[kFastThreads / 32][3];
  __shared__ float fred[kFastThreads / 32][2];
  const int warp = threadIdx.x >> 5;
  const int r0 = __reduce_add_sync(0xffffffffu, acc_r), r1 = __reduce_add_sync(0xffffffffu, acc_l1),
            r2 = __reduce_add_sync(0xffffffffu, acc_n);
  const float f0 = warp_sum(facc_r), f1 = warp_sum(facc_w);
  if (lane == 0) {
    ired[warp][0] = r0; ired[warp][1] = r1; ired[warp][2] = r2;
    fred[warp][0] = f0; fred[warp][1] = f1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int sr = 0, sl = 0, sn = 0;
    float fr = 0.f, fw = 0.f;
    for (int i = 0; i < kFastThreads / 32; ++i) { sr += ired[i][0]; sl += ired[i][1]; sn += ired[i][2]; fr += fred[i][0]; fw += fred[i][1]; }
    float* dst = partials + (size_t(b) * strips * xtiles + size_t(strip) * xtiles + xt) * 2 * kPartialSlots;
    // integers split into a multiple of 4096 and a 12-bit remainder: both exact in fp32, summed in double by finalize
    const int sr_lo = sr & 0xfff, sl_lo = sl & 0xfff;
    dst[0] = WEIGHTED ? fr * kScale : float(sr - sr_lo) * kScale;
    dst[1] = WEIGHTED ? fw : float(sn);
    dst[2] = float(sl - sl_lo) * kScale;
    dst[3] = 0.f;
    dst[4] = WEIGHTED ? 0.f : float(sr_lo) * kScale;
    dst[5] = 0.f;
    dst[6] = float(sl_lo) * kScale;
    dst[7] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ packed filter WS
// Unweighted KB / AVG beta_hat only (no l1): two pixels per 32-bit register in 16-bit lanes, plain integer adds (every
// intermediate fits: 8-neighbour sums <= 2040, biased residual in [8, 4088]), so one ALU instruction advances two
// pixels. Per lane the kernel accumulates (x - x_bar) * S*(x - x_hat) + 2048 (S = 4 for KB, 8 for AVG); lanes are
// flushed into 32-bit sums every 12 rows (12 * 4088 < 65536) and the bias is removed analytically at the end.
constexpr int kPackRows = 36;  // interior rows per CTA (3 flush groups of 12 = 6 chunks of 6)

__device__ __forceinline__ uint32_t pair_lo(uint32_t w, uint32_t sel) { return __byte_perm(w, 0u, sel); }

template <int KIND>
__global__ void __launch_bounds__(kFastThreads) filter_ws_packed_kernel(const uint8_t* __restrict__ img, int H, int W,
                                                                        float* __restrict__ partials, int strips, int xtiles) {
  const int xt = blockIdx.x % xtiles;
  const int strip = (blockIdx.x / xtiles) % strips;
  const int b = blockIdx.x / (xtiles * strips);
  const int lane = threadIdx.x & 31;
  const int x = xt * (kFastThreads * 4) + threadIdx.x * 4;
  const bool active = x < W;
  const int y0 = 1 + strip * kPackRows;
  const int yend = min(y0 + kPackRows, H - 1);
  const uint8_t* base = img + size_t(b) * H * W;
  // lane masks of the two output pairs (pixels x..x+1 and x+2..x+3): interior columns only
  auto in_col = [&](int c) { return active && (x + c >= 1) && (x + c <= W - 2); };
  const uint32_t m12 = (in_col(0) ? 0x0000ffffu : 0u) | (in_col(1) ? 0xffff0000u : 0u);
  const uint32_t m34 = (in_col(2) ? 0x0000ffffu : 0u) | (in_col(3) ? 0xffff0000u : 0u);
  const int valid_cols = int(in_col(0)) + int(in_col(1)) + int(in_col(2)) + int(in_col(3));

  // packed pairs of one image row: p[0]=(c0,c1) p[1]=(c1,c2) p[2]=(c2,c3) p[3]=(c3,c4) p[4]=(c4,c5); c1..c4 = own pixels.
  // Each thread loads its own aligned word and the words left and right of it (immediate offsets from one row pointer,
  // L1 hits): no shuffles, no warp-edge special cases.
  const bool has_l = active && x > 0, has_r = active && (x + 4 < W);
  const uint8_t* rp = base + size_t(y0 - 1) * W + (active ? x : 0);
  int yrow = y0 - 1;
  auto fetch = [&](uint32_t& w, uint32_t& lw, uint32_t& rw) {
    w = active ? __ldg(reinterpret_cast<const uint32_t*>(rp)) : 0u;
    lw = has_l ? __ldg(reinterpret_cast<const uint32_t*>(rp - 4)) : 0u;
    rw = has_r ? __ldg(reinterpret_cast<const uint32_t*>(rp + 4)) : 0u;
    if (yrow < H - 1) rp += W;  // rows past the image bottom re-read the last row (their results are never used)
    ++yrow;
  };
  auto unpack = [&](uint32_t w, uint32_t lw, uint32_t rw, uint32_t (&p)[5]) {
    const uint32_t wl = __funnelshift_r(lw, w, 24);  // bytes (c0, c1, c2, c3)
    const uint32_t wr = __funnelshift_r(w, rw, 8);   // bytes (c2, c3, c4, c5)
    p[0] = pair_lo(wl, 0x4140);
    p[1] = pair_lo(w, 0x4140);
    p[2] = pair_lo(w, 0x4241);
    p[3] = pair_lo(w, 0x4342);
    p[4] = pair_lo(wr, 0x4342);
  };
  uint32_t top[5], mid[5];
  {
    uint32_t w0, l0, r0, w1, l1, r1;
    fetch(w0, l0, r0);
    fetch(w1, l1, r1);
    unpack(w0, l0, r0, top);
    unpack(w1, l1, r1, mid);
  }
  constexpr uint32_t kBias = 0x08000800u, kTwoBias = 0x10001000u;
  long long total = 0;
  int rows_done = 0;
  for (int yg = y0; yg < yend; yg += 12) {
    uint32_t acc12 = 0, acc34 = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int yc = yg + half * 6;
      uint32_t w[6], le[6], re[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) fetch(w[r], le[r], re[r]);
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        uint32_t bot[5];
        unpack(w[r], le[r], re[r], bot);
        if (yc + r < yend) {
          uint32_t vs[5];
#pragma unroll
          for (int k = 0; k < 5; ++k) vs[k] = top[k] + bot[k];
#pragma unroll
          for (int o = 0; o < 2; ++o) {
            // output pair o: centre pair index 1 + 2*o, left-neighbour pair 2*o, right-neighbour pair 2 + 2*o
            const uint32_t C = mid[1 + 2 * o], sh = mid[2 * o] + mid[2 + 2 * o];
            uint32_t eb;
            if (KIND == PRED_KB) {
              const uint32_t t = vs[1 + 2 * o] + sh;                       // N+S+W+E
              const uint32_t P = 4u * C + (vs[2 * o] + vs[2 + 2 * o] + kBias);  // 4x + diagonals + bias
              eb = P - 2u * t;                                             // 4(x - x_hat) + 2048 per lane
            } else {
              const uint32_t s8 = vs[2 * o] + vs[1 + 2 * o] + vs[2 + 2 * o] + sh;
              eb = 8u * C + kBias - s8;                                    // 8(x - x_hat) + 2048 per lane
            }
            const uint32_t odd = (C & 0x00010001u) * 0xffffu;              // 0xffff in lanes with odd x
            const uint32_t deb = (eb & odd) | ((kTwoBias - eb) & ~odd);    // (x - x_bar) * residual + 2048
            if (o == 0) acc12 += deb; else acc34 += deb;   // border lanes are masked when the lanes are flushed
          }
          ++rows_done;
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) { top[k] = mid[k]; mid[k] = bot[k]; }
      }
    }
    acc12 &= m12;
    acc34 &= m34;
    total += (acc12 & 0xffffu) + (acc12 >> 16) + (acc34 & 0xffffu) + (acc34 >> 16);
  }
  total -= 2048ll * rows_done * valid_cols;
  const int npx = rows_done * valid_cols;
  // CTA reduction (exact integers; |total| per CTA < 2^31)
  __shared__ int ired[kFastThreads / 32][2];
  const int warp = threadIdx.x >> 5;
  const int r0 = __reduce_add_sync(0xffffffffu, int(total)), r1 = __reduce_add_sync(0xffffffffu, npx);
  if (lane == 0) { ired[warp][0] = r0; ired[warp][1] = r1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int sr = 0, sn = 0;
    for (int i = 0; i < kFastThreads / 32; ++i) { sr += ired[i][0]; sn += ired[i][1]; }
    constexpr float kScale = (KIND == PRED_KB) ? 0.25f : 0.125f;
    float* dst = partials + (size_t(b) * strips * xtiles + size_t(strip) * xtiles + xt) * 2 * kPartialSlots;
    const int sr_lo = sr & 0xfff;
    dst[0] = float(sr - sr_lo) * kScale; dst[1] = float(sn); dst[2] = 0.f; dst[3] = 0.f;
    dst[4] = float(sr_lo) * kScale;      dst[5] = 0.f;       dst[6] = 0.f; dst[7] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ adjoint filter WS
// Unweighted KB / AVG beta_hat through the adjoint of the predictor stencil. With s = x - x_bar = +-1 on the interior
// (0 elsewhere) and R = D * (x - x_hat) = K (*) x (K = 4*delta - 4*KB or 8*delta - 8*AVG, symmetric, sums to 0):
//     sum_interior s * R  =  sum_{q in image} x_q * (K (*) s)_q,
// so the stencil is applied to the PARITY plane instead of the pixels: e = s + 1 in {0, 1, 2} (1 outside the
// interior and outside the image; a constant is annihilated by K) fits four pixels per 32-bit register as bytes,
// both separable passes stay inside a byte (KB: T = e_l + e_r - 2e + 4 in [0,8], G' = T_u + T_d - 2T + 16 in [0,32];
// AVG: T = e_l + e + e_r in [0,6], G' = 9e + 16 - (T_u + T + T_d) in [0,32]) and the pixel side is two dp4a per word:
// sum x*G' and sum x (the +16 bias is removed as 16 * sum x). Exact integer arithmetic end to end; about 2.6 thread
// instructions per pixel instead of 13 in the packed kernel. A warp walks a 512-pixel-wide strip of kAdjRows rows
// (16 pixels = one 16-byte load per lane and row); neighbour words come from warp shuffles.
constexpr int kAdjRows = 64;    // G rows per warp task (T rows: + 2)
constexpr int kAdjWarps = 8;    // warps per CTA = images per CTA (all warps of a CTA share the row range)
constexpr int kAdjUnroll = 8;   // rows fetched per batch in the steady state
constexpr uint32_t kOnes = 0x01010101u;

template <int KIND>
struct AdjState {
  uint32_t M[4], C[4];                       // parity -> e: e = (w & M) * 2 + C
  uint32_t Tm1[4], Tm2[4], wm1[4], Em1[4];   // previous two T rows, previous pixel row, 9e+16 of the previous row (AVG)
  uint32_t acc = 0u, accx = 0u;
  int lane;
  // multipliers 2 and -2 arrive as kernel arguments: with immediates the compiler strength-reduces the multiply-adds
  // into shifts/adds on the (already busier) ALU pipe; as register operands they stay IMADs on the FMA pipe
  uint32_t two, neg2;

  // one image row: w = 16 pixels of this lane, (exl, exr) = e words of the strip's outer neighbours (lanes 0 / 31)
  template <bool kInterior, bool kEmit>
  __device__ __forceinline__ void row(const uint4& wv, uint32_t exl, uint32_t exr) {
    const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) e[i] = kInterior ? (w[i] & M[i]) * two + C[i] : kOnes;
    uint32_t eL = kOnes, eR = kOnes;
    if (kInterior) {
      eL = __shfl_up_sync(0xffffffffu, e[3], 1);
      eR = __shfl_down_sync(0xffffffffu, e[0], 1);
      if (lane == 0) eL = exl;
      if (lane == 31) eR = exr;
    }
    uint32_t T[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t el = __funnelshift_l(i ? e[i - 1] : eL, e[i], 8);      // e of the left neighbours
      const uint32_t er = __funnelshift_r(e[i], i < 3 ? e[i + 1] : eR, 8);  // e of the right neighbours
      T[i] = (KIND == PRED_KB) ? el + er + (e[i] * neg2 + 0x04040404u) : el + er + e[i];
    }
    if (kEmit) {  // G of the previous row is complete
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t G = (KIND == PRED_KB) ? Tm2[i] + T[i] + (Tm1[i] * neg2 + 0x10101010u)
                                             : Em1[i] - (Tm2[i] + Tm1[i] + T[i]);
        acc = __dp4a(wm1[i], G, acc);
        accx = __dp4a(wm1[i], kOnes, accx);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      Tm2[i] = Tm1[i]; Tm1[i] = T[i]; wm1[i] = w[i];
      if (KIND != PRED_KB) Em1[i] = 9u * e[i] + 0x10101010u;
    }
  }
};

template <int KIND, bool kMulti>
__global__ void __launch_bounds__(kAdjWarps * 32, 3) filter_ws_adjoint_kernel(const uint8_t* __restrict__ img, int B, int H,
                                                                              int W, float* __restrict__ partials,
                                                                              int rstrips, int cstrips, uint32_t two,
                                                                              uint32_t neg2) {
  const int lane = threadIdx.x & 31;
  const int cs = kMulti ? blockIdx.x % cstrips : 0;
  const int rs = (blockIdx.x / cstrips) % rstrips;
  const int b = (blockIdx.x / (cstrips * rstrips)) * kAdjWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int x = cs * 512 + lane * 16;
  const bool active = x < W;  // W % 16 == 0: a lane's 16 pixels are all inside or all outside
  const int r0 = rs * kAdjRows, r1 = min(r0 + kAdjRows, H);  // this CTA's warps own G rows [r0, r1) of their images

  AdjState<KIND> st;
  st.lane = lane;
  st.two = two;
  st.neg2 = neg2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    st.M[i] = active ? kOnes : 0u; st.C[i] = active ? 0u : kOnes;   // lanes right of the image hold e = 1, w = 0
    st.Tm1[i] = st.Tm2[i] = st.wm1[i] = st.Em1[i] = 0u;
  }
  if (active && x == 0) { st.M[0] &= ~0xffu; st.C[0] |= 0x01u; }                        // column 0
  if (active && x + 16 == W) { st.M[3] &= ~0xff000000u; st.C[3] |= 0x01000000u; }       // column W-1
  // strip edges inside the image (W > 512 only): lane 0 / lane 31 fetch the neighbouring word themselves
  const bool edge_l = kMulti && (lane == 0) && active && x > 0;
  const bool edge_r = kMulti && (lane == 31) && active && (x + 16 < W);
  const int edge_off = edge_l ? -4 : 16;

  const uint8_t* rp = img + size_t(b) * H * W + (active ? x : 0);   // row pointer of the next row to fetch
  auto fetch = [&](uint4& wv, uint32_t& ex) {
    wv = active ? __ldg(reinterpret_cast<const uint4*>(rp)) : make_uint4(0u, 0u, 0u, 0u);
    if (kMulti) ex = (edge_l || edge_r) ? (__ldg(reinterpret_cast<const uint32_t*>(rp + edge_off)) & kOnes) * 2u : kOnes;
    else ex = kOnes;
    rp += W;
  };
  // generic row (top / bottom of the strip, image border rows): runtime row class, one row at a time
  auto slow_row = [&](int y, bool emit) {
    uint4 wv = make_uint4(0u, 0u, 0u, 0u);
    uint32_t ex = kOnes;
    if (y >= 0 && y < H) fetch(wv, ex);
    const bool interior = (y > 0) && (y < H - 1);
    if (interior) { if (emit) st.template row<true, true>(wv, ex, ex); else st.template row<true, false>(wv, ex, ex); }
    else          { if (emit) st.template row<false, true>(wv, ex, ex); else st.template row<false, false>(wv, ex, ex); }
  };

  if (r0 > 0) rp += size_t(r0 - 1) * W;
  slow_row(r0 - 1, false);
  slow_row(r0, false);
  int y = r0 + 1;
  const int ylast = min(r1, H - 2);   // rows y..ylast are interior rows and each completes the G row above it
  for (; y + kAdjUnroll - 1 <= ylast; y += kAdjUnroll) {
    uint4 wv[kAdjUnroll];
    uint32_t ex[kAdjUnroll];
#pragma unroll
    for (int r = 0; r < kAdjUnroll; ++r) fetch(wv[r], ex[r]);
#pragma unroll
    for (int r = 0; r < kAdjUnroll; ++r) st.template row<true, true>(wv[r], ex[r], ex[r]);
  }
  for (; y <= r1; ++y) slow_row(y, true);

  // per lane: sum x*G' <= 16 px * 66 rows * 255 * 32 and 16 * sum x likewise; the warp total stays below 2^31
  const int mine = int(st.acc) - 16 * int(st.accx);
  const int sr = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) {
    constexpr float kScale = (KIND == PRED_KB) ? 0.25f : 0.125f;
    float* dst = partials + ((size_t(b) * rstrips + rs) * cstrips + cs) * 2 * kPartialSlots;
    const int sr_lo = sr & 0xfff;
    const int n = (rs == 0 && cs == 0) ? (H - 2) * (W - 2) : 0;   // the pixel count rides on the image's first task
    const int n_lo = n & 0xfff;
    dst[0] = float(sr - sr_lo) * kScale; dst[1] = float(n - n_lo); dst[2] = 0.f; dst[3] = 0.f;
    dst[4] = float(sr_lo) * kScale;      dst[5] = float(n_lo);     dst[6] = 0.f; dst[7] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ window (dp4a) filter WS
// Local-variance weighted (and / or L1-reporting) KB / AVG estimator. Same task shape as the adjoint kernel (a warp walks
// a 512-pixel-wide strip, 16 pixels per lane and row, neighbours through shuffles), but the per-pixel weights force
// per-pixel arithmetic. Every pixel's 3-byte row window (left, centre, right) sits in one 32-bit register, so each of
// the nine-point sums is three chained dp4a: S9 = sum x (coefficients 1,1,1), Q9 = sum x^2 (window against itself),
// R = D(x - x_hat) (KB: rows (1,-2,1), (-2,4,-2), (1,-2,1); AVG: 8x - S8). Then
//   -64 (5 + var) = S8^2 - 8 Q8 - 320  (exact integer, |.| < 2^23), S8 = S9 - x, Q8 = Q9 - x^2,
// and the weight is its reciprocal (MUFU.RCP, 1 ulp: the reference's own float32 weights are only good to 0.5 %,
// SURVEY.md 8c). Weights are carried with the opposite sign (no negation needed); the sign cancels in sum(w r)/sum(w).
// The parity sign is applied by accumulating all pixels and the odd pixels separately: sum s*R = 2 sum_odd - sum_all.
constexpr int kWinRows = 64;   // output rows per warp task
constexpr int kWinWarps = 8;

__device__ __forceinline__ int dp4a_us(uint32_t a, int b_s8x4, int c) {   // unsigned bytes of a times signed bytes of b
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_s8x4), "r"(c));
  return d;
}

template <int KIND, int WEIGHTED, bool kL1, bool kMulti>
__global__ void __launch_bounds__(kWinWarps * 32, 2) filter_ws_window_kernel(const uint8_t* __restrict__ img, int B, int H,
                                                                             int W, float* __restrict__ partials,
                                                                             int rstrips, int cstrips) {
  const int lane = threadIdx.x & 31;
  const int cs = kMulti ? blockIdx.x % cstrips : 0;
  const int rs = (blockIdx.x / cstrips) % rstrips;
  const int b = (blockIdx.x / (cstrips * rstrips)) * kWinWarps + (threadIdx.x >> 5);
  if (b >= B) return;
  const int x = cs * 512 + lane * 16;
  const bool active = x < W;
  const int r0 = 1 + rs * kWinRows, r1 = min(r0 + kWinRows, H - 1);   // interior rows [r0, r1) of this task
  const bool first_col = active && x == 0, last_col = active && (x + 16 == W);
  const bool edge_l = kMulti && (lane == 0) && active && x > 0;
  const bool edge_r = kMulti && (lane == 31) && active && (x + 16 < W);
  const int edge_off = edge_l ? -4 : 16;

  // masked 3-byte windows (left, centre, right, 0) of the 16 pixels of a row
  auto windows = [&](const uint4& v, uint32_t ex, uint32_t (&win)[16]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t wl = __shfl_up_sync(0xffffffffu, w[3], 1), wr = __shfl_down_sync(0xffffffffu, w[0], 1);
    if (kMulti) {
      if (lane == 0) wl = ex;
      if (lane == 31) wr = ex;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t prev = i ? w[i - 1] : wl, next = i < 3 ? w[i + 1] : wr;
      win[4 * i + 0] = __funnelshift_r(prev, w[i], 24) & 0x00ffffffu;
      win[4 * i + 1] = w[i] & 0x00ffffffu;
      win[4 * i + 2] = w[i] >> 8;
      win[4 * i + 3] = __funnelshift_r(w[i], next, 16) & 0x00ffffffu;
    }
  };
  const uint8_t* rp = img + size_t(b) * H * W + size_t(r0 - 1) * W + (active ? x : 0);
  auto fetch = [&](uint4& v, uint32_t& ex) {
    v = active ? __ldg(reinterpret_cast<const uint4*>(rp)) : make_uint4(0u, 0u, 0u, 0u);
    ex = 0u;
    if (kMulti && (edge_l || edge_r)) ex = __ldg(reinterpret_cast<const uint32_t*>(rp + edge_off));
    rp += W;
  };

  uint32_t ra[16], rb[16], rc[16];   // three row buffers; their roles (top, mid, bottom) rotate statically
  {
    uint4 v0, v1;
    uint32_t e0, e1;
    fetch(v0, e0);
    fetch(v1, e1);
    windows(v0, e0, ra);
    windows(v1, e1, rb);
  }
  int acc_all = 0, acc_odd = 0, acc_l1 = 0;
  float f_all = 0.f, f_odd = 0.f, f_w = 0.f;
  constexpr int kOnes3 = 0x00010101;
  constexpr int kEdgeRow = 0x0001fe01;    // ( 1, -2,  1, 0)
  constexpr int kMidRow = 0x00fe04fe;     // (-2,  4, -2, 0)

  uint4 vnext;
  uint32_t enext;
  fetch(vnext, enext);
  int y = r0;
  // one interior row: `bot` receives the windows of row y+1, then the 16 pixels of row y are accumulated
  auto row = [&](const uint32_t (&top)[16], const uint32_t (&mid)[16], uint32_t (&bot)[16]) {
    windows(vnext, enext, bot);
    if (y + 1 < r1) fetch(vnext, enext);   // the next row's load overlaps this row's arithmetic
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // columns 0 and W-1 are not estimated: their residual / weight is forced to 0 (only pixels 0 and 15 can be affected)
      const bool dead = (j == 0 && first_col) || (j == 15 && last_col);
      const int c = int(__byte_perm(mid[j], 0u, 0x4441));
      int R, s9 = 0;
      if (KIND == PRED_AVG || WEIGHTED != WS_UNWEIGHTED)
        s9 = dp4a_us(bot[j], kOnes3, dp4a_us(mid[j], kOnes3, dp4a_us(top[j], kOnes3, 0)));
      if (KIND == PRED_KB) R = dp4a_us(bot[j], kEdgeRow, dp4a_us(mid[j], kMidRow, dp4a_us(top[j], kEdgeRow, 0)));
      else R = 9 * c - s9;
      if ((j == 0 || j == 15) && dead) R = 0;
      if (WEIGHTED != WS_UNWEIGHTED) {
        const uint32_t q9 = __dp4a(bot[j], bot[j], __dp4a(mid[j], mid[j], __dp4a(top[j], top[j], 0u)));
        const int s8 = s9 - c;
        const int q8 = int(q9) - c * c;
        const int nd = s8 * s8 - 8 * q8 - 320;              // -64 (5 + var)
        const float fd = __int2float_rn(nd);
        float wgt = fd;
        if (WEIGHTED == WS_WEIGHTED) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(wgt) : "f"(fd));
        if ((j == 0 || j == 15) && dead) wgt = 0.f;
        const float fr = __int2float_rn(R);
        f_all = fmaf(wgt, fr, f_all);
        if (c & 1) f_odd = fmaf(wgt, fr, f_odd);
        f_w += wgt;
      } else {
        acc_all += R;
        acc_odd += R * (c & 1);
      }
      if (kL1) acc_l1 += abs(R);
    }
    ++y;
  };
  while (y < r1) {
    row(ra, rb, rc);
    if (y < r1) row(rb, rc, ra);
    if (y < r1) row(rc, ra, rb);
  }
  if (!active) { acc_all = acc_odd = acc_l1 = 0; f_all = f_odd = f_w = 0.f; }   // lanes right of the image saw zeros

  constexpr float kScale = (KIND == PRED_KB) ? 0.25f : 0.125f;
  const int rows = max(r1 - r0, 0);
  const int npx_lane = active ? rows * (16 - int(first_col) - int(last_col)) : 0;
  const int sr = __reduce_add_sync(0xffffffffu, 2 * acc_odd - acc_all);
  const int sl = __reduce_add_sync(0xffffffffu, acc_l1);
  const int sn = __reduce_add_sync(0xffffffffu, npx_lane);
  const float fr = warp_sum(2.f * f_odd - f_all), fw = warp_sum(f_w);
  if (lane == 0) {
    float* dst = partials + ((size_t(b) * rstrips + rs) * cstrips + cs) * 2 * kPartialSlots;
    const int sr_lo = sr & 0xfff, sl_lo = sl & 0xfff;
    dst[0] = WEIGHTED ? fr * kScale : float(sr - sr_lo) * kScale;
    dst[1] = WEIGHTED ? fw : float(sn);
    dst[2] = float(sl - sl_lo) * kScale;
    dst[3] = 0.f;
    dst[4] = WEIGHTED ? 0.f : float(sr_lo) * kScale;
    dst[5] = 0.f;
    dst[6] = float(sl_lo) * kScale;
    dst[7] = 0.f;
  }
}

// WS terms against a caller-supplied prediction (pixel units), grid-stride per image chunk.
template <bool kFloatIn>
__global__ void __launch_bounds__(256) ws_from_pred_kernel(const void* __restrict__ img, const float* __restrict__ xhat,
                                                           int xhat_cropped, const float* __restrict__ xbias, int B, int H,
                                                           int W, int weighted, int crop, float* __restrict__ partials,
                                                           int chunks) {
  const int b = blockIdx.x / chunks;
  const int chunk = blockIdx.x - b * chunks;
  const int h = crop ? H - 2 : H, w = crop ? W - 2 : W;
  const int npix = h * w;
  const int per = (npix + chunks - 1) / chunks;
  const int lo = chunk * per, hi = min(npix, lo + per);
  WsAcc acc;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const int r = i / w, cidx = i - r * w;
    const int y = r + crop, x = cidx + crop;
    const size_t pix = (size_t(b) * H + y) * W + x;
    float xv, xbar, s1 = 0.f, s2 = 0.f;
    if constexpr (kFloatIn) {
      const float* im = static_cast<const float*>(img);
      ws_load_f32(im[pix], xv, xbar);
      if (weighted != WS_UNWEIGHTED)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            if (!dy && !dx) continue;
            const float q = im[pix + dy * W + dx] * 255.f;
            s1 += q;
            s2 = fmaf(q, q, s2);
          }
    } else {
      const uint8_t* im = static_cast<const uint8_t*>(img);
      ws_load_u8(im[pix], xv, xbar);
      if (weighted != WS_UNWEIGHTED) {
        int i1 = 0, i2 = 0;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            if (!dy && !dx) continue;
            const int q = im[pix + dy * W + dx];
            i1 += q;
            i2 += q * q;
          }
        s1 = float(i1);
        s2 = float(i2);
      }
    }
    const size_t pidx = xhat_cropped ? (size_t(b) * (H - 2) + (y - 1)) * (W - 2) + (x - 1) : pix;
    const float wgt = ws_weight(weighted, s1, s2);
    ws_accumulate(acc, xv, xbar, xhat[pidx], wgt);
    if (xbias) acc.wb = fmaf(wgt * (xv - xbar), xbias[pidx], acc.wb);
  }
  __shared__ float red[8][kPartialSlots];
  const float v0 = warp_sum(acc.wr), v1 = warp_sum(acc.w), v2 = warp_sum(acc.l1), v3 = warp_sum(acc.wb);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = v0;
    red[warp][1] = v1;
    red[warp][2] = v2;
    red[warp][3] = v3;
  }
  __syncthreads();
  if (threadIdx.x < kPartialSlots) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    partials[(size_t(b) * chunks + chunk) * kPartialSlots + threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------ WS loss gradient
// beta_hat_b = (1/n) sum_{p in crop} (x_p - x_bar_p)(x_p - scale * o_p)  (WSLoss._error, src/_defs/losses.py:46-61, with
// scale = 255 for outputs in [0,1]) is linear in the prediction o: d beta_hat_b / d o_p = -scale (x_p - x_bar_p) / n.
// grad[b,p] = coef[b] * that inside the crop and 0 outside; coef[b] = dL/d beta_hat_b comes from the caller's autograd.
template <bool kFloatIn>
__global__ void __launch_bounds__(256) ws_grad_pred_kernel(const void* __restrict__ img, const float* __restrict__ coef,
                                                           float* __restrict__ grad, int H, int W, int crop, float scale,
                                                           float inv_n) {
  const int b = blockIdx.y;
  const float c = -coef[b] * scale * inv_n;
  const size_t px = size_t(H) * W;
  for (size_t i = size_t(blockIdx.x) * 256 + threadIdx.x; i < px; i += size_t(gridDim.x) * 256) {
    const int y = int(i / W), x = int(i - size_t(y) * W);
    float xv, xbar;
    if (kFloatIn) ws_load_f32(static_cast<const float*>(img)[b * px + i], xv, xbar);
    else ws_load_u8(static_cast<const uint8_t*>(img)[b * px + i], xv, xbar);
    const bool in = (y >= crop) && (y < H - crop) && (x >= crop) && (x < W - crop);
    grad[b * px + i] = in ? c * (xv - xbar) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ finalize
// One warp per image; partial records are summed in double in a fixed order => beta_hat is bit-reproducible
// across runs, batch composition and GPU count.
__global__ void finalize_kernel(const float* __restrict__ partials, int records, int B, float npix, int clip,
                                int correct_bias, float* __restrict__ beta_hat, float* __restrict__ l1) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31;
  double s[kPartialSlots] = {0, 0, 0, 0};
  const float* src = partials + size_t(b) * records * kPartialSlots;
  for (int r = lane; r < records; r += 32)
    for (int k = 0; k < kPartialSlots; ++k) s[k] += double(src[r * kPartialSlots + k]);
  for (int k = 0; k < kPartialSlots; ++k)
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  if (lane == 0) {
    float beta = float(s[0] / s[1]);
    if (clip) beta = fmaxf(beta, 0.f);
    if (correct_bias) beta -= beta * float(s[3] / s[1]);  // estimate.py:128
    beta_hat[b] = beta;
    if (l1) l1[b] = float(s[2] / double(npix));
  }
}

// ------------------------------------------------------------------------------------------------ UniformDropout blend
// src/unet/model/unet.py:32-42: every pixel of the dropped channels is kept with probability 1 - drop_rate (mask = 1) and
// otherwise replaced by its KB prediction over the reflect-padded channel: out = x * mask + KB(x) * (1 - mask).
// x: (B,C,H,W) float32 in [0,1] or uint8 pixels (scaled by 1/255 first, src/unet/evaluate.py:45); mask: (B,1,H,W) of 0 / 1;
// chan_mask bit c set = channel c is a dropped channel; other channels are copied.
template <bool kFloatIn>
__global__ void __launch_bounds__(256) kb_blend_kernel(const void* __restrict__ x, const float* __restrict__ mask,
                                                       float* __restrict__ out, int B, int C, int H, int W,
                                                       unsigned chan_mask) {
  const size_t n = size_t(B) * C * H * W;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int xx = int(i % W), yy = int((i / W) % H);
    const int c = int((i / (size_t(W) * H)) % C), b = int(i / (size_t(W) * H * C));
    const size_t plane = i - size_t(yy) * W - xx;
    auto at = [&](int y, int x_) {
      y = y < 0 ? -y : (y >= H ? 2 * H - 2 - y : y);       // reflect (padding_mode of F.pad(..., mode='reflect'))
      x_ = x_ < 0 ? -x_ : (x_ >= W ? 2 * W - 2 - x_ : x_);
      const size_t o = plane + size_t(y) * W + x_;
      if constexpr (kFloatIn) return static_cast<const float*>(x)[o];
      else return __fdiv_rn(float(static_cast<const uint8_t*>(x)[o]), 255.f);
    };
    const float v = at(yy, xx);
    if (!((chan_mask >> c) & 1u)) { out[i] = v; continue; }
    const float cross = (at(yy - 1, xx) + at(yy + 1, xx)) + (at(yy, xx - 1) + at(yy, xx + 1));
    const float diag = (at(yy - 1, xx - 1) + at(yy - 1, xx + 1)) + (at(yy + 1, xx - 1) + at(yy + 1, xx + 1));
    const float kb = 0.5f * cross - 0.25f * diag;
    const float m = mask[(size_t(b) * H + yy) * W + xx];
    out[i] = v * m + kb * (1.f - m);
  }
}

// ------------------------------------------------------------------------------------------------ matrix-form residuals
// src/filters/evaluate.py:53-76 get_filter_residuals: rows of the N x 9 neighbour matrix (src/_defs/filters.py:39-69; columns
// x00 x01 x02 x12 x22 x21 x20 x10 | x11) against an 8 x 1 float64 filter: resid = x11 - sum_k coef[k] * row[k], in float64.
template <typename T>
__global__ void __launch_bounds__(256) residual_matvec_kernel(const T* __restrict__ mat, const double* __restrict__ coef,
                                                              double* __restrict__ resid, long long n) {
  double k[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) k[j] = coef[j];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T* r = mat + i * 9;
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += double(r[j]) * k[j];     // same left-to-right order as a row-times-column product
    resid[i] = double(r[8]) - acc;
  }
}

// ------------------------------------------------------------------------------------------------ debug pack/unpack
__global__ void pack_kernel(const float* __restrict__ src, Act dst) {
  const size_t n = size_t(dst.B) * dst.C * dst.H * dst.W;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int c = int(i % dst.C);
    const int x = int((i / dst.C) % dst.W);
    const int y = int((i / (size_t(dst.C) * dst.W)) % dst.H);
    const int b = int(i / (size_t(dst.C) * dst.W * dst.H));
    const float v = src[((size_t(b) * dst.C + c) * dst.H + y) * dst.W + x];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    int ys[3], xs[3];
    const int ny = halo_targets(y, dst.H, ys), nx = halo_targets(x, dst.W, xs);
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nx; ++ix) {
        const size_t off = ((size_t(b) * (dst.H + 2) + ys[iy]) * (dst.W + 2) + xs[ix]) * dst.C + c;
        dst.base[off] = h;
        dst.base[dst.plane + off] = l;
      }
  }
}

// with_halo: dst is (B,C,H+2,W+2) and receives the stored border too (lets tests check the reflect halo)
__global__ void unpack_kernel(Act src, float* __restrict__ dst, int with_halo) {
  const int Ho = src.H + (with_halo ? 2 : 0), Wo = src.W + (with_halo ? 2 : 0);
  const size_t n = size_t(src.B) * src.C * Ho * Wo;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int x = int(i % Wo);
    const int y = int((i / Wo) % Ho);
    const int c = int((i / (size_t(Wo) * Ho)) % src.C);
    const int b = int(i / (size_t(Wo) * Ho * src.C));
    const int sy = with_halo ? y : y + 1, sx = with_halo ? x : x + 1;
    const size_t off = ((size_t(b) * (src.H + 2) + sy) * (src.W + 2) + sx) * src.C + c;
    if (src.fmt == ACT_F16) dst[i] = __half2float(reinterpret_cast<const __half*>(src.base)[off]);
    else if (src.fmt == ACT_F16F8) {
      // fp16 value + the stored residual: byte (c % 16) of the a2s run of 16-channel group c / 16 in plane 1
      const uint8_t* p1 = reinterpret_cast<const uint8_t*>(src.base + src.plane + (off - c));
      const __nv_fp8_e4m3 r = *reinterpret_cast<const __nv_fp8_e4m3*>(p1 + (c >> 4) * 32 + (c & 15));
      dst[i] = __half2float(reinterpret_cast<const __half*>(src.base)[off]) + float(r) * (1.f / kF8ScaleA2);
    }
    else dst[i] = __bfloat162float(src.base[off]) + __bfloat162float(src.base[src.plane + off]);
  }
}

}  // namespace

cudaError_t launch_first_conv(const void* img, int img_kind, int cin, const float* w, const float* bias, Act out,
                              cudaStream_t stream) {
  const bool img_is_float = img_kind == 1;
  if (cin == 1) {
    if (out.C != 64 || out.fmt == ACT_F16) return cudaErrorInvalidValue;   // e11 is a 64-channel full-resolution map
    const int segs = (out.W + kE11Seg - 1) / kE11Seg;
    const int groups = (out.H + kE11Rows - 1) / kE11Rows;
    const int grid = out.B * groups * segs;
    const bool f8 = out.fmt == ACT_F16F8;
#define WSU_E11(KIN)                                                                                         \
  do {                                                                                                       \
    if (f8) first_conv_kernel<KIN, ACT_F16F8><<<grid, 256, 0, stream>>>(img, w, bias, out, segs, groups);   \
    else first_conv_kernel<KIN, ACT_SPLIT><<<grid, 256, 0, stream>>>(img, w, bias, out, segs, groups);      \
  } while (0)
    if (img_kind == 1) WSU_E11(1);
    else if (img_kind == 2) WSU_E11(2);
    else WSU_E11(0);
#undef WSU_E11
    return cudaGetLastError();
  }
  if (img_kind == 2) return cudaErrorInvalidValue;   // the LSB-difference input exists for single-channel images only
  const size_t threads = size_t(out.B) * out.H * out.W * 8;
  const int grid = int((threads + 255) / 256);
  const size_t smem = (64 * cin * 9 + 64) * sizeof(float);
  if (img_is_float)
    first_conv_generic_kernel<true><<<grid, 256, smem, stream>>>(img, cin, w, bias, out);
  else
    first_conv_generic_kernel<false><<<grid, 256, smem, stream>>>(img, cin, w, bias, out);
  return cudaGetLastError();
}

int filter_ws_strips(int H) { return (H - 2 + kStripRows - 1) / kStripRows; }

bool filter_ws_fast_ok(const void* img, int img_is_float, int W, int kind, int want_bias, const float* xhat_out) {
  return !img_is_float && !want_bias && !xhat_out && (kind == PRED_KB || kind == PRED_AVG) && (W % 4 == 0) &&
         (reinterpret_cast<uintptr_t>(img) % 4 == 0);
}
int filter_ws_fast_records(int H, int W) {
  return ((H - 2 + kFastRows - 1) / kFastRows) * ((W + kFastThreads * 4 - 1) / (kFastThreads * 4)) * 2;
}

template <int KIND>
static void launch_fast_kind(const uint8_t* img, int B, int H, int W, int weighted, float* partials, int strips, int xtiles,
                             cudaStream_t stream) {
  const int grid = B * strips * xtiles;
  if (weighted == WS_UNWEIGHTED)
    filter_ws_fast_kernel<KIND, WS_UNWEIGHTED><<<grid, kFastThreads, 0, stream>>>(img, H, W, partials, strips, xtiles);
  else if (weighted == WS_WEIGHTED)
    filter_ws_fast_kernel<KIND, WS_WEIGHTED><<<grid, kFastThreads, 0, stream>>>(img, H, W, partials, strips, xtiles);
  else
    filter_ws_fast_kernel<KIND, WS_ANTIWEIGHTED><<<grid, kFastThreads, 0, stream>>>(img, H, W, partials, strips, xtiles);
}

int filter_ws_packed_records(int H, int W) {
  return ((H - 2 + kPackRows - 1) / kPackRows) * ((W + kFastThreads * 4 - 1) / (kFastThreads * 4)) * 2;
}
cudaError_t launch_filter_ws_packed(const void* img, int B, int H, int W, int kind, float* partials, cudaStream_t stream) {
  const int strips = (H - 2 + kPackRows - 1) / kPackRows;
  const int xtiles = (W + kFastThreads * 4 - 1) / (kFastThreads * 4);
  const uint8_t* im = static_cast<const uint8_t*>(img);
  const int grid = B * strips * xtiles;
  if (kind == PRED_KB) filter_ws_packed_kernel<PRED_KB><<<grid, kFastThreads, 0, stream>>>(im, H, W, partials, strips, xtiles);
  else filter_ws_packed_kernel<PRED_AVG><<<grid, kFastThreads, 0, stream>>>(im, H, W, partials, strips, xtiles);
  return cudaGetLastError();
}

bool filter_ws_adjoint_ok(const void* img, int H, int W) {
  return (W % 16 == 0) && (reinterpret_cast<uintptr_t>(img) % 16 == 0) && H >= 3 &&
         double(H - 2) * double(W - 2) < 2147483648.0;
}
int filter_ws_adjoint_records(int H, int W) { return ((H + kAdjRows - 1) / kAdjRows) * ((W + 511) / 512) * 2; }
cudaError_t launch_filter_ws_adjoint(const void* img, int B, int H, int W, int kind, float* partials, cudaStream_t stream) {
  const int rstrips = (H + kAdjRows - 1) / kAdjRows, cstrips = (W + 511) / 512;
  const long long ctas = (long long)((B + kAdjWarps - 1) / kAdjWarps) * rstrips * cstrips;
  if (ctas > 0x7fffffffll) return cudaErrorInvalidValue;
  const uint8_t* im = static_cast<const uint8_t*>(img);
  const int grid = int(ctas), thr = kAdjWarps * 32;
  if (kind == PRED_KB) {
    if (cstrips > 1) filter_ws_adjoint_kernel<PRED_KB, true><<<grid, thr, 0, stream>>>(im, B, H, W, partials, rstrips, cstrips, 2u, 0xfffffffeu);
    else filter_ws_adjoint_kernel<PRED_KB, false><<<grid, thr, 0, stream>>>(im, B, H, W, partials, rstrips, cstrips, 2u, 0xfffffffeu);
  } else {
    if (cstrips > 1) filter_ws_adjoint_kernel<PRED_AVG, true><<<grid, thr, 0, stream>>>(im, B, H, W, partials, rstrips, cstrips, 2u, 0xfffffffeu);
    else filter_ws_adjoint_kernel<PRED_AVG, false><<<grid, thr, 0, stream>>>(im, B, H, W, partials, rstrips, cstrips, 2u, 0xfffffffeu);
  }
  return cudaGetLastError();
}

bool filter_ws_window_ok(const void* img, int H, int W) {
  return (W % 16 == 0) && (reinterpret_cast<uintptr_t>(img) % 16 == 0) && H >= 3;
}
int filter_ws_window_records(int H, int W) { return ((H - 2 + kWinRows - 1) / kWinRows) * ((W + 511) / 512) * 2; }

template <int KIND, int WEIGHTED, bool kL1>
static void launch_window_t(const uint8_t* im, int B, int H, int W, float* partials, int rstrips, int cstrips, int grid,
                            cudaStream_t stream) {
  if (cstrips > 1)
    filter_ws_window_kernel<KIND, WEIGHTED, kL1, true><<<grid, kWinWarps * 32, 0, stream>>>(im, B, H, W, partials, rstrips, cstrips);
  else
    filter_ws_window_kernel<KIND, WEIGHTED, kL1, false><<<grid, kWinWarps * 32, 0, stream>>>(im, B, H, W, partials, rstrips, cstrips);
}
template <int KIND>
static void launch_window_kind(const uint8_t* im, int B, int H, int W, int weighted, int want_l1, float* partials, int rstrips,
                               int cstrips, int grid, cudaStream_t st) {
  if (weighted == WS_UNWEIGHTED) launch_window_t<KIND, WS_UNWEIGHTED, true>(im, B, H, W, partials, rstrips, cstrips, grid, st);
  else if (weighted == WS_WEIGHTED) {
    if (want_l1) launch_window_t<KIND, WS_WEIGHTED, true>(im, B, H, W, partials, rstrips, cstrips, grid, st);
    else launch_window_t<KIND, WS_WEIGHTED, false>(im, B, H, W, partials, rstrips, cstrips, grid, st);
  } else {
    if (want_l1) launch_window_t<KIND, WS_ANTIWEIGHTED, true>(im, B, H, W, partials, rstrips, cstrips, grid, st);
    else launch_window_t<KIND, WS_ANTIWEIGHTED, false>(im, B, H, W, partials, rstrips, cstrips, grid, st);
  }
}
cudaError_t launch_filter_ws_window(const void* img, int B, int H, int W, int kind, int weighted, int want_l1, float* partials,
                                    cudaStream_t stream) {
  const int rstrips = (H - 2 + kWinRows - 1) / kWinRows, cstrips = (W + 511) / 512;
  const long long ctas = (long long)((B + kWinWarps - 1) / kWinWarps) * rstrips * cstrips;
  if (ctas > 0x7fffffffll) return cudaErrorInvalidValue;
  const uint8_t* im = static_cast<const uint8_t*>(img);
  if (kind == PRED_KB) launch_window_kind<PRED_KB>(im, B, H, W, weighted, want_l1, partials, rstrips, cstrips, int(ctas), stream);
  else launch_window_kind<PRED_AVG>(im, B, H, W, weighted, want_l1, partials, rstrips, cstrips, int(ctas), stream);
  return cudaGetLastError();
}

cudaError_t launch_filter_ws_fast(const void* img, int B, int H, int W, int kind, int weighted, float* partials,
                                  cudaStream_t stream) {
  const int strips = (H - 2 + kFastRows - 1) / kFastRows;
  const int xtiles = (W + kFastThreads * 4 - 1) / (kFastThreads * 4);
  const uint8_t* im = static_cast<const uint8_t*>(img);
  if (kind == PRED_KB) launch_fast_kind<PRED_KB>(im, B, H, W, weighted, partials, strips, xtiles, stream);
  else launch_fast_kind<PRED_AVG>(im, B, H, W, weighted, partials, strips, xtiles, stream);
  return cudaGetLastError();
}

cudaError_t launch_filter_ws(const void* img, int img_is_float, int B, int H, int W, int kind, int weighted, int want_bias,
                             float* xhat_out, float* partials, cudaStream_t stream) {
  const int strips = filter_ws_strips(H);
  const size_t smem = size_t(kStripRows + 2) * W * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  if (img_is_float) {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(filter_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    filter_ws_kernel<true><<<B * strips, kFThreads, smem, stream>>>(img, B, H, W, kind, weighted, want_bias, xhat_out,
                                                                    partials, strips);
  } else {
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(filter_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    filter_ws_kernel<false><<<B * strips, kFThreads, smem, stream>>>(img, B, H, W, kind, weighted, want_bias, xhat_out,
                                                                     partials, strips);
  }
  return cudaGetLastError();
}

cudaError_t launch_ws_from_pred(const void* img, int img_is_float, const float* xhat, int xhat_cropped, const float* xbias,
                                int B, int H, int W, int weighted, int crop, float* partials, int chunks,
                                cudaStream_t stream) {
  if (img_is_float)
    ws_from_pred_kernel<true><<<B * chunks, 256, 0, stream>>>(img, xhat, xhat_cropped, xbias, B, H, W, weighted, crop,
                                                              partials, chunks);
  else
    ws_from_pred_kernel<false><<<B * chunks, 256, 0, stream>>>(img, xhat, xhat_cropped, xbias, B, H, W, weighted, crop,
                                                               partials, chunks);
  return cudaGetLastError();
}

cudaError_t launch_ws_grad_pred(const void* img, int img_is_float, const float* coef, float* grad, int B, int H, int W,
                                int crop, float scale, cudaStream_t stream) {
  const size_t px = size_t(H) * W;
  const dim3 grid(unsigned(std::min<size_t>((px + 255) / 256, 1024)), unsigned(B));
  const float inv_n = 1.f / (float(H - 2 * crop) * float(W - 2 * crop));
  if (img_is_float) ws_grad_pred_kernel<true><<<grid, 256, 0, stream>>>(img, coef, grad, H, W, crop, scale, inv_n);
  else ws_grad_pred_kernel<false><<<grid, 256, 0, stream>>>(img, coef, grad, H, W, crop, scale, inv_n);
  return cudaGetLastError();
}

cudaError_t launch_finalize(const float* partials, int records, int B, float npix, int clip, int correct_bias,
                            float* beta_hat, float* l1, cudaStream_t stream) {
  const int warps_per_block = 4;
  finalize_kernel<<<(B + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      partials, records, B, npix, clip, correct_bias, beta_hat, l1);
  return cudaGetLastError();
}

cudaError_t launch_kb_blend(const void* x, int x_is_float, const float* mask, float* out, int B, int C, int H, int W,
                            unsigned chan_mask, cudaStream_t stream) {
  const size_t n = size_t(B) * C * H * W;
  const int grid = int(std::min<size_t>((n + 255) / 256, 148 * 16));
  if (x_is_float) kb_blend_kernel<true><<<grid, 256, 0, stream>>>(x, mask, out, B, C, H, W, chan_mask);
  else kb_blend_kernel<false><<<grid, 256, 0, stream>>>(x, mask, out, B, C, H, W, chan_mask);
  return cudaGetLastError();
}

cudaError_t launch_residual_matvec(const void* mat, int mat_dtype, const double* coef, double* resid, long long n,
                                   cudaStream_t stream) {
  const int grid = int(std::min<long long>((n + 255) / 256, 148 * 16));
  if (mat_dtype == 0) residual_matvec_kernel<uint8_t><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(mat), coef, resid, n);
  else if (mat_dtype == 1) residual_matvec_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(mat), coef, resid, n);
  else residual_matvec_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(mat), coef, resid, n);
  return cudaGetLastError();
}

cudaError_t launch_pack(const float* src, Act dst, cudaStream_t stream) {
  pack_kernel<<<1024, 256, 0, stream>>>(src, dst);
  return cudaGetLastError();
}
cudaError_t launch_unpack(Act src, float* dst, int with_halo, cudaStream_t stream) {
  unpack_kernel<<<1024, 256, 0, stream>>>(src, dst, with_halo);
  return cudaGetLastError();
}

}  // namespace wsu
