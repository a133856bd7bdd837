// Parameter block + host launcher of the tcgen05 implicit-GEMM convolution (conv_mma.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "wsu_common.cuh"

namespace wsu {

enum : int { EPI_ACT = 0, EPI_HEAD = 1 };

// One launch = one layer of UNet.forward (src/unet/model/unet.py:141-189) over a micro-batch.
// n / d for 0 <= n with n * d < 2^32 by one multiply-high (the per-box index arithmetic of the persistent kernels ran two to
// three full integer divisions per box and warp: ~100 of the ~800 instructions an epilogue warp spends on a box)
struct FastDiv {
  uint32_t mul;    // floor(2^32 / d) + 1
  uint32_t mode;   // 0: multiply-high, 1: d == 1, 2: generic division (ranges too large for the shortcut)
  uint32_t d;
};
struct alignas(64) ConvParams {
  CUtensorMap tmapA0;  // first concat source  (upsampled path, unet.py:178 puts it first)
  CUtensorMap tmapA1;  // second concat source (skip); unused when cblocks == cblocks0
  CUtensorMap tmapH0;  // same sources with the haloed box (64 ch, 10, 18, 1 img, 2 planes) of the halo kernel
  CUtensorMap tmapH1;
  CUtensorMap tmapW;   // packed weights as rows of 128 B, box = 64 rows (CTA-pair kernels: each CTA loads its half of a tile)
  CUtensorMap tmapW32; // same, box = 32 rows (stacked Cout = 64 pair kernel: half of the 64-row Whi tile)
  CUtensorMap tmapOut; // destination map, box (32 ch, 8 px, 4 rows, 1 img, 1 plane), SWIZZLE_64B: TMA stores of the halo kernels
  int tma_store;       // 1: interior boxes leave through tmapOut
  FastDiv fd_sub_x, fd_sub_y, fd_n_tiles;   // divisions by sub_x, sub_y, n_tiles
  int w_resident;      // fp16 + fp8 layers with one input channel block keep all nine taps' weights in shared memory
  const uint8_t* wpack;  // packed + pre-swizzled split-bf16 weights, see pack_conv_weights()
  const float* bias;     // [Cout]
  int cblocks0;          // 64-channel blocks taken from source 0
  int cblocks;           // total 64-channel blocks (K = 64 * cblocks * ntaps)
  int ntaps;             // 9 (3x3 conv) or 1 (one phase of the 2x2 stride-2 transposed conv)
  int tap_dx[9], tap_dy[9];  // tap offsets in padded (halo) coordinates
  int npos;              // 1, or 4 output phases for the transposed conv
  int n_tiles;           // Cout / N_TILE
  int cout;              // Cout
  int B, H, W;           // logical input dims (= GEMM pixel grid)
  int TW, TH;            // sub-tile box (TW*TH == 128 pixels)
  int tiles_x, tiles_y;  // super-tiles (M_SUB sub-tiles side by side in x) per image
  int total_tiles;
  // halo kernel: 8x16-pixel boxes enumerated row-major over (image, box row, box column)
  int sub_x, sub_y;      // boxes per image row / column
  int total_sub;         // B * sub_y * sub_x
  int total_items;       // ceil(total_sub / M_SUB) * n_tiles
  // fused first layer (halo kernel, Cin = 64 only): when fuse_img != null the input boxes are relu(e11(image)) computed
  // in the kernel from the image and e11's fp32 weights [64][9] / bias [64] instead of being loaded from tmapH0
  const void* fuse_img;
  int fuse_img_is_float;
  const float* fuse_w;
  const float* fuse_b;
  int l2_prefetch;       // halo kernels: warm L2 with the boxes of the CTA's next work item
  int dbg;               // experiments only (env WSU_DBG): bit 0 skips the pooled output, bit 1 the main output stores, bit 3 the e4m3
                         // correction MMA of the fp16 + fp8 scheme
  int a_collector;       // Cout >= 128 layers: A_hi stays in the tensor core's A collector for its second product
  int terms;             // MMAs per MAC: 3 (split-bf16 inputs) or 2 / 1 (ONE fp16 input plane; CTA-pair kernel, Cout >= 128)
  int src0_f16;          // Cout = 64 halo kernels: the channel blocks of source 0 are ONE fp16 plane; 1: against fp16 (hi, lo) weights
                         // (one stacked N=128 MMA per K step), 2: against fp16 weights (one N=64 MMA, CTA-pair kernel only)
  int f8_blocks;         // Cout = 64 CTA-pair kernel: the remaining blocks are ACT_F16F8 maps (fp16 main + one e4m3 correction MMA)
  float corr_scale;      // what the correction columns [64,128) of a stacked accumulator are multiplied by (1 for split-bf16)
  // EPI_ACT
  int relu;
  int upsample;          // 1: write phase (pos>>1, pos&1) of a 2x upsampled map (ConvTranspose2d k=2,s=2)
  Act out;               // destination (dims = output dims)
  int do_pool;           // also write MaxPool2d(2,2) of the output (unet.py:144,149); needs TW == 16
  Act pool;
  // EPI_HEAD: 1x1 outconv + sigmoid (unet.py:189) and the WS reduction
  const void* img;       // [B][H][W] uint8 or float32 pixels the residual is taken against (may be null)
  int img_is_float;
  float wout[64];
  float bout;
  float* yhat;           // [B][H][W] sigmoid output in (0,1), may be null
  float* partials;       // [B][tiles_per_img][4 warps][kPartialSlots]
  int weighted;          // WS_* mode
  int crop;              // 1: interior only (estimate.py:113-114), 0: whole image (losses.py:57-60)
  int bias_pass;         // 1: this pass ran on the LSB-difference image; the head adds only sum w (x - x_bar) x_bias to slot 3 of
                         // the partial records the first pass wrote (estimate.py:126-128)
};

// Bytes of one packed weight chunk (hi + lo tile) for an N_TILE-wide tile and one 64-deep K block.
constexpr int wchunk_bytes(int n_tile) { return n_tile * 64 * 2 * 2; }

cudaError_t launch_conv_mma(const ConvParams& p, int n_tile, int epi, int num_sms, cudaStream_t stream);
// 3x3 convolutions only: one haloed TMA box per (box, channel block) feeds all 9 taps
cudaError_t launch_conv_halo(const ConvParams& p, int n_tile, int epi, int num_sms, cudaStream_t stream);
cudaError_t launch_conv_halo2(const ConvParams& p, int n_tile, int epi, int num_sms, cudaStream_t stream);  // CTA pairs
constexpr int halo_msub(int) { return 2; }
constexpr int kHaloTW = 8, kHaloTH = 16;
// Transposed convolution with shared-memory-resident weights: all 4 output phases stacked along N (N_TILE = 4 * co_t).
struct alignas(64) UpconvParams {
  CUtensorMap tmapA;      // source activation, box (64 ch, 16, 8, 1 img, 2 planes)
  const uint8_t* wres;    // [n_tiles][cblocks][hi tile | lo tile], rows r -> (phase r / co_t, channel nt*co_t + r % co_t)
  const float* bias;      // [Cout]
  int cblocks, n_tiles, co_t;
  int terms;              // MMAs per MAC: 3 (split-bf16 input) or 2 / 1 (one fp16 input plane)
  int B, H, W;            // input dims
  int tiles_x, tiles_y, total_boxes;
  Act out;                // (2H, 2W) destination
  CUtensorMap tmapOut;    // destination for TMA stores of one output phase: box (32 ch, 16 px, 2 rows), element strides (1, 2, 2)
  int tma_store;          // 1: interior boxes leave through tmapOut
  int zero_bias;          // 1: the bias was folded into the consuming layer (reduced plans): nothing to add
  FastDiv fd_tiles_x, fd_tiles_y, fd_co_t;   // divisions by tiles_x, tiles_y, co_t
};
cudaError_t launch_upconv_res(const UpconvParams& p, int n_tile, int num_sms, cudaStream_t stream);
constexpr int kUpconvResBytes = 131072;  // weight bytes that must fit: cblocks * N_TILE * 256
cudaError_t conv_mma_init();  // sets max dynamic shared memory on every instantiation

}  // namespace wsu
