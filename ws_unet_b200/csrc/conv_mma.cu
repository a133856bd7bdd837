// tcgen05 / TMEM / TMA implicit-GEMM convolution for the UNet pixel predictor.
//
// Replaces (from scratch, B200-native) the reference's nn.Conv2d(3x3, reflect) + F.relu blocks, the
// nn.ConvTranspose2d(k=2,s=2) up-convolutions, torch.cat skip concatenation, nn.MaxPool2d(2,2) and the
// 1x1 outconv + sigmoid head of UNet.forward (src/unet/model/unet.py:137-189), and fuses the WS residual
// reduction of src/unet/evaluate.py:128-132 / src/ws/estimate.py:113-121 into the last layer's epilogue.
//
// GEMM view: M = pixels (128-pixel TWxTH boxes), N = Cout tile, K = 64-channel block x tap.
//   A operand: one 5-D TMA box load {64 ch, TW, TH, 1 img, 2 planes} of the haloed split-bf16 activation at the
//              tap's offset -> two K-major SWIZZLE_128B tiles (hi, lo) of 128 rows x 128 B.
//   B operand: pre-swizzled packed weights, one 1-D bulk copy per (tap, channel block) -> (hi, lo) tiles.
//   D: fp32 in TMEM, M_SUB accumulators per stage (weights are reused across M_SUB pixel boxes), two stages so
//      the epilogue of tile i overlaps the MMAs of tile i+1.
//   Precision: D += Ahi*Whi + Alo*Whi + Ahi*Wlo  (3 bf16 MMAs per algorithmic MAC).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..5 = epilogue.
#include <type_traits>

#include "conv_mma.h"
#include "ptx.cuh"
#include "ws_math.cuh"

namespace wsu {

namespace {

constexpr int kThreads = 320;          // warps: 0 TMA, 1 MMA, 2..9 epilogue (two groups of four lane quadrants)
constexpr int kABytes = 2 * 128 * 128;  // hi + lo tile of 128 pixels x 64 channels
constexpr int kAccCols = 256;            // TMEM columns per accumulator stage
constexpr int kScratchPitch = 64;                       // dense 64-byte rows; the 16-byte piece index is XOR-swizzled with
                                                        // (row >> 1) & 3, which makes BOTH the lane-major writes and the
                                                        // 4-lanes-per-pixel reads bank-conflict free (an 80-byte padded
                                                        // pitch left the reads 2-way conflicted: the LSU data pipe, not
                                                        // DRAM, was what bounded the up-convolutions)
constexpr int kScratchPerWarp = 32 * kScratchPitch;     // 2048 B of epilogue staging per warp
constexpr int kScratchBytes = 8 * kScratchPerWarp;      // eight epilogue warps

template <int N_TILE>
struct Cfg {
  static constexpr int M_SUB = kAccCols / N_TILE;          // pixel boxes sharing one weight chunk
  static constexpr int W_BYTES = wchunk_bytes(N_TILE);
  static constexpr int SA = 4;                             // A ring depth
  static constexpr int SW = (N_TILE == 64) ? 3 : 2;        // W ring depth
  static constexpr int SMEM = SA * kABytes + SW * W_BYTES + kScratchBytes + 1024 /*align*/ + 4096 /*bias*/ + 256 /*barriers*/;
};

struct TileCoord {
  int b, y0, x0, nt, pos, tile_in_img;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int idx, int m_sub) {
  TileCoord t;
  t.pos = idx % p.npos;
  idx /= p.npos;
  t.nt = idx % p.n_tiles;
  idx /= p.n_tiles;
  const int tx = idx % p.tiles_x;
  idx /= p.tiles_x;
  const int ty = idx % p.tiles_y;
  t.b = idx / p.tiles_y;
  t.x0 = tx * p.TW * m_sub;
  t.y0 = ty * p.TH;
  t.tile_in_img = ty * p.tiles_x + tx;
  return t;
}

// store 32 channels (hi words h[16], lo words l[16]) of one pixel to every halo target of (oy, ox)
__device__ __forceinline__ void store_pixel32(const Act& o, int b, int oy, int ox, int c0, const uint32_t (&h)[16],
                                              const uint32_t (&l)[16]) {
  int ys[3], xs[3];
  const int ny = halo_targets(oy, o.H, ys);
  const int nx = halo_targets(ox, o.W, xs);
  for (int iy = 0; iy < ny; ++iy) {
    for (int ix = 0; ix < nx; ++ix) {
      const size_t off = ((size_t(b) * (o.H + 2) + ys[iy]) * (o.W + 2) + xs[ix]) * o.C + c0;
      uint4* ph = reinterpret_cast<uint4*>(o.base + off);
      uint4* pl = reinterpret_cast<uint4*>(o.base + o.plane + off);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ph[q] = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        pl[q] = make_uint4(l[4 * q], l[4 * q + 1], l[4 * q + 2], l[4 * q + 3]);
      }
    }
  }
}

// store 8 channels (one 16-byte piece per plane) of one pixel to every halo target of (oy, ox)
__device__ __forceinline__ void store_pixel8(const Act& o, int b, int oy, int ox, int c0, const uint32_t (&h)[4],
                                             const uint32_t (&l)[4], int fmt) {
  const uint4 vh = make_uint4(h[0], h[1], h[2], h[3]), vl = make_uint4(l[0], l[1], l[2], l[3]);
  const size_t off = ((size_t(b) * (o.H + 2) + (oy + 1)) * (o.W + 2) + (ox + 1)) * o.C + c0;
  const bool two = fmt != ACT_F16;   // fp16 maps have one plane (h holds the fp16 words)
  *reinterpret_cast<uint4*>(o.base + off) = vh;
  if (two) *reinterpret_cast<uint4*>(o.base + o.plane + off) = vl;
  if (oy == 1 || ox == 1 || oy == o.H - 2 || ox == o.W - 2) {   // reflect-halo duplicates (border pixels only)
    int ys[3], xs[3];
    const int ny = halo_targets(oy, o.H, ys), nx = halo_targets(ox, o.W, xs);
    for (int iy = 0; iy < ny; ++iy)
      for (int ix = 0; ix < nx; ++ix) {
        if (iy == 0 && ix == 0) continue;   // (oy+1, ox+1) itself was written above
        const size_t d = ((size_t(b) * (o.H + 2) + ys[iy]) * (o.W + 2) + xs[ix]) * o.C + c0;
        *reinterpret_cast<uint4*>(o.base + d) = vh;
        if (two) *reinterpret_cast<uint4*>(o.base + o.plane + d) = vl;
      }
  }
}

// Epilogue of one 128-pixel box: thread owns pixel (y, x) = TMEM lane; tbase addresses its accumulator columns.
// pool_xor: lane distance of the vertical 2x2-pool partner (= box width in pixels).
// Geometry of the 128-pixel box a warp is finishing: row r of the box <-> pixel (y0 + (r >> tw_shift), x0 + (r & mask)).
struct BoxGeo {
  int b, y0, x0, tw_shift, H, W;  // H, W: bounds of the GEMM pixel grid (input dims)
  int up, pos;                    // up = 1: output pixel (2y + pos/2, 2x + pos%2) of a 2x upsampled map
  __device__ __forceinline__ bool pixel(int row, int& oy, int& ox) const {
    const int y = y0 + (row >> tw_shift), x = x0 + (row & ((1 << tw_shift) - 1));
    oy = up ? 2 * y + (pos >> 1) : y;
    ox = up ? 2 * x + (pos & 1) : x;
    return (y < H) && (x < W);
  }
};

// Where this lane's four (pixel, 16-byte piece) stores of a chunk go: computed once per box, reused by every
// 32-channel chunk and both planes. Lane l re-reads pixel q = (l >> 2) + 8 i, piece l & 3 from the staging buffer.
// `border`: at least one of the warp's pixels also owns reflect-halo slots -> the generic path writes the duplicates.
struct StoreMap {
  uint32_t valid;  // bit i: pixel q = (lane >> 2) + 8 i of the warp's 32 lies inside the map
  bool border;
  bool full;       // every pixel of the warp's 32 lies inside the map (no ragged edge)
};
__device__ __forceinline__ StoreMap make_store_map(const Act& o, const BoxGeo& g, int lane, int row0) {
  StoreMap m;
  m.valid = 0;
  bool edge = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = (lane >> 2) + 8 * i;
    int oy, ox;
    const bool ok = g.pixel(row0 + q, oy, ox);
    m.valid |= uint32_t(ok) << i;
    edge |= ok && (oy == 1 || ox == 1 || oy == o.H - 2 || ox == o.W - 2);
  }
  m.border = __any_sync(0xffffffffu, edge);
  m.full = __all_sync(0xffffffffu, m.valid == 0xfu);
  return m;
}
// element offset of (pixel q of the warp, channel 0) in a plane - only the per-lane store path needs it (interior boxes leave
// through TMA stores), so it is formed there instead of once per box for every warp
__device__ __forceinline__ size_t store_offset(const Act& o, const BoxGeo& g, int row) {
  int oy, ox;
  g.pixel(row, oy, ox);
  return ((size_t(g.b) * (o.H + 2) + (oy + 1)) * (o.W + 2) + (ox + 1)) * o.C;
}

// MaxPool2d(2, 2) of a chunk whose fp16 plane sits in the staging buffer (rows = the warp's 4 x 8 pixels, 64 bytes each, pieces
// XOR-swizzled as written by store_chunk_coalesced): lane (pp = lane / 4, piece = lane % 4) reads the piece of the four source
// pixels of pooled pixel pp and takes the maximum on packed halves - max commutes with the fp16 rounding and with ReLU, so this
// equals converting the pooled fp32 values. 4 LDS + 12 HMNMX2 + 1 store per lane instead of 24 shuffles, 24 FMNMX and 48 selects.
__device__ __forceinline__ void pool_from_staging(const Act& pool, const uint8_t* scratch, int lane, int row0, int c0,
                                                  const BoxGeo& g) {
  const int pp = lane >> 2, piece = lane & 3;
  const int py = pp >> 2, px = pp & 3;
  uint4 m;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = (2 * py + (k >> 1)) * 8 + 2 * px + (k & 1);
    const uint4 v = *reinterpret_cast<const uint4*>(scratch + r * kScratchPitch + ((piece ^ ((r >> 1) & 3)) << 4));
    if (k == 0) {
      m = v;
    } else {
      auto hmax = [](uint32_t a, uint32_t b) {
        const __half2 r2 = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
        return *reinterpret_cast<const uint32_t*>(&r2);
      };
      m = make_uint4(hmax(m.x, v.x), hmax(m.y, v.y), hmax(m.z, v.z), hmax(m.w, v.w));
    }
  }
  const int y = g.y0 + (row0 >> 3) + 2 * py, x = g.x0 + 2 * px;   // top-left source pixel (H, W are even where a pool exists)
  if (y < g.H && x < g.W) {
    const uint32_t h4[4] = {m.x, m.y, m.z, m.w}, l4[4] = {0u, 0u, 0u, 0u};
    store_pixel8(pool, g.b, y >> 1, x >> 1, c0 + piece * 8, h4, l4, ACT_F16);
  }
}

// Warp-cooperative store of one 32-channel chunk of the warp's 32 pixels. Registers hold "my pixel, 32 channels";
// a 64-byte row per pixel is staged in shared memory and re-read so that 4 consecutive lanes write the 4 consecutive
// 16-byte pieces of one pixel: a store instruction then touches 8 half-lines instead of 32 different lines
// (uncoalesced 16-byte stores made the LSU, not the tensor pipe, the limiter of the store-heavy layers).
// `tm` (may be null): tensor map of the output with box {32 ch, 8 px, 4 rows, 1, 1} and SWIZZLE_64B, which is exactly the
// staging layout below. Interior boxes of the 8-pixel-wide halo kernels then leave through ONE TMA tensor store per plane,
// issued by one lane, instead of 4 shared-memory reads + 4 global stores per lane: a third less traffic on the SM's
// L1 / shared-memory path, which the MMA operand reads of these kernels already saturate.
__device__ __forceinline__ void store_chunk_coalesced(const Act& o, uint8_t* scratch, int lane, int row0, int c0,
                                                      const uint32_t (&h)[16], const uint32_t (&l)[16], const BoxGeo& g,
                                                      const StoreMap& map, const CUtensorMap* tm = nullptr, int dbg = 0,
                                                      int fmt = -1, const Act* pool = nullptr) {
  const int piece = lane & 3;
  const int planes = (fmt >= 0 ? fmt : o.fmt) == ACT_F16 ? 1 : 2;   // fp16 maps: h holds 32 fp16 channels = the same 64-byte row
  if ((dbg & 32) && !map.border && map.full) {   // WSU_DBG=32 (experiment): every lane stores its own pixel, no staging
    int oy, ox;
    g.pixel(row0 + lane, oy, ox);
    const size_t off = ((size_t(g.b) * (o.H + 2) + (oy + 1)) * (o.W + 2) + (ox + 1)) * o.C + c0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      *reinterpret_cast<uint4*>(o.base + off + q * 8) = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
      if (planes == 2) *reinterpret_cast<uint4*>(o.base + o.plane + off + q * 8) = make_uint4(l[4 * q], l[4 * q + 1], l[4 * q + 2], l[4 * q + 3]);
    }
    return;
  }
  // (the 8-pixel-wide boxes of the 3x3 kernels with the plain map; the 16-pixel-wide boxes of the up-convolution with the
  // element-strided map of one output phase)
  const bool use_tma = tm != nullptr && !map.border && map.full && (g.up ? g.tw_shift == 4 : g.tw_shift == 3);
#pragma unroll
  for (int plane = 0; plane < 2; ++plane) {
    if (plane >= planes) break;
    if (tm != nullptr && lane == 0) tma_store_wait_read();   // a previous TMA store out of this staging buffer has read it
    __syncwarp();
    uint4* mine = reinterpret_cast<uint4*>(scratch + lane * kScratchPitch);
    const int wsw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      mine[q ^ wsw] = plane ? make_uint4(l[4 * q], l[4 * q + 1], l[4 * q + 2], l[4 * q + 3])
                            : make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    if (use_tma) {
      fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) {
        // padded coordinates: pixel (y, x) lives at (y + 1, x + 1); the warp's 32 rows are box rows (row0 >> 3) .. + 3
        if (g.up) tma_store_5d(tm, scratch, c0, 2 * g.x0 + (g.pos & 1) + 1, 2 * (g.y0 + (row0 >> 4)) + (g.pos >> 1) + 1, g.b, plane);
        else tma_store_5d(tm, scratch, c0, g.x0 + 1, g.y0 + (row0 >> 3) + 1, g.b, plane);
        tma_store_commit();
      }
      if (plane == 0 && pool != nullptr) pool_from_staging(*pool, scratch, lane, row0, c0, g);
      continue;
    }
    __syncwarp();
    if (plane == 0 && pool != nullptr) pool_from_staging(*pool, scratch, lane, row0, c0, g);
    __nv_bfloat16* base = o.base + (plane ? o.plane : 0) + c0 + piece * 8;
    if (!map.border) {
      // interior box: one store per (pixel, piece), addresses from the per-box map
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = (lane >> 2) + 8 * i;
        const uint4 v = *reinterpret_cast<const uint4*>(scratch + q * kScratchPitch + ((piece ^ ((q >> 1) & 3)) << 4));
        if (((map.valid >> i) & 1) && !(dbg & 16)) *reinterpret_cast<uint4*>(base + store_offset(o, g, row0 + q)) = v;   // WSU_DBG=16: staging without the global stores
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int q = (lane >> 2) + 8 * i;
        const uint4 v = *reinterpret_cast<const uint4*>(scratch + q * kScratchPitch + ((piece ^ ((q >> 1) & 3)) << 4));
        int oy, ox;
        if (g.pixel(row0 + q, oy, ox)) {
          int ys[3], xs[3];
          const int ny = halo_targets(oy, o.H, ys), nx = halo_targets(ox, o.W, xs);
          for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix) {
              const size_t off = ((size_t(g.b) * (o.H + 2) + ys[iy]) * (o.W + 2) + xs[ix]) * o.C;
              *reinterpret_cast<uint4*>(base + off) = v;
            }
        }
      }
    }
  }
}

// STACKED (Cout = 64 layers of the halo kernel): the accumulator is 128 columns wide, columns [0,64) hold
// (Ahi + Alo) * Whi and columns [64,128) hold Ahi * Wlo of the same 64 output channels; they are summed here.
// Under the fp16 + fp8 scheme the second half holds the correction sum scaled by a power of two (corr_scale undoes it).
template <int N_TILE, int EPI, bool STACKED>
__device__ __forceinline__ void load_acc32(uint32_t taddr, float (&f)[32], float corr_scale) {
  uint32_t v[32];
  tmem_ld32(taddr, v);
  if constexpr (STACKED) {
    uint32_t w[32];
    tmem_ld32(taddr + 64, w);
    tmem_ld_wait();
    const uint64_t cs2 = pack_f32x2(corr_scale, corr_scale);
#pragma unroll
    for (int i = 0; i < 16; ++i)   // two columns per FFMA2 (the same IEEE fma per lane as the scalar form)
      unpack_f32x2(fma2(pack_u32x2(w[2 * i], w[2 * i + 1]), cs2, pack_u32x2(v[2 * i], v[2 * i + 1])), f[2 * i], f[2 * i + 1]);
  } else {
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
  }
}

// FMT >= 0: the output format (and ACT_F16 for the pooled map) is known at compile time - the kernels of the reduced plans
// write one format only, and the conversions of the others need not be compiled into them (registers, instruction cache)
template <int N_TILE, int EPI, bool STACKED = false, int FMT = -1>
__device__ __forceinline__ void epilogue_box(const ConvParams& p, const float* sBias, uint32_t tbase, int b, int y, int x,
                                       bool valid, int nt, int pos, int tx, int ty, int pool_xor, WsAcc& acc,
                                       const BoxGeo& geo, uint8_t* scratch, int lane, int row0) {
  if (p.dbg & 4) return;   // WSU_DBG=4: no epilogue work at all (timing experiment: what the MMA pipeline alone takes)
  if constexpr (EPI == EPI_ACT) {
    const StoreMap smap = make_store_map(p.out, geo, lane, row0);
    const int fmt = FMT >= 0 ? FMT : p.out.fmt;
    const int pool_fmt = FMT >= 0 ? int(ACT_F16) : p.pool.fmt;
#pragma unroll 1
    for (int cc = 0; cc < N_TILE / 32; ++cc) {
      const int n0 = nt * N_TILE + cc * 32;
      float f[32];
      load_acc32<N_TILE, EPI, STACKED>(tbase + cc * 32, f, (p.dbg & 8) ? 0.f : p.corr_scale);   // WSU_DBG=8: correction MMA off (diagnostic)
#pragma unroll
      for (int i = 0; i < 8; ++i) {   // 128-bit broadcast loads: 8 instead of 32 shared-memory wavefronts per chunk; packed adds
        const float4 bq = *reinterpret_cast<const float4*>(sBias + n0 + 4 * i);
        unpack_f32x2(add2(pack_f32x2(f[4 * i], f[4 * i + 1]), pack_f32x2(bq.x, bq.y)), f[4 * i], f[4 * i + 1]);
        unpack_f32x2(add2(pack_f32x2(f[4 * i + 2], f[4 * i + 3]), pack_f32x2(bq.z, bq.w)), f[4 * i + 2], f[4 * i + 3]);
      }
      // fp16 maps: ReLU rides on the conversion (cvt.rn.relu) - also for the pooled map, max-pool and ReLU commute
      const bool relu_in_cvt = p.relu && fmt == ACT_F16 && (!p.do_pool || pool_fmt == ACT_F16);
      if (p.relu && !relu_in_cvt) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
      }
      uint32_t h[16], l[16];
      if (fmt == ACT_F16) {
        if (relu_in_cvt) {
#pragma unroll
          for (int i = 0; i < 16; ++i) { h[i] = cvt_f16x2_relu(f[2 * i], f[2 * i + 1]); l[i] = 0u; }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { h[i] = cvt_f16x2(f[2 * i], f[2 * i + 1]); l[i] = 0u; }
        }
      } else if (fmt == ACT_F16F8) {
        pack_f16f8_32(f, h, l);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) split_pack2(f[2 * i], f[2 * i + 1], h[i], l[i]);
      }
      // fp16 / fp16 + e4m3 maps with an fp16 pooled map: the pool is taken from the staged fp16 plane inside the store
      const bool staged_pool = p.do_pool && !(p.dbg & 3) && pool_fmt == ACT_F16 && (fmt == ACT_F16 || fmt == ACT_F16F8) && p.relu &&
                               geo.tw_shift == 3 && !geo.up;
      if (!(p.dbg & 2)) store_chunk_coalesced(p.out, scratch, lane, row0, n0, h, l, geo, smap, p.tma_store ? &p.tmapOut : nullptr, p.dbg, fmt,
                                              staged_pool ? &p.pool : nullptr);
      if (p.do_pool && !(p.dbg & 1) && !staged_pool) {
        // 2x2 max over (x^1, y^1): with TW == 16 both partners live in this warp (lane^1, lane^16). Each exchange moves
        // only the half the partner will keep, so the four lanes of a quad end up with 8 channels each of the pooled
        // pixel (24 shuffles per chunk instead of 64) and every lane stores one 16-byte piece per plane.
        const bool ox1 = tx & 1, oy1 = ty & 1;
        float m16[16], m8[8];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float keep = ox1 ? f[16 + i] : f[i], send = ox1 ? f[i] : f[16 + i];
          m16[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float keep = oy1 ? m16[8 + i] : m16[i], send = oy1 ? m16[i] : m16[8 + i];
          m8[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, pool_xor));
        }
        if (valid) {
          uint32_t h4[4], l4[4];
          if (pool_fmt == ACT_F16) {   // max commutes with the (monotone) rounding: pool(fp16(v)) == fp16(pool(v))
            if (relu_in_cvt) {
#pragma unroll
              for (int i = 0; i < 4; ++i) { h4[i] = cvt_f16x2_relu(m8[2 * i], m8[2 * i + 1]); l4[i] = 0u; }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) { h4[i] = cvt_f16x2(m8[2 * i], m8[2 * i + 1]); l4[i] = 0u; }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) split_pack2(m8[2 * i], m8[2 * i + 1], h4[i], l4[i]);
          }
          store_pixel8(p.pool, b, y >> 1, x >> 1, n0 + 16 * int(ox1) + 8 * int(oy1), h4, l4, pool_fmt);
        }
      }
    }
  } else {
    // 1x1 outconv over the 64 ReLU'd channels of d42, sigmoid, WS residual terms
    float z = p.bout;
#pragma unroll 1
    for (int cc = 0; cc < N_TILE / 32; ++cc) {
      float f[32];
      load_acc32<N_TILE, EPI, STACKED>(tbase + cc * 32, f, p.corr_scale);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 bq = *reinterpret_cast<const float4*>(sBias + cc * 32 + 4 * i);
        z = fmaf(fmaxf(f[4 * i] + bq.x, 0.f), p.wout[cc * 32 + 4 * i], z);
        z = fmaf(fmaxf(f[4 * i + 1] + bq.y, 0.f), p.wout[cc * 32 + 4 * i + 1], z);
        z = fmaf(fmaxf(f[4 * i + 2] + bq.z, 0.f), p.wout[cc * 32 + 4 * i + 2], z);
        z = fmaf(fmaxf(f[4 * i + 3] + bq.w, 0.f), p.wout[cc * 32 + 4 * i + 3], z);
      }
    }
    if (valid) {
      const float s = 1.f / (1.f + expf(-z));
      const size_t pix = (size_t(b) * p.H + y) * p.W + x;
      if (p.yhat) p.yhat[pix] = s;
      const bool inside = p.crop ? (y >= 1 && y < p.H - 1 && x >= 1 && x < p.W - 1) : true;
      if (p.img && inside) {
        const float xhat = s * 255.f;
        float xv, xbar, s1 = 0.f, s2 = 0.f;
        if (p.img_is_float) {
          const float* im = static_cast<const float*>(p.img);
          ws_load_f32(im[pix], xv, xbar);
          if (p.weighted != WS_UNWEIGHTED) {
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx) {
                if (dy == 0 && dx == 0) continue;
                const float q = im[pix + dy * p.W + dx] * 255.f;
                s1 += q;
                s2 = fmaf(q, q, s2);
              }
          }
        } else {
          const uint8_t* im = static_cast<const uint8_t*>(p.img);
          ws_load_u8(im[pix], xv, xbar);
          if (p.weighted != WS_UNWEIGHTED) {
            int i1 = 0, i2 = 0;
            for (int dy = -1; dy <= 1; ++dy)
              for (int dx = -1; dx <= 1; ++dx) {
                if (dy == 0 && dx == 0) continue;
                const int q = im[pix + dy * p.W + dx];
                i1 += q;
                i2 += q * q;
              }
            s1 = float(i1);
            s2 = float(i2);
          }
        }
        const float wgt = ws_weight(p.weighted, s1, s2);
        if (p.bias_pass) acc.wb = fmaf(wgt * (xv - xbar), xhat, acc.wb);   // xhat = pixel_estimator(x_bar - x) here
        else ws_accumulate(acc, xv, xbar, xhat, wgt);
      }
    }
  }
}

template <int N_TILE, int EPI>
__global__ void __launch_bounds__(kThreads, 1) conv_mma_kernel(const __grid_constant__ ConvParams p) {
  using C = Cfg<N_TILE>;
  constexpr int M_SUB = C::M_SUB;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // SA x 32 KB
  uint8_t* sW = smem + C::SA * kABytes;                 // SW x W_BYTES
  uint8_t* sScratch = sW + C::SW * C::W_BYTES;          // epilogue store staging
  float* sBias = reinterpret_cast<float*>(sScratch + kScratchBytes);  // up to 1024 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + 4096);
  uint64_t* a_full = bars;                 // [SA]
  uint64_t* a_empty = a_full + C::SA;      // [SA]
  uint64_t* w_full = a_empty + C::SA;      // [SW]
  uint64_t* w_empty = w_full + C::SW;      // [SW]
  uint64_t* acc_full = w_empty + C::SW;    // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int KB = p.cblocks * p.ntaps;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmapA0);
    prefetch_tmap(&p.tmapA1);
    for (int i = 0; i < C::SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::SW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.cout && i < 1024; i += kThreads) sBias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (elect_one()) {
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(p, tile, M_SUB);
        const uint8_t* wsrc = p.wpack + size_t(t.pos * p.n_tiles + t.nt) * KB * C::W_BYTES;
        for (int kb = 0; kb < KB; ++kb) {
          const int c = kb / p.ntaps;
          const int tap = kb - c * p.ntaps;
          mbar_wait(&w_empty[ws], wph ^ 1);
          mbar_arrive_expect_tx(&w_full[ws], C::W_BYTES);
          bulk_load(sW + ws * C::W_BYTES, wsrc + size_t(kb) * C::W_BYTES, C::W_BYTES, &w_full[ws]);
          if (++ws == C::SW) { ws = 0; wph ^= 1; }
          const bool src0 = c < p.cblocks0;
          const CUtensorMap* tm = src0 ? &p.tmapA0 : &p.tmapA1;
          const int ch = (src0 ? c : c - p.cblocks0) * 64;
#pragma unroll 1
          for (int j = 0; j < M_SUB; ++j) {
            mbar_wait(&a_empty[as], aph ^ 1);
            mbar_arrive_expect_tx(&a_full[as], kABytes);
            tma_load_5d(sA + as * kABytes, tm, &a_full[as], ch, t.x0 + j * p.TW + p.tap_dx[tap],
                        t.y0 + p.tap_dy[tap], t.b, 0);
            if (++as == C::SA) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (single thread)
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(N_TILE);
      int as = 0, ws = 0, acs = 0;
      uint32_t aph = 0, wph = 0, acph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acs], acph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&w_full[ws], wph);
          const uint32_t w_hi = smem_u32(sW + ws * C::W_BYTES);
          const uint32_t w_lo = w_hi + N_TILE * 128;
#pragma unroll 1
          for (int j = 0; j < M_SUB; ++j) {
            mbar_wait(&a_full[as], aph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(sA + as * kABytes);
            const uint32_t a_lo = a_hi + 128 * 128;
            const uint32_t d = tmem_base + uint32_t(acs * kAccCols + j * N_TILE);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da_hi = make_sw128_desc(a_hi + k * 32);
              const uint64_t da_lo = make_sw128_desc(a_lo + k * 32);
              const uint64_t dw_hi = make_sw128_desc(w_hi + k * 32);
              const uint64_t dw_lo = make_sw128_desc(w_lo + k * 32);
              umma_bf16(d, da_hi, dw_hi, idesc, (kb | k) != 0);
              umma_bf16(d, da_lo, dw_hi, idesc, 1);
              umma_bf16(d, da_hi, dw_lo, idesc, 1);
            }
            umma_commit(&a_empty[as]);  // frees the A slot when these MMAs retire
            if (++as == C::SA) { as = 0; aph ^= 1; }
          }
          umma_commit(&w_empty[ws]);
          if (++ws == C::SW) { ws = 0; wph ^= 1; }
        }
        umma_commit(&acc_full[acs]);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue warps (TMEM -> registers -> HBM)
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read
    const int grp = (warp - 2) >> 2;            // epilogue group: boxes j = grp, grp + 2, ...
    const int row = quad * 32 + lane;           // pixel row of the 128-pixel box
    const int ty = row / p.TW, tx = row - ty * p.TW;
    int acs = 0;
    uint32_t acph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(p, tile, M_SUB);
      mbar_wait(&acc_full[acs], acph);
      tc_fence_after();
      const int y = t.y0 + ty;
      WsAcc acc;
#pragma unroll 1
      for (int j = grp; j < M_SUB; j += 2) {
        const int x = t.x0 + j * p.TW + tx;
        const bool valid = (y < p.H) && (x < p.W);
        const uint32_t tbase = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acs * kAccCols + j * N_TILE);
        const BoxGeo geo{t.b, t.y0, t.x0 + j * p.TW, 4, p.H, p.W, p.upsample, t.pos};
        epilogue_box<N_TILE, EPI>(p, sBias, tbase, t.b, y, x, valid, t.nt, t.pos, tx, ty, p.TW, acc, geo,
                                  sScratch + (warp - 2) * kScratchPerWarp, lane, quad * 32);
      }
      // all TMEM reads of this stage are complete -> hand the accumulators back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acs]);
      if (++acs == 2) { acs = 0; acph ^= 1; }
      if constexpr (EPI == EPI_HEAD) {
        if (p.partials) {
          const float wr = warp_sum(acc.wr), w = warp_sum(acc.w), l1 = warp_sum(acc.l1), wb = warp_sum(acc.wb);
          if (elect_one()) {
            const size_t tiles_per_img = size_t(p.tiles_x) * p.tiles_y;
            float* dst = p.partials + ((size_t(t.b) * tiles_per_img + t.tile_in_img) * 8 + (warp - 2)) * kPartialSlots;
            if (p.bias_pass) dst[3] = wb;
            else { dst[0] = wr; dst[1] = w; dst[2] = l1; dst[3] = 0.f; }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int N_TILE, int EPI>
cudaError_t launch_t(const ConvParams& p, int num_sms, cudaStream_t stream) {
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  conv_mma_kernel<N_TILE, EPI><<<grid, kThreads, Cfg<N_TILE>::SMEM, stream>>>(p);
  return cudaGetLastError();
}


// ================================================================================================ halo kernel
// 3x3 convolution with ONE activation load per (box, 64-channel block): the TMA box is the 8x16-pixel output box plus
// its 1-pixel halo (10 x 18 pixels x 64 channels x {hi, lo}); tap (dy, dx) is the same shared-memory tile read through
// a descriptor whose start is shifted by (dy*10 + dx) rows and whose 8-row atoms are 10 rows (1280 B) apart.
// L2 -> SM activation traffic drops from 9 x 32 KB to 45 KB per box and channel block.
// Warp roles: 0 = activation (TMA box) producer, 1 = MMA issuer, 2..5 = epilogue, 6 = weight producer.
constexpr int kHaloThreads = 352;  // warps: 0 box producer, 1 MMA, 2..9 epilogue (group g owns box g), 10 weight producer
constexpr int kHaloRows = (kHaloTW + 2) * (kHaloTH + 2);  // 180 rows of 128 B per plane
constexpr int kHaloABytes = 2 * kHaloRows * 128;           // 46080
constexpr uint32_t kHaloSBO = (kHaloTW + 2) * 128;         // 1280

// Operand-bandwidth note: an M=128 x N x K=16 MMA takes N/2 tensor cycles but reads 4 KB (A) + N*32 B (B) from shared
// memory at 128 B/cycle. For Cout = 64 three N=64 MMAs (hi*hi, lo*hi, hi*lo) need 18 KB = 144 smem cycles for 96 tensor
// cycles. Stacking [Whi; Wlo] as ONE 128-row B tile turns hi*hi and hi*lo into a single N=128 MMA (A read once):
// 14 KB = 112 smem cycles per 96 tensor cycles. The accumulator is then 128 columns wide and the epilogue adds halves.
template <int N_TILE>
struct HCfg {
  static constexpr bool STACKED = (N_TILE == 64);
  static constexpr int M_SUB = halo_msub(N_TILE);
  static constexpr int ACC_W = 128;                          // TMEM columns per box (stacked 2 x 64, or 128)
  static constexpr int SA = 3;                               // M_SUB resident boxes + 1 prefetch slot
  static constexpr int W_BYTES = STACKED ? 2 * N_TILE * 128  // ring slot: whole (hi, lo) chunk, contiguous
                                         : N_TILE * 128;     //            or one plane of one tap
  static constexpr int W_PER_TAP = STACKED ? 1 : 2;          // ring slots consumed per tap
  static constexpr int SW = 4;
  static constexpr int BIAS_BYTES = (N_TILE == 64) ? 256 : 4096;
  static constexpr int SMEM = SA * kHaloABytes + SW * W_BYTES + kScratchBytes + 1024 + BIAS_BYTES + 256 + 1024 /*patch*/;
};

struct BoxCoord {
  int b, y0, x0, sub_in_img;
};
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
  if (f.mode == 0) return int(__umulhi(uint32_t(n), f.mul));
  if (f.mode == 1) return n;
  return n / int(f.d);
}
__device__ __forceinline__ BoxCoord decode_box(const ConvParams& p, int s) {
  BoxCoord c;
  const int r = fast_div(s, p.fd_sub_x);
  const int tx = s - r * p.sub_x;
  c.b = fast_div(r, p.fd_sub_y);
  const int ty = r - c.b * p.sub_y;
  c.x0 = tx * kHaloTW;
  c.y0 = ty * kHaloTH;
  c.sub_in_img = ty * p.sub_x + tx;
  return c;
}

// FUSE: the layer's input is the first convolution e11 (Cin = 1, K = 9) of the image itself. Three extra warps compute
// each haloed box of relu(e11(image)) straight into the activation ring (split-bf16, swizzled exactly as TMA would have
// written it), so e11's 67 MB/image feature map is never written to or read from HBM.
constexpr int kFuseWarps = 3, kFuseThreads = kFuseWarps * 32;

__device__ __forceinline__ int reflect_clamp(int v, int n) {
  v = v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v);
  return min(max(v, 0), n - 1);
}

template <int N_TILE, int EPI, bool FUSE = false>
__global__ void __launch_bounds__(kHaloThreads + (FUSE ? kFuseThreads : 0), 1)
conv_halo_kernel(const __grid_constant__ ConvParams p) {
  using C = HCfg<N_TILE>;
  constexpr int M_SUB = C::M_SUB;
  constexpr int kThreadsAll = kHaloThreads + (FUSE ? kFuseThreads : 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = smem + C::SA * kHaloABytes;
  uint8_t* sScratch = sW + C::SW * C::W_BYTES;
  float* sBias = reinterpret_cast<float*>(sScratch + kScratchBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + C::BIAS_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::SA;
  uint64_t* w_full = a_empty + C::SA;
  uint64_t* w_empty = w_full + C::SW;
  uint64_t* acc_full = w_empty + C::SW;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* sPatch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // FUSE: 20 x 12 image patch

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmapH0);
    prefetch_tmap(&p.tmapH1);
    for (int i = 0; i < C::SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::SW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.cout && i < C::BIAS_BYTES / 4; i += kThreadsAll) sBias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (FUSE && warp >= 11) {
    // ===================================================== fused e11 producer (3 warps): compute the haloed box in place
    const int t = threadIdx.x - 11 * 32;
    const int q = t & 7;                       // this thread's 8 output channels = one 16-byte piece of a pixel row
    uint64_t wr[4][9], bs[4];   // fp32 pairs: one FFMA2 advances two output channels
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bs[i] = pack_f32x2(p.fuse_b[q * 8 + 2 * i], p.fuse_b[q * 8 + 2 * i + 1]);
#pragma unroll
      for (int k = 0; k < 9; ++k) wr[i][k] = pack_f32x2(p.fuse_w[(q * 8 + 2 * i) * 9 + k], p.fuse_w[(q * 8 + 2 * i + 1) * 9 + k]);
    }
    const int H = p.H, W = p.W;
    int as = 0;
    uint32_t aph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int s0 = (item / p.n_tiles) * M_SUB;
      const int nsub = min(M_SUB, p.total_sub - s0);
      for (int j = 0; j < nsub; ++j) {
        const BoxCoord bc = decode_box(p, s0 + j);
        mbar_wait(&a_empty[as], aph ^ 1);
        // image patch rows y0-2 .. y0+17, cols x0-2 .. x0+9 (reflect-resolved, scaled to [0,1] like e11 does)
        for (int i = t; i < 240; i += kFuseThreads) {
          const int yy = reflect_clamp(bc.y0 - 2 + i / 12, H), xx = reflect_clamp(bc.x0 - 2 + i % 12, W);
          const size_t o = (size_t(bc.b) * H + yy) * W + xx;
          sPatch[i] = p.fuse_img_is_float ? static_cast<const float*>(p.fuse_img)[o]
                                          : __fdiv_rn(float(static_cast<const uint8_t*>(p.fuse_img)[o]), 255.f);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kFuseThreads) : "memory");
        uint8_t* slot = sA + as * kHaloABytes;
#pragma unroll 1
        for (int task = t; task < kHaloRows * 8; task += kFuseThreads) {
          const int pix = task >> 3;
          const int py = pix / (kHaloTW + 2), px = pix - py * (kHaloTW + 2);
          // halo pixels take the value of their mirror image (reflect padding of e11's output, unet.py:73)
          const int yy = reflect_clamp(bc.y0 - 1 + py, H), xx = reflect_clamp(bc.x0 - 1 + px, W);
          const int ry = min(max(yy - (bc.y0 - 2), 1), 18), rx = min(max(xx - (bc.x0 - 2), 1), 10);
          const float* win = sPatch + (ry - 1) * 12 + (rx - 1);
          uint64_t acc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = bs[i];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              const float v = win[dy * 12 + dx];
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fma2_bcast(v, wr[i][dy * 3 + dx], acc[i]);
            }
          uint32_t h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a0, a1;
            unpack_f32x2(acc[i], a0, a1);
            split_pack2(fmaxf(a0, 0.f), fmaxf(a1, 0.f), h[i], l[i]);
          }
          // SWIZZLE_128B on absolute address bits: the lo plane starts 180 rows in, i.e. at row phase (pix + 4) & 7
          *reinterpret_cast<uint4*>(slot + pix * 128 + ((q ^ (pix & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(slot + kHaloRows * 128 + pix * 128 + ((q ^ ((pix + 4) & 7)) << 4)) =
              make_uint4(l[0], l[1], l[2], l[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        asm volatile("bar.sync 1, %0;" ::"n"(kFuseThreads) : "memory");
        if (t == 0) mbar_arrive(&a_full[as]);
        if (++as == C::SA) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 0) {
    // ===================================================== activation producer: one haloed box per (box, channel block)
    if (!FUSE && elect_one()) {
      int as = 0;
      uint32_t aph = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int s0 = (item / p.n_tiles) * M_SUB;
        const int nsub = min(M_SUB, p.total_sub - s0);
        // the boxes of this CTA's NEXT item are pulled into L2 now: under heavy HBM write traffic a cold 45 KB box load
        // takes longer than the ring can cover
        const int item_n = item + gridDim.x;
        const int s0_n = (item_n / p.n_tiles) * M_SUB;
        const int nsub_n = (p.l2_prefetch && item_n < p.total_items) ? min(M_SUB, p.total_sub - s0_n) : 0;
        for (int c = 0; c < p.cblocks; ++c) {
          const bool src0 = c < p.cblocks0;
          const CUtensorMap* tm = src0 ? &p.tmapH0 : &p.tmapH1;
          const int ch = (src0 ? c : c - p.cblocks0) * 64;
          for (int j = 0; j < nsub_n; ++j) {
            const BoxCoord bn = decode_box(p, s0_n + j);
            tma_prefetch_5d(tm, ch, bn.x0, bn.y0, bn.b, 0);
          }
          const uint32_t box_bytes = (src0 && p.src0_f16) ? kHaloABytes / 2 : kHaloABytes;   // one fp16 plane or (hi, lo)
          for (int j = 0; j < nsub; ++j) {
            const BoxCoord bc = decode_box(p, s0 + j);
            mbar_wait(&a_empty[as], aph ^ 1);
            mbar_arrive_expect_tx(&a_full[as], box_bytes);
            // padded coords of pixel (y, x) are (y+1, x+1): the box with its halo starts at (y0, x0)
            tma_load_5d(sA + as * kHaloABytes, tm, &a_full[as], ch, bc.x0, bc.y0, bc.b, 0);
            if (++as == C::SA) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 10) {
    // ===================================================== weight producer: ring slots of every (block, tap)
    if (elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      const int halves = p.cblocks * 9 * C::W_PER_TAP;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int nt = item % p.n_tiles;
        const uint8_t* wsrc = p.wpack + size_t(nt) * halves * C::W_BYTES;
        for (int hh = 0; hh < halves; ++hh) {
          mbar_wait(&w_empty[ws], wph ^ 1);
          mbar_arrive_expect_tx(&w_full[ws], C::W_BYTES);
          bulk_load(sW + ws * C::W_BYTES, wsrc + size_t(hh) * C::W_BYTES, C::W_BYTES, &w_full[ws]);
          if (++ws == C::SW) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(N_TILE);
      int as = 0, ws = 0, acs = 0;
      uint32_t aph = 0, wph = 0, acph = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
        const int s0 = (item / p.n_tiles) * M_SUB;
        const int nsub = min(M_SUB, p.total_sub - s0);
        mbar_wait(&acc_empty[acs], acph ^ 1);
        tc_fence_after();
        for (int c = 0; c < p.cblocks; ++c) {
          uint32_t a_slot[M_SUB];
          if constexpr (C::STACKED) {
            // Taps are issued in groups of three, box-major inside a group: box 0 finishes its last tap three tap-steps
            // (not one) before the item ends, so its ring slot is refilled in time for the next item's second box.
            // (Tap-major order left only ~770 cycles for a 45 KB box load with these short N=64 MMAs.)
            // F16 blocks (source 0 of a decoder layer under a reduced precision plan: the up-convolution's output stored as
            // ONE fp16 plane, weights as an fp16 (hi, lo) pair): one N=128 MMA per K=16 step instead of two.
            auto issue_block = [&](auto f16_c) {
              constexpr bool F16 = decltype(f16_c)::value;
              for (int g = 0; g < 3; ++g) {
                int wslot[3];
#pragma unroll
                for (int j = 0; j < M_SUB; ++j) {
                  if (j < nsub) {
#pragma unroll
                    for (int tt = 0; tt < 3; ++tt) {
                      const int tap = 3 * g + tt;
                      if (j == 0) {
                        wslot[tt] = ws;
                        mbar_wait(&w_full[ws], wph);
                        if (++ws == C::SW) { ws = 0; wph ^= 1; }
                      }
                      if (tap == 0) {
                        mbar_wait(&a_full[as], aph);
                        a_slot[j] = uint32_t(as);
                        if (++as == C::SA) { as = 0; aph ^= 1; }
                      }
                      tc_fence_after();
                      const uint64_t wd = make_sw128_desc(smem_u32(sW + wslot[tt] * C::W_BYTES));
                      const uint64_t ad = make_sw128_desc(smem_u32(sA + a_slot[j] * kHaloABytes), kHaloSBO);
                      const uint32_t tap_off = uint32_t(g * (kHaloTW + 2) + tt) * 128;
                      const uint32_t d = tmem_base + uint32_t(acs * kAccCols + j * C::ACC_W);
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        const uint64_t da_hi = desc_at(desc_lo(ad), desc_hi(ad), tap_off + k * 32);
                        const uint64_t dw_hi = desc_at(desc_lo(wd), desc_hi(wd), k * 32);
                        if constexpr (F16) {
                          // columns [0,64) += A*Whi, [64,128) += A*Wlo (fp16 operands)
                          umma_bf16(d, da_hi, dw_hi, make_idesc_f16_m(128, 128), (c | tap | k) != 0);
                        } else {
                          const uint64_t da_lo = desc_at(desc_lo(ad), desc_hi(ad), tap_off + kHaloRows * 128 + k * 32);
                          // B tile of 128 rows = [Whi; Wlo]: columns [0,64) += Ahi*Whi, [64,128) += Ahi*Wlo
                          umma_bf16(d, da_hi, dw_hi, make_idesc_bf16(128), (c | tap | k) != 0);
                          umma_bf16(d, da_lo, dw_hi, make_idesc_bf16(64), 1);   // columns [0,64) += Alo*Whi
                        }
                      }
                      if (j == nsub - 1) umma_commit(&w_empty[wslot[tt]]);
                      if (tap == 8) umma_commit(&a_empty[a_slot[j]]);
                    }
                  }
                }
              }
            };
            if (p.src0_f16 && c < p.cblocks0) issue_block(std::true_type{});
            else issue_block(std::false_type{});
          } else {
            for (int tap = 0; tap < 9; ++tap) {
              const int ws_hi = ws;
              mbar_wait(&w_full[ws], wph);
              if (++ws == C::SW) { ws = 0; wph ^= 1; }
              const int ws_lo = ws;
              mbar_wait(&w_full[ws], wph);
              if (++ws == C::SW) { ws = 0; wph ^= 1; }
              const uint32_t w_hi = smem_u32(sW + ws_hi * C::W_BYTES);
              const uint32_t w_lo = smem_u32(sW + ws_lo * C::W_BYTES);
              const uint32_t tap_off = uint32_t((tap / 3) * (kHaloTW + 2) + (tap % 3)) * 128;
#pragma unroll
              for (int j = 0; j < M_SUB; ++j) {
                if (j < nsub) {
                  if (tap == 0) {
                    mbar_wait(&a_full[as], aph);
                    a_slot[j] = uint32_t(as);
                    if (++as == C::SA) { as = 0; aph ^= 1; }
                  }
                  tc_fence_after();
                  const uint64_t ad = make_sw128_desc(smem_u32(sA + a_slot[j] * kHaloABytes), kHaloSBO);
                  const uint64_t whd = make_sw128_desc(w_hi), wld = make_sw128_desc(w_lo);
                  const uint32_t d = tmem_base + uint32_t(acs * kAccCols + j * C::ACC_W);
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint64_t da_hi = desc_at(desc_lo(ad), desc_hi(ad), tap_off + k * 32);
                    const uint64_t da_lo = desc_at(desc_lo(ad), desc_hi(ad), tap_off + kHaloRows * 128 + k * 32);
                    const uint64_t dw_hi = desc_at(desc_lo(whd), desc_hi(whd), k * 32);
                    const uint64_t dw_lo = desc_at(desc_lo(wld), desc_hi(wld), k * 32);
                    // (keeping A_hi in the A collector for hi*hi, hi*lo measured 4-8 % SLOWER in this single-CTA kernel,
                    //  profiles/r01_layer_profile_a_collector.log; the CTA-pair kernel does use it)
                    umma_bf16(d, da_hi, dw_hi, idesc, (c | tap | k) != 0);
                    umma_bf16(d, da_lo, dw_hi, idesc, 1);
                    umma_bf16(d, da_hi, dw_lo, idesc, 1);
                  }
                  if (tap == 8) umma_commit(&a_empty[a_slot[j]]);
                }
              }
              umma_commit(&w_empty[ws_hi]);
              umma_commit(&w_empty[ws_lo]);
            }
          }
        }
        umma_commit(&acc_full[acs]);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    // ===================================================== epilogue warps
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;  // group g finishes box g of every work item
    const int row = quad * 32 + lane;
    const int ty = row / kHaloTW, tx = row % kHaloTW;
    int acs = 0;
    uint32_t acph = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int nt = item % p.n_tiles;
      const int s0 = (item / p.n_tiles) * M_SUB;
      const int nsub = min(M_SUB, p.total_sub - s0);
      mbar_wait(&acc_full[acs], acph);
      tc_fence_after();
#pragma unroll 1
      for (int j = grp; j < nsub; j += 2) {
        const BoxCoord bc = decode_box(p, s0 + j);
        const int y = bc.y0 + ty, x = bc.x0 + tx;
        const bool valid = (y < p.H) && (x < p.W);
        const uint32_t tbase = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acs * kAccCols + j * C::ACC_W);
        WsAcc acc;
        const BoxGeo geo{bc.b, bc.y0, bc.x0, 3, p.H, p.W, 0, 0};
        epilogue_box<N_TILE, EPI, C::STACKED>(p, sBias, tbase, bc.b, y, x, valid, nt, 0, tx, ty, kHaloTW, acc, geo,
                                              sScratch + (warp - 2) * kScratchPerWarp, lane, quad * 32);
        if constexpr (EPI == EPI_HEAD) {
          if (p.partials) {
            const float wr = warp_sum(acc.wr), w = warp_sum(acc.w), l1 = warp_sum(acc.l1), wb = warp_sum(acc.wb);
            if (elect_one()) {
              const size_t subs_per_img = size_t(p.sub_x) * p.sub_y;
              float* dst = p.partials + ((size_t(bc.b) * subs_per_img + bc.sub_in_img) * 4 + quad) * kPartialSlots;
              if (p.bias_pass) dst[3] = wb;
              else { dst[0] = wr; dst[1] = w; dst[2] = l1; dst[3] = 0.f; }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acs]);
      if (++acs == 2) { acs = 0; acph ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();   // TMA stores out of this warp's staging buffer are complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int N_TILE, int EPI>
cudaError_t launch_halo_t(const ConvParams& p, int num_sms, cudaStream_t stream) {
  const int grid = p.total_items < num_sms ? p.total_items : num_sms;
  if constexpr (N_TILE == 64 && EPI == EPI_ACT) {
    if (p.fuse_img) {
      if (p.cblocks != 1) return cudaErrorInvalidValue;
      conv_halo_kernel<64, EPI_ACT, true><<<grid, kHaloThreads + kFuseThreads, HCfg<64>::SMEM, stream>>>(p);
      return cudaGetLastError();
    }
  }
  conv_halo_kernel<N_TILE, EPI><<<grid, kHaloThreads, HCfg<N_TILE>::SMEM, stream>>>(p);
  return cudaGetLastError();
}


// ================================================================================================ CTA-pair halo kernel
// Same algorithm as conv_halo_kernel, issued as tcgen05.mma.cta_group::2: a cluster of two CTAs computes M = 256
// (each CTA its own 8x16 pixel box, accumulators in its own TMEM) and each CTA stages only HALF of the B rows, so the
// shared-memory operand traffic per SM drops from 14/24 KB to 11/18 KB per K=16 step (Cout=64 stacked / Cout>=128) and
// the weight ring shrinks. Leader CTA (rank 0) issues all MMAs; activation boxes are loaded by each CTA with the
// cta_group::2 TMA form that credits the leader's mbarrier; weight half-tiles arrive on a local barrier and the idle
// MMA warp of the peer CTA relays their completion to the leader.
// TERMS = MMAs per algorithmic MAC. 3: split-bf16 activations (hi, lo planes) x split-bf16 weights, hi*hi + hi*lo + lo*hi.
// 2 / 1: the input map is ONE fp16 plane (ACT_F16) against fp16 (hi, lo) / fp16 hi-only weights - the layers the precision
// plan (api.cu, option "precision") runs below three terms. Half-size boxes: the ring holds six of them.
// RES: the layer's whole weight set stays in shared memory (fp16 + fp8 layers with ONE input channel block, e12 / d42:
// 9 taps x (4 KB fp16 + 4 KB e4m3) per CTA = 72 KB, about what the ring took). These layers stream 46 KB of boxes and 36 KB
// of weights per box from L2 (8.2 TB/s for the loads alone, WSU_DBG timings) and everything crosses the SM's shared-memory
// data path, which bounds them: keeping the weights removes 44 % of the L2 -> SM bytes and 7 % of the shared-memory traffic.
template <int N_TILE, int TERMS = 3, bool RES = false>
struct H2Cfg {
  static constexpr bool STACKED = (N_TILE == 64);
  static constexpr int M_SUB = 2;                 // box slots per item (x 2 CTAs = 4 boxes share one weight pass)
  static constexpr int ACC_W = 128;
  static constexpr int A_BYTES = TERMS == 3 ? kHaloABytes : kHaloABytes / 2;
  static constexpr int SA = TERMS == 3 ? 3 : ((TERMS == 2 || RES) ? 4 : 6);   // one-term + resident weights: four boxes leave room for the nine 8 KB taps   // fp16 boxes are 22.5 KB: an even count keeps the weight ring 1024-byte aligned
  static constexpr int W_SLOT = (TERMS == 1 || RES) ? 8192 : 16384;   // [X tile 8 KB][Y tile 4 KB (stacked) | Z tile 8 KB]; RES: [fp16 4 KB][e4m3 4 KB]
  static constexpr int W_BYTES = STACKED ? 12288 : (TERMS == 1 ? 8192 : 16384);
  // a tap's weights are consumed in 2 x 4 MMAs: ~1500 / 1000 / 540 cycles with 3 / 2 / 1 terms, against an L2 -> shared
  // latency of ~2500 cycles per bulk copy: the ring has to be deeper the fewer terms a tap takes
  static constexpr int SW = RES ? 9 : (TERMS == 3 ? 4 : (TERMS == 2 ? 6 : 8));
  static_assert((SA * A_BYTES) % 1024 == 0, "pre-swizzled weight tiles need a 1024-byte aligned ring");
  static constexpr int BIAS_BYTES = (N_TILE == 64) ? 256 : 4096;
  static constexpr int SMEM = SA * A_BYTES + SW * W_SLOT + kScratchBytes + 1024 + BIAS_BYTES + 512 /*barriers*/;
};

// How one 64-channel block of a stacked (Cout = 64) layer is multiplied. SPLIT3: split-bf16 planes, hi*[Whi;Wlo] (N=128) +
// lo*Whi (N=64). F16_2 / F16_1: ONE fp16 plane (source 0 of a decoder layer under a reduced plan) against fp16 [Whi;Wlo]
// (N=128) / Whi (N=64). F8: ACT_F16F8 planes - fp16 main product (N=64, columns [0,64)) + ONE e4m3 MMA over the two
// correction operands (N=64, columns [64,128), scaled; the epilogue multiplies by ConvParams::corr_scale).
enum : int { MODE_SPLIT3 = 0, MODE_F16_2 = 1, MODE_F16_1 = 2, MODE_F8 = 3 };
__device__ __forceinline__ int block_mode(const ConvParams& p, int c) {
  if (p.src0_f16 && c < p.cblocks0) return p.src0_f16 == 2 ? MODE_F16_1 : MODE_F16_2;
  return p.f8_blocks ? MODE_F8 : MODE_SPLIT3;
}

// Tried and dropped (profiles/r02_ab_epilogue_warps_tma_store.log): sixteen epilogue warps per CTA (two per box quarter, half
// the chunks each, 1 KB of staging per warp) - bit-identical and 13 % slower over the chain: these kernels are bound by the
// shared-memory data path, which more warps do not widen, and 608 threads leave 96 registers per thread (spills).
// W8 (Cout = 64 kernel): 0 = every block mode behind runtime switches (plans 0-2); 1 = the fp16 + fp8 plan with streamed
// weights (d41: only MODE_F16_1 / MODE_F8 and the fp16 + e4m3 output format are compiled in - the generic kernel is 91 KB of
// code, which the unrolled single-thread issue loop and the epilogue warps fight over in the instruction cache); 2 = the same
// with resident weights (e12, d42).
template <int N_TILE, int EPI, bool COLL = true, int TERMS = 3, int W8 = 0>
__global__ void __launch_bounds__(kHaloThreads, 1) conv_halo2_kernel(const __grid_constant__ ConvParams p) {
  constexpr bool RES = W8 == 2;                       // all nine taps of the (single) channel block stay in shared memory
  constexpr bool F8ONLY = W8 != 0 && N_TILE == 64;
  static_assert(TERMS == 3 || N_TILE == 128, "the one- and two-term variants exist for the Cout >= 128 layers only");
  static_assert(W8 == 0 || (N_TILE == 64 && TERMS == 3) || (N_TILE == 128 && TERMS == 1 && W8 == 2),
                "W8: fp16 + fp8 variants of the Cout = 64 kernel, or the one-term Cout = 128 kernel with resident weights (e21)");
  using C = H2Cfg<N_TILE, TERMS, RES>;
  constexpr int M_SUB = C::M_SUB;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = smem + C::SA * C::A_BYTES;
  uint8_t* sScratch = sW + C::SW * C::W_SLOT;
  float* sBias = reinterpret_cast<float*>(sScratch + kScratchBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + C::BIAS_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + C::SA;
  uint64_t* w_full = a_empty + C::SA;
  uint64_t* w_empty = w_full + C::SW;
  uint64_t* acc_full = w_empty + C::SW;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int items = p.total_items;   // pair items: ceil(total_sub / 4) * n_tiles (host fills this for the pair kernel)

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmapH0);
    prefetch_tmap(&p.tmapH1);
    prefetch_tmap(&p.tmapW);
    prefetch_tmap(&p.tmapW32);
    // a_full / w_full of the LEADER collect the bytes of both CTAs' TMA loads (cta_group::2 loads credit the leader's barrier)
    // after ONE arrival, the leader's own arrive.expect_tx for twice the per-CTA bytes: the peer issues its loads without
    // signalling first. (A remote mbarrier.arrive.release.cluster costs a GPU-scope fence, ~900 cycles: paid once per tap by
    // the weight relay it had made every layer with fewer than three MMAs per MAC wait on the relay, not on the tensor pipe.)
    for (int i = 0; i < C::SA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < C::SW; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.cout && i < C::BIAS_BYTES / 4; i += kHaloThreads) sBias[i] = p.bias[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / peer-credited TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================== activation boxes: this CTA's box of every slot
    if (elect_one()) {
      int as = 0;
      uint32_t aph = 0;
      for (int item = pair; item < items; item += npairs) {
        const int s0 = fast_div(item, p.fd_n_tiles) * (2 * M_SUB);
        const int nslot = min(M_SUB, (p.total_sub - s0 + 1) / 2);
        const int item_n = item + npairs;
        const int s0_n = fast_div(item_n, p.fd_n_tiles) * (2 * M_SUB);
        const int nslot_n = (p.l2_prefetch && item_n < items) ? min(M_SUB, (p.total_sub - s0_n + 1) / 2) : 0;
        for (int c = 0; c < p.cblocks; ++c) {
          const bool src0 = c < p.cblocks0;
          const CUtensorMap* tm = src0 ? &p.tmapH0 : &p.tmapH1;
          const int ch = (src0 ? c : c - p.cblocks0) * 64;
          for (int j = 0; j < nslot_n; ++j) {   // L2 warm-up of this CTA's boxes of the pair's next item
            const BoxCoord bn = decode_box(p, min(s0_n + 2 * j + int(rank), p.total_sub - 1));
            tma_prefetch_5d(tm, ch, bn.x0, bn.y0, bn.b, 0);
          }
          for (int j = 0; j < nslot; ++j) {
            const BoxCoord bc = decode_box(p, min(s0 + 2 * j + int(rank), p.total_sub - 1));
            mbar_wait(&a_empty[as], aph ^ 1);
            const uint32_t full_leader = mapa_u32(smem_u32(&a_full[as]), 0);
            // source-0 blocks of a decoder layer under a reduced plan are ONE fp16 plane (half a box)
            if (leader) mbar_arrive_expect_tx(&a_full[as], (src0 && p.src0_f16) ? C::A_BYTES : 2 * C::A_BYTES);
            tma_load_5d_2sm(sA + as * C::A_BYTES, tm, full_leader, ch, bc.x0, bc.y0, bc.b, 0);
            if (++as == C::SA) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 10) {
    // ===================================================== weight half-tiles of this CTA
    if (elect_one()) {
      int ws = 0;
      uint32_t wph = 0;
      const int chunks = p.cblocks * 9;
      if constexpr (RES) {
        // all nine taps once: this CTA's 32 rows of the fp16 tile and of the e4m3 tile of every tap (n_tiles == 1, cblocks == 1)
        if (pair < items) {
          for (int q = 0; q < 9; ++q) {
            uint8_t* dst = sW + q * C::W_SLOT;
            const uint32_t full_leader = mapa_u32(smem_u32(&w_full[q]), 0);
            const int row0 = q * 2 * N_TILE;
            if (leader) mbar_arrive_expect_tx(&w_full[q], 2 * 8192);
            if constexpr (N_TILE == 64) {
              tma_load_2d_2sm(dst, &p.tmapW32, full_leader, 0, row0 + int(rank) * 32);
              tma_load_2d_2sm(dst + 4096, &p.tmapW32, full_leader, 0, row0 + 64 + int(rank) * 32);
            } else {
              tma_load_2d_2sm(dst, &p.tmapW, full_leader, 0, row0 + int(rank) * 64);   // this CTA's 64 rows of the fp16 tile
            }
          }
        }
      } else
      for (int item = pair; item < items; item += npairs) {
        const int nt = item - fast_div(item, p.fd_n_tiles) * p.n_tiles;
        for (int q = 0; q < chunks; ++q) {
          uint8_t* dst = sW + ws * C::W_SLOT;
          mbar_wait(&w_empty[ws], wph ^ 1);
          // the packed weights seen as rows of 128 B: 64- (or 32-) row boxes, raw copy (they are stored pre-swizzled)
          if constexpr (C::STACKED) {
            const uint32_t full_leader = mapa_u32(smem_u32(&w_full[ws]), 0);
            const int row0 = (nt * chunks + q) * 2 * N_TILE;                        // chunk = two 64-row tiles of 8 KB
            const int mode = block_mode(p, q / 9);
            if (mode == MODE_F8) {
              // [main fp16 tile][correction e4m3 tile]: this CTA's 32 rows of each (the pair's B operand has N = 64)
              if (leader) mbar_arrive_expect_tx(&w_full[ws], 2 * 8192);
              tma_load_2d_2sm(dst, &p.tmapW32, full_leader, 0, row0 + int(rank) * 32);
              tma_load_2d_2sm(dst + 8192, &p.tmapW32, full_leader, 0, row0 + 64 + int(rank) * 32);
            } else if (mode == MODE_F16_1) {
              if (leader) mbar_arrive_expect_tx(&w_full[ws], 2 * 4096);             // this CTA's 32 rows of the fp16 tile
              tma_load_2d_2sm(dst, &p.tmapW32, full_leader, 0, row0 + int(rank) * 32);
            } else {
              if (leader) mbar_arrive_expect_tx(&w_full[ws], 2 * C::W_BYTES);
              tma_load_2d_2sm(dst, &p.tmapW, full_leader, 0, row0 + int(rank) * 64);           // X: rank 0 Whi, rank 1 Wlo: [Whi; Wlo] over the pair
              tma_load_2d_2sm(dst + 8192, &p.tmapW32, full_leader, 0, row0 + int(rank) * 32);  // Y: this CTA's half of Whi
            }
          } else {
            const uint32_t full_leader = mapa_u32(smem_u32(&w_full[ws]), 0);
            const int row_hi = (nt * chunks + q) * 2 * N_TILE + int(rank) * 64;
            if (leader) mbar_arrive_expect_tx(&w_full[ws], 2 * C::W_BYTES);
            tma_load_2d_2sm(dst, &p.tmapW, full_leader, 0, row_hi);                                      // X: this CTA's 64 rows of Whi
            if constexpr (TERMS != 1) tma_load_2d_2sm(dst + 8192, &p.tmapW, full_leader, 0, row_hi + N_TILE);   // Z: ... of Wlo
          }
          if (++ws == C::SW) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      if (leader) {
        // ===================================================== leader: MMA issuer for the pair
        constexpr uint32_t idesc = TERMS == 3 ? make_idesc_bf16_m(256, N_TILE) : make_idesc_f16_m(256, N_TILE);
        const int f8_first = p.src0_f16 ? p.cblocks0 : 0;   // first channel block that accumulates into the correction columns
        int as = 0, ws = 0, acs = 0;
        uint32_t aph = 0, wph = 0, acph = 0;
        for (int item = pair; item < items; item += npairs) {
          const int s0 = fast_div(item, p.fd_n_tiles) * (2 * M_SUB);
          const int nslot = min(M_SUB, (p.total_sub - s0 + 1) / 2);
          mbar_wait(&acc_empty[acs], acph ^ 1);
          tc_fence_after();
          for (int c = 0; c < p.cblocks; ++c) {
            uint32_t a_slot[M_SUB];
            // F16 blocks (stacked Cout = 64 kernel only): source 0 of a decoder layer under a reduced plan, one fp16 plane
            // against the fp16 [Whi; Wlo] tile = a single N=128 MMA per K step
            auto issue_block = [&](auto mode_c) {
              constexpr int MODE = decltype(mode_c)::value;
              // Three-term layers hold 45 KB boxes in a ring of three (two in use, one prefetched): in tap-major order both
              // boxes of an item are released at its very end and the second box of the next item / channel block arrives
              // ~2 500 cycles late (e12: 10.4k cycles per item for 7.8k of MMAs). In groups of three taps, box-major inside
              // a group, box 0 is released 1/6 item early and box 1 of the successor is first needed one group into it, which
              // covers the load. Reduced-term layers (22.5 KB boxes, ring of 4 / 6) stay tap-major.
              constexpr int GT = (TERMS == 3) ? 3 : 1;     // taps per group
#pragma unroll 1
              for (int g = 0; g < 9 / GT; ++g) {
                int wslot[GT];
#pragma unroll
                for (int j = 0; j < M_SUB; ++j) {
                  if (j < nslot) {
#pragma unroll
                    for (int tt = 0; tt < GT; ++tt) {
                      const int tap = GT * g + tt;
                      if constexpr (RES) {
                        wslot[tt] = tap;
                        if (j == 0 && item == pair) mbar_wait(&w_full[tap], 0);   // the tap's weights have landed (first item only)
                      } else if (j == 0) {
                        wslot[tt] = ws;
                        mbar_wait(&w_full[ws], wph);
                        if (++ws == C::SW) { ws = 0; wph ^= 1; }
                      }
                      const uint32_t w_x = smem_u32(sW + wslot[tt] * C::W_SLOT), w_y = w_x + (RES ? 4096 : 8192);
                      const uint32_t tap_off = uint32_t((tap / 3) * (kHaloTW + 2) + (tap % 3)) * 128;
                      if (tap == 0) {
                        mbar_wait(&a_full[as], aph);
                        a_slot[j] = uint32_t(as);
                        if (++as == C::SA) { as = 0; aph ^= 1; }
                      }
                      tc_fence_after();
                      const uint64_t ad = make_sw128_desc(smem_u32(sA + a_slot[j] * C::A_BYTES), kHaloSBO);
                      const uint64_t wxd = make_sw128_desc(w_x), wyd = make_sw128_desc(w_y);
                      const uint32_t d = tmem_base + uint32_t(acs * kAccCols + j * C::ACC_W);
#pragma unroll
                      for (int k = 0; k < 4; ++k) {
                        const uint64_t da_hi = desc_at(desc_lo(ad), desc_hi(ad), tap_off + k * 32);
                        const uint64_t da_lo = desc_at(desc_lo(ad), desc_hi(ad), tap_off + kHaloRows * 128 + k * 32);
                        const uint64_t dw_x = desc_at(desc_lo(wxd), desc_hi(wxd), k * 32);
                        const uint64_t dw_y = desc_at(desc_lo(wyd), desc_hi(wyd), k * 32);
                        if constexpr (MODE == MODE_F16_2) {         // fp16 A x fp16 [Whi; Wlo]
                          umma_bf16_2sm(d, da_hi, dw_x, make_idesc_f16_m(256, 128), (c | tap | k) != 0);
                        } else if constexpr (MODE == MODE_F16_1) {  // fp16 A x fp16 Whi
                          umma_bf16_2sm(d, da_hi, dw_x, make_idesc_f16_m(256, 64), (c | tap | k) != 0);
                        } else if constexpr (MODE == MODE_F8) {     // fp16 main product, then both corrections in one e4m3 MMA
                          umma_bf16_2sm(d, da_hi, dw_x, make_idesc_f16_m(256, 64), (c | tap | k) != 0);
                          if (!(p.dbg & 8)) umma_f8_2sm(d + 64, da_lo, dw_y, make_idesc_f16_m(256, 64), ((c - f8_first) | tap | k) != 0);
                        } else if constexpr (TERMS == 2) {   // fp16 A x (Whi, Wlo): A read from shared memory once
                          umma_bf16_2sm_a_fill(d, da_hi, dw_x, idesc, (c | tap | k) != 0);
                          umma_bf16_2sm_a_lastuse(d, da_hi, dw_y, idesc, 1);
                        } else if constexpr (TERMS == 1) {   // fp16 A x fp16 W
                          umma_bf16_2sm(d, da_hi, dw_x, idesc, (c | tap | k) != 0);
                        } else if constexpr (C::STACKED) {
                          umma_bf16_2sm(d, da_hi, dw_x, make_idesc_bf16_m(256, 128), (c | tap | k) != 0);  // Ahi * [Whi; Wlo]
                          umma_bf16_2sm(d, da_lo, dw_y, make_idesc_bf16_m(256, 64), 1);                    // Alo * Whi
                        } else {
                          if constexpr (COLL) {   // A_hi read once for its two products (compile-time: a runtime branch in this
                                                  // single-thread issue loop costs ~10 % of the layer)
                            umma_bf16_2sm_a_fill(d, da_hi, dw_x, idesc, (c | tap | k) != 0);
                            umma_bf16_2sm_a_lastuse(d, da_hi, dw_y, idesc, 1);
                            umma_bf16_2sm(d, da_lo, dw_x, idesc, 1);
                          } else {
                            umma_bf16_2sm(d, da_hi, dw_x, idesc, (c | tap | k) != 0);
                            umma_bf16_2sm(d, da_lo, dw_x, idesc, 1);
                            umma_bf16_2sm(d, da_hi, dw_y, idesc, 1);
                          }
                        }
                      }
                      if constexpr (!RES) {
                        if (j == nslot - 1) umma_commit_2sm(&w_empty[wslot[tt]], 3);
                      }
                      if (tap == 8) umma_commit_2sm(&a_empty[a_slot[j]], 3);
                    }
                  }
                }
              }
            };
            if constexpr (RES && N_TILE == 64) {
              issue_block(std::integral_constant<int, MODE_F8>{});
            } else if constexpr (F8ONLY) {
              if (block_mode(p, c) == MODE_F8) issue_block(std::integral_constant<int, MODE_F8>{});
              else issue_block(std::integral_constant<int, MODE_F16_1>{});
            } else if constexpr (C::STACKED) {
              const int mode = block_mode(p, c);
              if (mode == MODE_F8) issue_block(std::integral_constant<int, MODE_F8>{});
              else if (mode == MODE_F16_1) issue_block(std::integral_constant<int, MODE_F16_1>{});
              else if (mode == MODE_F16_2) issue_block(std::integral_constant<int, MODE_F16_2>{});
              else issue_block(std::integral_constant<int, MODE_SPLIT3>{});
            } else {
              issue_block(std::integral_constant<int, MODE_SPLIT3>{});
            }
          }
          umma_commit_2sm(&acc_full[acs], 3);
          if (++acs == 2) { acs = 0; acph ^= 1; }
        }
      }
    }
  } else {
    // ===================================================== epilogue warps (each CTA finishes its own boxes)
    const int quad = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int ty = row / kHaloTW, tx = row % kHaloTW;
    int acs = 0;
    uint32_t acph = 0;
    for (int item = pair; item < items; item += npairs) {
      const int nt = item - fast_div(item, p.fd_n_tiles) * p.n_tiles;
      const int s0 = fast_div(item, p.fd_n_tiles) * (2 * M_SUB);
      const int nslot = min(M_SUB, (p.total_sub - s0 + 1) / 2);
      mbar_wait(&acc_full[acs], acph);
      tc_fence_after();
#pragma unroll 1
      for (int j = grp; j < nslot; j += 2) {
        const int s = s0 + 2 * j + int(rank);
        if (s < p.total_sub) {
          const BoxCoord bc = decode_box(p, s);
          const int y = bc.y0 + ty, x = bc.x0 + tx;
          const bool valid = (y < p.H) && (x < p.W);
          const uint32_t tbase = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acs * kAccCols + j * C::ACC_W);
          WsAcc acc;
          const BoxGeo geo{bc.b, bc.y0, bc.x0, 3, p.H, p.W, 0, 0};
          // the one- and two-term kernels write fp16 maps only, the resident-weight kernel fp16 + e4m3 maps (launch_halo2_t checks)
          constexpr int FMT = TERMS != 3 ? int(ACT_F16) : (F8ONLY ? int(ACT_F16F8) : -1);
          epilogue_box<N_TILE, EPI, C::STACKED, FMT>(p, sBias, tbase, bc.b, y, x, valid, nt, 0, tx, ty, kHaloTW, acc, geo,
                                                     sScratch + (warp - 2) * kScratchPerWarp, lane, quad * 32);
          if constexpr (EPI == EPI_HEAD) {
            if (p.partials) {
              const float wr = warp_sum(acc.wr), w = warp_sum(acc.w), l1 = warp_sum(acc.l1), wb = warp_sum(acc.wb);
              if (lane == 0) {
                const size_t subs_per_img = size_t(p.sub_x) * p.sub_y;
                float* dst = p.partials + ((size_t(bc.b) * subs_per_img + bc.sub_in_img) * 4 + quad) * kPartialSlots;
                if (p.bias_pass) dst[3] = wb;
                else { dst[0] = wr; dst[1] = w; dst[2] = l1; dst[3] = 0.f; }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&acc_empty[acs]), 0));
      if (++acs == 2) { acs = 0; acph ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();   // TMA stores out of this warp's staging buffer are complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int N_TILE, int EPI>
cudaError_t launch_halo2_t(const ConvParams& p, int num_sms, cudaStream_t stream) {
  int pairs = num_sms / 2;
  if (pairs > p.total_items) pairs = p.total_items;
  const bool fmt_f16 = p.out.fmt == ACT_F16 && (!p.do_pool || p.pool.fmt == ACT_F16);
  const bool fmt_f16f8 = EPI == EPI_HEAD || (p.out.fmt == ACT_F16F8 && (!p.do_pool || p.pool.fmt == ACT_F16));
  const bool res = N_TILE == 64 && p.w_resident && p.f8_blocks && p.cblocks == 1 && !p.src0_f16 && p.n_tiles == 1 && fmt_f16f8;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kHaloThreads);
  cfg.dynamicSmemBytes = (N_TILE == 128 && p.terms == 2) ? H2Cfg<128, 2>::SMEM : (N_TILE == 128 && p.terms == 1) ? H2Cfg<128, 1>::SMEM : H2Cfg<N_TILE>::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if constexpr (N_TILE == 64) {
    if (res) {
      if (p.terms != 3) return cudaErrorInvalidValue;
      cfg.dynamicSmemBytes = H2Cfg<64, 3, true>::SMEM;
      return cudaLaunchKernelEx(&cfg, conv_halo2_kernel<64, EPI, true, 3, 2>, p);
    }
    if constexpr (EPI == EPI_ACT) {   // the fp16 + fp8 plan with streamed weights (d41: fp16 source-0 blocks, one MMA per MAC)
      if (!(p.dbg & 64) && p.f8_blocks && (p.src0_f16 == 0 || p.src0_f16 == 2) && fmt_f16f8 && p.terms == 3)   // WSU_DBG=64: generic kernel (A/B)
        return cudaLaunchKernelEx(&cfg, conv_halo2_kernel<64, EPI_ACT, true, 3, 1>, p);
    }
  }
  if (p.terms != 3 && !fmt_f16) return cudaErrorInvalidValue;   // the reduced-term kernels are compiled for fp16 outputs
  if constexpr (N_TILE == 128 && EPI == EPI_ACT) {
    if (p.terms == 1 && p.w_resident && p.cblocks == 1 && p.n_tiles == 1) {   // e21: 72 KB of weights per CTA, loaded once
      cfg.dynamicSmemBytes = H2Cfg<128, 1, true>::SMEM;
      return cudaLaunchKernelEx(&cfg, conv_halo2_kernel<128, EPI_ACT, true, 1, 2>, p);
    }
  }
  if constexpr (N_TILE == 128 && EPI == EPI_ACT) {
    if (p.terms == 2) return cudaLaunchKernelEx(&cfg, conv_halo2_kernel<128, EPI_ACT, true, 2>, p);
    if (p.terms == 1) return cudaLaunchKernelEx(&cfg, conv_halo2_kernel<128, EPI_ACT, true, 1>, p);
  }
  if (p.terms != 3) return cudaErrorInvalidValue;
  if (N_TILE == 128 && !p.a_collector) return cudaLaunchKernelEx(&cfg, conv_halo2_kernel<N_TILE, EPI, false>, p);
  return cudaLaunchKernelEx(&cfg, conv_halo2_kernel<N_TILE, EPI, true>, p);
}

// ================================================================================================ resident-weight upconv
// ConvTranspose2d(k=2, s=2) (unet.py:177,183): out[2y+dy][2x+dx][co] = b[co] + sum_ci in[y][x][ci] * w[ci][co][dy][dx].
// The four phases are stacked along N and the whole (hi, lo) weight set of the CTA's N tile (128 KB) stays in shared
// memory, so each activation box is read exactly once per N tile and the kernel is bound by its output writes.
constexpr int kUpThreads = 320;          // three-term: warps 0 TMA, 1 MMA, 2..9 epilogue (two groups of four lane quadrants)
constexpr int kUpThreadsWide = 576;      // one / two terms: 16 epilogue warps (four groups). With a third of the MMAs per box the
                                         // epilogue (TMEM -> bias -> fp16/split pack -> staged 2x2 pixel-shuffle stores) is what a
                                         // box waits for: 7 000 cycles of epilogue against 1 000 of MMAs with 8 warps.

// TERMS as in the CTA-pair convolution: 3 = split-bf16 input (hi, lo planes), 2 / 1 = ONE fp16 input plane against fp16
// (hi, lo) / hi-only weights. The output format (p.out.fmt) is independent of it.
template <int N_TILE, int TERMS = 3>
__global__ void __launch_bounds__(TERMS == 3 ? kUpThreads : kUpThreadsWide, 1) upconv_res_kernel(const __grid_constant__ UpconvParams p) {
  constexpr int kThreadsUp = TERMS == 3 ? kUpThreads : kUpThreadsWide;
  constexpr int kEpiGroups = (kThreadsUp / 32 - 2) / 4;        // 2 or 4 groups of four epilogue warps
  constexpr int kBoxBytes = TERMS == 3 ? kABytes : kABytes / 2;  // hi + lo tile (32 KB) or one fp16 tile of 128 pixels x 64 channels
  constexpr int kUpSA = TERMS == 3 ? 2 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;                       // cblocks x (hi | lo) x N_TILE rows x 128 B  (<= 128 KB)
  uint8_t* sA = smem + kUpconvResBytes;     // kUpSA x 32 KB
  uint8_t* sScratch = sA + kUpSA * kBoxBytes;
  float* sBias = reinterpret_cast<float*>(sScratch + (kEpiGroups / 2) * kScratchBytes);   // co_t floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + 256);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + kUpSA;
  uint64_t* acc_full = a_empty + kUpSA;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt = blockIdx.x % p.n_tiles;
  const int first_box = blockIdx.x / p.n_tiles;
  const int box_step = gridDim.x / p.n_tiles;
  const uint32_t w_tile_bytes = N_TILE * 128;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmapA);
    for (int i = 0; i < kUpSA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4 * kEpiGroups); }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < p.co_t; i += kThreadsUp) sBias[i] = p.bias[nt * p.co_t + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // weights of this N tile: resident for the whole kernel
      const uint32_t wbytes = uint32_t(p.cblocks) * 2 * w_tile_bytes;
      const uint8_t* wsrc = p.wres + size_t(nt) * wbytes;
      mbar_arrive_expect_tx(w_bar, TERMS == 1 ? wbytes / 2 : wbytes);
      for (int i = 0; i < p.cblocks * 2; ++i) {
        if (TERMS == 1 && (i & 1)) continue;   // hi tiles only
        bulk_load(sW + i * w_tile_bytes, wsrc + size_t(i) * w_tile_bytes, w_tile_bytes, w_bar);
      }
      int as = 0;
      uint32_t aph = 0;
      for (int box = first_box; box < p.total_boxes; box += box_step) {
        const int r = fast_div(box, p.fd_tiles_x), tx = box - r * p.tiles_x;
        const int b = fast_div(r, p.fd_tiles_y), ty = r - b * p.tiles_y;
        for (int c = 0; c < p.cblocks; ++c) {
          mbar_wait(&a_empty[as], aph ^ 1);
          mbar_arrive_expect_tx(&a_full[as], kBoxBytes);
          tma_load_5d(sA + as * kBoxBytes, &p.tmapA, &a_full[as], c * 64, tx * 16 + 1, ty * 8 + 1, b, 0);
          if (++as == kUpSA) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = TERMS == 3 ? make_idesc_bf16(N_TILE) : make_idesc_f16_m(128, N_TILE);
      int as = 0, acs = 0;
      uint32_t aph = 0, acph = 0;
      mbar_wait(w_bar, 0);
      for (int box = first_box; box < p.total_boxes; box += box_step) {
        mbar_wait(&acc_empty[acs], acph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + uint32_t(acs * kAccCols);
        for (int c = 0; c < p.cblocks; ++c) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(sA + as * kBoxBytes), a_lo = a_hi + 128 * 128;
          const uint32_t w_hi = smem_u32(sW) + uint32_t(c) * 2 * w_tile_bytes, w_lo = w_hi + w_tile_bytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da_hi = make_sw128_desc(a_hi + k * 32), da_lo = make_sw128_desc(a_lo + k * 32);
            const uint64_t dw_hi = make_sw128_desc(w_hi + k * 32), dw_lo = make_sw128_desc(w_lo + k * 32);
            umma_bf16(d, da_hi, dw_hi, idesc, (c | k) != 0);
            if constexpr (TERMS == 3) umma_bf16(d, da_lo, dw_hi, idesc, 1);
            if constexpr (TERMS >= 2) umma_bf16(d, da_hi, dw_lo, idesc, 1);
          }
          umma_commit(&a_empty[as]);
          if (++as == kUpSA) { as = 0; aph ^= 1; }
        }
        umma_commit(&acc_full[acs]);
        if (++acs == 2) { acs = 0; acph ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3, grp = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int ty_in = row >> 4, tx_in = row & 15;
    int acs = 0;
    uint32_t acph = 0;
    constexpr int kChunks = N_TILE / 32;
    for (int box = first_box; box < p.total_boxes; box += box_step) {
      const int r = fast_div(box, p.fd_tiles_x), tx = box - r * p.tiles_x;
      const int b = fast_div(r, p.fd_tiles_y), ty = r - b * p.tiles_y;
      const int y = ty * 8 + ty_in, x = tx * 16 + tx_in;
      const bool valid = (y < p.H) && (x < p.W);
      mbar_wait(&acc_full[acs], acph);
      tc_fence_after();
      const uint32_t tbase = tmem_base + (uint32_t(quad * 32) << 16) + uint32_t(acs * kAccCols);
#pragma unroll 1
      for (int cc = grp * (kChunks / kEpiGroups); cc < (grp + 1) * (kChunks / kEpiGroups); ++cc) {
        uint32_t v[32];
        tmem_ld32(tbase + cc * 32, v);
        tmem_ld_wait();
        const int col0 = cc * 32;
        const int pos = fast_div(col0, p.fd_co_t), cl = col0 - pos * p.co_t;
        uint32_t h[16], l[16];
        if (p.out.fmt == ACT_F16 && p.zero_bias) {   // reduced plans: the bias lives in the consuming layer (api.cu, commit)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            h[i] = cvt_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            l[i] = 0u;
          }
        } else if (p.out.fmt == ACT_F16) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bq = *reinterpret_cast<const float4*>(sBias + cl + 4 * i);
            h[2 * i] = cvt_f16x2(__uint_as_float(v[4 * i]) + bq.x, __uint_as_float(v[4 * i + 1]) + bq.y);
            h[2 * i + 1] = cvt_f16x2(__uint_as_float(v[4 * i + 2]) + bq.z, __uint_as_float(v[4 * i + 3]) + bq.w);
            l[2 * i] = l[2 * i + 1] = 0u;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bq = *reinterpret_cast<const float4*>(sBias + cl + 4 * i);
            split_pack2(__uint_as_float(v[4 * i]) + bq.x, __uint_as_float(v[4 * i + 1]) + bq.y, h[2 * i], l[2 * i]);
            split_pack2(__uint_as_float(v[4 * i + 2]) + bq.z, __uint_as_float(v[4 * i + 3]) + bq.w, h[2 * i + 1], l[2 * i + 1]);
          }
        }
        const BoxGeo geo{b, ty * 8, tx * 16, 4, p.H, p.W, 1, pos};
        const StoreMap smap = make_store_map(p.out, geo, lane, quad * 32);   // per chunk: the output phase changes with it
        store_chunk_coalesced(p.out, sScratch + (warp - 2) * kScratchPerWarp, lane, quad * 32, nt * p.co_t + cl, h, l, geo,
                              smap, p.tma_store ? &p.tmapOut : nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acs]);
      if (++acs == 2) { acs = 0; acph ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();   // TMA stores out of this warp's staging buffer are complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}
constexpr int kUpSmem = kUpconvResBytes + 2 * kABytes + kScratchBytes + 1024 + 256 + 256;
constexpr int kUpSmemWide = kUpconvResBytes + 2 * kABytes + 2 * kScratchBytes + 1024 + 256 + 256;   // 16 epilogue warps' staging

}  // namespace

cudaError_t conv_mma_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(conv_mma_kernel<64, EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_mma_kernel<128, EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_mma_kernel<64, EPI_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo_kernel<64, EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo_kernel<64, EPI_ACT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo_kernel<128, EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCfg<128>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo_kernel<64, EPI_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, HCfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<64, EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<128, EPI_ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<128>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<128, EPI_ACT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<128>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<128, EPI_ACT, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<128, 2>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<128, EPI_ACT, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<128, 1>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<64, EPI_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  static_assert(H2Cfg<64, 3, true>::SMEM <= 232448, "resident-weight variant exceeds the shared memory of an SM");
  e = cudaFuncSetAttribute(conv_halo2_kernel<64, EPI_ACT, true, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<64, 3, true>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<64, EPI_HEAD, true, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<64, 3, true>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo2_kernel<64, EPI_ACT, true, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<64>::SMEM);
  if (e != cudaSuccess) return e;
  static_assert(H2Cfg<128, 1, true>::SMEM <= 232448, "resident-weight variant exceeds the shared memory of an SM");
  e = cudaFuncSetAttribute(conv_halo2_kernel<128, EPI_ACT, true, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, H2Cfg<128, 1, true>::SMEM);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upconv_res_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upconv_res_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upconv_res_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmemWide);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upconv_res_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmemWide);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upconv_res_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmemWide);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(upconv_res_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmemWide);
  return e;
}

cudaError_t launch_conv_mma(const ConvParams& p, int n_tile, int epi, int num_sms, cudaStream_t stream) {
  if (p.TW * p.TH != 128) return cudaErrorInvalidValue;
  if (p.do_pool && p.TW != 16) return cudaErrorInvalidValue;
  if (epi == EPI_HEAD) {
    if (n_tile != 64 || p.n_tiles != 1) return cudaErrorInvalidValue;
    return launch_t<64, EPI_HEAD>(p, num_sms, stream);
  }
  if (n_tile == 64) return launch_t<64, EPI_ACT>(p, num_sms, stream);
  if (n_tile == 128) return launch_t<128, EPI_ACT>(p, num_sms, stream);
  return cudaErrorInvalidValue;
}

}  // namespace wsu

namespace wsu {
cudaError_t launch_conv_halo(const ConvParams& p, int n_tile, int epi, int num_sms, cudaStream_t stream) {
  if (p.ntaps != 9 || p.npos != 1 || p.upsample) return cudaErrorInvalidValue;
  if (epi == EPI_HEAD) {
    if (n_tile != 64 || p.n_tiles != 1) return cudaErrorInvalidValue;
    return launch_halo_t<64, EPI_HEAD>(p, num_sms, stream);
  }
  if (n_tile == 64) return p.n_tiles == 1 ? launch_halo_t<64, EPI_ACT>(p, num_sms, stream) : cudaErrorInvalidValue;
  if (n_tile == 128) return launch_halo_t<128, EPI_ACT>(p, num_sms, stream);
  return cudaErrorInvalidValue;
}
}  // namespace wsu

namespace wsu {
cudaError_t launch_upconv_res(const UpconvParams& p, int n_tile, int num_sms, cudaStream_t stream) {
  if (p.cblocks * n_tile * 256 > kUpconvResBytes || n_tile != 4 * p.co_t) return cudaErrorInvalidValue;
  int grid = (num_sms / p.n_tiles) * p.n_tiles;
  const int max_useful = p.total_boxes * p.n_tiles;
  if (grid > max_useful) grid = max_useful;
  if (grid <= 0) return cudaErrorInvalidValue;
  if (n_tile != 256 && n_tile != 128) return cudaErrorInvalidValue;
  if (p.terms == 3) {
    if (n_tile == 256) upconv_res_kernel<256><<<grid, kUpThreads, kUpSmem, stream>>>(p);
    else upconv_res_kernel<128><<<grid, kUpThreads, kUpSmem, stream>>>(p);
  } else if (p.terms == 2) {
    if (n_tile == 256) upconv_res_kernel<256, 2><<<grid, kUpThreadsWide, kUpSmemWide, stream>>>(p);
    else upconv_res_kernel<128, 2><<<grid, kUpThreadsWide, kUpSmemWide, stream>>>(p);
  } else if (p.terms == 1) {
    if (n_tile == 256) upconv_res_kernel<256, 1><<<grid, kUpThreadsWide, kUpSmemWide, stream>>>(p);
    else upconv_res_kernel<128, 1><<<grid, kUpThreadsWide, kUpSmemWide, stream>>>(p);
  } else return cudaErrorInvalidValue;
  return cudaGetLastError();
}
}  // namespace wsu

namespace wsu {
// CTA-pair variant: p.total_items must be ceil(total_sub / 4) * n_tiles
cudaError_t launch_conv_halo2(const ConvParams& p, int n_tile, int epi, int num_sms, cudaStream_t stream) {
  if (p.ntaps != 9 || p.npos != 1 || p.upsample) return cudaErrorInvalidValue;
  if (epi == EPI_HEAD) {
    if (n_tile != 64 || p.n_tiles != 1) return cudaErrorInvalidValue;
    return launch_halo2_t<64, EPI_HEAD>(p, num_sms, stream);
  }
  if (n_tile == 64) return p.n_tiles == 1 ? launch_halo2_t<64, EPI_ACT>(p, num_sms, stream) : cudaErrorInvalidValue;
  if (n_tile == 128) return launch_halo2_t<128, EPI_ACT>(p, num_sms, stream);
  return cudaErrorInvalidValue;
}
}  // namespace wsu
