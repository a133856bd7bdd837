// C ABI of libwsunet (include/wsunet.h): handle, weight packing, per-shape plan (buffers + TMA tensor maps),
// layer chain of UNet.forward (src/unet/model/unet.py:137-189) and the fused / stand-alone WS estimators.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/wsunet.h"
#include "conv_mma.h"
#include "stencil.h"

using namespace wsu;

namespace {

thread_local std::string g_err;
thread_local int64_t g_launches = 0;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(WSU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
  } while (0)
#define LAUNCH_TRY(expr)                                                                           \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    ++g_launches;                                                                                  \
    if (_e != cudaSuccess)                                                                         \
      return fail(WSU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));               \
  } while (0)

// Every entry point runs on the device it was asked for and leaves the calling thread's current device as it found it
// (a process that drives several GPUs from one thread must not have its later torch allocations land elsewhere).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) err = cudaSetDevice(device);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_SCOPE(dev)      \
  DeviceGuard _dev_guard(dev); \
  CUDA_TRY(_dev_guard.err)

// ---------------------------------------------------------------------------------------------- bf16 on the host
uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return uint16_t((u >> 16) | 0x40);  // NaN
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return uint16_t((u + r) >> 16);
}
float bf2f(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
// fp16 on the host (round to nearest even through the toolkit's own conversion)
uint16_t f2h(float f) {
  const __half_raw r = __half_raw(__float2half_rn(f));
  return r.x;
}
float h2f(uint16_t v) {
  __half_raw r;
  r.x = v;
  return __half2float(__half(r));
}
// e4m3 on the host (round to nearest even, saturating to +-448), the toolkit's own conversion
uint8_t f2e4m3(float f) { return uint8_t(__nv_cvt_float_to_fp8(f, __NV_SATFINITE, __NV_E4M3)); }
inline size_t sw128_off(int row, int k) {  // byte offset of bf16 element (row, k) inside a K-major SWIZZLE_128B tile
  return size_t(row) * 128 + size_t(((k >> 3) ^ (row & 7)) << 4) + size_t(k & 7) * 2;
}

// ---------------------------------------------------------------------------------------------- driver entry point
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 5-D map over a split-bf16 NHWC activation: dims (C, W+2, H+2, B, plane), box (64, TW, TH, 1, 2)
int make_act_tmap(CUtensorMap* m, const Act& a, int TW, int TH) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(WSU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t Wp = a.W + 2, Hp = a.H + 2;
  const cuuint32_t planes = a.fmt == ACT_F16 ? 1 : 2;   // 2-byte elements either way (bf16 pair or one fp16 plane)
  cuuint64_t dims[5] = {cuuint64_t(a.C), Wp, Hp, cuuint64_t(a.B), planes};
  cuuint64_t strides[4] = {cuuint64_t(a.C) * 2, Wp * a.C * 2, Hp * Wp * a.C * 2, cuuint64_t(a.plane) * 2};
  cuuint32_t box[5] = {64, cuuint32_t(TW), cuuint32_t(TH), 1, planes};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, a.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(WSU_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(int(r)));
  return WSU_OK;
}

// 2-D map over a layer's packed weights seen as rows of 128 B (64 two-byte elements): box = 64 rows = one CTA's half of a
// 128-row tile. No swizzle: the tiles are stored in the order the tensor core reads them.
int make_w_tmap(CUtensorMap* m, const void* wpack, size_t bytes, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(WSU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {64, cuuint64_t(bytes / 128)};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, cuuint32_t(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wpack), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(WSU_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed with code " + std::to_string(int(r)));
  return WSU_OK;
}

// 5-D map for TMA STORES of a 32-pixel x 32-channel chunk: dims as make_act_tmap, box (32 ch, 8 px, 4 rows, 1, 1); the
// 64-byte inner rows are XOR-swizzled (SWIZZLE_64B) exactly like the epilogue's staging buffer.
// up = true: the map of ONE output phase of a 2x up-convolution - a warp's chunk is 2 input rows x 16 input pixels, whose
// outputs of phase (dy, dx) sit at every second pixel of every second row: box (32 ch, 16 px, 2 rows) traversed with element
// strides (1, 2, 2); the start coordinate carries the phase.
int make_out_tmap(CUtensorMap* m, const Act& a, bool up = false) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(WSU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t Wp = a.W + 2, Hp = a.H + 2;
  const cuuint32_t planes = a.fmt == ACT_F16 ? 1 : 2;
  cuuint64_t dims[5] = {cuuint64_t(a.C), Wp, Hp, cuuint64_t(a.B), planes};
  cuuint64_t strides[4] = {cuuint64_t(a.C) * 2, Wp * a.C * 2, Hp * Wp * a.C * 2, cuuint64_t(a.plane) * 2};
  cuuint32_t box[5] = {32, 8, 4, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (up) {   // boxDim counts traversed positions: ceil(32 / 2) = 16 pixels, ceil(4 / 2) = 2 rows
    box[1] = 32; box[2] = 4;
    estr[1] = 2; estr[2] = 2;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, a.base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(WSU_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed with code " + std::to_string(int(r)));
  return WSU_OK;
}

// ---------------------------------------------------------------------------------------------- model description
struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> dims;
};

struct LayerW {  // device-side packed weights of one tensor-core layer
  uint8_t* wpack = nullptr;
  float* bias = nullptr;
  std::vector<float> bias_host;   // the uploaded bias (all zero when it was folded into the consuming layer)
  int cin = 0, cout = 0, n_tile = 0, ntaps = 0, npos = 0;
  int terms = 3;            // MMAs per MAC: 3 = split-bf16 weights against split-bf16 inputs; 2 / 1 = fp16 (hi, lo) / hi weights
                            // against ONE fp16 input plane (precision plan, see wsu_context::precision)
  size_t wpack_bytes = 0;
  int src0_terms = 2;       // MMAs per MAC of the fp16 source-0 blocks (f16_cblocks0 > 0): 2, or 1 under the fp16 + fp8 plan
  bool f8 = false;          // the blocks after f16_cblocks0 are ACT_F16F8 maps: fp16 main tile + e4m3 correction tile
  float corr_scale = 1.f;   // 1 / (kF8ScaleA2 * pw): undoes the scaling of the correction accumulator
  int f16_cblocks0 = 0;     // decoder layer at level 0 under a reduced plan: its first channel blocks (the up-convolution's
                            // output) are fp16 pairs against ONE fp16 plane, the skip half stays three-term
  uint8_t* wres = nullptr;  // transposed conv only: phase-stacked tiles for the resident-weight kernel (may stay null)
  int res_ntile = 0, res_cot = 0;
};

struct Plan {  // everything that depends on (micro-batch, H, W)
  int mb = 0, H = 0, W = 0;
  std::map<std::string, Act> acts;
  std::vector<void*> allocs;
  float* partials = nullptr;
  size_t arena_bytes = 0;
  int tiles_per_img = 0;
  std::vector<std::pair<ConvParams, std::pair<int, int>>> convs;  // params, (n_tile, epi); head is the last entry
  std::vector<std::string> conv_names;
  std::vector<UpconvParams> ups;   // resident-weight up-convolutions
  std::vector<int> up_idx;         // per conv step: index into ups or -1
};

}  // namespace

struct wsu_context {
  int device = 0, nsteps = 0, in_ch = 1, out_ch = 1, num_sms = 148;
  int64_t micro_batch = 0;
  std::map<std::string, HostTensor> host_w;
  bool committed = false;
  std::map<std::string, LayerW> layers;
  float* e11_w = nullptr;
  float* e11_b = nullptr;
  float wout[64];
  float bout = 0.f;
  std::unique_ptr<Plan> plan;
  // host-path staging
  uint8_t* stage_img[2] = {nullptr, nullptr};
  float* stage_out = nullptr;
  size_t stage_imgs = 0, stage_px = 0;
  cudaStream_t s_copy = nullptr, s_comp = nullptr;
  // The activation buffers of the plan are shared by every call on this handle. Calls may arrive on different streams
  // (torch's current stream, the host-buffer path's own compute stream): each call records ev_chain when it is done and
  // a call on another stream waits for it first, so two passes never overlap in the buffers.
  cudaEvent_t ev_chain = nullptr;
  cudaStream_t chain_stream = nullptr;
  bool chain_pending = false;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  // optional per-layer timing of the last micro-batch ("profile" option)
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<std::string> prof_names;
  int prof_n = 0;
  int last_nimg = 0;  // images in the last micro-batch that ran
  // compute e11 inside e12's producer warps (option "fuse_e11"). Bit-identical, saves 134 MB/image of HBM traffic, but the
  // three CUDA-core producer warps cannot keep up with the tensor pipe (e12 2.66 ms vs 0.56 + 1.83 ms per 32 images), so it
  // is off by default.
  bool fuse_e11 = false;
  int use_pair = 2;  // 3x3 layers as CTA pairs (tcgen05 cta_group::2): 0 never, 1 Cout>=128 layers only, 2 all (default: with the weight
                     // tiles loaded by cta_group::2 TMA the Cout=64 layers gain 3-6 % too; they lost before, behind the per-tap relay fence)
  int dbg = 0;            // env WSU_DBG: knock-out switches for timing experiments (results are wrong when set)
  int a_collector = 1;    // option "a_collector" (default on, +0.4 % measured): Cout >= 128 layers reuse A_hi from the A collector (hi*hi, hi*lo, lo*hi order)
  int l2_prefetch = 0;    // halo kernels prefetch the next item's boxes into L2 (option "l2_prefetch"); measured 1 % slower
  // Precision plan (option "precision"). 0: every layer three-term split-bf16 (16 significand bits per operand; 2e-5 px).
  // 1 / 2: the layers whose INPUT lives at level >= 1 of the UNet (e21.., d3x, up-convolutions) read one fp16 plane against
  // fp16 (hi, lo) / fp16 hi-only weights = 2 / 1 MMAs per MAC and half the activation bytes; the full-resolution layers
  // (e12, d41, d42), whose rounding reaches the output directly, stay three-term. Whether a plan keeps a given model inside
  // the 1e-3 px bar depends on its weights: UNet.calibrate_precision() measures it against plan 0 and picks.
  int w_resident = 1;         // option "w_resident": e12 / d42 under the fp16 + fp8 plan keep their 72 KB of weights in shared memory
  int tma_store = 1;          // option "tma_store": interior boxes of the 3x3 halo kernels are written by TMA tensor stores out of
                              // the staging buffer (one cp.async.bulk.tensor per 32 x 32 chunk and plane instead of 4 LDS + 4 STG per
                              // lane). Bit-identical. Measured with the variants interleaved pass by pass: -1.3 % of the chain under
                              // the three-term plan, -3.1 % under fp16x1, -1.8 % under fp16x1_f8 (e12 -5..7 %, e21 -10 %); the
                              // first A/B of this option ran the variants one after the other and read as neutral because the
                              // clocks sag during a process.
  bool alias_buffers = true;  // option "alias_buffers": feature maps with disjoint lifetimes share arena bytes
  // 3: plan 2, and the full-resolution 3x3 layers (e12, d41's skip half, d42) read ACT_F16F8 maps: one fp16 MMA for the main
  // product and ONE e4m3 MMA for both correction terms (two MMA times instead of three at ~15 significant bits - this part
  // does not depend on the weights, only the deep-layer part of the plan does).
  int precision = 0;
  int precision_active = 0;   // what commit could honour (needs resident up-convolutions: unet_1, unet_2)
  bool use_upres = true;  // transposed convs through upconv_res_kernel (option "upconv_resident")
  bool use_halo = true;  // 3x3 layers through conv_halo_kernel (option "halo"; 0 = per-tap reload kernel, for A/B runs)
};

namespace {

std::string enc_name(int level, int idx) { return "e" + std::to_string(level + 1) + std::to_string(idx); }
std::string dec_name(int level, int idx) { return "d" + std::to_string(4 - level) + std::to_string(idx); }
std::string up_name(int level) { return "upconv" + std::to_string(4 - level); }
int chan(int level) { return 64 << level; }

void free_plan(Plan* p) {
  if (!p) return;
  for (void* q : p->allocs) cudaFree(q);
  p->allocs.clear();
}

// Feature maps are carved out of ONE arena. A map lives from the step that writes it to the last step that reads it; maps
// whose lifetimes do not overlap share bytes (first-fit over the maps alive at the newcomer's birth). With "alias_buffers"
// off (tests that inspect every layer after a pass) every map keeps its own range.
struct ActDecl {
  std::string name;
  Act a;
  size_t bytes = 0, offset = 0;
  int birth = 0, death = 0;
};

int declare_act(std::vector<ActDecl>& decls, const std::string& name, int B, int H, int W, int C, int fmt) {
  ActDecl d;
  d.name = name;
  d.a.B = B; d.a.H = H; d.a.W = W; d.a.C = C;
  d.a.fmt = fmt;
  d.a.plane = act_plane_elems(B, H, W, C);
  d.a.base = nullptr;
  d.bytes = ((d.a.plane * (fmt == ACT_F16 ? 1 : 2) * sizeof(__nv_bfloat16)) + 1023) & ~size_t(1023);   // TMA / swizzle friendly
  d.birth = -1;
  d.death = -1;
  decls.push_back(d);
  return WSU_OK;
}

size_t layout_acts(std::vector<ActDecl>& decls, bool alias) {
  std::vector<size_t> order(decls.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return decls[x].birth < decls[y].birth; });
  size_t total = 0;
  std::vector<size_t> placed;
  for (size_t oi : order) {
    ActDecl& d = decls[oi];
    size_t off = 0;
    if (!alias) {
      off = total;
    } else {
      // lowest offset whose range is free of every map still alive when this one is written (a step's inputs are alive
      // while it writes: death >= birth counts as overlap)
      bool moved = true;
      while (moved) {
        moved = false;
        for (size_t pj : placed) {
          const ActDecl& q = decls[pj];
          const bool time_overlap = !(q.death < d.birth || d.death < q.birth);
          if (time_overlap && off < q.offset + q.bytes && q.offset < off + d.bytes) { off = q.offset + q.bytes; moved = true; }
        }
      }
    }
    d.offset = off;
    total = std::max(total, off + d.bytes);
    placed.push_back(oi);
  }
  return total;
}

int place_acts(Plan& pl, std::vector<ActDecl>& decls, bool alias, cudaStream_t st) {
  const size_t total = layout_acts(decls, alias);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, total);
  if (e != cudaSuccess) return fail(WSU_ERR_CUDA, "cudaMalloc(feature-map arena, " + std::to_string(total) + " B): " + cudaGetErrorString(e));
  pl.allocs.push_back(p);
  pl.arena_bytes = total;
  // overhanging reads stay finite. Zero-filled ON THE CALLER'S STREAM: the chain kernels run there, and a non-blocking
  // stream does not order itself behind a legacy-stream memset.
  e = cudaMemsetAsync(p, 0, total, st);
  if (e != cudaSuccess) return fail(WSU_ERR_CUDA, std::string("cudaMemsetAsync(arena): ") + cudaGetErrorString(e));
  for (ActDecl& d : decls) {
    d.a.base = reinterpret_cast<__nv_bfloat16*>(static_cast<uint8_t*>(p) + d.offset);
    pl.acts[d.name] = d.a;
  }
  return WSU_OK;
}

// Build a conv step. src1 may be null (no concat).
int add_conv(wsu_context* h, Plan& pl, const std::string& lname, const LayerW& lw, const Act& src0, const Act* src1, const Act* out, const Act* pool,
             bool upsample, bool relu, int epi) {
  ConvParams p;
  std::memset(static_cast<void*>(&p), 0, sizeof(p));
  const int TW = 16, TH = 8;
  const int m_sub = 256 / lw.n_tile;
  int rc = make_act_tmap(&p.tmapA0, src0, TW, TH);
  if (rc) return rc;
  rc = make_act_tmap(&p.tmapA1, src1 ? *src1 : src0, TW, TH);
  if (rc) return rc;
  rc = make_act_tmap(&p.tmapH0, src0, kHaloTW + 2, kHaloTH + 2);
  if (rc) return rc;
  rc = make_act_tmap(&p.tmapH1, src1 ? *src1 : src0, kHaloTW + 2, kHaloTH + 2);
  if (rc) return rc;
  p.wpack = lw.wpack;
  if ((rc = make_w_tmap(&p.tmapW, lw.wpack, lw.wpack_bytes, 64))) return rc;
  if ((rc = make_w_tmap(&p.tmapW32, lw.wpack, lw.wpack_bytes, 32))) return rc;
  p.bias = lw.bias;
  p.cblocks0 = src0.C / 64;
  p.cblocks = lw.cin / 64;
  p.ntaps = lw.ntaps;
  for (int t = 0; t < 9; ++t) {
    p.tap_dx[t] = (lw.ntaps == 9) ? t % 3 : 1;
    p.tap_dy[t] = (lw.ntaps == 9) ? t / 3 : 1;
  }
  p.npos = lw.npos;
  p.terms = lw.terms;
  p.src0_f16 = lw.f16_cblocks0 > 0 ? (lw.src0_terms == 1 ? 2 : 1) : 0;
  p.f8_blocks = lw.f8 ? 1 : 0;
  p.corr_scale = lw.corr_scale;
  {
    const bool in0_f16 = src0.fmt == ACT_F16, in1_f16 = src1 ? src1->fmt == ACT_F16 : in0_f16;
    const int rest_fmt = p.src0_f16 ? (src1 ? src1->fmt : -1) : src0.fmt;   // format of the blocks that are not fp16 source-0 blocks
    bool ok = p.src0_f16 ? (lw.terms == 3 && in0_f16 && !in1_f16 && lw.f16_cblocks0 == src0.C / 64 && lw.n_tile == 64)
                         : ((lw.terms != 3) == in0_f16 && in1_f16 == in0_f16);
    ok = ok && (lw.f8 ? (rest_fmt == ACT_F16F8 && lw.n_tile == 64) : rest_fmt != ACT_F16F8);
    if (!ok) return fail(WSU_ERR_STATE, "internal: layer " + lname + " and its input maps disagree about the precision plan");
  }
  p.n_tiles = lw.cout / lw.n_tile;
  p.cout = lw.cout;
  p.B = src0.B; p.H = src0.H; p.W = src0.W;
  p.TW = TW; p.TH = TH;
  p.tiles_x = (p.W + TW * m_sub - 1) / (TW * m_sub);
  p.tiles_y = (p.H + TH - 1) / TH;
  p.total_tiles = p.B * p.tiles_y * p.tiles_x * p.n_tiles * p.npos;
  p.sub_x = (p.W + kHaloTW - 1) / kHaloTW;
  p.sub_y = (p.H + kHaloTH - 1) / kHaloTH;
  {
    const uint64_t nmax = uint64_t(p.B) * p.sub_x * p.sub_y * std::max(1, p.n_tiles) + 8;   // largest dividend any kernel forms
    auto fd = [&](int d) {
      FastDiv f;
      f.d = uint32_t(d);
      f.mode = d == 1 ? 1u : (nmax * uint64_t(d) < (uint64_t(1) << 32) ? 0u : 2u);
      f.mul = d > 1 ? uint32_t((uint64_t(1) << 32) / uint64_t(d)) + 1u : 0u;
      return f;
    };
    p.fd_sub_x = fd(p.sub_x);
    p.fd_sub_y = fd(p.sub_y);
    p.fd_n_tiles = fd(p.n_tiles);
  }
  p.total_sub = p.B * p.sub_x * p.sub_y;
  p.total_items = ((p.total_sub + halo_msub(lw.n_tile) - 1) / halo_msub(lw.n_tile)) * p.n_tiles;
  p.relu = relu;
  p.upsample = upsample;
  if (out) p.out = *out;
  if (out && !upsample) {
    if ((rc = make_out_tmap(&p.tmapOut, *out))) return rc;
  }
  if (pool) { p.do_pool = 1; p.pool = *pool; }
  if (epi == EPI_HEAD) {
    std::memcpy(p.wout, h->wout, sizeof(p.wout));
    p.bout = h->bout;
    // partial records per image: 8 per super-tile (per-tap kernel) or 4 per box (halo kernels), whichever is larger
    pl.tiles_per_img = std::max(2 * p.tiles_x * p.tiles_y, p.sub_x * p.sub_y);
  }
  pl.convs.push_back({p, {lw.n_tile, epi}});
  pl.conv_names.push_back(lname);
  int ui = -1;
  if (upsample && lw.wres && out) {
    UpconvParams u;
    std::memset(static_cast<void*>(&u), 0, sizeof(u));
    rc = make_act_tmap(&u.tmapA, src0, 16, 8);
    if (rc) return rc;
    u.wres = lw.wres;
    u.bias = lw.bias;
    u.cblocks = lw.cin / 64;
    u.terms = lw.terms;
    u.co_t = lw.res_cot;
    u.n_tiles = lw.cout / lw.res_cot;
    u.B = src0.B; u.H = src0.H; u.W = src0.W;
    u.tiles_x = (u.W + 15) / 16;
    u.tiles_y = (u.H + 7) / 8;
    u.total_boxes = u.B * u.tiles_x * u.tiles_y;
    {
      const uint64_t nmax = uint64_t(u.total_boxes) + 1024;
      auto fd = [&](int d) {
        FastDiv f;
        f.d = uint32_t(d);
        f.mode = d == 1 ? 1u : (nmax * uint64_t(d) < (uint64_t(1) << 32) ? 0u : 2u);
        f.mul = d > 1 ? uint32_t((uint64_t(1) << 32) / uint64_t(d)) + 1u : 0u;
        return f;
      };
      u.fd_tiles_x = fd(u.tiles_x);
      u.fd_tiles_y = fd(u.tiles_y);
      u.fd_co_t = fd(u.co_t);
    }
    u.out = *out;
    if ((rc = make_out_tmap(&u.tmapOut, *out, true))) return rc;
    u.zero_bias = std::all_of(lw.bias_host.begin(), lw.bias_host.end(), [](float v) { return v == 0.f; }) ? 1 : 0;
    ui = int(pl.ups.size());
    pl.ups.push_back(u);
  }
  pl.up_idx.push_back(ui);
  return WSU_OK;
}

int build_plan_impl(wsu_context* h, int mb, int H, int W, cudaStream_t st);

int build_plan(wsu_context* h, int mb, int H, int W, cudaStream_t st) {
  if (h->plan && h->plan->mb >= mb && h->plan->H == H && h->plan->W == W) return WSU_OK;
  if (h->plan) { cudaDeviceSynchronize(); free_plan(h->plan.get()); }
  h->plan.reset(new Plan());
  const int rc = build_plan_impl(h, mb, H, W, st);
  if (rc != WSU_OK) {  // never keep a half-built plan around
    free_plan(h->plan.get());
    h->plan.reset();
  }
  return rc;
}

struct Step { std::string layer, src0, src1, out, pool; bool up, relu; int epi; };

// feature maps (with lifetimes) and layer steps of one pass over `mb` images of H x W
void describe_chain(const wsu_context* h, int mb, int H, int W, std::vector<ActDecl>& decls, std::vector<Step>& steps) {
  const int n = h->nsteps;
  // activations. Under a reduced-precision plan every map at level >= 1 (and every up-convolution output) is read only by
  // one-/two-term layers or channel blocks and is stored as ONE fp16 plane; e11, e12, d41 feed three-term layers and stay
  // split-bf16.
  auto fmt_at = [&](int level) {
    if (h->precision_active != 0 && level >= 1) return int(ACT_F16);
    return h->precision_active == 3 ? int(ACT_F16F8) : int(ACT_SPLIT);   // level-0 maps read by the full-resolution 3x3 layers
  };
  for (int l = 0; l <= n; ++l) {
    const int hh = H >> l, ww = W >> l;
    declare_act(decls, enc_name(l, 1), mb, hh, ww, chan(l), fmt_at(l));
    if (!(n == 0)) declare_act(decls, enc_name(l, 2), mb, hh, ww, chan(l), fmt_at(l));
    if (l < n) declare_act(decls, "p" + std::to_string(l + 1), mb, hh / 2, ww / 2, chan(l), fmt_at(l + 1));
  }
  for (int l = n - 1; l >= 0; --l) {
    const int hh = H >> l, ww = W >> l;
    // up-convolution outputs are fp16 at every level (u4 feeds d41's two-term source-0 blocks); their bias is folded
    // into the consuming layer's bias at commit, which keeps the stored values small and their rounding harmless
    declare_act(decls, "u" + std::to_string(4 - l), mb, hh, ww, chan(l), fmt_at(l + 1));
    declare_act(decls, dec_name(l, 1), mb, hh, ww, chan(l), fmt_at(l));
    if (l > 0) declare_act(decls, dec_name(l, 2), mb, hh, ww, chan(l), fmt_at(l));
  }
  // layer chain (first conv = step 0 is launched separately); "" = no such tensor
  for (int l = 0; l <= n; ++l) {
    if (l > 0) steps.push_back({enc_name(l, 1), "p" + std::to_string(l), "", enc_name(l, 1), "", false, true, EPI_ACT});
    if (n == 0) steps.push_back({enc_name(0, 2), enc_name(0, 1), "", "", "", false, true, EPI_HEAD});
    else steps.push_back({enc_name(l, 2), enc_name(l, 1), "", enc_name(l, 2), (l < n) ? "p" + std::to_string(l + 1) : "", false, true, EPI_ACT});
  }
  for (int l = n - 1; l >= 0; --l) {
    const std::string below = (l + 1 == n) ? enc_name(l + 1, 2) : dec_name(l + 1, 2);
    const std::string u = "u" + std::to_string(4 - l);
    steps.push_back({up_name(l), below, "", u, "", true, false, EPI_ACT});
    steps.push_back({dec_name(l, 1), u, enc_name(l, 2), dec_name(l, 1), "", false, true, EPI_ACT});
    if (l > 0) steps.push_back({dec_name(l, 2), dec_name(l, 1), "", dec_name(l, 2), "", false, true, EPI_ACT});
    else steps.push_back({dec_name(l, 2), dec_name(l, 1), "", "", "", false, true, EPI_HEAD});
  }
  auto decl_of = [&](const std::string& nm) -> ActDecl* {
    for (ActDecl& d : decls) if (d.name == nm) return &d;
    return nullptr;
  };
  auto touch = [&](const std::string& nm, int t, bool write) {
    if (nm.empty()) return;
    ActDecl* d = decl_of(nm);
    if (write && d->birth < 0) d->birth = t;
    d->death = std::max(d->death, t);
  };
  touch(enc_name(0, 1), 0, true);
  for (size_t i = 0; i < steps.size(); ++i) {
    const int t = int(i) + 1;
    touch(steps[i].src0, t, false);
    touch(steps[i].src1, t, false);
    touch(steps[i].out, t, true);
    touch(steps[i].pool, t, true);
  }
  for (ActDecl& d : decls) if (d.birth < 0) { d.birth = 0; d.death = int(steps.size()) + 1; }   // never written: keep apart
}

// feature-map bytes of one image under the current options (what the micro-batch budget divides)
size_t per_image_bytes(const wsu_context* h, int H, int W) {
  std::vector<ActDecl> decls;
  std::vector<Step> steps;
  describe_chain(h, 1, H, W, decls, steps);
  return layout_acts(decls, h->alias_buffers);
}

int build_plan_impl(wsu_context* h, int mb, int H, int W, cudaStream_t st) {
  Plan& pl = *h->plan;
  pl.mb = mb; pl.H = H; pl.W = W;
  int rc;
  std::vector<ActDecl> decls;
  std::vector<Step> steps;
  describe_chain(h, mb, H, W, decls, steps);
  if ((rc = place_acts(pl, decls, h->alias_buffers, st))) return rc;
  auto A = [&](const std::string& s) -> Act& { return pl.acts.at(s); };
  for (const Step& sp : steps) {
    if ((rc = add_conv(h, pl, sp.layer, h->layers.at(sp.layer), A(sp.src0), sp.src1.empty() ? nullptr : &A(sp.src1),
                       sp.out.empty() ? nullptr : &A(sp.out), sp.pool.empty() ? nullptr : &A(sp.pool), sp.up, sp.relu, sp.epi)))
      return rc;
  }
  void* pp = nullptr;
  CUDA_TRY(cudaMalloc(&pp, size_t(mb) * pl.tiles_per_img * 4 * kPartialSlots * sizeof(float)));
  pl.allocs.push_back(pp);
  pl.partials = static_cast<float*>(pp);
  return WSU_OK;
}

int check_shape(wsu_context* h, int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return fail(WSU_ERR_INVALID, "B, H, W must be positive");
  const int div = 1 << h->nsteps;
  if (H % div || W % div)
    return fail(WSU_ERR_INVALID, "H and W must be divisible by 2^nsteps (torch.cat would raise in the reference, unet.py:178)");
  if ((H >> h->nsteps) < 2 || (W >> h->nsteps) < 2)
    return fail(WSU_ERR_INVALID, "reflect padding needs >= 2 pixels per dimension at the deepest level");
  return WSU_OK;
}

int pick_micro_batch(wsu_context* h, int B, int H, int W) {
  if (h->micro_batch > 0) return int(std::min<int64_t>(h->micro_batch, B));
  const size_t budget = size_t(20) << 30;
  int mb = int(std::max<size_t>(1, budget / per_image_bytes(h, H, W)));
  mb = std::min(std::min(mb, 64), B);
  // even passes: avoid a short ragged tail pass (e.g. B=256 -> 8 x 32 instead of 7 x 33 + 25)
  const int passes = (B + mb - 1) / mb;
  return (B + passes - 1) / passes;
}

// one micro-batch through the layer chain. img points at this micro-batch's first image.
// bias_pass: the predictor runs on the LSB-difference image (x_bar - x) / 255 of `img` and the head adds sum w (x - x_bar) x_bias
// to the partial records of the preceding normal pass (src/ws/estimate.py:126-128)
int run_chain(wsu_context* h, const void* img, int img_dtype, int nimg, const void* ws_img, int ws_dtype, float* yhat,
              int weighted, int crop, cudaStream_t st, bool bias_pass = false) {
  Plan& pl = *h->plan;
  h->last_nimg = nimg;
  Act first = pl.acts.at(enc_name(0, 1));
  auto mark = [&](size_t i) {
    if (!h->profile) return;
    while (h->prof_ev.size() <= i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      h->prof_ev.push_back(e);
    }
    cudaEventRecord(h->prof_ev[i], st);
  };
  if (h->profile) {
    h->prof_names.assign(1, enc_name(0, 1));
    h->prof_names.insert(h->prof_names.end(), pl.conv_names.begin(), pl.conv_names.end());
    h->prof_n = int(h->prof_names.size());
  }
  mark(0);
  // a short last micro-batch reuses the plan: only the first `nimg` images of each buffer are live
  bool fuse_first = false;
  if (!pl.convs.empty()) {
    const ConvParams& c0 = pl.convs[0].first;
    const int nt0 = pl.convs[0].second.first, epi0 = pl.convs[0].second.second;
    fuse_first = h->fuse_e11 && h->use_halo && h->in_ch == 1 && epi0 == EPI_ACT && nt0 == 64 &&
                 c0.cblocks == 1 && c0.ntaps == 9 && !bias_pass && !c0.f8_blocks;
  }
  if (!fuse_first)
    LAUNCH_TRY(launch_first_conv(img, bias_pass ? 2 : (img_dtype == WSU_F32 ? 1 : 0), h->in_ch, h->e11_w, h->e11_b,
                                 Act{first.base, first.plane, nimg, first.H, first.W, first.C, first.fmt}, st));
  for (size_t i = 0; i < pl.convs.size(); ++i) {
    ConvParams p = pl.convs[i].first;
    const int n_tile = pl.convs[i].second.first, epi = pl.convs[i].second.second;
    const bool halo = h->use_halo && p.ntaps == 9;
    p.l2_prefetch = h->l2_prefetch;
    p.a_collector = h->a_collector;
    p.dbg = h->dbg;
    p.tma_store = (h->tma_store && halo && epi == EPI_ACT && !p.upsample) ? 1 : 0;
    p.w_resident = h->w_resident;
    if (nimg != pl.mb) {
      p.B = nimg;
      p.total_tiles = nimg * p.tiles_y * p.tiles_x * p.n_tiles * p.npos;
      p.total_sub = nimg * p.sub_x * p.sub_y;
      p.total_items = ((p.total_sub + halo_msub(n_tile) - 1) / halo_msub(n_tile)) * p.n_tiles;
    }
    if (i == 0 && fuse_first) {
      p.fuse_img = img;
      p.fuse_img_is_float = (img_dtype == WSU_F32);
      p.fuse_w = h->e11_w;
      p.fuse_b = h->e11_b;
    }
    if (epi == EPI_HEAD) {
      p.img = ws_img;
      p.img_is_float = (ws_dtype == WSU_F32);
      p.yhat = yhat;
      p.partials = ws_img ? pl.partials : nullptr;
      p.weighted = weighted;
      p.crop = crop;
      p.bias_pass = bias_pass ? 1 : 0;
    }
    mark(i + 1);
    if (i == 0 && fuse_first) {
      LAUNCH_TRY(launch_conv_halo(p, n_tile, epi, h->num_sms, st));   // e11 computed by e12's producer warps (single-CTA kernel)
    } else if (p.f8_blocks || p.src0_f16 == 2) {   // fp16 + fp8 blocks / one-term source-0 blocks exist in the CTA-pair kernel only
      p.total_items = ((p.total_sub + 3) / 4) * p.n_tiles;
      LAUNCH_TRY(launch_conv_halo2(p, n_tile, epi, h->num_sms, st));
    } else if (p.src0_f16) {
      if (h->use_pair == 2) {
        p.total_items = ((p.total_sub + 3) / 4) * p.n_tiles;
        LAUNCH_TRY(launch_conv_halo2(p, n_tile, epi, h->num_sms, st));
      } else {
        LAUNCH_TRY(launch_conv_halo(p, n_tile, epi, h->num_sms, st));
      }
    } else if (pl.up_idx[i] >= 0 && (h->use_upres || p.terms != 3)) {
      UpconvParams u = pl.ups[pl.up_idx[i]];
      if (nimg != pl.mb) { u.B = nimg; u.total_boxes = nimg * u.tiles_x * u.tiles_y; }
      u.tma_store = h->tma_store;
      LAUNCH_TRY(launch_upconv_res(u, 4 * u.co_t, h->num_sms, st));
    } else if (p.ntaps == 9 && (p.terms != 3 || (halo && (h->use_pair == 2 || (h->use_pair == 1 && n_tile == 128))))) {
      p.total_items = ((p.total_sub + 3) / 4) * p.n_tiles;   // pair items: 2 slots x 2 CTAs = 4 boxes
      LAUNCH_TRY(launch_conv_halo2(p, n_tile, epi, h->num_sms, st));
    } else if (halo)
      LAUNCH_TRY(launch_conv_halo(p, n_tile, epi, h->num_sms, st));
    else
      LAUNCH_TRY(launch_conv_mma(p, n_tile, epi, h->num_sms, st));
  }
  mark(pl.convs.size() + 1);
  return WSU_OK;
}

int unet_ws_device(wsu_context* h, const void* img, int dtype, int B, int H, int W, bool want_ws, int weighted, int clip,
                   int crop, int correct_bias, float* beta, float* l1, float* yhat, cudaStream_t st) {
  if (!h) return fail(WSU_ERR_INVALID, "null handle");
  if (!h->committed) return fail(WSU_ERR_STATE, "weights not committed (call wsu_commit_weights)");
  int rc = check_shape(h, B, H, W);
  if (rc) return rc;
  if (dtype != WSU_U8 && dtype != WSU_F32) return fail(WSU_ERR_INVALID, "dtype must be WSU_U8 or WSU_F32");
  if (want_ws) {
    if (h->in_ch != 1) return fail(WSU_ERR_INVALID, "the WS estimator is defined for single-channel images");
    if (weighted < -1 || weighted > 1) return fail(WSU_ERR_INVALID, "weighted must be -1, 0 or 1");
    if (weighted != 0 && !crop)
      return fail(WSU_ERR_INVALID, "local-variance weights exist only on the interior (crop=1), estimate.py:94-96");
    if (crop && (H < 3 || W < 3)) return fail(WSU_ERR_INVALID, "crop=1 needs H, W >= 3");
    if (correct_bias && dtype != WSU_U8)
      return fail(WSU_ERR_INVALID, "correct_bias needs uint8 images (the second pass runs on the integer difference x_bar - x, estimate.py:127)");
  }
  DEVICE_SCOPE(h->device);
  if (!h->ev_chain) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_chain, cudaEventDisableTiming));
  if (h->chain_pending && h->chain_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chain, 0));
  const int mb = pick_micro_batch(h, B, H, W);
  if ((rc = build_plan(h, mb, H, W, st))) return rc;
  const size_t px = size_t(H) * W;
  const size_t esz = dtype == WSU_F32 ? 4 : 1;
  struct ChainDone {   // records the end of this call's work on every exit path
    wsu_context* h;
    cudaStream_t st;
    ~ChainDone() {
      if (cudaEventRecord(h->ev_chain, st) == cudaSuccess) { h->chain_stream = st; h->chain_pending = true; }
    }
  } chain_done{h, st};
  for (int b0 = 0; b0 < B; b0 += mb) {
    const int nimg = std::min(mb, B - b0);
    const uint8_t* im = static_cast<const uint8_t*>(img) + size_t(b0) * h->in_ch * px * esz;
    if ((rc = run_chain(h, im, dtype, nimg, want_ws ? im : nullptr, dtype, yhat ? yhat + size_t(b0) * px : nullptr, weighted, crop,
                        st)))
      return rc;
    if (want_ws && correct_bias) {   // second predictor pass on the difference image; only slot 3 of the partial records changes
      if ((rc = run_chain(h, im, dtype, nimg, im, dtype, nullptr, weighted, crop, st, true))) return rc;
    }
    if (want_ws) {
      const float npix = crop ? float(H - 2) * float(W - 2) : float(H) * float(W);
      const ConvParams& hp = h->plan->convs.back().first;
      const int records = h->use_halo ? hp.sub_x * hp.sub_y * 4 : hp.tiles_x * hp.tiles_y * 8;  // one per epilogue warp
      LAUNCH_TRY(launch_finalize(h->plan->partials, records, nimg, npix, clip, correct_bias ? 1 : 0, beta + b0,
                                 l1 ? l1 + b0 : nullptr, st));
    }
  }
  return WSU_OK;
}

int filter_common_check(const void* img, int dtype, int kind, int B, int H, int W) {
  if (!img) return fail(WSU_ERR_INVALID, "null image pointer");
  if (dtype != WSU_U8 && dtype != WSU_F32) return fail(WSU_ERR_INVALID, "dtype must be WSU_U8 or WSU_F32");
  if (kind < WSU_PRED_KB || kind > WSU_PRED_ID) return fail(WSU_ERR_INVALID, "unknown predictor kind");
  if (B <= 0 || H < 3 || W < 3) return fail(WSU_ERR_INVALID, "need B >= 1 and H, W >= 3 ('valid' 3x3)");
  return WSU_OK;
}

}  // namespace

// ================================================================================================ exported
extern "C" {

const char* wsu_last_error(void) { return g_err.c_str(); }
int wsu_version(void) { return 200; }
int64_t wsu_launch_count(int reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int wsu_create(wsu_handle* out, int device, int nsteps, int in_channels, int out_channels) {
  if (!out) return fail(WSU_ERR_INVALID, "null out pointer");
  if (nsteps < 0 || nsteps > 4) return fail(WSU_ERR_INVALID, "nsteps must be in 0..4 (unet.py:66,99-132)");
  if (in_channels < 1 || in_channels > 16) return fail(WSU_ERR_INVALID, "in_channels must be in 1..16");
  if (out_channels != 1) return fail(WSU_ERR_INVALID, "only out_channels=1 is implemented (the reference never uses another value)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(WSU_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libwsunet has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(WSU_ERR_INVALID, "device index out of range");
  DEVICE_SCOPE(device);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(WSU_ERR_CUDA, "libwsunet is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
  CUDA_TRY(conv_mma_init());
  wsu_context* h = new wsu_context();
  h->device = device;
  h->nsteps = nsteps;
  h->in_ch = in_channels;
  h->out_ch = out_channels;
  h->num_sms = prop.multiProcessorCount;
  // experiment switches (profiling runs): same meaning as the wsu_set_option keys
  if (const char* e = std::getenv("WSU_CTA_PAIR")) h->use_pair = std::atoi(e);
  if (const char* e = std::getenv("WSU_FUSE_E11")) h->fuse_e11 = std::atoi(e) != 0;
  if (const char* e = std::getenv("WSU_HALO")) h->use_halo = std::atoi(e) != 0;
  if (const char* e = std::getenv("WSU_A_COLLECTOR")) h->a_collector = std::atoi(e) != 0;
  if (const char* e = std::getenv("WSU_DBG")) h->dbg = std::atoi(e);
  if (const char* e = std::getenv("WSU_PRECISION")) h->precision = std::max(0, std::min(3, std::atoi(e)));
  *out = h;
  return WSU_OK;
}

int wsu_destroy(wsu_handle h) {
  if (!h) return WSU_OK;
  DeviceGuard guard(h->device);
  cudaDeviceSynchronize();
  free_plan(h->plan.get());
  for (auto& kv : h->layers) { cudaFree(kv.second.wpack); cudaFree(kv.second.bias); cudaFree(kv.second.wres); }
  cudaFree(h->e11_w);
  cudaFree(h->e11_b);
  for (int i = 0; i < 2; ++i) {
    cudaFree(h->stage_img[i]);
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_free[i]) cudaEventDestroy(h->ev_free[i]);
  }
  cudaFree(h->stage_out);
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->s_comp) cudaStreamDestroy(h->s_comp);
  if (h->ev_chain) cudaEventDestroy(h->ev_chain);
  delete h;
  return WSU_OK;
}

int wsu_set_option(wsu_handle h, const char* key, int64_t value) {
  if (!h || !key) return fail(WSU_ERR_INVALID, "null argument");
  if (!std::strcmp(key, "micro_batch")) {
    if (value < 0) return fail(WSU_ERR_INVALID, "micro_batch must be >= 0");
    h->micro_batch = value;
    return WSU_OK;
  }
  if (!std::strcmp(key, "fuse_e11")) {
    h->fuse_e11 = value != 0;
    return WSU_OK;
  }
  if (!std::strcmp(key, "a_collector")) {
    h->a_collector = value != 0;
    return WSU_OK;
  }
  if (!std::strcmp(key, "l2_prefetch")) {
    h->l2_prefetch = value != 0;
    return WSU_OK;
  }
  if (!std::strcmp(key, "cta_pair")) {
    h->use_pair = int(value);
    return WSU_OK;
  }
  if (!std::strcmp(key, "upconv_resident")) {
    h->use_upres = value != 0;
    return WSU_OK;
  }
  if (!std::strcmp(key, "halo")) {
    h->use_halo = value != 0;
    return WSU_OK;
  }
  if (!std::strcmp(key, "precision")) {
    if (value < 0 || value > 3)
      return fail(WSU_ERR_INVALID, "precision must be 0 (three-term), 1 / 2 (two- / one-term deep layers) or 3 (2 + fp16/fp8 full-resolution layers)");
    if (h->precision != int(value)) {
      h->precision = int(value);
      if (h->committed) return wsu_commit_weights(h);   // weights are packed per plan (bf16 or fp16 pairs); drops the shape plan too
    }
    return WSU_OK;
  }
  if (!std::strcmp(key, "dbg")) {   // the WSU_DBG switches as an option, for interleaved A/B timing (tools/option_ab.py)
    h->dbg = int(value);
    return WSU_OK;
  }
  if (!std::strcmp(key, "w_resident")) {
    h->w_resident = value != 0;
    return WSU_OK;
  }
  if (!std::strcmp(key, "tma_store")) {
    h->tma_store = value != 0;
    return WSU_OK;
  }
  if (!std::strcmp(key, "alias_buffers")) {
    if (h->alias_buffers != (value != 0)) {
      h->alias_buffers = value != 0;
      if (h->plan) { cudaDeviceSynchronize(); free_plan(h->plan.get()); h->plan.reset(); }
    }
    return WSU_OK;
  }
  if (!std::strcmp(key, "profile")) {
    h->profile = value != 0;
    return WSU_OK;
  }
  return fail(WSU_ERR_INVALID, std::string("unknown option ") + key);
}

int wsu_load_weights(wsu_handle h, const char* name, const float* data, const int64_t* dims, int ndims) {
  if (!h || !name || !data || !dims || ndims < 1 || ndims > 4) return fail(WSU_ERR_INVALID, "bad argument");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndims; ++i) {
    if (dims[i] <= 0) return fail(WSU_ERR_INVALID, "non-positive dim");
    t.dims.push_back(dims[i]);
    n *= size_t(dims[i]);
  }
  t.data.assign(data, data + n);
  h->host_w[name] = std::move(t);
  h->committed = false;
  return WSU_OK;
}

static int expect_dims(wsu_context* h, const std::string& name, std::vector<int64_t> want, const HostTensor** out) {
  auto it = h->host_w.find(name);
  if (it == h->host_w.end()) return fail(WSU_ERR_STATE, "missing state_dict entry '" + name + "'");
  if (it->second.dims != want) {
    std::string s = "size mismatch for " + name + ": got (";
    for (auto d : it->second.dims) s += std::to_string(d) + ",";
    s += ") expected (";
    for (auto d : want) s += std::to_string(d) + ",";
    return fail(WSU_ERR_INVALID, s + ")");
  }
  *out = &it->second;
  return WSU_OK;
}

// f16_cblocks0: leading 64-channel blocks packed as fp16 pairs although the layer is three-term (decoder source 0);
// bias_override: bias to upload instead of the state_dict's (up-convolution bias folded away / folded in)
static int upload_layer(wsu_context* h, const std::string& name, int cin, int cout, bool transposed, int terms,
                        int f16_cblocks0 = 0, const std::vector<float>* bias_override = nullptr, bool f8 = false,
                        int src0_terms = 2) {
  const HostTensor *w, *b;
  int rc;
  if (transposed) {
    if ((rc = expect_dims(h, name + ".weight", {cin, cout, 2, 2}, &w))) return rc;
  } else {
    if ((rc = expect_dims(h, name + ".weight", {cout, cin, 3, 3}, &w))) return rc;
  }
  if ((rc = expect_dims(h, name + ".bias", {cout}, &b))) return rc;
  LayerW lw;
  lw.cin = cin; lw.cout = cout;
  lw.terms = terms;
  lw.f16_cblocks0 = f16_cblocks0;
  lw.f8 = f8;
  lw.src0_terms = src0_terms;
  // fp16 + fp8 blocks: power-of-two scales. w * pw fills e4m3's range; (w - fp16(w)) * sw with sw = A2 * pw / A1 makes both
  // correction products carry the same total scale A2 * pw, which corr_scale removes in the epilogue.
  float pw = 1.f, sw = 1.f;
  if (f8) {
    float wmax = 0.f;
    for (int co = 0; co < cout; ++co)
      for (int ci = f16_cblocks0 * 64; ci < cin; ++ci)
        for (int t = 0; t < 9; ++t) wmax = std::max(wmax, std::fabs(w->data[(size_t(co) * cin + ci) * 9 + t]));
    if (wmax > 0.f) pw = std::exp2(std::floor(std::log2(256.f / wmax)));
    sw = kF8ScaleA2 * pw / kF8ScaleA1;
    lw.corr_scale = 1.f / (kF8ScaleA2 * pw);
  }
  // (hi, lo) pair of one weight in its channel block's operand type: bf16 for the three-term scheme, fp16 under a reduced plan
  auto split16 = [terms, f16_cblocks0](float v, int cblock, uint16_t& vh, uint16_t& vl) {
    if (terms == 3 && cblock >= f16_cblocks0) { vh = f2bf(v); vl = f2bf(v - bf2f(vh)); }
    else { vh = f2h(v); vl = f2h(v - h2f(vh)); }
  };
  lw.n_tile = cout == 64 ? 64 : 128;
  lw.ntaps = transposed ? 1 : 9;
  lw.npos = transposed ? 4 : 1;
  const int n_tiles = cout / lw.n_tile, cblocks = cin / 64, KB = cblocks * lw.ntaps;
  const size_t chunk = size_t(wchunk_bytes(lw.n_tile));
  std::vector<uint8_t> pack(size_t(lw.npos) * n_tiles * KB * chunk, 0);
  for (int pos = 0; pos < lw.npos; ++pos)
    for (int nt = 0; nt < n_tiles; ++nt)
      for (int c = 0; c < cblocks; ++c)
        for (int tap = 0; tap < lw.ntaps; ++tap) {
          uint8_t* hi = pack.data() + (size_t(pos * n_tiles + nt) * KB + size_t(c) * lw.ntaps + tap) * chunk;
          uint8_t* lo = hi + size_t(lw.n_tile) * 128;
          for (int n = 0; n < lw.n_tile; ++n)
            for (int k = 0; k < 64; ++k) {
              const int co = nt * lw.n_tile + n, ci = c * 64 + k;
              float v;
              if (transposed)
                v = w->data[((size_t(ci) * cout + co) * 2 + (pos >> 1)) * 2 + (pos & 1)];
              else
                v = w->data[((size_t(co) * cin + ci) * 3 + tap / 3) * 3 + tap % 3];
              if (f8 && c >= f16_cblocks0) {
                // main tile: fp16(w); correction tile: per 16-channel group the 32 bytes [e4m3(w pw) x16 | e4m3((w - fp16 w) sw) x16]
                const uint16_t vm = f2h(v);
                std::memcpy(hi + sw128_off(n, k), &vm, 2);
                const int g = k >> 4, i = k & 15;
                uint8_t* row = lo + size_t(n) * 128;
                auto byte_at = [&](int b) -> uint8_t& { return row[size_t((((b >> 4) ^ (n & 7)) << 4) + (b & 15))]; };   // SWIZZLE_128B on bytes
                byte_at(32 * g + i) = f2e4m3(v * pw);
                byte_at(32 * g + 16 + i) = f2e4m3((v - h2f(vm)) * sw);
                continue;
              }
              uint16_t vh, vl;
              split16(v, c, vh, vl);
              std::memcpy(hi + sw128_off(n, k), &vh, 2);
              std::memcpy(lo + sw128_off(n, k), &vl, 2);
            }
        }
  if (transposed) {
    // phase-stacked resident layout: largest co_t in {64, 32} whose (hi, lo) tiles fit kUpconvResBytes
    for (int cot : {64, 32}) {
      const int ntile = 4 * cot;
      if (cout % cot == 0 && cblocks * ntile * 256 <= kUpconvResBytes) { lw.res_cot = cot; lw.res_ntile = ntile; break; }
    }
    if (lw.res_cot) {
      const int rt = cout / lw.res_cot;
      const size_t tile = size_t(lw.res_ntile) * 128;
      std::vector<uint8_t> res(size_t(rt) * cblocks * 2 * tile, 0);
      for (int nt = 0; nt < rt; ++nt)
        for (int c = 0; c < cblocks; ++c) {
          uint8_t* hi = res.data() + (size_t(nt) * cblocks + c) * 2 * tile;
          uint8_t* lo = hi + tile;
          for (int r = 0; r < lw.res_ntile; ++r)
            for (int k = 0; k < 64; ++k) {
              const int pos = r / lw.res_cot, co = nt * lw.res_cot + r % lw.res_cot, ci = c * 64 + k;
              const float v = w->data[((size_t(ci) * cout + co) * 2 + (pos >> 1)) * 2 + (pos & 1)];
              uint16_t vh, vl;
              split16(v, c, vh, vl);
              std::memcpy(hi + sw128_off(r, k), &vh, 2);
              std::memcpy(lo + sw128_off(r, k), &vl, 2);
            }
        }
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&lw.wres), res.size()));
      CUDA_TRY(cudaMemcpy(lw.wres, res.data(), res.size(), cudaMemcpyHostToDevice));
    }
  }
  auto old = h->layers.find(name);
  if (old != h->layers.end()) { cudaFree(old->second.wpack); cudaFree(old->second.bias); cudaFree(old->second.wres); }
  lw.wpack_bytes = pack.size();
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&lw.wpack), pack.size()));
  CUDA_TRY(cudaMemcpy(lw.wpack, pack.data(), pack.size(), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&lw.bias), size_t(cout) * 4));
  CUDA_TRY(cudaMemcpy(lw.bias, bias_override ? bias_override->data() : b->data.data(), size_t(cout) * 4, cudaMemcpyHostToDevice));
  lw.bias_host.assign(bias_override ? bias_override->data() : b->data.data(), (bias_override ? bias_override->data() : b->data.data()) + cout);
  h->layers[name] = lw;
  return WSU_OK;
}

int wsu_commit_weights(wsu_handle h) {
  if (!h) return fail(WSU_ERR_INVALID, "null handle");
  DEVICE_SCOPE(h->device);
  CUDA_TRY(cudaDeviceSynchronize());
  int rc;
  const int n = h->nsteps;
  // e11: CUDA-core first layer keeps fp32 weights
  const HostTensor *w, *b;
  if ((rc = expect_dims(h, "e11.weight", {64, h->in_ch, 3, 3}, &w))) return rc;
  if ((rc = expect_dims(h, "e11.bias", {64}, &b))) return rc;
  cudaFree(h->e11_w);
  cudaFree(h->e11_b);
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->e11_w), w->data.size() * 4));
  CUDA_TRY(cudaMemcpy(h->e11_w, w->data.data(), w->data.size() * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->e11_b), 64 * 4));
  CUDA_TRY(cudaMemcpy(h->e11_b, b->data.data(), 64 * 4, cudaMemcpyHostToDevice));
  // A reduced-precision plan needs every up-convolution's weight set resident in shared memory (upconv_res_kernel is the
  // only fp16-input transposed convolution): 64 * N_TILE * cblocks * 4 B <= 128 KB, i.e. unet_1 and unet_2.
  h->precision_active = h->precision;
  for (int l = n - 1; l >= 0 && h->precision_active; --l) {
    const int cb = chan(l + 1) / 64;
    bool fits = false;
    for (int cot : {64, 32}) fits = fits || (chan(l) % cot == 0 && cb * 4 * cot * 256 <= kUpconvResBytes);
    if (!fits) h->precision_active = 0;
  }
  if (n == 0) h->precision_active = 0;
  if (h->precision_active == 3 && h->in_ch != 1) h->precision_active = 2;   // the ACT_F16F8 first layer exists for one input channel
  const bool f8 = h->precision_active == 3;
  const int deep_terms = h->precision_active >= 2 ? 1 : h->precision_active == 1 ? 2 : 3;
  auto terms_at = [&](int input_level) { return input_level >= 1 ? deep_terms : 3; };
  for (int l = 0; l <= n; ++l) {
    if (l > 0 && (rc = upload_layer(h, enc_name(l, 1), chan(l - 1), chan(l), false, terms_at(l)))) return rc;
    if ((rc = upload_layer(h, enc_name(l, 2), chan(l), chan(l), false, terms_at(l), 0, nullptr, f8 && l == 0))) return rc;
  }
  for (int l = n - 1; l >= 0; --l) {
    if (!h->precision_active) {
      if ((rc = upload_layer(h, up_name(l), chan(l + 1), chan(l), true, 3))) return rc;
      if ((rc = upload_layer(h, dec_name(l, 1), 2 * chan(l), chan(l), false, 3))) return rc;
    } else {
      // The up-convolution's output is stored as one fp16 plane. Its bias b_up is a per-channel constant, and the reflect
      // padding of a constant is that constant, so conv3x3(cat[u + b_up, skip]) = conv3x3(cat[u, skip]) + sum_{ci,tap}
      // W[co][ci][tap] * b_up[ci] exactly: the up-convolution stores u WITHOUT its bias (small values, small absolute
      // rounding) and the consuming layer's bias absorbs the constant (computed in double).
      const int cu = chan(l), co = chan(l);
      const HostTensor *wd, *bd, *bu;
      if ((rc = expect_dims(h, dec_name(l, 1) + ".weight", {co, 2 * cu, 3, 3}, &wd))) return rc;
      if ((rc = expect_dims(h, dec_name(l, 1) + ".bias", {co}, &bd))) return rc;
      if ((rc = expect_dims(h, up_name(l) + ".bias", {cu}, &bu))) return rc;
      std::vector<float> zero(size_t(cu), 0.f), folded(size_t(co), 0.f);
      for (int o = 0; o < co; ++o) {
        double acc = bd->data[o];
        for (int ci = 0; ci < cu; ++ci) {
          double ws = 0;
          for (int t = 0; t < 9; ++t) ws += wd->data[(size_t(o) * 2 * cu + ci) * 9 + t];
          acc += ws * double(bu->data[ci]);
        }
        folded[o] = float(acc);
      }
      if ((rc = upload_layer(h, up_name(l), chan(l + 1), chan(l), true, deep_terms, 0, &zero))) return rc;
      if (l >= 1) rc = upload_layer(h, dec_name(l, 1), 2 * chan(l), chan(l), false, deep_terms, 0, &folded);
      else rc = upload_layer(h, dec_name(l, 1), 2 * chan(l), chan(l), false, 3, cu / 64, &folded, f8, f8 ? 1 : 2);
      if (rc) return rc;
    }
    if ((rc = upload_layer(h, dec_name(l, 2), chan(l), chan(l), false, terms_at(l), 0, nullptr, f8 && l == 0))) return rc;
  }
  if ((rc = expect_dims(h, "outconv.weight", {1, 64, 1, 1}, &w))) return rc;
  if ((rc = expect_dims(h, "outconv.bias", {1}, &b))) return rc;
  std::memcpy(h->wout, w->data.data(), 64 * 4);
  h->bout = b->data[0];
  if (h->plan) { free_plan(h->plan.get()); h->plan.reset(); }  // plans embed wout/bout and weight pointers
  h->committed = true;
  return WSU_OK;
}

int wsu_unet_forward(wsu_handle h, const void* x_dev, int x_dtype, float* y_dev, int B, int H, int W, void* stream) {
  if (!x_dev || !y_dev) return fail(WSU_ERR_INVALID, "null tensor pointer");
  return unet_ws_device(h, x_dev, x_dtype, B, H, W, false, 0, 0, 0, 0, nullptr, nullptr, y_dev, static_cast<cudaStream_t>(stream));
}

int wsu_unet_ws_estimate(wsu_handle h, const void* img_dev, int img_dtype, int B, int H, int W, int weighted, int clip,
                         int crop, int correct_bias, float* beta_dev, float* l1_dev, float* yhat_dev, void* stream) {
  if (!img_dev || !beta_dev) return fail(WSU_ERR_INVALID, "null tensor pointer");
  return unet_ws_device(h, img_dev, img_dtype, B, H, W, true, weighted, clip, crop, correct_bias, beta_dev, l1_dev, yhat_dev,
                        static_cast<cudaStream_t>(stream));
}

int wsu_unet_ws_estimate_host(wsu_handle h, const uint8_t* img_host, int B, int H, int W, int weighted, int clip, int crop,
                              int correct_bias, float* beta_host, float* l1_host) {
  if (!h) return fail(WSU_ERR_INVALID, "null handle");
  if (!img_host || !beta_host) return fail(WSU_ERR_INVALID, "null host pointer");
  int rc = check_shape(h, B, H, W);
  if (rc) return rc;
  DEVICE_SCOPE(h->device);
  const int mb = pick_micro_batch(h, B, H, W);
  const size_t px = size_t(H) * W;
  if (!h->s_copy) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->s_comp, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
    }
  }
  if (h->stage_px < size_t(mb) * px) {
    for (int i = 0; i < 2; ++i) {
      cudaFree(h->stage_img[i]);
      h->stage_img[i] = nullptr;
      CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->stage_img[i]), size_t(mb) * px));
    }
    h->stage_px = size_t(mb) * px;
  }
  if (h->stage_imgs < size_t(B)) {
    cudaFree(h->stage_out);
    h->stage_out = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&h->stage_out), size_t(B) * 2 * sizeof(float)));
    h->stage_imgs = size_t(B);
  }
  float* beta_dev = h->stage_out;
  float* l1_dev = h->stage_out + B;
  int slot = 0;
  for (int b0 = 0; b0 < B; b0 += mb, slot ^= 1) {
    const int nimg = std::min(mb, B - b0);
    CUDA_TRY(cudaStreamWaitEvent(h->s_copy, h->ev_free[slot], 0));  // slot's previous consumer finished
    CUDA_TRY(cudaMemcpyAsync(h->stage_img[slot], img_host + size_t(b0) * px, size_t(nimg) * px, cudaMemcpyHostToDevice,
                             h->s_copy));
    CUDA_TRY(cudaEventRecord(h->ev_in[slot], h->s_copy));
    CUDA_TRY(cudaStreamWaitEvent(h->s_comp, h->ev_in[slot], 0));
    rc = unet_ws_device(h, h->stage_img[slot], WSU_U8, nimg, H, W, true, weighted, clip, crop, correct_bias, beta_dev + b0,
                        l1_dev + b0, nullptr, h->s_comp);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->ev_free[slot], h->s_comp));
  }
  CUDA_TRY(cudaMemcpyAsync(beta_host, beta_dev, size_t(B) * 4, cudaMemcpyDeviceToHost, h->s_comp));
  if (l1_host) CUDA_TRY(cudaMemcpyAsync(l1_host, l1_dev, size_t(B) * 4, cudaMemcpyDeviceToHost, h->s_comp));
  CUDA_TRY(cudaStreamSynchronize(h->s_comp));
  return WSU_OK;
}

int wsu_filter_predict(int device, const void* img_dev, int img_dtype, int kind, float* xhat_dev, int B, int H, int W,
                       void* stream) {
  int rc = filter_common_check(img_dev, img_dtype, kind, B, H, W);
  if (rc) return rc;
  if (!xhat_dev) return fail(WSU_ERR_INVALID, "null output pointer");
  DEVICE_SCOPE(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partials = nullptr;
  const int strips = filter_ws_strips(H);
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&partials), size_t(B) * strips * kPartialSlots * 4, st));
  LAUNCH_TRY(launch_filter_ws(img_dev, img_dtype == WSU_F32, B, H, W, kind, 0, 0, xhat_dev, partials, st));
  CUDA_TRY(cudaFreeAsync(partials, st));
  return WSU_OK;
}

int wsu_filter_ws_estimate(int device, const void* img_dev, int img_dtype, int kind, int weighted, int clip,
                           int correct_bias, float* beta_dev, float* l1_dev, int B, int H, int W, void* stream) {
  int rc = filter_common_check(img_dev, img_dtype, kind, B, H, W);
  if (rc) return rc;
  if (!beta_dev) return fail(WSU_ERR_INVALID, "null output pointer");
  if (weighted < -1 || weighted > 1) return fail(WSU_ERR_INVALID, "weighted must be -1, 0 or 1");
  DEVICE_SCOPE(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partials = nullptr;
  int records = filter_ws_strips(H);
  const bool fast = filter_ws_fast_ok(img_dev, img_dtype == WSU_F32, W, kind, correct_bias, nullptr);
  const bool packed = fast && weighted == WSU_UNWEIGHTED && l1_dev == nullptr;
  // WSU_EST_KERNEL=packed keeps the 16-bit-lane kernel for A/B runs; default is the adjoint (parity-plane) kernel
  static const bool prefer_packed = [] { const char* e = std::getenv("WSU_EST_KERNEL"); return e && !std::strcmp(e, "packed"); }();
  const bool adjoint = packed && !prefer_packed && filter_ws_adjoint_ok(img_dev, H, W);
  // weighted and/or L1-reporting requests: dp4a window kernel (WSU_EST_KERNEL=fast keeps the scalar sliding-window kernel)
  static const bool prefer_fast = [] { const char* e = std::getenv("WSU_EST_KERNEL"); return e && !std::strcmp(e, "fast"); }();
  const bool window = fast && !packed && !prefer_fast && !prefer_packed && filter_ws_window_ok(img_dev, H, W);
  if (fast) records = adjoint ? filter_ws_adjoint_records(H, W) : packed ? filter_ws_packed_records(H, W)
                    : window ? filter_ws_window_records(H, W) : filter_ws_fast_records(H, W);
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&partials), size_t(B) * records * kPartialSlots * 4, st));
  if (adjoint)
    LAUNCH_TRY(launch_filter_ws_adjoint(img_dev, B, H, W, kind, partials, st));
  else if (window)
    LAUNCH_TRY(launch_filter_ws_window(img_dev, B, H, W, kind, weighted, l1_dev != nullptr, partials, st));
  else if (packed)
    LAUNCH_TRY(launch_filter_ws_packed(img_dev, B, H, W, kind, partials, st));
  else if (fast)
    LAUNCH_TRY(launch_filter_ws_fast(img_dev, B, H, W, kind, weighted, partials, st));
  else
    LAUNCH_TRY(launch_filter_ws(img_dev, img_dtype == WSU_F32, B, H, W, kind, weighted, correct_bias, nullptr, partials, st));
  LAUNCH_TRY(launch_finalize(partials, records, B, float(H - 2) * float(W - 2), clip, correct_bias, beta_dev, l1_dev, st));
  CUDA_TRY(cudaFreeAsync(partials, st));
  return WSU_OK;
}

namespace {
// Host-buffer path of the filter estimators: per (calling thread, device) two streams, a ring of three device staging
// slots and a result buffer, all created once and reused by every call (no cudaMalloc / cudaFree on the call path:
// allocating and unmapping a 2.6 GB batch buffer per call had capped this path at 18 GB/s of host-link traffic).
struct HostPath {
  static constexpr int kSlots = 3;
  cudaStream_t s[2] = {nullptr, nullptr};
  cudaEvent_t slot_free[kSlots] = {nullptr, nullptr, nullptr};
  uint8_t* slot[kSlots] = {nullptr, nullptr, nullptr};
  size_t slot_bytes = 0;
  float* dout = nullptr;
  size_t dout_n = 0;
  ~HostPath() {   // thread exit: the context may already be gone, errors are ignored
    for (int i = 0; i < kSlots; ++i) { if (slot[i]) cudaFree(slot[i]); if (slot_free[i]) cudaEventDestroy(slot_free[i]); }
    if (dout) cudaFree(dout);
    for (int i = 0; i < 2; ++i) if (s[i]) cudaStreamDestroy(s[i]);
  }
};
thread_local std::map<int, HostPath> g_host_paths;   // keyed by device index

int host_path_prepare(HostPath& hp, size_t slot_bytes, size_t n_out) {
  for (int i = 0; i < 2; ++i)
    if (!hp.s[i]) CUDA_TRY(cudaStreamCreateWithFlags(&hp.s[i], cudaStreamNonBlocking));
  for (int i = 0; i < HostPath::kSlots; ++i)
    if (!hp.slot_free[i]) CUDA_TRY(cudaEventCreateWithFlags(&hp.slot_free[i], cudaEventDisableTiming));
  if (hp.slot_bytes < slot_bytes) {
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamSynchronize(hp.s[i]));
    for (int i = 0; i < HostPath::kSlots; ++i) {
      if (hp.slot[i]) { cudaFree(hp.slot[i]); hp.slot[i] = nullptr; }
    }
    hp.slot_bytes = 0;
    for (int i = 0; i < HostPath::kSlots; ++i) CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&hp.slot[i]), slot_bytes));
    hp.slot_bytes = slot_bytes;
  }
  if (hp.dout_n < n_out) {
    for (int i = 0; i < 2; ++i) CUDA_TRY(cudaStreamSynchronize(hp.s[i]));
    if (hp.dout) { cudaFree(hp.dout); hp.dout = nullptr; }
    hp.dout_n = 0;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&hp.dout), n_out * sizeof(float)));
    hp.dout_n = n_out;
  }
  return WSU_OK;
}
}  // namespace

int wsu_filter_ws_estimate_host(int device, const uint8_t* img_host, int kind, int weighted, int clip, int correct_bias,
                                float* beta_host, float* l1_host, int B, int H, int W) {
  if (!img_host || !beta_host) return fail(WSU_ERR_INVALID, "null host pointer");
  int rc = filter_common_check(img_host, WSU_U8, kind, B, H, W);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(WSU_ERR_INVALID, "device index out of range");
  DEVICE_SCOPE(device);
  const size_t px = size_t(H) * W;
  // chunks of <= 64 MB alternate between two streams and rotate through three staging slots: the H2D copy of chunk
  // i+1 overlaps the kernels of chunk i; results stay on the device until one D2H copy at the end (a per-chunk copy
  // into pageable host memory would block the host and serialise the pipeline)
  const int chunk = std::max(1, std::min(B, int((size_t(64) << 20) / px)));
  HostPath& hp = g_host_paths[device];
  if ((rc = host_path_prepare(hp, size_t(chunk) * px, size_t(B) * 2))) return rc;
  float* dout = hp.dout;
  int i = 0;
  for (int b0 = 0; b0 < B; b0 += chunk, ++i) {
    const int n = std::min(chunk, B - b0), k = i & 1, sl = i % HostPath::kSlots;
    cudaError_t e = cudaStreamWaitEvent(hp.s[k], hp.slot_free[sl], 0);   // the slot's previous kernel (other stream) is done
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(hp.slot[sl], img_host + size_t(b0) * px, size_t(n) * px, cudaMemcpyHostToDevice, hp.s[k]);
    if (e != cudaSuccess) { rc = fail(WSU_ERR_CUDA, std::string("host staging: ") + cudaGetErrorString(e)); break; }
    rc = wsu_filter_ws_estimate(device, hp.slot[sl], WSU_U8, kind, weighted, clip, correct_bias, dout + b0,
                                l1_host ? dout + B + b0 : nullptr, n, H, W, hp.s[k]);
    if (rc) break;
    cudaEventRecord(hp.slot_free[sl], hp.s[k]);
  }
  cudaError_t e0 = cudaStreamSynchronize(hp.s[0]), e1 = cudaStreamSynchronize(hp.s[1]);
  if (rc) return rc;
  CUDA_TRY(e0);
  CUDA_TRY(e1);
  CUDA_TRY(cudaMemcpy(beta_host, dout, size_t(B) * 4, cudaMemcpyDeviceToHost));
  if (l1_host) CUDA_TRY(cudaMemcpy(l1_host, dout + B, size_t(B) * 4, cudaMemcpyDeviceToHost));
  return WSU_OK;
}

int wsu_ws_grad_prediction(int device, const void* img_dev, int img_dtype, const float* coef_dev, int crop, float scale,
                           float* grad_dev, int B, int H, int W, void* stream) {
  if (!img_dev || !coef_dev || !grad_dev) return fail(WSU_ERR_INVALID, "null tensor pointer");
  if (img_dtype != WSU_U8 && img_dtype != WSU_F32) return fail(WSU_ERR_INVALID, "dtype must be WSU_U8 or WSU_F32");
  if (crop < 0 || crop > 1) return fail(WSU_ERR_INVALID, "crop must be 0 or 1");
  if (B <= 0 || B > 65535 || H < 1 + 2 * crop || W < 1 + 2 * crop) return fail(WSU_ERR_INVALID, "need 1 <= B <= 65535 and a non-empty crop");
  DEVICE_SCOPE(device);
  LAUNCH_TRY(launch_ws_grad_pred(img_dev, img_dtype == WSU_F32, coef_dev, grad_dev, B, H, W, crop, scale,
                                 static_cast<cudaStream_t>(stream)));
  return WSU_OK;
}

int wsu_ws_from_prediction(int device, const void* img_dev, int img_dtype, const float* xhat_dev, int xhat_cropped,
                           const float* xbias_dev, int weighted, int clip, int crop, float* beta_dev, float* l1_dev, int B,
                           int H, int W, void* stream) {
  if (!img_dev || !xhat_dev || !beta_dev) return fail(WSU_ERR_INVALID, "null tensor pointer");
  if (img_dtype != WSU_U8 && img_dtype != WSU_F32) return fail(WSU_ERR_INVALID, "dtype must be WSU_U8 or WSU_F32");
  if (B <= 0 || H < 3 || W < 3) return fail(WSU_ERR_INVALID, "need B >= 1 and H, W >= 3");
  if (weighted < -1 || weighted > 1) return fail(WSU_ERR_INVALID, "weighted must be -1, 0 or 1");
  if (weighted != 0 && !crop) return fail(WSU_ERR_INVALID, "local-variance weights exist only on the interior (crop=1)");
  if (xhat_cropped && !crop) return fail(WSU_ERR_INVALID, "cropped predictions need crop=1");
  DEVICE_SCOPE(device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunks = 32;
  float* partials = nullptr;
  CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&partials), size_t(B) * chunks * kPartialSlots * 4, st));
  LAUNCH_TRY(launch_ws_from_pred(img_dev, img_dtype == WSU_F32, xhat_dev, xhat_cropped, xbias_dev, B, H, W, weighted, crop,
                                 partials, chunks, st));
  const float npix = crop ? float(H - 2) * float(W - 2) : float(H) * float(W);
  LAUNCH_TRY(launch_finalize(partials, chunks, B, npix, clip, xbias_dev != nullptr, beta_dev, l1_dev, st));
  CUDA_TRY(cudaFreeAsync(partials, st));
  return WSU_OK;
}

int wsu_uniform_dropout(int device, const void* x_dev, int x_dtype, const float* mask_dev, float* out_dev, int B, int C,
                        int H, int W, uint32_t channel_mask, void* stream) {
  if (!x_dev || !mask_dev || !out_dev) return fail(WSU_ERR_INVALID, "null tensor pointer");
  if (x_dtype != WSU_U8 && x_dtype != WSU_F32) return fail(WSU_ERR_INVALID, "dtype must be WSU_U8 or WSU_F32");
  if (B <= 0 || C <= 0 || C > 32 || H < 2 || W < 2) return fail(WSU_ERR_INVALID, "need B >= 1, 1 <= C <= 32 and H, W >= 2 (reflect padding)");
  DEVICE_SCOPE(device);
  LAUNCH_TRY(launch_kb_blend(x_dev, x_dtype == WSU_F32, mask_dev, out_dev, B, C, H, W, channel_mask, static_cast<cudaStream_t>(stream)));
  return WSU_OK;
}

int wsu_filter_residual_rows(int device, const void* mat_dev, int mat_dtype, const double* coef_dev, double* resid_dev,
                             int64_t n_rows, void* stream) {
  if (!mat_dev || !coef_dev || !resid_dev) return fail(WSU_ERR_INVALID, "null tensor pointer");
  if (mat_dtype < 0 || mat_dtype > 2) return fail(WSU_ERR_INVALID, "mat_dtype must be 0 (uint8), 1 (float32) or 2 (float64)");
  if (n_rows <= 0) return fail(WSU_ERR_INVALID, "n_rows must be positive");
  DEVICE_SCOPE(device);
  LAUNCH_TRY(launch_residual_matvec(mat_dev, mat_dtype, coef_dev, resid_dev, n_rows, static_cast<cudaStream_t>(stream)));
  return WSU_OK;
}

int wsu_get_info(wsu_handle h, const char* key, int64_t* out) {
  if (!h || !key || !out) return fail(WSU_ERR_INVALID, "null argument");
  if (!std::strcmp(key, "micro_batch")) *out = h->plan ? h->plan->mb : 0;
  else if (!std::strcmp(key, "last_images")) *out = h->last_nimg;
  else if (!std::strcmp(key, "num_sms")) *out = h->num_sms;
  else if (!std::strcmp(key, "precision")) *out = h->precision_active;
  else if (!std::strcmp(key, "bytes_per_image")) *out = h->plan ? int64_t(h->plan->arena_bytes / size_t(h->plan->mb)) : 0;
  else if (!std::strcmp(key, "layers")) *out = h->plan ? int64_t(h->plan->convs.size()) + 1 : 0;
  else return fail(WSU_ERR_INVALID, std::string("unknown info key ") + key);
  return WSU_OK;
}

int wsu_profile_read(wsu_handle h, float* ms_out, int cap) {
  if (!h || !ms_out) return fail(WSU_ERR_INVALID, "null argument");
  if (!h->profile || h->prof_n == 0 || int(h->prof_ev.size()) < h->prof_n + 1) return fail(WSU_ERR_STATE, "no profile recorded");
  DEVICE_SCOPE(h->device);
  CUDA_TRY(cudaEventSynchronize(h->prof_ev[h->prof_n]));
  const int n = std::min(cap, h->prof_n);
  for (int i = 0; i < n; ++i) CUDA_TRY(cudaEventElapsedTime(&ms_out[i], h->prof_ev[i], h->prof_ev[i + 1]));
  return h->prof_n;
}

const char* wsu_profile_name(wsu_handle h, int i) {
  if (!h || i < 0 || i >= h->prof_n) return "";
  return h->prof_names[i].c_str();
}

int wsu_debug_layer(wsu_handle h, const char* name, float* dst_dev, size_t cap, int with_halo, int64_t* dims_out,
                    void* stream) {
  if (!h || !name || !dst_dev) return fail(WSU_ERR_INVALID, "null argument");
  if (!h->plan) return fail(WSU_ERR_STATE, "no forward pass has run yet");
  auto it = h->plan->acts.find(name);
  if (it == h->plan->acts.end()) return fail(WSU_ERR_INVALID, std::string("unknown layer ") + name);
  const Act& a = it->second;
  const int Ho = a.H + (with_halo ? 2 : 0), Wo = a.W + (with_halo ? 2 : 0);
  if (size_t(a.B) * a.C * Ho * Wo > cap) return fail(WSU_ERR_INVALID, "destination too small");
  if (dims_out) { dims_out[0] = a.B; dims_out[1] = a.C; dims_out[2] = Ho; dims_out[3] = Wo; }
  DEVICE_SCOPE(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->chain_pending && h->chain_stream != st) CUDA_TRY(cudaStreamWaitEvent(st, h->ev_chain, 0));
  LAUNCH_TRY(launch_unpack(a, dst_dev, with_halo, st));
  if (cudaEventRecord(h->ev_chain, st) == cudaSuccess) { h->chain_stream = st; h->chain_pending = true; }  // the next pass must not overwrite what this copy reads
  return WSU_OK;
}

}  // extern "C"
