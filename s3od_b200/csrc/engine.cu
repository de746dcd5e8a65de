// Context, launch plan and C ABI of the B200-native S3OD forward path (see include/s3od_b200.h).
//
// Data layout in HBM (nb = images of the current micro-batch, g = S/16, P = g*g patches, ntok = P + 5):
//   patches  bf16 [B*P, 768]            im2col of the letterboxed, normalised input (whole batch)
//   x        fp32 [nb*ntok, D]          residual stream          xn / ctx bf16 [nb*ntok, D]   LN output / attention output
//   q, k, v  bf16 [nb*H, ntok, 64]      (head-major; V is read as an MN-major MMA operand)     hmid bf16 [nb*ntok, I]
//   taps     bf16 [nb*P, D] x 4         == NHWC (nb, g, g, D): token-major IS channels-last, no permute (model.py:206)
//   head     bf16 NHWC everywhere; mask logits fp32 planar (B, K, S, S) straight into the caller's buffer.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/s3od_b200.h"
#include "launch.h"

using namespace s3od;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(expr)                                                                                        \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return fail(S3OD_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + std::to_string(__LINE__)); \
  } while (0)

// ---------------------------------------------------------------------------------------------- TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor map with 128B swizzle; dims[0] is the contiguous dimension, strides in BYTES for dims 1..rank-1.
bool make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    g_err = "cuTensorMapEncodeTiled entry point not available";
    return false;
  }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    g_err = "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

// row-major [rows, cols] bf16 matrix, box = 64 columns x box_rows rows
bool tmap_matrix(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {cols * 2};
  const uint32_t box[2] = {64, box_rows};
  return make_tmap(m, base, 2, dims, strides, box);
}

// NHWC (B, H, W, C) activation as the 5-D (C, W, 1, H, B) view, box = 64 ch x 16 x 1 x 8 x 1
bool tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, 1, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[5] = {64, kTileW, 1, kTileH, 1};
  return make_tmap(m, base, 5, dims, strides, box);
}

// NHWC (B, H, W, 64) activation for the row-streaming convolution: box = 64 ch x 130 pixels of one row
bool tmap_nhwc_row(CUtensorMap* m, const void* base, int B, int H, int W, int C) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, 1, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[5] = {64, kRowPx + 2, 1, 1, 1};
  return make_tmap(m, base, 5, dims, strides, box);
}
// ... and its bf16 NHWC output: box = 64 ch x 32 pixels (one epilogue warp)
bool tmap_nhwc_row_out(CUtensorMap* m, const void* base, int B, int H, int W, int C) {
  const uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, 1, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[5] = {64, 32, 1, 1, 1};
  return make_tmap(m, base, 5, dims, strides, box);
}

// transposed-convolution rows: input (B, H, W, 128) fetched as 64-channel halves, box = 64 ch x 130 pixels of one row
bool tmap_nhwc_row128(CUtensorMap* m, const void* base, int B, int H, int W) {
  const uint64_t dims[5] = {128, (uint64_t)W, 1, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[4] = {256, (uint64_t)W * 256, (uint64_t)W * 256, (uint64_t)H * W * 256};
  const uint32_t box[5] = {64, kRowPx + 2, 1, 1, 1};
  return make_tmap(m, base, 5, dims, strides, box);
}
// ... and its output (B, 2H, 2W, 64) seen as (64 ch, column phase, W, 2H, B): box = 64 ch x 32 same-phase pixels
bool tmap_convt_out(CUtensorMap* m, const void* base, int B, int H, int W) {
  const uint64_t dims[5] = {64, 2, (uint64_t)W, (uint64_t)2 * H, (uint64_t)B};
  const uint64_t strides[4] = {128, 256, (uint64_t)2 * W * 128, (uint64_t)2 * H * 2 * W * 128};
  const uint32_t box[5] = {64, 1, 32, 1, 1};
  return make_tmap(m, base, 5, dims, strides, box);
}

bool make_convt_rows(ConvTRowParams* p, const void* in, const void* wr, const float* bias, void* out, int B, int H, int W, int relu) {
  if (!tmap_nhwc_row128(&p->tma_in, in, B, H, W)) return false;
  if (!tmap_matrix(&p->tma_w, wr, 4 * 256, 128, 256)) return false;
  if (!tmap_convt_out(&p->tma_out, out, B, H, W)) return false;
  p->H = H; p->W = W;
  p->strips_x = W / kRowPx;
  p->strips_y = std::max(1, H / kCtPairsPerStrip);
  p->bias = bias;
  p->relu = relu;
  return true;
}

// the same tensor viewed as (2C, W/2, 2, H/2, B): element (c + px*C, w2, py, h2, b) = in[b, 2*h2+py, 2*w2+px, c]
bool tmap_nhwc_s2(CUtensorMap* m, const void* base, int B, int H, int W, int C) {
  const uint64_t dims[5] = {(uint64_t)2 * C, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)B};
  const uint64_t strides[4] = {(uint64_t)2 * C * 2, (uint64_t)W * C * 2, (uint64_t)2 * W * C * 2, (uint64_t)H * W * C * 2};
  const uint32_t box[5] = {64, kTileW, 1, kTileH, 1};
  return make_tmap(m, base, 5, dims, strides, box);
}

ConvGeom geom_3x3(int H, int W, int cin) {
  ConvGeom g{};
  g.H = H; g.W = W;
  g.tiles_h = (H + kTileH - 1) / kTileH;
  g.tiles_w = (W + kTileW - 1) / kTileW;
  g.cin_blocks = cin / 64;
  g.ntaps = 9;
  for (int t = 0; t < 9; ++t) {
    g.dc[t] = 0; g.dp[t] = 0;
    g.dh[t] = t / 3 - 1;
    g.dw[t] = t % 3 - 1;
  }
  return g;
}

// 3x3 stride 2 pad 1 on the (2C, W/2, 2, H/2, B) view: input row 2*oh + ky - 1 -> (h2, py) = (oh-1,1), (oh,0), (oh,1)
ConvGeom geom_3x3_s2(int OH, int OW, int cin) {
  ConvGeom g{};
  g.H = OH; g.W = OW;
  g.tiles_h = (OH + kTileH - 1) / kTileH;
  g.tiles_w = (OW + kTileW - 1) / kTileW;
  g.cin_blocks = cin / 64;
  g.ntaps = 9;
  const int d2[3] = {-1, 0, 0}, par[3] = {1, 0, 1};
  for (int t = 0; t < 9; ++t) {
    const int ky = t / 3, kx = t % 3;
    g.dh[t] = d2[ky]; g.dp[t] = par[ky];
    g.dw[t] = d2[kx]; g.dc[t] = par[kx] * cin;
  }
  return g;
}

// ConvTranspose2d k4 s2 p1 as four 2x2 sub-pixel convolutions: output (2i+a, 2j+b); for a = 0 the two row taps are
// input rows (i, i-1) with kernel rows (1, 3); for a = 1 rows (i+1, i) with kernel rows (0, 2); same for columns.
// weights.py packs tap (r, c) = r*2 + c in the same order.
ConvGeom geom_convt_phase(int H, int W, int cin, int a, int b) {
  ConvGeom g{};
  g.H = H; g.W = W;
  g.tiles_h = (H + kTileH - 1) / kTileH;
  g.tiles_w = (W + kTileW - 1) / kTileW;
  g.cin_blocks = cin / 64;
  g.ntaps = 4;
  const int off[2][2] = {{0, -1}, {1, 0}};
  for (int t = 0; t < 4; ++t) {
    g.dc[t] = 0; g.dp[t] = 0;
    g.dh[t] = off[a][t / 2];
    g.dw[t] = off[b][t % 2];
  }
  return g;
}

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

}  // namespace

// ================================================================================================ context
struct s3od_ctx {
  int device = 0, arch = 0, K = 3, S = 1024, max_batch = 1, mb = 1;
  int g = 64, P = 4096, ntok = 4101, D = 768, H = 12, I = 3072, L = 11, vt_pitch = 4104;
  int taps[4] = {2, 5, 8, 11};
  int oc[4] = {256, 512, 1024, 1024};
  int num_sms = 148;
  int pool_blocks = 1;
  bool finalized = false;
  long long launches = 0;
  std::unordered_map<std::string, DevBuf> w;     // packed weights by name
  std::unordered_map<std::string, DevBuf> act;   // activations by name
  std::vector<void*> allocs;
  bool pre_affine = false;                       // "pre.affine" reproduces every entry of "pre.lut" (checked in s3od_finalize)
  float pre_ab[6] = {0, 0, 0, 0, 0, 0};
  ImageDesc* d_img = nullptr;
  PostDesc* d_post = nullptr;
  int last_nb = 0;
  // optional per-op timing with CUDA events on the launch stream (bench.py roofline numbers)
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<std::pair<int, int>> ev_ops;        // (op index, nb) for every recorded event pair
  // launch plan: every op takes (nb images, first image index b0, mask-logit out, iou-logit out, stream)
  using Op = std::function<cudaError_t(int, int, float*, float*, cudaStream_t)>;
  std::vector<std::pair<std::string, Op>> plan;
};

namespace {

}  // namespace
namespace s3od {
int train_fail(int code, const std::string& msg) { return fail(code, msg); }     // train.cu shares the thread-local error string
}
namespace {

template <class T>
T* wptr(s3od_ctx* c, const std::string& name) {
  auto it = c->w.find(name);
  return it == c->w.end() ? nullptr : reinterpret_cast<T*>(it->second.p);
}

bool alloc_act(s3od_ctx* c, const std::string& name, size_t bytes) {
  void* p = nullptr;
  bytes = (bytes + 255) & ~size_t(255);
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    g_err = "cudaMalloc failed for activation " + name + " (" + std::to_string(bytes) + " bytes)";
    return false;
  }
  cudaMemset(p, 0, bytes);
  c->allocs.push_back(p);
  c->act[name] = DevBuf{p, bytes};
  return true;
}

template <class T>
T* aptr(s3od_ctx* c, const std::string& name) {
  return reinterpret_cast<T*>(c->act.at(name).p);
}

using bf16 = __nv_bfloat16;

// ---- plan builders -----------------------------------------------------------------------------------------------
// kSmallN > 0: a second, one-CTA configuration with 128 x kSmallN tiles for the same GEMM, chosen at launch time when it fills the
// SMs better than the 256 x 256 CTA-pair tiles.  With M = 4101 rows (one image) and N = 768 the pair kernel has 17 x 3 = 51 work items
// for 74 SM pairs; 128 x 192 tiles give 33 x 4 = 132 tiles for 148 SMs in ONE round of 3/4-size tiles.  Per-element arithmetic does
// not depend on the tile shape (same k-block order), so results stay bit-identical across batch sizes.
template <int BN, class Epi, int EW, int kSmallN = 0>
bool add_linear(s3od_ctx* c, const std::string& label, const void* a, uint64_t a_rows_total, int rows_per_image, int Kdim,
                const void* bw, int N, typename Epi::Params ep, bool a_is_batch_window = false, bool per_image_tiles = false,
                std::function<void(typename Epi::Params&, int, int, float*, float*)> patch = nullptr) {
  if (N % BN != 0 || Kdim % 64 != 0) {
    g_err = "bad GEMM shape for " + label;
    return false;
  }
  GemmParams<Epi> p{};
  if (!tmap_matrix(&p.tma_a, a, a_rows_total, Kdim, kBM)) return false;
  if (!tmap_matrix(&p.tma_b, bw, N, Kdim, b_box_rows<BN>())) return false;
  p.n_tiles = N / BN;
  p.num_k_blocks = Kdim / 64;
  p.b_row_offset = 0;
  p.a_row_offset = 0;
  p.epi = ep;
  GemmParams<Epi> ps = p;                           // the small-tile configuration (only used when kSmallN > 0)
  bool have_small = false;
  if constexpr (kSmallN > 0) {
    if (N % kSmallN == 0 && !per_image_tiles) {
      if (!tmap_matrix(&ps.tma_b, bw, N, Kdim, kSmallN)) return false;
      ps.n_tiles = N / kSmallN;
      have_small = true;
    }
  }
  const int sms = c->num_sms;
  c->plan.emplace_back(label, [=](int nb, int b0, float* mo, float* io, cudaStream_t st) mutable -> cudaError_t {
    GemmParams<Epi> q = p;
    q.M = nb * rows_per_image;
    q.m_tiles = (q.M + kBM - 1) / kBM;
    if (per_image_tiles) {                      // token GEMMs: tiles never straddle two images
      q.rows_per_image = rows_per_image;
      q.tiles_per_image = (rows_per_image + kBM - 1) / kBM;
      q.m_tiles = nb * q.tiles_per_image;
    }
    q.a_row_offset = a_is_batch_window ? b0 * rows_per_image : 0;
    if (patch) patch(q.epi, nb, b0, mo, io);
    if constexpr (kSmallN > 0) {
      if (have_small) {
        // rounds x relative tile size of the two configurations
        const int items_pair = ((q.m_tiles + 1) / 2) * q.n_tiles, pairs = sms / 2;
        const int tiles_small = q.m_tiles * ps.n_tiles;
        const double t_pair = static_cast<double>((items_pair + pairs - 1) / pairs);
        // a pair item keeps each of its two SMs busy with 128 x BN of output; a small tile is 128 x kSmallN on one SM
        const double t_small = static_cast<double>((tiles_small + sms - 1) / sms) * (static_cast<double>(kSmallN) / BN);
        if (t_small < 0.9 * t_pair) {
          GemmParams<Epi> r = ps;
          r.M = q.M; r.m_tiles = q.m_tiles; r.a_row_offset = q.a_row_offset; r.epi = q.epi;
          return launch_gemm<kSmallN, A_LINEAR, Epi, EW>(r, sms, st);
        }
      }
    }
    return launch_gemm<BN, A_LINEAR, Epi, EW>(q, sms, st);
  });
  return true;
}

template <int BN, class Epi, int EW>
bool add_conv(s3od_ctx* c, const std::string& label, const CUtensorMap& tma_a, const ConvGeom& geom, const void* bw,
              int b_rows_total, int b_row_offset, int Kdim, int N, typename Epi::Params ep,
              std::function<void(typename Epi::Params&, int, int, float*, float*)> patch = nullptr) {
  if (N % BN != 0 || Kdim % 64 != 0 || Kdim != geom.ntaps * geom.cin_blocks * 64) {
    g_err = "bad conv shape for " + label;
    return false;
  }
  GemmParams<Epi> p{};
  p.tma_a = tma_a;
  if (!tmap_matrix(&p.tma_b, bw, b_rows_total, Kdim, b_box_rows<BN>())) return false;
  p.n_tiles = N / BN;
  p.num_k_blocks = Kdim / 64;
  p.b_row_offset = b_row_offset;
  p.geom = geom;
  p.epi = ep;
  const int sms = c->num_sms;
  c->plan.emplace_back(label, [=](int nb, int b0, float* mo, float* io, cudaStream_t st) mutable -> cudaError_t {
    GemmParams<Epi> q = p;
    q.M = 0;
    q.m_tiles = nb * geom.tiles_h * geom.tiles_w;
    if (patch) patch(q.epi, nb, b0, mo, io);
    return launch_gemm<BN, A_CONV, Epi, EW>(q, sms, st);
  });
  return true;
}

// 3x3 / stride 1 / pad 1 convolution with Cin = 64 and a small Cout on the row-streaming kernel (conv_rows.cuh);
// needs W % 128 == 0, otherwise the caller falls back to the generic implicit GEMM.
template <int NOUT, class Epi>
bool add_conv_rows(s3od_ctx* c, const std::string& label, const bf16* in, int Hs, int Ws, const void* bw, typename Epi::Params ep,
                   std::function<void(typename Epi::Params&, int, int, float*, float*)> patch = nullptr) {
  RowConvParams<Epi> p{};
  if (Ws % kRowPx != 0) {
    g_err = "row convolution needs W % 128 == 0 for " + label;
    return false;
  }
  if (!tmap_nhwc_row(&p.tma_in, in, c->mb, Hs, Ws, 64)) return false;
  if (!tmap_matrix(&p.tma_w, bw, NOUT, 9 * 64, NOUT)) return false;
  p.H = Hs; p.W = Ws;
  p.strips_x = Ws / kRowPx;
  p.strips_y = (Hs + kRowsPerStrip - 1) / kRowsPerStrip;
  p.epi = ep;
  if constexpr (std::is_same_v<Epi, EpiConv>) {
    if (ep.res1 != nullptr || ep.res2 != nullptr || ep.out_relu != nullptr || ep.up != 1) {
      g_err = "row convolution supports bias + relu only: " + label;
      return false;
    }
    if (!tmap_nhwc_row_out(&p.tma_out, ep.out, c->mb, Hs, Ws, 64)) return false;
  }
  const int sms = c->num_sms;
  c->plan.emplace_back(label, [=](int nb, int b0, float* mo, float* io, cudaStream_t st) mutable -> cudaError_t {
    RowConvParams<Epi> q = p;
    q.num_strips = nb * q.strips_x * q.strips_y;
    if (patch) patch(q.epi, nb, b0, mo, io);
    return launch_conv_rows<NOUT, Epi>(q, sms, st);
  });
  return true;
}

EpiConv::Params conv_epi(bf16* out, bf16* out_relu, const float* bias, const bf16* res1, const bf16* res2, int relu, int cout,
                         int oh, int ow) {
  EpiConv::Params e{};
  e.out = out; e.out_relu = out_relu; e.bias = bias; e.res1 = res1; e.res2 = res2; e.relu = relu;
  e.linear = 0; e.hin = 0; e.win = 0; e.up = 1; e.ph_h = 0; e.ph_w = 0; e.cout = cout; e.oh = oh; e.ow = ow;
  return e;
}

// 3x3 / stride 1 / pad 1 convolution on an NHWC activation of c->mb images
template <int BN, int EW>
bool add_conv3x3(s3od_ctx* c, const std::string& label, const bf16* in, int Hs, int Ws, int cin, const std::string& wname,
                 int cout, EpiConv::Params ep) {
  CUtensorMap ta;
  if (!tmap_nhwc(&ta, in, c->mb, Hs, Ws, cin)) return false;
  const bf16* wt = wptr<bf16>(c, wname);
  return add_conv<BN, EpiConv, EW>(c, label, ta, geom_3x3(Hs, Ws, cin), wt, cout, 0, 9 * cin, cout, ep);
}

#ifndef S3OD_POOL_MIN
#define S3OD_POOL_MIN 128
#endif
#ifndef S3OD_FLAT_TILES
#define S3OD_FLAT_TILES 1
#endif
constexpr bool kFlatTiles = S3OD_FLAT_TILES != 0;
#ifndef S3OD_SWAP128
#define S3OD_SWAP128 1
#endif
#ifndef S3OD_ROWCONV
#define S3OD_ROWCONV 1
#endif
constexpr bool kSwap128 = S3OD_SWAP128 != 0;      // compile-time A/B switches (tools/build_variants.sh); never read from the environment
constexpr bool kRowConv = S3OD_ROWCONV != 0;

bool build_plan(s3od_ctx* c) {
  const int mb = c->mb, g = c->g, P = c->P, ntok = c->ntok, D = c->D, H = c->H, I = c->I, K = c->K, S = c->S;
  const size_t MT = static_cast<size_t>(mb) * ntok;
  // flat tiling: the epilogue sees ONE "image" of nb * ntok rows (RowInfo.b = 0, RowInfo.t = global row)
  std::function<void(EpiResidual::Params&, int, int, float*, float*)> flat_residual;
  std::function<void(EpiGelu::Params&, int, int, float*, float*)> flat_gelu;
  if (kFlatTiles) {
    flat_residual = [ntok](EpiResidual::Params& e, int nb, int, float*, float*) { e.ntok = nb * ntok; };
    flat_gelu = [ntok](EpiGelu::Params& e, int nb, int, float*, float*) { e.ntok = nb * ntok; };
  }
  // ---- activations
  bool ok = true;
  ok = ok && alloc_act(c, "patches", static_cast<size_t>(c->max_batch) * P * 768 * 2);
  ok = ok && alloc_act(c, "x", MT * D * 4);
  ok = ok && alloc_act(c, "dx", MT * D * 2);
  ok = ok && alloc_act(c, "xn", MT * D * 2);
  ok = ok && alloc_act(c, "ctx", MT * D * 2);
  ok = ok && alloc_act(c, "q", MT * D * 2);
  ok = ok && alloc_act(c, "k", MT * D * 2);
  ok = ok && alloc_act(c, "v", MT * D * 2);
  ok = ok && alloc_act(c, "hmid", MT * I * 2);
  for (int j = 0; j < 4; ++j) ok = ok && alloc_act(c, "tap" + std::to_string(j), static_cast<size_t>(mb) * P * D * 2);
  const int R[5] = {0, 4 * g, 2 * g, g, g / 2};     // resolution of layer_k_rn, k = 1..4
  for (int j = 0; j < 4; ++j) ok = ok && alloc_act(c, "f" + std::to_string(j), static_cast<size_t>(mb) * P * c->oc[j] * 2);
  ok = ok && alloc_act(c, "r0", static_cast<size_t>(mb) * R[1] * R[1] * c->oc[0] * 2);
  ok = ok && alloc_act(c, "r1", static_cast<size_t>(mb) * R[2] * R[2] * c->oc[1] * 2);
  ok = ok && alloc_act(c, "r3", static_cast<size_t>(mb) * R[4] * R[4] * c->oc[3] * 2);
  for (int k = 1; k <= 4; ++k) {
    const size_t n = static_cast<size_t>(mb) * R[k] * R[k] * 256 * 2;
    ok = ok && alloc_act(c, "l" + std::to_string(k), n);
    ok = ok && alloc_act(c, "l" + std::to_string(k) + "r", n);
  }
  const size_t tmax = static_cast<size_t>(mb) * R[1] * R[1] * 256 * 2;
  for (const char* nm : {"tA", "tB", "tC", "tD"}) ok = ok && alloc_act(c, nm, tmax);
  ok = ok && alloc_act(c, "p4", static_cast<size_t>(mb) * R[3] * R[3] * 256 * 2);
  ok = ok && alloc_act(c, "p3", static_cast<size_t>(mb) * R[2] * R[2] * 256 * 2);
  ok = ok && alloc_act(c, "p2", static_cast<size_t>(mb) * R[1] * R[1] * 256 * 2);
  const int R0 = 8 * g;                              // path_1 resolution (S/2)
  ok = ok && alloc_act(c, "p1", static_cast<size_t>(mb) * R0 * R0 * 256 * 2);
  ok = ok && alloc_act(c, "mh1", static_cast<size_t>(mb) * R0 * R0 * 128 * 2);
  ok = ok && alloc_act(c, "feat0", static_cast<size_t>(mb) * S * S * 64 * 2);
  ok = ok && alloc_act(c, "feat", static_cast<size_t>(mb) * S * S * 64 * 2);
  // blocks per image of the pooled (last) up-sampling level: enough for 8 blocks per SM at a full micro-batch, and at least
  // S3OD_POOL_MIN per image so that the small tail chunks of the batch API (4 images) still fill the GPU; never more than
  // the work units of one image
  c->pool_blocks = std::max(1, std::min(std::max(8 * c->num_sms / std::max(1, mb) + 1, S3OD_POOL_MIN),
                                        ((R0 / 2 + 7) / 8) * ((R0 / 2 + 31) / 32)));
  ok = ok && alloc_act(c, "pool", static_cast<size_t>(mb) * c->pool_blocks * 256 * 4);
  if (!ok) return false;

  float* x = aptr<float>(c, "x");
  bf16* dx = aptr<bf16>(c, "dx");
  bf16* xn = aptr<bf16>(c, "xn");
  bf16* actx = aptr<bf16>(c, "ctx");
  bf16* hmid = aptr<bf16>(c, "hmid");
  const int sms = c->num_sms;

  // ---- encoder ---------------------------------------------------------------------------------------------------
  {
    EpiPatch::Params e{x, wptr<float>(c, "patch.b"), P, ntok, D};
    if (!add_linear<256, EpiPatch, 8>(c, "patch_embed", aptr<bf16>(c, "patches"), static_cast<uint64_t>(c->max_batch) * P, P, 768,
                                      wptr<bf16>(c, "patch.w"), D, e, /*a_is_batch_window=*/true, /*per_image_tiles=*/true))
      return false;
    const float* prefix = wptr<float>(c, "prefix");
    c->plan.emplace_back("prefix_tokens", [=](int nb, int, float*, float*, cudaStream_t st) {
      return launch_fill_prefix(x, prefix, ntok, D, nb, st);
    });
  }
  AttnParams ap{};
  {
    const uint64_t BH = static_cast<uint64_t>(mb) * H;
    const uint64_t dq[3] = {64, (uint64_t)ntok, BH};
    const uint64_t sq[2] = {128, (uint64_t)ntok * 128};
    const uint32_t bq[3] = {64, kAttnTile, 1}, bkv[3] = {64, kAttnKvTile, 1};
    if (!make_tmap(&ap.tma_q, aptr<bf16>(c, "q"), 3, dq, sq, bq)) return false;
    if (!make_tmap(&ap.tma_k, aptr<bf16>(c, "k"), 3, dq, sq, bkv)) return false;
    if (!make_tmap(&ap.tma_v, aptr<bf16>(c, "v"), 3, dq, sq, bkv)) return false;
    ap.out = actx;
    ap.ntok = ntok;
    ap.heads = H;
    ap.kv_tiles = (ntok + kAttnKvTile - 1) / kAttnKvTile;
  }
  for (int l = 0; l < c->L; ++l) {
    const std::string pre = "enc." + std::to_string(l) + ".";
    const float *ln1w = wptr<float>(c, pre + "ln1.w"), *ln1b = wptr<float>(c, pre + "ln1.b");
    const float *ln2w = wptr<float>(c, pre + "ln2.w"), *ln2b = wptr<float>(c, pre + "ln2.b");
    {
      // x += dx of the previous layer's MLP; hidden_states[l] is complete here, so this is where its tap is taken
      bf16* tap = nullptr;
      for (int j = 0; j < 4; ++j)
        if (c->taps[j] == l) tap = aptr<bf16>(c, "tap" + std::to_string(j));
      const bf16* dprev = l > 0 ? dx : nullptr;
      c->plan.emplace_back(pre + "ln1", [=](int nb, int, float*, float*, cudaStream_t st) {
        return launch_layernorm(x, dprev, ln1w, ln1b, xn, tap, nb * ntok, ntok, D, 1e-5f, st);
      });
    }
    {
      EpiQKV::Params e{};
      e.q = aptr<bf16>(c, "q"); e.k = aptr<bf16>(c, "k"); e.v = aptr<bf16>(c, "v");
      e.bias = wptr<float>(c, pre + "qkv.b");
      e.ntok = ntok; e.heads = H; e.D = D;
      e.grid_w = g; e.inv_gh = 1.0f / g; e.inv_gw = 1.0f / g;
      e.qscale = 0.125f * 1.4426950408889634f;
      if (!add_linear<256, EpiQKV, 8>(c, pre + "qkv", xn, MT, ntok, D, wptr<bf16>(c, pre + "qkv.w"), 3 * D, e, false, true)) return false;
    }
    c->plan.emplace_back(pre + "attention", [=](int nb, int, float*, float*, cudaStream_t st) {
      return launch_attention(ap, (ntok + 127) / 128, nb * H, st);
    });
    {
      EpiResidual::Params e{dx, wptr<float>(c, pre + "o.b"), wptr<float>(c, pre + "ls1"), ntok, D};
      // flat M tiling (S3OD_FLAT_TILES): dx is a plain [nb * ntok, D] matrix, so the tiles may straddle images - per-image
      // tiling spends one tile in 33 on the 5-row remainder of every image (ntok = 4101 = 32 * 128 + 5)
      if (!add_linear<256, EpiResidual, 8, 192>(c, pre + "o_proj", actx, MT, ntok, D, wptr<bf16>(c, pre + "o.w"), D, e, false, !kFlatTiles,
                                                flat_residual)) return false;
    }
    c->plan.emplace_back(pre + "ln2", [=](int nb, int, float*, float*, cudaStream_t st) {
      return launch_layernorm(x, dx, ln2w, ln2b, xn, nullptr, nb * ntok, ntok, D, 1e-5f, st);
    });
    {
      EpiGelu::Params e{hmid, wptr<float>(c, pre + "up.b"), I, ntok};
      if (!add_linear<256, EpiGelu, 8>(c, pre + "up_proj", xn, MT, ntok, D, wptr<bf16>(c, pre + "up.w"), I, e, false, !kFlatTiles,
                                       flat_gelu)) return false;
    }
    {
      EpiResidual::Params e{dx, wptr<float>(c, pre + "down.b"), wptr<float>(c, pre + "ls2"), ntok, D};
      if (!add_linear<256, EpiResidual, 8, 192>(c, pre + "down_proj", hmid, MT, ntok, I, wptr<bf16>(c, pre + "down.w"), D, e, false, !kFlatTiles,
                                                flat_residual)) return false;
    }
  }

  {  // x += dx of the last needed layer; its output is the last tap (no LayerNorm: the final encoder.norm is dead code, F3)
    bf16* tap = nullptr;
    for (int j = 0; j < 4; ++j)
      if (c->taps[j] == c->L) tap = aptr<bf16>(c, "tap" + std::to_string(j));
    c->plan.emplace_back("enc.final_residual", [=](int nb, int, float*, float*, cudaStream_t st) {
      return launch_layernorm(x, dx, nullptr, nullptr, nullptr, tap, nb * ntok, ntok, D, 1e-5f, st);
    });
  }

  // ---- DPT head: projects + resize layers (model.py:193-211) --------------------------------------------------------
  const uint64_t MP = static_cast<uint64_t>(mb) * P;
  for (int j = 0; j < 4; ++j) {
    const std::string sj = std::to_string(j);
    EpiConv::Params e = conv_epi(aptr<bf16>(c, "f" + sj), nullptr, wptr<float>(c, "head.proj" + sj + ".b"), nullptr, nullptr, 0,
                                 c->oc[j], g, g);
    e.linear = 1; e.hin = g; e.win = g;
    if (!add_linear<256, EpiConv, 8>(c, "head.proj" + sj, aptr<bf16>(c, "tap" + sj), MP, P, D, wptr<bf16>(c, "head.proj" + sj + ".w"),
                                     c->oc[j], e))
      return false;
  }
  {  // ConvTranspose k4 s4 (256 -> 256): one GEMM, N = 16 phases x 256, depth-to-space in the epilogue
    EpiConv::Params e = conv_epi(aptr<bf16>(c, "r0"), nullptr, wptr<float>(c, "head.rs0.b"), nullptr, nullptr, 0, c->oc[0], R[1], R[1]);
    e.linear = 1; e.hin = g; e.win = g; e.up = 4;
    if (!add_linear<256, EpiConv, 8>(c, "head.rs0", aptr<bf16>(c, "f0"), MP, P, c->oc[0], wptr<bf16>(c, "head.rs0.w"), 16 * c->oc[0], e))
      return false;
  }
  {  // ConvTranspose k2 s2 (512 -> 512)
    EpiConv::Params e = conv_epi(aptr<bf16>(c, "r1"), nullptr, wptr<float>(c, "head.rs1.b"), nullptr, nullptr, 0, c->oc[1], R[2], R[2]);
    e.linear = 1; e.hin = g; e.win = g; e.up = 2;
    if (!add_linear<256, EpiConv, 8>(c, "head.rs1", aptr<bf16>(c, "f1"), MP, P, c->oc[1], wptr<bf16>(c, "head.rs1.w"), 4 * c->oc[1], e))
      return false;
  }
  {  // Conv 3x3 stride 2 (1024 -> 1024) on f3
    CUtensorMap ta;
    if (!tmap_nhwc_s2(&ta, aptr<bf16>(c, "f3"), mb, g, g, c->oc[3])) return false;
    EpiConv::Params e = conv_epi(aptr<bf16>(c, "r3"), nullptr, wptr<float>(c, "head.rs3.b"), nullptr, nullptr, 0, c->oc[3], R[4], R[4]);
    if (!add_conv<256, EpiConv, 8>(c, "head.rs3", ta, geom_3x3_s2(R[4], R[4], c->oc[3]), wptr<bf16>(c, "head.rs3.w"), c->oc[3], 0,
                                   9 * c->oc[3], c->oc[3], e))
      return false;
  }
  // layer{k}_rn: 3x3, no bias, -> 256; the relu'd copy feeds the first conv of the residual unit that consumes it
  const char* rn_in[5] = {"", "r0", "r1", "f2", "r3"};
  for (int k = 1; k <= 4; ++k) {
    const std::string sk = std::to_string(k);
    EpiConv::Params e = conv_epi(aptr<bf16>(c, "l" + sk), aptr<bf16>(c, "l" + sk + "r"), nullptr, nullptr, nullptr, 0, 256, R[k], R[k]);
    if (!add_conv3x3<256, 8>(c, "head.rn" + sk, aptr<bf16>(c, rn_in[k]), R[k], R[k], c->oc[k - 1], "head.rn" + sk + ".w", 256, e))
      return false;
  }
  // ---- fusion blocks (model.py:383-405), BN folded into the conv weights (eval mode) -------------------------------
  bf16 *tA = aptr<bf16>(c, "tA"), *tB = aptr<bf16>(c, "tB"), *tC = aptr<bf16>(c, "tC"), *tD = aptr<bf16>(c, "tD");
  const char* pname[5] = {"p1", "p1", "p2", "p3", "p4"};    // output of refinenet k is p_k (index by k)
  for (int k = 4; k >= 1; --k) {
    const std::string rk = "head.ref" + std::to_string(k) + ".";
    const int Rk = R[k];
    bf16* lk = aptr<bf16>(c, "l" + std::to_string(k));
    bf16* lkr = aptr<bf16>(c, "l" + std::to_string(k) + "r");
    const bf16* sum_in;       // input of resConfUnit2 and its relu'd copy
    const bf16* sum_in_relu;
    if (k == 4) {             // refinenet4 is called with one input: resConfUnit1 unused
      sum_in = lk;
      sum_in_relu = lkr;
    } else {
      // output = p_{k+1} + resConfUnit1(l_k):   a = relu(bn1(conv1(relu(l_k))));  s = bn2(conv2(a)) + l_k + p_{k+1}
      if (!add_conv3x3<256, 8>(c, rk + "rcu1.c1", lkr, Rk, Rk, 256, rk + "rcu1.c1.w", 256,
                               conv_epi(tA, nullptr, wptr<float>(c, rk + "rcu1.c1.b"), nullptr, nullptr, 1, 256, Rk, Rk)))
        return false;
      if (!add_conv3x3<256, 8>(c, rk + "rcu1.c2", tA, Rk, Rk, 256, rk + "rcu1.c2.w", 256,
                               conv_epi(tB, tC, wptr<float>(c, rk + "rcu1.c2.b"), lk, aptr<bf16>(c, pname[k + 1]), 0, 256, Rk, Rk)))
        return false;
      sum_in = tB;
      sum_in_relu = tC;
    }
    // resConfUnit2
    if (!add_conv3x3<256, 8>(c, rk + "rcu2.c1", sum_in_relu, Rk, Rk, 256, rk + "rcu2.c1.w", 256,
                             conv_epi(tA, nullptr, wptr<float>(c, rk + "rcu2.c1.b"), nullptr, nullptr, 1, 256, Rk, Rk)))
      return false;
    if (!add_conv3x3<256, 8>(c, rk + "rcu2.c2", tA, Rk, Rk, 256, rk + "rcu2.c2.w", 256,
                             conv_epi(tD, nullptr, wptr<float>(c, rk + "rcu2.c2.b"), sum_in, nullptr, 0, 256, Rk, Rk)))
      return false;
    // out_conv (1x1) at low resolution, then 2x bilinear up-sampling
    {
      EpiConv::Params e = conv_epi(tA, nullptr, wptr<float>(c, rk + "out.b"), nullptr, nullptr, 0, 256, Rk, Rk);
      e.linear = 1; e.hin = Rk; e.win = Rk;
      if (!add_linear<256, EpiConv, 8>(c, rk + "out_conv", tD, static_cast<uint64_t>(mb) * Rk * Rk, Rk * Rk, 256,
                                       wptr<bf16>(c, rk + "out.w"), 256, e))
        return false;
    }
    {
      bf16* pk = aptr<bf16>(c, pname[k]);
      float* pool = (k == 1) ? aptr<float>(c, "pool") : nullptr;
      const int pb = c->pool_blocks;
      c->plan.emplace_back(rk + "upsample", [=](int nb, int, float*, float*, cudaStream_t st) {
        return launch_upsample2x(tA, pk, pool, pb, nb, Rk, Rk, sms, st);
      });
    }
  }
  // ---- IoU head (model.py:185-191)
  {
    const float* pool = aptr<float>(c, "pool");
    const int pb = c->pool_blocks;
    const float inv = 1.0f / (static_cast<float>(R0) * R0);
    const float *w1 = wptr<float>(c, "head.cls.w1"), *b1 = wptr<float>(c, "head.cls.b1");
    const float *w2 = wptr<float>(c, "head.cls.w2"), *b2 = wptr<float>(c, "head.cls.b2");
    c->plan.emplace_back("head.iou", [=](int nb, int b0, float*, float* io, cudaStream_t st) {
      return launch_iou_head(pool, pb, inv, w1, b1, w2, b2, io + static_cast<size_t>(b0) * K, K, nb, st);
    });
  }
  // ---- mask head (model.py:455-467)
  bf16 *p1 = aptr<bf16>(c, "p1"), *mh1 = aptr<bf16>(c, "mh1"), *feat0 = aptr<bf16>(c, "feat0"), *feat = aptr<bf16>(c, "feat");
  // 256 -> 128 channels: operand-swapped kernel (conv_swap.cuh) when the tile grid pairs up, else the 128-wide pair GEMM
  // (-DS3OD_SWAP128=0 builds the library without the swapped kernel: A/B measurements only)
  const ConvGeom g_c1 = geom_3x3(R0, R0, 256);
  if (kSwap128 && (g_c1.tiles_h * g_c1.tiles_w) % 2 == 0) {
    ConvSwapParams sp{};
    if (!tmap_nhwc(&sp.tma_x, p1, mb, R0, R0, 256)) return false;
    if (!tmap_matrix(&sp.tma_w, wptr<bf16>(c, "head.mh.c1.w"), 128, 9 * 256, 128)) return false;
    sp.geom = g_c1;
    sp.num_k_blocks = 9 * 256 / 64;
    sp.out = mh1;
    sp.bias = wptr<float>(c, "head.mh.c1.b");
    sp.relu = 0;
    const int sms = c->num_sms, per_img = g_c1.tiles_h * g_c1.tiles_w;
    c->plan.emplace_back("head.mh.c1", [=](int nb, int, float*, float*, cudaStream_t st) mutable -> cudaError_t {
      ConvSwapParams q = sp;
      q.m_tiles = nb * per_img;
      return launch_conv_swap128(q, sms, st);
    });
  } else if (!add_conv3x3<128, 8>(c, "head.mh.c1", p1, R0, R0, 256, "head.mh.c1.w", 128,
                                  conv_epi(mh1, nullptr, wptr<float>(c, "head.mh.c1.b"), nullptr, nullptr, 0, 128, R0, R0))) {
    return false;
  }
  const bool rows_ok = kRowConv && (S % kRowPx == 0);
  if (rows_ok && R0 % kRowPx == 0) {
    // ConvTranspose2d k4 s2 p1 + ReLU on the row-streaming kernel, one launch per output-column phase
    ConvTRowParams p{};
    if (!make_convt_rows(&p, mh1, wptr<bf16>(c, "head.mh.up.wr"), wptr<float>(c, "head.mh.up.b"), feat0, mb, R0, R0, 1)) return false;
    const int sms = c->num_sms;
    for (int b = 0; b < 2; ++b) {
      p.phase_b = b;
      c->plan.emplace_back("head.mh.up.rows" + std::to_string(b), [=](int nb, int, float*, float*, cudaStream_t st) -> cudaError_t {
        ConvTRowParams q = p;
        q.num_strips = nb * q.strips_x * q.strips_y;
        return launch_convt_rows(q, sms, st);
      });
    }
  } else {
    CUtensorMap ta;
    if (!tmap_nhwc(&ta, mh1, mb, R0, R0, 128)) return false;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        EpiConv::Params e = conv_epi(feat0, nullptr, wptr<float>(c, "head.mh.up.b"), nullptr, nullptr, 1, 64, S, S);
        e.up = 2; e.ph_h = a; e.ph_w = b;
        if (!add_conv<64, EpiConv, 4>(c, "head.mh.up." + std::to_string(a) + std::to_string(b), ta, geom_convt_phase(R0, R0, 128, a, b),
                                      wptr<bf16>(c, "head.mh.up.w"), 4 * 64, (a * 2 + b) * 64, 4 * 128, 64, e))
          return false;
      }
  }
  if (rows_ok) {
    if (!add_conv_rows<64, EpiConv>(c, "head.mh.c2", feat0, S, S, wptr<bf16>(c, "head.mh.c2.w"),
                                    conv_epi(feat, nullptr, wptr<float>(c, "head.mh.c2.b"), nullptr, nullptr, 1, 64, S, S)))
      return false;
  } else {
    CUtensorMap ta;
    if (!tmap_nhwc(&ta, feat0, mb, S, S, 64)) return false;
    if (!add_conv<64, EpiConv, 4>(c, "head.mh.c2", ta, geom_3x3(S, S, 64), wptr<bf16>(c, "head.mh.c2.w"), 64, 0, 9 * 64, 64,
                                  conv_epi(feat, nullptr, wptr<float>(c, "head.mh.c2.b"), nullptr, nullptr, 1, 64, S, S)))
      return false;
  }
  {
    CUtensorMap ta;
    if (!tmap_nhwc(&ta, feat, mb, S, S, 64)) return false;
    EpiMask::Params e{nullptr, wptr<float>(c, "head.mh.heads.b"), wptr<float>(c, "head.mh.heads.w2"), wptr<float>(c, "head.mh.heads.b2"), S, K};
    auto patch = [S, K](EpiMask::Params& q, int, int b0, float* mo, float*) { q.out = mo + static_cast<size_t>(b0) * K * S * S; };
    bool r;
    if (rows_ok && K == 3)
      r = add_conv_rows<96, EpiMask>(c, "head.mh.heads", feat, S, S, wptr<bf16>(c, "head.mh.heads.w"), e, patch);
    else if (rows_ok && K == 1)
      r = add_conv_rows<32, EpiMask>(c, "head.mh.heads", feat, S, S, wptr<bf16>(c, "head.mh.heads.w"), e, patch);
    else if (K == 3)
      r = add_conv<96, EpiMask, 4>(c, "head.mh.heads", ta, geom_3x3(S, S, 64), wptr<bf16>(c, "head.mh.heads.w"), 96, 0, 9 * 64, 96, e, patch);
    else
      r = add_conv<32, EpiMask, 4>(c, "head.mh.heads", ta, geom_3x3(S, S, 64), wptr<bf16>(c, "head.mh.heads.w"), 32, 0, 9 * 64, 32, e, patch);
    if (!r) return false;
  }
  return true;
}

// Every tensor the plan reads, with the exact byte size the architecture implies (D / I / heads / out_channels / K): the
// launch plan builds TMA maps and pointers from these dimensions alone, so a checkpoint of another architecture must be
// refused here - the reference's strict load_state_dict raises in the same situation (predictor.py:76).
std::vector<std::pair<std::string, size_t>> required_tensors(const s3od_ctx* c) {
  const size_t D = c->D, I = c->I, K = c->K, F = 256;
  std::vector<std::pair<std::string, size_t>> r = {
      {"patch.w", D * 768 * 2}, {"patch.b", D * 4}, {"prefix", 5 * D * 4}, {"pre.lut", 768 * 2}};
  for (int l = 0; l < c->L; ++l) {
    const std::string p = "enc." + std::to_string(l) + ".";
    r.push_back({p + "ln1.w", D * 4}); r.push_back({p + "ln1.b", D * 4});
    r.push_back({p + "qkv.w", 3 * D * D * 2}); r.push_back({p + "qkv.b", 3 * D * 4});
    r.push_back({p + "o.w", D * D * 2}); r.push_back({p + "o.b", D * 4}); r.push_back({p + "ls1", D * 4});
    r.push_back({p + "ln2.w", D * 4}); r.push_back({p + "ln2.b", D * 4});
    r.push_back({p + "up.w", I * D * 2}); r.push_back({p + "up.b", I * 4});
    r.push_back({p + "down.w", D * I * 2}); r.push_back({p + "down.b", D * 4}); r.push_back({p + "ls2", D * 4});
  }
  for (int j = 0; j < 4; ++j) {
    const size_t oc = c->oc[j];
    r.push_back({"head.proj" + std::to_string(j) + ".w", oc * D * 2});
    r.push_back({"head.proj" + std::to_string(j) + ".b", oc * 4});
    r.push_back({"head.rn" + std::to_string(j + 1) + ".w", F * 9 * oc * 2});
  }
  const size_t oc0 = c->oc[0], oc1 = c->oc[1], oc3 = c->oc[3];
  r.push_back({"head.rs0.w", 16 * oc0 * oc0 * 2}); r.push_back({"head.rs0.b", oc0 * 4});
  r.push_back({"head.rs1.w", 4 * oc1 * oc1 * 2}); r.push_back({"head.rs1.b", oc1 * 4});
  r.push_back({"head.rs3.w", oc3 * 9 * oc3 * 2}); r.push_back({"head.rs3.b", oc3 * 4});
  r.push_back({"head.mh.c1.w", 128 * 9 * F * 2}); r.push_back({"head.mh.c1.b", 128 * 4});
  r.push_back({"head.mh.up.w", 4 * 64 * 4 * 128 * 2}); r.push_back({"head.mh.up.wr", 16 * 64 * 128 * 2});
  r.push_back({"head.mh.up.b", 64 * 4});
  r.push_back({"head.mh.c2.w", 64 * 9 * 64 * 2}); r.push_back({"head.mh.c2.b", 64 * 4});
  r.push_back({"head.mh.heads.w", 32 * K * 9 * 64 * 2}); r.push_back({"head.mh.heads.b", 32 * K * 4});
  r.push_back({"head.mh.heads.w2", K * 32 * 4}); r.push_back({"head.mh.heads.b2", K * 4});
  r.push_back({"head.cls.w1", 64 * F * 4}); r.push_back({"head.cls.b1", 64 * 4});
  r.push_back({"head.cls.w2", K * 64 * 4}); r.push_back({"head.cls.b2", K * 4});
  for (int k = 1; k <= 4; ++k) {
    const std::string p = "head.ref" + std::to_string(k) + ".";
    r.push_back({p + "out.w", F * F * 2});
    r.push_back({p + "out.b", F * 4});
    for (int u = (k == 4 ? 2 : 1); u <= 2; ++u)
      for (int cc = 1; cc <= 2; ++cc) {
        r.push_back({p + "rcu" + std::to_string(u) + ".c" + std::to_string(cc) + ".w", F * 9 * F * 2});
        r.push_back({p + "rcu" + std::to_string(u) + ".c" + std::to_string(cc) + ".b", F * 4});
      }
  }
  return r;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

const char* s3od_last_error(void) { return g_err.c_str(); }
const char* s3od_version(void) { return "s3od_b200 0.1 (sm_100a)"; }

int s3od_create(s3od_ctx** out, int device, int arch, int num_outputs, int image_size, int max_batch, int micro_batch) {
  if (out == nullptr) return fail(S3OD_ERR_ARG, "ctx out pointer is null");
  if (arch != S3OD_ARCH_VITB && arch != S3OD_ARCH_VITL) return fail(S3OD_ERR_ARG, "unknown arch");
  if (num_outputs != 1 && num_outputs != 3) return fail(S3OD_ERR_ARG, "num_outputs must be 1 or 3");
  if (image_size < 32 || image_size % 32 != 0) return fail(S3OD_ERR_ARG, "image_size must be a positive multiple of 32");
  if (max_batch < 1 || micro_batch < 1) return fail(S3OD_ERR_ARG, "batch sizes must be >= 1");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(S3OD_ERR_STATE, std::string("s3od_b200 needs an sm_100 GPU, found ") + prop.name);
  s3od_ctx* c = new s3od_ctx();
  c->device = device; c->arch = arch; c->K = num_outputs; c->S = image_size;
  c->max_batch = max_batch; c->mb = micro_batch < max_batch ? micro_batch : max_batch;
  c->g = image_size / 16; c->P = c->g * c->g; c->ntok = c->P + 5;
  c->vt_pitch = (c->ntok + 7) & ~7;
  if (arch == S3OD_ARCH_VITB) {
    c->D = 768; c->H = 12; c->I = 3072; c->L = 11;
    const int t[4] = {2, 5, 8, 11};
    memcpy(c->taps, t, sizeof(t));
  } else {
    c->D = 1024; c->H = 16; c->I = 4096; c->L = 23;
    const int t[4] = {4, 11, 17, 23};
    memcpy(c->taps, t, sizeof(t));
  }
  c->num_sms = prop.multiProcessorCount;
  *out = c;
  return S3OD_OK;
}

int s3od_set_tensor(s3od_ctx* c, const char* name, const void* host_data, size_t bytes) {
  if (c == nullptr || name == nullptr || host_data == nullptr || bytes == 0) return fail(S3OD_ERR_ARG, "bad argument to s3od_set_tensor");
  if (c->finalized) return fail(S3OD_ERR_STATE, "context already finalized");
  CK(cudaSetDevice(c->device));
  void* p = nullptr;
  CK(cudaMalloc(&p, (bytes + 255) & ~size_t(255)));
  CK(cudaMemcpy(p, host_data, bytes, cudaMemcpyHostToDevice));
  auto it = c->w.find(name);
  if (it != c->w.end()) cudaFree(it->second.p);
  c->w[name] = DevBuf{p, bytes};
  return S3OD_OK;
}

int s3od_finalize(s3od_ctx* c) {
  if (c == nullptr) return fail(S3OD_ERR_ARG, "null ctx");
  if (c->finalized) return S3OD_OK;
  CK(cudaSetDevice(c->device));
  for (const auto& req : required_tensors(c)) {
    auto it = c->w.find(req.first);
    if (it == c->w.end()) return fail(S3OD_ERR_MISSING, "missing tensor: " + req.first);
    if (it->second.bytes != req.second)
      return fail(S3OD_ERR_ARG, "size mismatch for " + req.first + ": got " + std::to_string(it->second.bytes) + " bytes, this architecture needs " +
                                    std::to_string(req.second) + " (checkpoint of another model?)");
  }
  CK(cudaMalloc(reinterpret_cast<void**>(&c->d_img), sizeof(ImageDesc) * c->max_batch));
  CK(cudaMalloc(reinterpret_cast<void**>(&c->d_post), sizeof(PostDesc) * c->max_batch));
  // Optional "pre.affine" = {a[3], b[3]} fp32: the preprocess kernel evaluates bf16(fma(v, a, b)) instead of the table
  // look-up when - and only when - that reproduces all 3 x 256 entries of "pre.lut" bit for bit (fmaf and the
  // round-to-nearest-even bf16 conversion are the same IEEE operations on the host and on the device).
  c->pre_affine = false;
  auto aff = c->w.find("pre.affine");
  if (aff != c->w.end() && aff->second.bytes == 6 * sizeof(float) && c->w["pre.lut"].bytes == 768 * sizeof(uint16_t)) {
    uint16_t lut[768];
    CK(cudaMemcpy(lut, c->w["pre.lut"].p, sizeof(lut), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c->pre_ab, aff->second.p, sizeof(c->pre_ab), cudaMemcpyDeviceToHost));
    bool same = true;
    for (int ch = 0; ch < 3 && same; ++ch)
      for (int v = 0; v < 256 && same; ++v) {
        const float f = fmaf(static_cast<float>(v), c->pre_ab[ch], c->pre_ab[3 + ch]);
        uint32_t u;
        memcpy(&u, &f, 4);
        const uint16_t bf = static_cast<uint16_t>((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);     // RNE (finite values)
        same = bf == lut[ch * 256 + v];
      }
    c->pre_affine = same;
  }
  if (!build_plan(c)) return S3OD_ERR_CUDA;
  c->finalized = true;
  return S3OD_OK;
}

static_assert(sizeof(s3od_image) == sizeof(ImageDesc), "s3od_image must mirror ImageDesc");
static_assert(sizeof(s3od_post) == sizeof(PostDesc), "s3od_post must mirror PostDesc");

// CUDA-event pair around one launch outside the forward plan (op index plan.size() = preprocess, +1 = postprocess)
static int profile_mark(s3od_ctx* c, int op, int nb, cudaStream_t st, bool begin) {
  if (!c->profile) return S3OD_OK;
  if (begin) {
    while (c->ev_pool.size() < c->ev_used + 2) {
      cudaEvent_t ev;
      CK(cudaEventCreate(&ev));
      c->ev_pool.push_back(ev);
    }
    c->ev_ops.emplace_back(op, nb);
    CK(cudaEventRecord(c->ev_pool[c->ev_used], st));
    c->ev_used += 2;
  } else {
    CK(cudaEventRecord(c->ev_pool[c->ev_used - 1], st));
  }
  return S3OD_OK;
}

int s3od_preprocess_u8(s3od_ctx* c, const s3od_image* images, int batch, s3od_stream stream) {
  if (c == nullptr || !c->finalized) return fail(S3OD_ERR_STATE, "context not finalized");
  if (images == nullptr || batch < 1 || batch > c->max_batch) return fail(S3OD_ERR_ARG, "bad batch for s3od_preprocess_u8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CK(cudaMemcpyAsync(c->d_img, images, sizeof(ImageDesc) * batch, cudaMemcpyHostToDevice, st));
  if (profile_mark(c, static_cast<int>(c->plan.size()), batch, st, true) != S3OD_OK) return S3OD_ERR_CUDA;
  int common_mode = images[0].mode;
  for (int i = 0; i < batch; ++i) {
    const s3od_image& im = images[i];
    if (im.mode != common_mode) common_mode = 2;
    // the vector paths read the source with 8- / 16-byte loads relative to d_src
    if (im.d_src == nullptr || (reinterpret_cast<uintptr_t>(im.d_src) & 15) != 0)
      return fail(S3OD_ERR_ARG, "s3od_preprocess_u8: d_src must be a 16-byte aligned device pointer");
    if (im.mode < 0 || im.mode > 2 || im.h < 1 || im.w < 1 || im.new_h < 1 || im.new_w < 1 || im.pad_h < 0 || im.pad_w < 0)
      return fail(S3OD_ERR_ARG, "s3od_preprocess_u8: bad image geometry / resize mode");
    if (im.mode == 0 && (im.new_h != im.h || im.new_w != im.w)) return fail(S3OD_ERR_ARG, "s3od_preprocess_u8: mode 0 needs new size == size");
    if (im.mode == 1 && (2 * im.new_h != im.h || 2 * im.new_w != im.w))
      return fail(S3OD_ERR_ARG, "s3od_preprocess_u8: mode 1 needs size == 2 x new size");
    if (im.mode == 2 && (im.d_xtab == nullptr || im.d_ytab == nullptr))
      return fail(S3OD_ERR_ARG, "s3od_preprocess_u8: mode 2 needs the coefficient tables");
  }
  CK(launch_preprocess(c->d_img, wptr<bf16>(c, "pre.lut"), c->pre_affine ? c->pre_ab : nullptr, aptr<bf16>(c, "patches"), c->S, batch,
                       common_mode, st));
  if (profile_mark(c, static_cast<int>(c->plan.size()), batch, st, false) != S3OD_OK) return S3OD_ERR_CUDA;
  c->launches += 1;
  return S3OD_OK;
}

int s3od_pack_input_f32(s3od_ctx* c, const float* d_x, int batch, s3od_stream stream) {
  if (c == nullptr || !c->finalized) return fail(S3OD_ERR_STATE, "context not finalized");
  if (d_x == nullptr || batch < 1 || batch > c->max_batch) return fail(S3OD_ERR_ARG, "bad batch for s3od_pack_input_f32");
  CK(launch_pack_input(d_x, aptr<bf16>(c, "patches"), c->S, batch, static_cast<cudaStream_t>(stream)));
  c->launches += 1;
  return S3OD_OK;
}

int s3od_forward(s3od_ctx* c, int batch, float* d_mask_logits, float* d_iou_logits, s3od_stream stream) {
  if (c == nullptr || !c->finalized) return fail(S3OD_ERR_STATE, "context not finalized");
  if (batch < 1 || batch > c->max_batch || d_mask_logits == nullptr || d_iou_logits == nullptr)
    return fail(S3OD_ERR_ARG, "bad argument to s3od_forward");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int b0 = 0; b0 < batch; b0 += c->mb) {
    const int nb = std::min(c->mb, batch - b0);
    for (size_t oi = 0; oi < c->plan.size(); ++oi) {
      auto& op = c->plan[oi];
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (c->profile) {
        while (c->ev_pool.size() < c->ev_used + 2) {
          cudaEvent_t ev;
          CK(cudaEventCreate(&ev));
          c->ev_pool.push_back(ev);
        }
        e0 = c->ev_pool[c->ev_used++];
        e1 = c->ev_pool[c->ev_used++];
        c->ev_ops.emplace_back(static_cast<int>(oi), nb);
        CK(cudaEventRecord(e0, st));
      }
      cudaError_t e = op.second(nb, b0, d_mask_logits, d_iou_logits, st);
      if (e != cudaSuccess) return fail(S3OD_ERR_CUDA, "launch of '" + op.first + "' failed: " + cudaGetErrorString(e));
      if (c->profile) CK(cudaEventRecord(e1, st));
      c->launches += 1;
    }
    c->last_nb = nb;
  }
  return S3OD_OK;
}

int s3od_postprocess(s3od_ctx* c, const float* d_mask_logits, const float* d_iou_logits, const s3od_post* images, int batch,
                     float* d_ious, int32_t* d_best_idx, s3od_stream stream) {
  if (c == nullptr || !c->finalized) return fail(S3OD_ERR_STATE, "context not finalized");
  if (images == nullptr || batch < 1 || batch > c->max_batch || d_mask_logits == nullptr || d_iou_logits == nullptr ||
      d_ious == nullptr || d_best_idx == nullptr)
    return fail(S3OD_ERR_ARG, "bad argument to s3od_postprocess");
  if ((reinterpret_cast<uintptr_t>(d_mask_logits) & 15) != 0) return fail(S3OD_ERR_ARG, "s3od_postprocess: d_mask_logits must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int maxH = 0, maxW = 0;
  bool mult4 = true;
  // Tile kernel (up-sampling or identity, <= 3 taps per axis): the largest input region one R x 512 output tile reads.
  // ATen's first tap is trunc(scale * (i + 0.5) - support + 0.5) clamped at 0, so n outputs span <= scale * (n - 1) + 1 + k inputs.
  // R = 32 output rows when that region fits in 24 KB of shared memory (2x up-sampling: 20 x 260 floats), else 16.
  bool tile_ok = true, identity = true, twice = true;
  for (int i = 0; i < batch; ++i) {
    const s3od_post& im = images[i];
    maxH = std::max(maxH, im.H);
    maxW = std::max(maxW, im.W);
    mult4 = mult4 && (im.W % 4 == 0);
    const int in_h = c->S - 2 * im.pad_h, in_w = c->S - 2 * im.pad_w;
    if (im.H < 1 || im.W < 1 || in_h < 1 || in_w < 1 || im.pad_h < 0 || im.pad_w < 0 || im.ky < 1 || im.kx < 1)
      return fail(S3OD_ERR_ARG, "bad image geometry in s3od_postprocess");
    if (im.d_src == nullptr || im.d_all_masks == nullptr || im.d_rgba == nullptr || im.d_ystart == nullptr || im.d_yw == nullptr ||
        im.d_xstart == nullptr || im.d_xw == nullptr)
      return fail(S3OD_ERR_ARG, "null pointer in s3od_postprocess image descriptor");
    // 128-bit stores of the mask planes / RGBA pixels and 32-bit loads of the RGB source
    if (((reinterpret_cast<uintptr_t>(im.d_all_masks) | reinterpret_cast<uintptr_t>(im.d_rgba)) & 15) != 0 ||
        (reinterpret_cast<uintptr_t>(im.d_src) & 3) != 0)
      return fail(S3OD_ERR_ARG, "s3od_postprocess: d_all_masks / d_rgba must be 16-byte aligned, d_src 4-byte aligned");
    tile_ok = tile_ok && im.ky <= 3 && im.kx <= 3 && im.H >= in_h && im.W >= in_w;
    // scale exactly 1: ATen's table is (first tap i, weights 1, 0), the resize is a copy of the cropped mask
    identity = identity && im.H == in_h && im.W == in_w && im.pad_w % 4 == 0 && im.ky <= 2 && im.kx <= 2;
    // scale exactly 1/2: ATen's table is the plain 2x bilinear one (taps .25 / .75, weight 1 at the borders)
    twice = twice && im.H == 2 * in_h && im.W == 2 * in_w && im.pad_w % 2 == 0 && im.ky <= 3 && im.kx <= 3;
  }
  int tile_out_rows = 0, tile_rows = 0, tile_cols = 0;
  for (int R : {32, 16}) {
    if (!tile_ok) break;
    tile_out_rows = R; tile_rows = 0; tile_cols = 0;
    for (int i = 0; i < batch; ++i) {
      const s3od_post& im = images[i];
      const int in_h = c->S - 2 * im.pad_h, in_w = c->S - 2 * im.pad_w;
      const long long rows = (static_cast<long long>(R) * in_h + im.H - 1) / im.H + im.ky + 1;
      const long long cols = (512LL * in_w + im.W - 1) / im.W + im.kx + 1;
      tile_rows = std::max(tile_rows, static_cast<int>(std::min<long long>(rows, c->S)));
      tile_cols = std::max(tile_cols, static_cast<int>(std::min<long long>(cols, c->S)));
    }
    if (static_cast<size_t>(tile_rows) * (tile_cols + tile_cols / 32 + 1) * sizeof(float) <= 24 * 1024) break;
  }
  if (identity && mult4) tile_out_rows = -1;
  else if (twice && mult4) tile_out_rows = -2;
  CK(cudaMemcpyAsync(c->d_post, images, sizeof(PostDesc) * batch, cudaMemcpyHostToDevice, st));
  if (profile_mark(c, static_cast<int>(c->plan.size()) + 1, batch, st, true) != S3OD_OK) return S3OD_ERR_CUDA;
  CK(launch_postprocess(c->d_post, d_mask_logits, d_iou_logits, d_ious, d_best_idx, c->S, c->K, batch, maxH, maxW, mult4, tile_out_rows,
                        tile_rows, tile_cols, st));
  if (profile_mark(c, static_cast<int>(c->plan.size()) + 1, batch, st, false) != S3OD_OK) return S3OD_ERR_CUDA;
  c->launches += 1;
  return S3OD_OK;
}

int s3od_get_stage(s3od_ctx* c, const char* name, void** d_ptr, size_t* bytes) {
  if (c == nullptr || name == nullptr || d_ptr == nullptr || bytes == nullptr) return fail(S3OD_ERR_ARG, "bad argument to s3od_get_stage");
  auto it = c->act.find(name);
  if (it == c->act.end()) return fail(S3OD_ERR_MISSING, std::string("unknown stage: ") + name);
  *d_ptr = it->second.p;
  *bytes = it->second.bytes;
  return S3OD_OK;
}

int s3od_read_stage(s3od_ctx* c, const char* name, void* d_dst, size_t bytes, s3od_stream stream) {
  if (c == nullptr || name == nullptr || d_dst == nullptr) return fail(S3OD_ERR_ARG, "bad argument to s3od_read_stage");
  auto it = c->act.find(name);
  if (it == c->act.end()) return fail(S3OD_ERR_MISSING, std::string("unknown stage: ") + name);
  if (bytes > it->second.bytes) return fail(S3OD_ERR_ARG, std::string("stage ") + name + " is smaller than the requested size");
  CK(cudaMemcpyAsync(d_dst, it->second.p, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_profile_enable(s3od_ctx* c, int on) {
  if (c == nullptr) return fail(S3OD_ERR_ARG, "null ctx");
  c->profile = on != 0;
  c->ev_used = 0;
  c->ev_ops.clear();
  return S3OD_OK;
}

// Writes "label<TAB>launches<TAB>images<TAB>total_ms\n" per plan entry, for all forwards since the last enable/read.
int s3od_profile_read(s3od_ctx* c, char* buf, size_t buf_bytes) {
  if (c == nullptr || buf == nullptr || buf_bytes == 0) return fail(S3OD_ERR_ARG, "bad argument to s3od_profile_read");
  const size_t nops = c->plan.size() + 2;          // + preprocess, postprocess
  std::vector<double> ms(nops, 0.0);
  std::vector<long long> cnt(nops, 0), imgs(nops, 0);
  if (c->ev_used > 0) CK(cudaEventSynchronize(c->ev_pool[c->ev_used - 1]));
  for (size_t i = 0; i < c->ev_ops.size(); ++i) {
    float t = 0.0f;
    CK(cudaEventElapsedTime(&t, c->ev_pool[2 * i], c->ev_pool[2 * i + 1]));
    ms[c->ev_ops[i].first] += t;
    cnt[c->ev_ops[i].first] += 1;
    imgs[c->ev_ops[i].first] += c->ev_ops[i].second;
  }
  std::string out;
  for (size_t i = 0; i < nops; ++i) {
    char line[256];
    const char* label = i < c->plan.size() ? c->plan[i].first.c_str() : (i == c->plan.size() ? "preprocess" : "postprocess");
    snprintf(line, sizeof(line), "%s\t%lld\t%lld\t%.6f\n", label, cnt[i], imgs[i], ms[i]);
    out += line;
  }
  c->ev_used = 0;
  c->ev_ops.clear();
  if (out.size() + 1 > buf_bytes) return fail(S3OD_ERR_ARG, "profile buffer too small: need " + std::to_string(out.size() + 1));
  memcpy(buf, out.c_str(), out.size() + 1);
  return S3OD_OK;
}

long long s3od_launch_count(s3od_ctx* c) { return c == nullptr ? 0 : c->launches; }

int s3od_preprocess_mode(s3od_ctx* c) {
  if (c == nullptr || !c->finalized) return fail(S3OD_ERR_STATE, "context not finalized");
  return c->pre_affine ? 1 : 0;
}

void s3od_destroy(s3od_ctx* c) {
  if (c == nullptr) return;
  cudaSetDevice(c->device);
  for (auto& kv : c->w) cudaFree(kv.second.p);
  for (void* p : c->allocs) cudaFree(p);
  if (c->d_img) cudaFree(c->d_img);
  if (c->d_post) cudaFree(c->d_post);
  for (cudaEvent_t ev : c->ev_pool) cudaEventDestroy(ev);
  delete c;
}

// ------------------------------------------------------------------------------------------ kernel-level entry points
int s3od_op_gemm_f32(const void* d_a, const void* d_b, float* d_c, int M, int N, int K, s3od_stream stream) {
  return s3od_op_gemm_f32_bias(d_a, d_b, nullptr, d_c, M, N, K, stream);
}

int s3od_op_gemm_f32_bias(const void* d_a, const void* d_b, const float* d_bias, float* d_c, int M, int N, int K, s3od_stream stream) {
  if (N % 128 != 0 || K % 64 != 0 || M < 1) return fail(S3OD_ERR_ARG, "s3od_op_gemm_f32 needs N % 128 == 0 and K % 64 == 0");
  if ((reinterpret_cast<uintptr_t>(d_bias) & 15) != 0) return fail(S3OD_ERR_ARG, "s3od_op_gemm_f32_bias needs a 16-byte aligned bias");
  GemmParams<EpiStoreF32> p{};
  if (!tmap_matrix(&p.tma_a, d_a, M, K, kBM)) return S3OD_ERR_CUDA;
  if (!tmap_matrix(&p.tma_b, d_b, N, K, b_box_rows<128>())) return S3OD_ERR_CUDA;
  p.M = M; p.m_tiles = (M + kBM - 1) / kBM; p.n_tiles = N / 128; p.num_k_blocks = K / 64;
  p.epi = EpiStoreF32::Params{d_c, N, 0, d_bias};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (N % 256 == 0) {                               // the 256-wide tile configuration every encoder GEMM uses (CTA-pair kernel)
    if (!tmap_matrix(&p.tma_b, d_b, N, K, b_box_rows<256>())) return S3OD_ERR_CUDA;
    p.n_tiles = N / 256;
    CK((launch_gemm<256, A_LINEAR, EpiStoreF32, 8>(p, sms, static_cast<cudaStream_t>(stream))));
    return S3OD_OK;
  }
  CK((launch_gemm<128, A_LINEAR, EpiStoreF32, 8>(p, sms, static_cast<cudaStream_t>(stream))));
  return S3OD_OK;
}

// c[i] = sum over z of partial[z][i]   (the k-splits of s3od_op_gemm_f32_splitk; n a multiple of 4, 16-byte aligned)
__global__ void __launch_bounds__(256) sum_k_splits_kernel(const float4* __restrict__ partial, int splits, long long n4, float4* __restrict__ c) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    float4 acc = partial[i];
    for (int z = 1; z < splits; ++z) {
      const float4 v = partial[z * n4 + i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    c[i] = acc;
  }
}

int s3od_op_gemm_f32_splitk(const void* d_a, const void* d_b, float* d_c, int M, int N, int K, int splits, float* d_workspace, s3od_stream stream) {
  if (splits <= 1) return s3od_op_gemm_f32(d_a, d_b, d_c, M, N, K, stream);
  if (!use_pair_kernel()) return fail(S3OD_ERR_ARG, "s3od_op_gemm_f32_splitk needs the CTA-pair GEMM (library built with S3OD_PAIR=0)");
  if (N % 128 != 0 || M < 1 || splits > 64 || K % (64 * splits) != 0 || d_workspace == nullptr || (reinterpret_cast<uintptr_t>(d_c) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(d_workspace) & 15) != 0)
    return fail(S3OD_ERR_ARG, "s3od_op_gemm_f32_splitk needs N % 128 == 0, K % (64 * splits) == 0, splits <= 64 and a 16-byte aligned workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmParams<EpiStoreF32> p{};
  const bool wide = N % 256 == 0;
  if (!tmap_matrix(&p.tma_a, d_a, M, K, kBM)) return S3OD_ERR_CUDA;
  if (!tmap_matrix(&p.tma_b, d_b, N, K, wide ? b_box_rows<256>() : b_box_rows<128>())) return S3OD_ERR_CUDA;
  p.M = M; p.m_tiles = (M + kBM - 1) / kBM; p.n_tiles = N / (wide ? 256 : 128);
  p.num_k_blocks = K / 64 / splits;
  p.k_splits = splits;
  p.epi = EpiStoreF32::Params{d_workspace, N, static_cast<long long>(M) * N};
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (wide) CK((launch_gemm<256, A_LINEAR, EpiStoreF32, 8>(p, sms, st)));
  else CK((launch_gemm<128, A_LINEAR, EpiStoreF32, 8>(p, sms, st)));
  const long long n4 = static_cast<long long>(M) * N / 4;
  const int grid = static_cast<int>(std::min<long long>((n4 + 255) / 256, 148LL * 8));
  sum_k_splits_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(d_workspace), splits, n4, reinterpret_cast<float4*>(d_c));
  CK(cudaGetLastError());
  return S3OD_OK;
}

int s3od_op_wgrad_gemm_f32(const void* d_a, int lda, const void* d_b, int ldb, float* d_c, int M, int N, int K, int splits, float* d_workspace,
                           s3od_stream stream) {
  if (d_a == nullptr || d_b == nullptr || d_c == nullptr || M < 64 || N < 64 || K < 1 || M % 64 != 0 || N % 64 != 0 || lda < M || ldb < N || lda % 8 != 0 ||
      ldb % 8 != 0 || (reinterpret_cast<uintptr_t>(d_a) & 15) != 0 || (reinterpret_cast<uintptr_t>(d_b) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(d_c) & 15) != 0)
    return fail(S3OD_ERR_ARG, "s3od_op_wgrad_gemm_f32 needs M % 64 == 0, N % 64 == 0, row pitches that are multiples of 8 elements and 16-byte aligned buffers");
  GemmTnParams p{};
  p.M = M; p.N = N;
  p.m_tiles = (M + 127) / 128; p.n_tiles = (N + 255) / 256;
  p.k_blocks = (K + 63) / 64;
  if (splits < 1) splits = 1;
  if (splits > p.k_blocks) splits = p.k_blocks;
  p.k_blocks_per_split = (p.k_blocks + splits - 1) / splits;
  p.splits = (p.k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;          // no split without a k-block
  if (p.splits > 1 && (d_workspace == nullptr || (reinterpret_cast<uintptr_t>(d_workspace) & 15) != 0))
    return fail(S3OD_ERR_ARG, "s3od_op_wgrad_gemm_f32: splits > 1 need a 16-byte aligned workspace of splits * M * N floats");
  p.out = p.splits > 1 ? d_workspace : d_c;
  const uint64_t da[3] = {64, (uint64_t)K, (uint64_t)(M / 64)}, sa[2] = {(uint64_t)lda * 2, 128};
  const uint64_t db[3] = {64, (uint64_t)K, (uint64_t)(N / 64)}, sb[2] = {(uint64_t)ldb * 2, 128};
  const uint32_t ba[3] = {64, 64, 2}, bb[3] = {64, 64, 4};
  if (!make_tmap(&p.tma_a, d_a, 3, da, sa, ba) || !make_tmap(&p.tma_b, d_b, 3, db, sb, bb)) return S3OD_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_gemm_tn(p, sms, st));
  if (p.splits > 1) {
    const long long n4 = static_cast<long long>(M) * N / 4;
    const int grid = static_cast<int>(std::min<long long>((n4 + 255) / 256, 148LL * 8));
    sum_k_splits_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(d_workspace), p.splits, n4, reinterpret_cast<float4*>(d_c));
    CK(cudaGetLastError());
  }
  return S3OD_OK;
}

int s3od_op_conv3x3_wgrad_f32(const void* d_dy, const void* d_x, float* d_dw, int batch, int h, int w, int cin, int cout, int splits,
                              float* d_workspace, s3od_stream stream) {
  if (d_dy == nullptr || d_x == nullptr || d_dw == nullptr || batch < 1 || h < 1 || w < 1 || cin < 64 || cout < 64 || cin % 64 != 0 || cout % 64 != 0 ||
      (reinterpret_cast<uintptr_t>(d_dy) & 15) != 0 || (reinterpret_cast<uintptr_t>(d_x) & 15) != 0 || (reinterpret_cast<uintptr_t>(d_dw) & 15) != 0)
    return fail(S3OD_ERR_ARG, "s3od_op_conv3x3_wgrad_f32 needs cin % 64 == 0, cout % 64 == 0 and 16-byte aligned buffers");
  GemmTnParams p{};
  p.conv = 1;
  p.M = cout; p.N = 9 * cin;
  p.m_tiles = (p.M + 127) / 128; p.n_tiles = (p.N + 255) / 256;
  p.patches_w = (w + 15) / 16; p.patches_h = (h + 3) / 4; p.cin_blocks = cin / 64;
  p.k_blocks = batch * p.patches_h * p.patches_w;
  if (splits < 1) splits = 1;
  if (splits > p.k_blocks) splits = p.k_blocks;
  p.k_blocks_per_split = (p.k_blocks + splits - 1) / splits;
  p.splits = (p.k_blocks + p.k_blocks_per_split - 1) / p.k_blocks_per_split;
  if (p.splits > 1 && (d_workspace == nullptr || (reinterpret_cast<uintptr_t>(d_workspace) & 15) != 0))
    return fail(S3OD_ERR_ARG, "s3od_op_conv3x3_wgrad_f32: splits > 1 need a 16-byte aligned workspace of splits * cout * 9 * cin floats");
  p.out = p.splits > 1 ? d_workspace : d_dw;
  const uint64_t W = w, H = h, B = batch;
  const uint64_t da[5] = {64, W, H, B, (uint64_t)(cout / 64)};
  const uint64_t sa[4] = {(uint64_t)cout * 2, W * cout * 2, H * W * cout * 2, 128};
  const uint32_t ba[5] = {64, 16, 4, 1, 2};
  const uint64_t db[4] = {(uint64_t)cin, W, H, B};
  const uint64_t sb[3] = {(uint64_t)cin * 2, W * cin * 2, H * W * cin * 2};
  const uint32_t bb[4] = {64, 16, 4, 1};
  if (!make_tmap(&p.tma_a, d_dy, 5, da, sa, ba) || !make_tmap(&p.tma_b, d_x, 4, db, sb, bb)) return S3OD_ERR_CUDA;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_gemm_tn(p, sms, st));
  if (p.splits > 1) {
    const long long n4 = static_cast<long long>(p.M) * p.N / 4;
    const int grid = static_cast<int>(std::min<long long>((n4 + 255) / 256, 148LL * 8));
    sum_k_splits_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(d_workspace), p.splits, n4, reinterpret_cast<float4*>(d_dw));
    CK(cudaGetLastError());
  }
  return S3OD_OK;
}

int s3od_op_layernorm(const float* d_x, const float* d_w, const float* d_b, void* d_y, int M, int D, float eps, s3od_stream stream) {
  CK(launch_layernorm(const_cast<float*>(d_x), nullptr, d_w, d_b, static_cast<bf16*>(d_y), nullptr, M, M, D, eps,
                      static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_op_attention(const void* d_q, const void* d_k, const void* d_v, void* d_out, int batch, int heads, int ntok,
                      s3od_stream stream) {
  if (batch < 1 || heads < 1 || ntok < 1) return fail(S3OD_ERR_ARG, "bad shape for s3od_op_attention");
  AttnParams ap{};
  const uint64_t BH = static_cast<uint64_t>(batch) * heads;
  const uint64_t dq[3] = {64, (uint64_t)ntok, BH};
  const uint64_t sq[2] = {128, (uint64_t)ntok * 128};
  const uint32_t bq[3] = {64, kAttnTile, 1}, bkv[3] = {64, kAttnKvTile, 1};
  if (!make_tmap(&ap.tma_q, d_q, 3, dq, sq, bq)) return S3OD_ERR_CUDA;
  if (!make_tmap(&ap.tma_k, d_k, 3, dq, sq, bkv)) return S3OD_ERR_CUDA;
  if (!make_tmap(&ap.tma_v, d_v, 3, dq, sq, bkv)) return S3OD_ERR_CUDA;
  ap.out = static_cast<bf16*>(d_out);
  ap.ntok = ntok; ap.heads = heads; ap.kv_tiles = (ntok + kAttnKvTile - 1) / kAttnKvTile;
  CK(launch_attention(ap, (ntok + kAttnTile - 1) / kAttnTile, static_cast<int>(BH), static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_train_attention_forward(const void* d_q, const void* d_k, const void* d_v, void* d_out, float* d_lse, int batch, int heads, int ntok,
                                 int ntok_padded, s3od_stream stream) {
  if (d_q == nullptr || d_k == nullptr || d_v == nullptr || d_out == nullptr || d_lse == nullptr || batch < 1 || heads < 1 || ntok < 1 ||
      ntok_padded < ntok || ntok_padded % 384 != 0)
    return fail(S3OD_ERR_ARG, "bad argument to s3od_train_attention_forward (ntok_padded must be a multiple of 384)");
  AttnParams ap{};
  const uint64_t BH = static_cast<uint64_t>(batch) * heads;
  const uint64_t dq[3] = {64, (uint64_t)ntok, BH};                 // rows >= ntok of a box are zero-filled by the TMA unit
  const uint64_t sq[2] = {128, (uint64_t)ntok_padded * 128};
  const uint32_t bq[3] = {64, kAttnTile, 1}, bkv[3] = {64, kAttnKvTile, 1};
  if (!make_tmap(&ap.tma_q, d_q, 3, dq, sq, bq)) return S3OD_ERR_CUDA;
  if (!make_tmap(&ap.tma_k, d_k, 3, dq, sq, bkv)) return S3OD_ERR_CUDA;
  if (!make_tmap(&ap.tma_v, d_v, 3, dq, sq, bkv)) return S3OD_ERR_CUDA;
  ap.out = static_cast<bf16*>(d_out);
  ap.ntok = ntok; ap.heads = heads; ap.kv_tiles = (ntok + kAttnKvTile - 1) / kAttnKvTile;
  ap.lse = d_lse; ap.lse_stride = ntok_padded;
  CK(launch_attention(ap, (ntok + kAttnTile - 1) / kAttnTile, static_cast<int>(BH), static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_train_attention_backward(const void* d_q, const void* d_k, const void* d_v, const void* d_dout, const float* d_lse, const float* d_delta,
                                  float* d_dq, float* d_dk, float* d_dv, int batch, int heads, int ntok_padded, s3od_stream stream) {
  if (d_q == nullptr || d_k == nullptr || d_v == nullptr || d_dout == nullptr || d_lse == nullptr || d_delta == nullptr || d_dq == nullptr ||
      d_dk == nullptr || d_dv == nullptr || batch < 1 || heads < 1 || ntok_padded < 384 || ntok_padded % 384 != 0)
    return fail(S3OD_ERR_ARG, "bad argument to s3od_train_attention_backward (ntok_padded must be a multiple of 384)");
  const uint64_t BH = static_cast<uint64_t>(batch) * heads;
  const uint64_t dq[3] = {64, (uint64_t)ntok_padded, BH};
  const uint64_t sq[2] = {128, (uint64_t)ntok_padded * 128};
  const uint32_t b128[3] = {64, kAttnTile, 1}, b96[3] = {64, kAttnKvTile, 1};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AttnBwdParams p{};
  p.lse = d_lse; p.delta = d_delta; p.npad = ntok_padded; p.scale_ds = 1.0f;
  // dQ: the CTA owns 128 query rows and walks the keys
  if (!make_tmap(&p.tma_x, d_q, 3, dq, sq, b128) || !make_tmap(&p.tma_y, d_dout, 3, dq, sq, b128) || !make_tmap(&p.tma_u, d_k, 3, dq, sq, b96) ||
      !make_tmap(&p.tma_w, d_v, 3, dq, sq, b96))
    return S3OD_ERR_CUDA;
  p.out_ds = d_dq; p.out_p = nullptr;
  CK(launch_attention_backward(p, false, static_cast<int>(BH), st));
  // dK, dV: the CTA owns 128 key rows and walks the queries (the transposed problem)
  if (!make_tmap(&p.tma_x, d_k, 3, dq, sq, b128) || !make_tmap(&p.tma_y, d_v, 3, dq, sq, b128) || !make_tmap(&p.tma_u, d_q, 3, dq, sq, b96) ||
      !make_tmap(&p.tma_w, d_dout, 3, dq, sq, b96))
    return S3OD_ERR_CUDA;
  p.out_ds = d_dk; p.out_p = d_dv;
  CK(launch_attention_backward(p, true, static_cast<int>(BH), st));
  return S3OD_OK;
}

int s3od_op_conv3x3(const void* d_in, const void* d_w, const float* d_bias, void* d_out, int batch, int h, int w, int cin, int cout,
                    int relu, s3od_stream stream) {
  if (cin % 64 != 0 || cout % 128 != 0) return fail(S3OD_ERR_ARG, "s3od_op_conv3x3 needs cin % 64 == 0 and cout % 128 == 0");
  const bool wide = cout % 256 == 0;                 // 256-wide tiles, else the 128-wide configuration (the mask head's 256 -> 128)
  GemmParams<EpiConv> p{};
  if (!tmap_nhwc(&p.tma_a, d_in, batch, h, w, cin)) return S3OD_ERR_CUDA;
  if (!tmap_matrix(&p.tma_b, d_w, cout, 9 * cin, wide ? b_box_rows<256>() : b_box_rows<128>())) return S3OD_ERR_CUDA;
  p.geom = geom_3x3(h, w, cin);
  p.m_tiles = batch * p.geom.tiles_h * p.geom.tiles_w;
  p.n_tiles = cout / (wide ? 256 : 128);
  p.num_k_blocks = 9 * cin / 64;
  p.epi = conv_epi(static_cast<bf16*>(d_out), nullptr, d_bias, nullptr, nullptr, relu, cout, h, w);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (wide) CK((launch_gemm<256, A_CONV, EpiConv, 8>(p, sms, static_cast<cudaStream_t>(stream))));
  else CK((launch_gemm<128, A_CONV, EpiConv, 8>(p, sms, static_cast<cudaStream_t>(stream))));
  return S3OD_OK;
}

int s3od_op_conv3x3_rows(const void* d_in, const void* d_w, const float* d_bias, void* d_out, int batch, int h, int w, int relu,
                         s3od_stream stream) {
  if (w % kRowPx != 0 || batch < 1 || h < 1) return fail(S3OD_ERR_ARG, "s3od_op_conv3x3_rows needs w % 128 == 0");
  RowConvParams<EpiConv> p{};
  if (!tmap_nhwc_row(&p.tma_in, d_in, batch, h, w, 64)) return S3OD_ERR_CUDA;
  if (!tmap_matrix(&p.tma_w, d_w, 64, 9 * 64, 64)) return S3OD_ERR_CUDA;
  p.H = h; p.W = w;
  p.strips_x = w / kRowPx;
  p.strips_y = (h + kRowsPerStrip - 1) / kRowsPerStrip;
  p.num_strips = batch * p.strips_x * p.strips_y;
  p.epi = conv_epi(static_cast<bf16*>(d_out), nullptr, d_bias, nullptr, nullptr, relu, 64, h, w);
  if (!tmap_nhwc_row_out(&p.tma_out, d_out, batch, h, w, 64)) return S3OD_ERR_CUDA;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK((launch_conv_rows<64, EpiConv>(p, sms, static_cast<cudaStream_t>(stream))));
  return S3OD_OK;
}

int s3od_op_convt_rows(const void* d_in, const void* d_wr, const float* d_bias, void* d_out, int batch, int h, int w, int relu,
                       s3od_stream stream) {
  if (w % kRowPx != 0 || batch < 1 || h < 1) return fail(S3OD_ERR_ARG, "s3od_op_convt_rows needs w % 128 == 0");
  ConvTRowParams p{};
  if (!make_convt_rows(&p, d_in, d_wr, d_bias, d_out, batch, h, w, relu)) return S3OD_ERR_CUDA;
  p.num_strips = batch * p.strips_x * p.strips_y;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int b = 0; b < 2; ++b) {
    p.phase_b = b;
    CK(launch_convt_rows(p, sms, static_cast<cudaStream_t>(stream)));
  }
  return S3OD_OK;
}

// ---- saliency metrics: device reductions behind EvaluationMetrics.step (metrics.py:213-421)
int s3od_metrics_stats(const float* d_pred, const float* d_mask, int h, int w, const float* d_thresholds, void* d_stats, size_t stats_bytes,
                       s3od_stream stream) {
  if (d_pred == nullptr || d_mask == nullptr || d_thresholds == nullptr || d_stats == nullptr || h < 0 || w < 0 ||
      stats_bytes < sod_stats_bytes())
    return fail(S3OD_ERR_ARG, "bad argument for s3od_metrics_stats (stats buffer needs s3od_metrics_stats_bytes() bytes)");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_sod_stats(d_pred, d_mask, h, w, d_thresholds, d_stats, sms, static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_metrics_region(const float* d_pred, const float* d_mask, int h, int w, int x_split, int y_split, void* d_region,
                        size_t region_bytes, s3od_stream stream) {
  if (d_pred == nullptr || d_mask == nullptr || d_region == nullptr || h < 0 || w < 0 || region_bytes < sod_region_bytes())
    return fail(S3OD_ERR_ARG, "bad argument for s3od_metrics_region");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_sod_region(d_pred, d_mask, h, w, x_split, y_split, d_region, sms, static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

size_t s3od_metrics_weighted_f_workspace_bytes(int h, int w) { return h < 0 || w < 0 ? 0 : sod_wfm_workspace_bytes(h, w); }

int s3od_metrics_weighted_f(const float* d_pred, const float* d_mask, int h, int w, void* d_workspace, size_t workspace_bytes, void* d_sums,
                            s3od_stream stream) {
  if (d_pred == nullptr || d_mask == nullptr || d_workspace == nullptr || d_sums == nullptr || h < 1 || w < 1 || w > 12000 ||
      workspace_bytes < sod_wfm_workspace_bytes(h, w))
    return fail(S3OD_ERR_ARG, "bad argument for s3od_metrics_weighted_f (workspace needs s3od_metrics_weighted_f_workspace_bytes(h, w) bytes)");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_sod_wfm(d_pred, d_mask, h, w, d_workspace, d_sums, sms, static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

size_t s3od_metrics_stats_bytes(void) { return sod_stats_bytes(); }
size_t s3od_metrics_region_bytes(void) { return sod_region_bytes(); }

// ---- visualisation: device-resident composites of visualizer.py and the pair counts behind is_ambiguous
int s3od_threshold_f32(const float* d_in, float* d_out, size_t n, float threshold, s3od_stream stream) {
  if ((d_in == nullptr || d_out == nullptr) && n > 0) return fail(S3OD_ERR_ARG, "bad argument for s3od_threshold_f32");
  if ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15)
    return fail(S3OD_ERR_ARG, "s3od_threshold_f32 needs 16-byte aligned buffers");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_threshold(d_in, d_out, n, threshold, sms, static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_vis_composite(const uint8_t* d_image, const float* d_mask, uint8_t* d_out, int h, int w, int bg_r, int bg_g, int bg_b,
                       s3od_stream stream) {
  if (d_image == nullptr || d_mask == nullptr || d_out == nullptr || h < 0 || w < 0) return fail(S3OD_ERR_ARG, "bad argument for s3od_vis_composite");
  CK(launch_composite(d_image, d_mask, d_out, static_cast<size_t>(h) * w, static_cast<float>(bg_r & 255), static_cast<float>(bg_g & 255),
                      static_cast<float>(bg_b & 255), static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_vis_mask_grid(const uint8_t* d_image, const float* d_masks, int num_masks, uint8_t* d_out, int h, int w, s3od_stream stream) {
  if (d_image == nullptr || d_masks == nullptr || d_out == nullptr || num_masks < 1 || h < 0 || w < 0)
    return fail(S3OD_ERR_ARG, "bad argument for s3od_vis_mask_grid");
  const int gw = num_masks < 4 ? num_masks : 4;
  CK(launch_mask_grid(d_image, d_masks, d_out, num_masks, h, w, gw, static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

int s3od_mask_pair_counts(const float* d_masks, int num_masks, int h, int w, unsigned long long* d_counts, s3od_stream stream) {
  if (d_masks == nullptr || d_counts == nullptr || num_masks < 1 || num_masks > 4 || h < 0 || w < 0)
    return fail(S3OD_ERR_ARG, "s3od_mask_pair_counts takes 1..4 masks");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_mask_pair_counts(d_masks, num_masks, static_cast<size_t>(h) * w, d_counts, sms, static_cast<cudaStream_t>(stream)));
  return S3OD_OK;
}

}  // extern "C"
