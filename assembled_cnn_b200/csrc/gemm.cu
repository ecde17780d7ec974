// tcgen05 implicit-GEMM convolutions for sm_100a.
//
//   conv_gemm_kernel : y[M, Cout] = im2col(x)[M, K] * w[Cout, K]^T          (fprop, dgrad, dense)
//                      A tiles arrive by TMA (tiled 2-D for 1x1/s1, im2col 4-D otherwise) in the
//                      128B/64B/32B-swizzled K-major layout tcgen05.mma reads directly; the fp32
//                      accumulator lives in TMEM; a 4-warp epilogue drains it with tcgen05.ld and
//                      fuses bias / gradient-accumulate / ReLU-mask / batch-norm column sums.
//   wgrad_gemm_kernel: dw[Cout, K] += dy[P, Cout]^T * im2col(x)[P, K]       (split-K over pixels)
//                      both operands are MN-major (the pixel index is the GEMM K dimension).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp_idx % 4).
//
// Reference semantics: nets/model_helper.py:67-78 (conv2d_fixed_padding), tf.gradients backward.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace acnn {

// ------------------------------------------------------------------------------------------
// Tensor-map creation (driver entry points fetched lazily: libacnn.so does not link libcuda).
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;
static int g_driver_version = 0;

static int load_driver_fns() {
  if (g_encode_tiled && g_encode_im2col) return ACNN_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return ACNN_ERR_CUDA;
  }
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) {
    set_error("cuTensorMapEncodeIm2col entry point unavailable (%s)", cudaGetErrorString(e));
    return ACNN_ERR_CUDA;
  }
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  cudaDriverGetVersion(&g_driver_version);
  return ACNN_OK;
}

static CUtensorMapSwizzle swizzle_enum(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                      : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// bf16 matrix [rows][cols] (cols contiguous, row pitch ld elements); box = [box_rows][box_cols].
static int make_map_2d(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld,
                       int box_rows, int box_cols) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                              dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_enum(box_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rows=%lld cols=%lld ld=%lld box=%dx%d", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_rows, box_cols);
    return ACNN_ERR_CUDA;
  }
  return ACNN_OK;
}

// Element pitches of the input tensor: dense NHWC unless the geometry overrides them (used by the
// space-to-depth stem, whose "pixels" are overlapping 4-pixel windows of a W-padded image).
static void input_pitches(const acnn_conv_geom& g, int64_t* pix, int64_t* row, int64_t* img) {
  *pix = g.x_pix_stride > 0 ? g.x_pix_stride : g.Cin;
  *row = g.x_row_pitch > 0 ? g.x_row_pitch : (int64_t)g.W * *pix;
  *img = g.x_img_pitch > 0 ? g.x_img_pitch : (int64_t)g.H * *row;
}

static bool is_plain(const acnn_conv_geom& g) {
  return g.kh == 1 && g.kw == 1 && g.stride == 1 && g.pad_h_lo == 0 && g.pad_w_lo == 0 &&
         g.pad_h_hi == 0 && g.pad_w_hi == 0 && g.x_pix_stride <= 0 && g.x_row_pitch <= 0 &&
         g.x_img_pitch <= 0;
}

// bf16 NHWC tensor [B][H][W][C]; one load = `pixels` consecutive output pixels x `cw` channels of
// one filter tap.  The bounding box of base pixels is [-pad_lo, dim + pad_hi - (k-1)).
static int make_map_im2col(CUtensorMap* m, const void* base, const acnn_conv_geom& g, int cw,
                           int pixels) {
  int64_t pix, row, img;
  input_pitches(g, &pix, &row, &img);
  cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.B};
  cuuint64_t strides[3] = {(cuuint64_t)pix * 2, (cuuint64_t)row * 2, (cuuint64_t)img * 2};
  int lower[2] = {-g.pad_w_lo, -g.pad_h_lo};
  int upper[2] = {g.pad_w_hi - (g.kw - 1), g.pad_h_hi - (g.kh - 1)};
  cuuint32_t estr[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  CUresult r = g_encode_im2col(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base),
                               dims, strides, lower, upper, (cuuint32_t)cw, (cuuint32_t)pixels,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(cw * 2),
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d): B=%d H=%d W=%d C=%d k=%dx%d s=%d cw=%d px=%d "
              "pitches=%lld/%lld/%lld", (int)r, g.B, g.H, g.W, g.Cin, g.kh, g.kw, g.stride, cw,
              pixels, (long long)pix, (long long)row, (long long)img);
    return ACNN_ERR_CUDA;
  }
  // Driver <= 13.1 mis-encodes im2col maps of tensors smaller than 128 KiB (same fix-up the
  // CUTLASS im2col descriptor builder applies): clear bit 21 of the second descriptor word.
  if (g_driver_version <= 13010 && (int64_t)g.B * img * 2 < 131072) {
    reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  return ACNN_OK;
}

// ------------------------------------------------------------------------------------------
// fprop / dgrad / dense kernel (persistent, warp-specialised)
// ------------------------------------------------------------------------------------------
struct ConvGemmParams {
  int M;          // output pixels B*Ho*Wo
  int Cout;       // GEMM N
  int Cin;        // channels per filter tap
  int Ktot;       // kh*kw*Cin
  int kw;         // taps per filter row
  int HoWo, Wo;   // to decompose a row index into (n, p, q)
  int stride, pad_h_lo, pad_w_lo;
  int b_sw_bytes; // swizzle span of the weight tile (128 unless Ktot < 64)
  int stages;     // smem pipeline depth
  int m_tiles, n_tiles;
  void* y;
  float* ch_part;   // [gridDim.x / n_tiles][2][Cout] per-CTA partial (sum, sum of squares) rows
  const float* bias;
  int has_add, has_mask;
  int out_f32;
  int out_bufs;     // 1 or 2 output staging half tiles (2: a half's TMA store drains under the next)
  int split_epi;    // two independent 4-warp epilogue groups alternate tiles (implies out_bufs = 2
                    // and two sets of add / mask staging tiles)
};

constexpr int kBM = 128;        // output pixels per CTA tile == UMMA M == TMEM lanes
constexpr int kStageK = 64;     // K elements per pipeline stage
constexpr int kThreads = 192;          // wgrad kernel: TMA warp + MMA warp + 4 epilogue warps
constexpr int kConvThreads = 320;      // conv kernel: TMA warp + MMA warp + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 224 * 1024;   // dynamic smem (227 KiB max per CTA, ~0.2 KiB static)

// MT = 128-row M tiles per CTA tile (1 or 2).  With MT = 2 a CTA computes two output tiles that
// share every weight stage: twice the tensor work per pipeline round-trip (the single-warp issue
// loops, not bandwidth, bound the N <= 128 layers) and half the weight traffic per FLOP.
// (a-plane, b-plane) of the t-th cross product of two 3-plane operands, smallest magnitude first:
// (2,0) (1,1) (0,2) [2^-16]  (1,0) (0,1) [2^-8]  (0,0)
__host__ __device__ constexpr int plane_term_a(int t) { return t == 0 ? 2 : ((t == 1 || t == 3) ? 1 : 0); }
__host__ __device__ constexpr int plane_term_b(int t) { return t == 2 ? 2 : ((t == 1 || t == 4) ? 1 : 0); }

// NP = operand planes: 1 = bf16 operands; 3 = fp32 operands split into three bf16 planes
// (x = hi + mid + lo, 24 mantissa bits) whose six significant cross products are accumulated in the
// same fp32 TMEM accumulator -- the fp32 parity mode (acnn.h ACNN_F32) on the same TMA / im2col /
// descriptor / epilogue code as the bf16 path.
static inline int fprop_stage_bytes(int bn, int mt, int np, bool pair = false) {
  return np * (mt * kBM * kStageK * 2 + (pair ? bn / 2 : bn) * kStageK * 2);
}
static inline int fprop_stages(int bn, int mt, int np, bool has_add, bool has_mask, bool out_f32,
                               bool pair = false, int out_bufs = 1, int aux_sets = 1) {
  const int half_n = bn > 128 ? 128 : bn;
  const int tile = kBM * half_n * 2;
  const int fixed = 1024 + (out_f32 ? 0 : out_bufs * tile) + (has_add ? aux_sets * tile : 0) +
                    (has_mask ? aux_sets * tile : 0);
  int st = (kSmemBudget - fixed) / fprop_stage_bytes(bn, mt, np, pair);
  if (st > kMaxStages) st = kMaxStages;
  if (st < 2) st = 2;
  return st;
}

// CG2: a CTA PAIR (tcgen05 cta_group::2) computes one 256 x BN tile: each CTA stages its own 128 rows
// of A and only HALF of the weight tile, the leader issues one M = 256 MMA that reads both halves.
// Shared-memory ingest per SM per k-block drops from A + B to A + B/2 -- what bounds the N = 256
// tiles.
template <int BN, int MT, int NP = 1, bool CG2 = false>
struct FpropCfg {
  static constexpr int kAHalfBytes = kBM * kStageK * 2;   // 16 KiB per M tile
  static constexpr int kABytes = MT * kAHalfBytes;        // one plane
  static constexpr int kBRows = CG2 ? BN / 2 : BN;        // weight rows staged by this CTA
  static constexpr int kBBytes = kBRows * kStageK * 2;    // one plane
  static constexpr int kStageBytes = NP * (kABytes + kBBytes);
  static_assert(2 * kStageBytes + 1024 <= kSmemBudget, "two pipeline stages must fit");
  // the epilogue handles the accumulator in column halves of <= 128 (one staging buffer each for
  // the output, add and mask tiles), so a 128 x 256 tile needs no more staging than 128 x 128
  static constexpr int kHalfN = BN > 128 ? 128 : BN;
  static constexpr int kNHalf = BN / kHalfN;
  static constexpr int kSubW = kHalfN < 64 ? kHalfN : 64;  // staging sub-tile width (TMA box)
  static constexpr int kTileBytes = kBM * kHalfN * 2;      // one staged half tile (bf16)
  // accumulator columns of one stage: MT tiles of BN columns; NP == 3 keeps TWO accumulators per
  // tile -- the hi*hi products and the five small cross terms separately (added in the epilogue) --
  // because the tensor core truncates an addend below the accumulator's ulp: small terms summed
  // into the big accumulator lose ~6 bits each (measured 2e-5 relative at K = 4608)
  static constexpr int kAccCols = (NP == 3 ? 2 : MT) * BN;
  static constexpr int kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;   // two accumulator stages
  static_assert(2 * kAccCols <= 512 && (NP == 1 || MT == 1), "TMEM holds 512 columns");
  static_assert(!CG2 || (MT == 1 && NP == 1), "CTA pairs: one M tile per CTA, bf16 operands");
  // smem: [stages x (A|B)] [out staging] [add staging] [mask staging]
  static int stages_for(bool has_add, bool has_mask, bool out_f32, int out_bufs = 1,
                        int aux_sets = 1) {
    return fprop_stages(BN, MT, NP, has_add, has_mask, out_f32, CG2, out_bufs, aux_sets);
  }
  static int smem_bytes(int stages, bool has_add, bool has_mask, bool out_f32, int out_bufs = 1,
                        int aux_sets = 1) {
    return 1024 + stages * kStageBytes + (out_f32 ? 0 : out_bufs * kTileBytes) +
           (has_add ? aux_sets * kTileBytes : 0) + (has_mask ? aux_sets * kTileBytes : 0);
  }
};

// Split epilogue of conv_gemm_kernel (one M tile per CTA tile, bf16 output, no bias): the 8 epilogue
// warps form two independent groups of four (one warp per TMEM lane quarter); group g drains
// accumulator stage g of the tiles g, g + 2, ... with its own output / add / mask staging tiles, named
// barrier (1 + g, 128 threads), aux mbarrier and TMA store queue, so the chain of one tile (barrier ->
// TMEM load -> convert -> stage -> barrier -> store -> statistics, per column half) overlaps the next
// tile's.  That chain -- not the tensor pipe, not HBM -- paces tiles with one or two k-blocks (the
// small-K 1x1 convolutions).  Same arithmetic per element as the joint epilogue; the statistics are
// summed in a different (fixed) order.
// MT == 2 (two M tiles per CTA tile): the two groups drain the two M tiles of EVERY tile concurrently
// (group g = M tile g, both waiting on the same accumulator barrier; the "accumulator empty" barrier
// then counts two arrivals) instead of alternating tiles.
template <int BN, int MT>
__device__ __forceinline__ void conv_epilogue_split(
    const ConvGemmParams& p, const CUtensorMap* tmC, const CUtensorMap* tmAdd,
    const CUtensorMap* tmMask, uint32_t tmem_base, uint8_t* s_out0, uint8_t* s_add, uint8_t* s_mask,
    uint64_t* tfull_bar, uint64_t* tempty_bar, uint64_t* aux_bars, int n0, int m_first, int m_step,
    int my_tiles) {
  using Cfg = FpropCfg<BN, MT, 1, false>;
  constexpr bool kPerMt = MT == 2;                     // groups = M tiles of one CTA tile
  constexpr int kTileM = MT * kBM;
  constexpr int kHalfN = Cfg::kHalfN;
  constexpr int kNHalf = Cfg::kNHalf;
  constexpr int kSubW = Cfg::kSubW;
  constexpr int kRowBytes = kSubW * 2;
  constexpr int kSubBytes = kBM * kRowBytes;
  constexpr int kNSub = kHalfN / kSubW;
  constexpr int kNChunk = kHalfN / 8;                  // 16-byte column chunks of a half tile
  constexpr int kNRg = 128 / kNChunk;                  // row groups inside one epilogue group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;
  const int grp = (warp - 2) >> 2;
  const int r = quarter * 32 + lane;                   // row inside the tile == TMEM lane
  const int gt = static_cast<int>(threadIdx.x) - 64 - grp * 128;   // thread index inside the group
  const bool gleader = gt == 0;
  const bool stats = p.ch_part != nullptr;
  const bool has_aux = p.has_add || p.has_mask;
  const int swz = (kRowBytes == 128) ? (r & 7) : ((r >> 1) & 3);
  uint8_t* so_ = s_out0 + grp * Cfg::kTileBytes;
  uint8_t* sa_ = s_add + grp * Cfg::kTileBytes;
  uint8_t* sm_ = s_mask + grp * Cfg::kTileBytes;
  uint64_t* abar = &aux_bars[grp];
  const int st_chunk = gt % kNChunk, st_rg = gt / kNChunk;
  float acc_s[kNHalf][8], acc_q[kNHalf][8];
#pragma unroll
  for (int hf = 0; hf < kNHalf; ++hf)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc_s[hf][e] = acc_q[hf][e] = 0.f;
  uint32_t aux_n = 0;
  auto group_sync = [&]() {
    if (grp == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
    else asm volatile("bar.sync 2, 128;\n" ::: "memory");
  };
  auto issue_aux = [&](int am0, int anh) {
    mbar_expect_tx(abar, (p.has_add ? Cfg::kTileBytes : 0) + (p.has_mask ? Cfg::kTileBytes : 0));
#pragma unroll
    for (int sub = 0; sub < kNSub; ++sub) {
      if (p.has_add) tma_load_2d(sa_ + sub * kSubBytes, tmAdd, abar, anh + sub * kSubW, am0);
      if (p.has_mask) tma_load_2d(sm_ + sub * kSubBytes, tmMask, abar, anh + sub * kSubW, am0);
    }
  };
  const int it0 = kPerMt ? 0 : grp, it_step = kPerMt ? 1 : 2;
  const int m_off = kPerMt ? grp * kBM : 0;            // this group's M tile inside the CTA tile
  if (gleader && has_aux && it0 < my_tiles) issue_aux((m_first + it0 * m_step) * kTileM + m_off, n0);

  for (int it = it0; it < my_tiles; it += it_step) {
    const int m0 = (m_first + it * m_step) * kTileM + m_off;
    const int acc = kPerMt ? (it & 1) : grp;           // TMEM accumulator stage of this tile
#pragma unroll
    for (int hf = 0; hf < kNHalf; ++hf) {
      const int nh = n0 + hf * kHalfN;
      if (gleader) tma_store_wait_read();              // this group's previous store has read so_
      group_sync();
      if (hf == 0) {
        mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
        tc_fence_after();
      }
      if (has_aux) {
        mbar_wait(abar, aux_n & 1);
        ++aux_n;
      }
#pragma unroll
      for (int c = 0; c < kHalfN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * Cfg::kAccCols +
                          (kPerMt ? grp * BN : 0) + hf * kHalfN + c * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
        const int sub = (c * 32) / kSubW;
        const int j0 = ((c * 32) % kSubW) / 8;
        const int soff = sub * kSubBytes + r * kRowBytes;
        if (p.has_add) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = *reinterpret_cast<const uint4*>(sa_ + soff + (((j0 + q) ^ swz) << 4));
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              bf16x2_add(f[q * 8 + e * 2 + 0], f[q * 8 + e * 2 + 1], w4[e]);
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
        if (p.has_mask) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = *reinterpret_cast<const uint4*>(sm_ + soff + (((j0 + q) ^ swz) << 4));
            pk[q * 4 + 0] &= bf16x2_gt0_mask(u.x);
            pk[q * 4 + 1] &= bf16x2_gt0_mask(u.y);
            pk[q * 4 + 2] &= bf16x2_gt0_mask(u.z);
            pk[q * 4 + 3] &= bf16x2_gt0_mask(u.w);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(so_ + soff + (((j0 + q) ^ swz) << 4)) =
              make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
      }
      if (hf == kNHalf - 1) tc_fence_before();
      fence_proxy_async();
      group_sync();
      if (gleader) {
        if (hf == kNHalf - 1) mbar_arrive(&tempty_bar[acc]);
#pragma unroll
        for (int sub = 0; sub < kNSub; ++sub)
          tma_store_2d(tmC, so_ + sub * kSubBytes, nh + sub * kSubW, m0);
        tma_store_commit();
        if (has_aux) {
          if (hf + 1 < kNHalf) issue_aux(m0, nh + kHalfN);
          else if (it + it_step < my_tiles)
            issue_aux((m_first + (it + it_step) * m_step) * kTileM + m_off, n0);
        }
      }
      if (stats) {
        // column sums of the half tile as stored (bf16-rounded); rows past M were computed from
        // zero-filled operands and contribute zero
        const int sub = st_chunk / (kSubW / 8), jj = st_chunk % (kSubW / 8);
#pragma unroll 4
        for (int rr = st_rg; rr < kBM; rr += kNRg) {
          const int sw = (kRowBytes == 128) ? (rr & 7) : ((rr >> 1) & 3);
          const uint4 u = *reinterpret_cast<const uint4*>(so_ + sub * kSubBytes + rr * kRowBytes +
                                                          ((jj ^ sw) << 4));
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            bf16x2_sum_sq(acc_s[hf][2 * e], acc_q[hf][2 * e], acc_s[hf][2 * e + 1],
                          acc_q[hf][2 * e + 1], w4[e]);
        }
      }
    }
  }
  if (gleader) tma_store_wait_all();
  if (stats && my_tiles > 0) {
    // cross-group / cross-row-group reduction in the (now idle) staging tiles
    float* red_sum = reinterpret_cast<float*>(s_out0);
    float* red_sq = red_sum + 2 * kNRg * BN;
    static_assert(2 * 2 * kNRg * BN * 4 <= 2 * Cfg::kTileBytes, "staging tiles too small for stats");
    asm volatile("bar.sync 3, 256;\n" ::: "memory");    // both groups' stores have drained
#pragma unroll
    for (int hf = 0; hf < kNHalf; ++hf)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red_sum[(grp * kNRg + st_rg) * BN + hf * kHalfN + st_chunk * 8 + e] = acc_s[hf][e];
        red_sq[(grp * kNRg + st_rg) * BN + hf * kHalfN + st_chunk * 8 + e] = acc_q[hf][e];
      }
    asm volatile("bar.sync 3, 256;\n" ::: "memory");
    for (int col = static_cast<int>(threadIdx.x) - 64; col < BN; col += kEpiThreads) {
      float ss = 0.f, qq = 0.f;
      for (int g2 = 0; g2 < 2 * kNRg; ++g2) {
        ss += red_sum[g2 * BN + col];
        qq += red_sq[g2 * BN + col];
      }
      float* row = p.ch_part + static_cast<size_t>(m_first) * 2 * p.Cout;
      row[n0 + col] = ss;
      row[p.Cout + n0 + col] = qq;
    }
  }
}

// One CTA per SM walks output tiles (fixed N tile, M tiles strided by the grid).  The TMA producer
// runs ahead across tiles through the smem ring; the MMA issuer alternates between two TMEM
// accumulators; the 8 epilogue warps (two per TMEM lane quarter, alternating 32-column chunks)
// drain accumulator i while the tensor core fills i^1.
template <int BN, int CW, bool IM2COL, int MT, int NP, bool CG2>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAdd,
                 const __grid_constant__ CUtensorMap tmMask, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmB2, const ConvGemmParams p) {
  using Cfg = FpropCfg<BN, MT, NP, CG2>;
  constexpr int kTileM = (CG2 ? 2 : MT) * kBM;  // output pixels per CTA tile (per CTA pair with CG2)
  constexpr int kChunks = kStageK / CW;        // A chunks (one filter tap each when Cin < 64)
  constexpr int kChunkBytes = kBM * CW * 2;
  constexpr int kKSteps = CW / 16;             // UMMA K = 16 bf16
  constexpr uint32_t kIdesc = make_idesc_bf16_m(CG2 ? 256 : 128, BN, false, false);
  constexpr int kHalfN = Cfg::kHalfN;
  constexpr int kNHalf = Cfg::kNHalf;
  constexpr int kSubW = Cfg::kSubW;
  constexpr int kRowBytes = kSubW * 2;
  constexpr int kSubBytes = kBM * kRowBytes;
  constexpr int kNSub = kHalfN / kSubW;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint64_t aux_bars[2];
  __shared__ uint32_t tmem_base_smem;
  uint64_t& aux_bar = aux_bars[0];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int kStages = p.stages;
  uint8_t* s_out0 = smem + kStages * Cfg::kStageBytes;
  uint8_t* s_add = s_out0 + (p.out_f32 ? 0 : p.out_bufs * Cfg::kTileBytes);
  uint8_t* s_mask = s_add + (p.has_add ? (p.split_epi ? 2 : 1) * Cfg::kTileBytes : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.Ktot + kStageK - 1) / kStageK;
  // static tile schedule: this CTA (CG2: this CTA pair) owns N tile n_tile and M tiles m_first,
  // m_first + m_step, ...; with CG2 the CTA of rank r in its pair computes rows r*128.. of the tile
  const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;
  const int unit = CG2 ? (blockIdx.x >> 1) : blockIdx.x;
  const int units = CG2 ? (gridDim.x >> 1) : gridDim.x;
  const int n_tile = unit % p.n_tiles;
  const int m_first = unit / p.n_tiles;
  const int m_step = units / p.n_tiles;
  const int n0 = n_tile * BN;
  const int my_tiles = m_first < p.m_tiles ? (p.m_tiles - m_first + m_step - 1) / m_step : 0;
  const int m_rank_off = CG2 ? static_cast<int>(cta_rank) * kBM : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (NP == 3) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB1);
      tma_prefetch_desc(&tmB2);
    }
    if (!p.out_f32) tma_prefetch_desc(&tmC);
    if (p.has_add) tma_prefetch_desc(&tmAdd);
    if (p.has_mask) tma_prefetch_desc(&tmMask);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      // CG2: the epilogues of both CTAs release it; split epilogue with two M tiles: both groups
      mbar_init(&tempty_bar[s], (CG2 || (MT == 2 && p.split_epi)) ? 2 : 1);
    }
    mbar_init(&aux_bars[0], 1);
    mbar_init(&aux_bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG2) tmem_alloc_pair<Cfg::kTmemCols>(&tmem_base_smem);
    else tmem_alloc<Cfg::kTmemCols>(&tmem_base_smem);
  }
  tc_fence_before();
  if (CG2) cluster_sync();      // the peer's barriers are initialised before any remote arrive
  else __syncthreads();
  tc_fence_after();
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of
  // the preceding kernel; its outputs are read only after this point
  pdl_entry();
  const uint32_t tmem_base = tmem_base_smem;

  // 32-bit shared-window addresses, computed once: the issue loops below run on one warp each and
  // are instruction-latency bound (every instruction saved per k-block is ~1% of a small-N layer)
  const uint32_t smem_a0 = smem_u32(smem);
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const uint32_t stage_wrap = static_cast<uint32_t>(kStages) * Cfg::kStageBytes;
  // chunks of the last k-block (a partial stage exists only when a filter tap is narrower than
  // a stage, i.e. Cin < 64 with an odd tap count)
  const int tail_chunks = kChunks == 1 ? 1 : ((p.Ktot - (num_kb - 1) * kStageK + CW - 1) / CW);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp stays converged (so the uniform-datapath TMA/MMA instructions need no
    // election loops) and every lane tracks the same loop state; one elected lane issues.
    uint32_t soff = 0, sbar = 0, phase = 0;             // stage byte offset / barrier offset
    const uint32_t b_bytes = Cfg::kBRows * (p.b_sw_bytes < 128 ? p.b_sw_bytes : 128);
    // CG2: both CTAs' loads complete on the LEADER's full barrier (cluster address of rank 0)
    const uint32_t full_bytes = (CG2 ? 2 : 1) * NP * (MT * kChunks * kChunkBytes + b_bytes);
    const uint32_t tail_bytes = (CG2 ? 2 : 1) * NP * (MT * tail_chunks * kChunkBytes + b_bytes);
    const uint32_t full0_lead = CG2 ? mapa_shared(full0, 0) : full0;
    const int nb0 = n0 + (CG2 ? static_cast<int>(cta_rank) * Cfg::kBRows : 0);
    const int Cin = p.Cin, fkw = p.kw;
    for (int it = 0; it < my_tiles; ++it) {
      const int m0 = (m_first + it * m_step) * kTileM + m_rank_off;
      int img[MT], h0[MT], w0[MT];
#pragma unroll
      for (int h = 0; h < MT; ++h) {
        img[h] = h0[h] = w0[h] = 0;
        if (IM2COL) {
          const int mh = m0 + h * kBM;
          img[h] = mh / p.HoWo;
          const int rem = mh - img[h] * p.HoWo;
          const int po = rem / p.Wo;
          const int qo = rem - po * p.Wo;
          h0[h] = po * p.stride - p.pad_h_lo;
          w0[h] = qo * p.stride - p.pad_w_lo;
        }
      }
      // filter tap (tr, ts) and channel offset tc of the next chunk, advanced incrementally
      int tr = 0, ts = 0, tc = 0, k0 = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool last = kb == num_kb - 1;
        const int nch = (kChunks > 1 && last) ? tail_chunks : kChunks;
        mbar_wait_a(empty0 + sbar, phase ^ 1);
        const bool leader = elect_one();
        const uint32_t fb = CG2 ? full0_lead + sbar : full0 + sbar;
        const uint32_t sa = smem_a0 + soff;
        if (leader) {
          if (!CG2 || cta_rank == 0)
            mbar_expect_tx_a(full0 + sbar, (kChunks > 1 && last) ? tail_bytes : full_bytes);
          if (CG2) {
            tma_load_2d_pair(sa + Cfg::kABytes, &tmB, fb, k0, nb0);
          } else {
#pragma unroll
            for (int pl = 0; pl < NP; ++pl)
              tma_load_2d_a(sa + NP * Cfg::kABytes + pl * Cfg::kBBytes,
                            pl == 0 ? &tmB : (pl == 1 ? &tmB1 : &tmB2), fb, k0, n0);
          }
        }
#pragma unroll
        for (int j = 0; j < kChunks; ++j) {
          if (j < nch) {
            if (leader) {
#pragma unroll
              for (int pl = 0; pl < NP; ++pl) {
                const CUtensorMap* mA = pl == 0 ? &tmA : (pl == 1 ? &tmA1 : &tmA2);
#pragma unroll
                for (int h = 0; h < MT; ++h) {
                  const uint32_t dst =
                      sa + pl * Cfg::kABytes + h * Cfg::kAHalfBytes + j * kChunkBytes;
                  if (IM2COL) {
                    if (CG2)
                      tma_load_im2col_4d_pair(dst, mA, fb, tc, w0[h], h0[h], img[h], (uint16_t)ts,
                                              (uint16_t)tr);
                    else
                      tma_load_im2col_4d_a(dst, mA, fb, tc, w0[h], h0[h], img[h], (uint16_t)ts,
                                           (uint16_t)tr);
                  } else {
                    if (CG2) tma_load_2d_pair(dst, mA, fb, k0 + j * CW, m0 + h * kBM);
                    else tma_load_2d_a(dst, mA, fb, k0 + j * CW, m0 + h * kBM);
                  }
                }
              }
            }
            tc += CW;
            if (tc >= Cin) {
              tc = 0;
              if (++ts == fkw) { ts = 0; ++tr; }
            }
          }
        }
        __syncwarp();
        k0 += kStageK;
        soff += Cfg::kStageBytes;
        sbar += 8;
        if (soff == stage_wrap) { soff = 0; sbar = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    uint32_t soff16 = 0, sbar = 0, phase = 0;            // stage offset in descriptor units (16 B)
    const uint32_t stage_wrap16 = stage_wrap >> 4;
    const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    // descriptors of stage 0 / chunk 0; everything else is an add on the 14-bit address field
    const uint64_t a_desc0 = make_smem_desc(smem_a0, 16, 8 * CW * 2, swizzle_layout_type(CW * 2));
    const uint64_t b_desc0 = make_smem_desc(smem_a0 + NP * Cfg::kABytes, 16, 8 * p.b_sw_bytes,
                                            swizzle_layout_type(p.b_sw_bytes));
    // CG2: only the leader CTA of the pair issues MMAs (they span both CTAs' operands and TMEM)
    for (int it = 0; it < ((!CG2 || cta_rank == 0) ? my_tiles : 0); ++it) {
      const uint32_t acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait_a(tempty0 + acc * 8, acc_phase ^ 1);   // epilogue(s) have drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * Cfg::kAccCols;
      for (int kb = 0; kb < num_kb; ++kb) {
        const bool last = kb == num_kb - 1;
        const int nch = (kChunks > 1 && last) ? tail_chunks : kChunks;
        mbar_wait_a(full0 + sbar, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da0 = a_desc0 + soff16;
          const uint64_t db0 = b_desc0 + soff16;
#pragma unroll
          for (int j = 0; j < kChunks; ++j) {
            if (j < nch) {
#pragma unroll
              for (int h = 0; h < MT; ++h) {
#pragma unroll
                for (int ks = 0; ks < kKSteps; ++ks) {
                  // NP == 3: the six cross products of the (hi, mid, lo) planes that are
                  // significant at fp32 precision, smallest first
#pragma unroll
                  for (int t = 0; t < (NP == 3 ? 6 : 1); ++t) {
                    const int pa = NP == 3 ? plane_term_a(t) : 0;
                    const int pb = NP == 3 ? plane_term_b(t) : 0;
                    // NP == 3: t = 5 is hi*hi -> accumulator 0; t < 5 -> the small-term
                    // accumulator at column offset BN (its first MMA of a tile is t = 0)
                    const bool small = NP == 3 && t < 5;
                    const uint32_t d_col = tmem_d + h * BN + (small ? BN : 0);
                    const uint64_t da =
                        da0 + ((pa * Cfg::kABytes + h * Cfg::kAHalfBytes + j * kChunkBytes + ks * 32) >> 4);
                    const uint64_t db = db0 + ((pb * Cfg::kBBytes + (j * kKSteps + ks) * 32) >> 4);
                    const uint32_t accum =
                        (j | ks | (small ? t : 0)) ? 1u : static_cast<uint32_t>(kb != 0);
                    if (CG2) umma_bf16_pair(d_col, da, db, kIdesc, accum);
                    else umma_bf16(d_col, da, db, kIdesc, accum);
                  }
                }
              }
            }
          }
          if (CG2) {      // frees the stage / publishes the accumulator in BOTH CTAs
            umma_commit_pair(empty0 + sbar);
            if (last) umma_commit_pair(tfull0 + acc * 8);
          } else {
            umma_commit_a(empty0 + sbar);
            if (last) umma_commit_a(tfull0 + acc * 8);
          }
        }
        __syncwarp();
        soff16 += Cfg::kStageBytes >> 4;
        sbar += 8;
        if (soff16 == stage_wrap16) { soff16 = 0; sbar = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    if constexpr (NP == 1 && !CG2) {
      if (p.split_epi) {
        conv_epilogue_split<BN, MT>(p, &tmC, &tmAdd, &tmMask, tmem_base, s_out0, s_add, s_mask,
                                    tfull_bar, tempty_bar, aux_bars, n0, m_first, m_step, my_tiles);
        goto epilogue_done;
      }
    }
    {
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
    const int egrp = (warp - 2) >> 2;                  // 0 / 1: even / odd 32-column chunks
    const int r = quarter * 32 + lane;                 // row inside the tile == TMEM lane
    const bool leader = (warp == 2 && lane == 0);
    const bool stats = p.ch_part != nullptr;
    const bool has_aux = p.has_add || p.has_mask;
    const int swz = (kRowBytes == 128) ? (r & 7) : ((r >> 1) & 3);
    // batch-norm statistics: thread t owns the 8 columns of 16-byte chunk `st_chunk` (of every
    // column half) over the rows st_rg, st_rg + kNRg, ... of every tile, read back from the
    // staged bf16 tile
    constexpr int kNChunk = kHalfN / 8;
    constexpr int kStatThreads = (BN == 32) ? 128 : kEpiThreads;
    constexpr int kNRg = kStatThreads / kNChunk;       // row groups; kBM / kNRg rows per thread
    const int st_t = threadIdx.x - 64;
    const bool st_on = st_t < kStatThreads;
    const int st_chunk = st_t % kNChunk, st_rg = st_t / kNChunk;
    float acc_s[kNHalf][8], acc_q[kNHalf][8];
#pragma unroll
    for (int hf = 0; hf < kNHalf; ++hf)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc_s[hf][e] = acc_q[hf][e] = 0.f;
    uint32_t aux_n = 0;                                // completed aux-barrier phases
    // output staging: with two buffers the TMA store of a half tile keeps draining (its shared-
    // memory reads are paced by the memory system accepting the writes) under the next half's TMEM
    // reads instead of stalling all 8 warps -- what bounded the store-heavy small-K tiles
    const bool two_out = p.out_bufs == 2;
    uint32_t ebuf = 0;
    // add / mask half tiles are fetched ONE HALF AHEAD: the loads of the next half are issued as
    // soon as every thread has consumed the current one (they then overlap this half's TMA store,
    // the statistics pass and the next half's TMEM reads instead of stalling all 8 warps for an
    // L2 / HBM round trip per half -- what bounded the small-K dgrad layers)
    auto issue_aux = [&](int am0, int anh) {
      mbar_expect_tx(&aux_bar, (p.has_add ? Cfg::kTileBytes : 0) +
                                   (p.has_mask ? Cfg::kTileBytes : 0));
#pragma unroll
      for (int sub = 0; sub < kNSub; ++sub) {
        if (p.has_add)
          tma_load_2d(s_add + sub * kSubBytes, &tmAdd, &aux_bar, anh + sub * kSubW, am0);
        if (p.has_mask)
          tma_load_2d(s_mask + sub * kSubBytes, &tmMask, &aux_bar, anh + sub * kSubW, am0);
      }
    };
    if (leader && has_aux && my_tiles > 0) issue_aux(m_first * kTileM + m_rank_off, n0);

    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
#pragma unroll
      for (int mh = 0; mh < MT; ++mh) {
      const int m0 = (m_first + it * m_step) * kTileM + mh * kBM + m_rank_off;
      const uint32_t acc_col = acc * Cfg::kAccCols + mh * BN;
      const int row = m0 + r;
      const bool row_ok = row < p.M;
#pragma unroll
      for (int hf = 0; hf < kNHalf; ++hf) {
        const int nh = n0 + hf * kHalfN;               // first output column of this half
        uint8_t* s_out = s_out0 + ebuf * Cfg::kTileBytes;
        if (leader && !p.out_f32) {
          // the TMA store that last used this staging buffer must have finished READING it
          if (two_out) tma_store_wait_read1();
          else tma_store_wait_read();
        }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");   // staging buffers free for everyone
        if (hf == 0 && mh == 0) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
        }
        if (has_aux) {
          mbar_wait(&aux_bar, aux_n & 1);
          ++aux_n;
        }
#pragma unroll
        for (int c2 = 0; c2 < (kHalfN / 32 + 1) / 2; ++c2) {
          const int c = c2 * 2 + egrp;                   // this warp group's chunk
          if (c >= kHalfN / 32) break;                   // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc_col +
                            hf * kHalfN + c * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (NP == 3) {   // + the small-term accumulator
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc_col + BN +
                              hf * kHalfN + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += __uint_as_float(v[i]);
          }
          const int col0 = nh + c * 32;
          const int sub = (c * 32) / kSubW;
          const int j0 = ((c * 32) % kSubW) / 8;
          const int soff = sub * kSubBytes + r * kRowBytes;
          if (p.bias) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] += __ldg(p.bias + col0 + i);
          }
          if (p.has_add) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 u =
                  *reinterpret_cast<const uint4*>(s_add + soff + (((j0 + q) ^ swz) << 4));
              const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                bf16x2_add(f[q * 8 + e * 2 + 0], f[q * 8 + e * 2 + 1], w4[e]);
            }
          }
          // ReLU mask of the destination tensor: 0xffff per bf16 half that is > 0, applied to the
          // packed bf16 output below (zeroing the half == zeroing the fp32 value before rounding)
          if (p.out_f32) {
            if (p.has_mask) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u =
                    *reinterpret_cast<const uint4*>(s_mask + soff + (((j0 + q) ^ swz) << 4));
                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t k2 = bf16x2_gt0_mask(w4[e]);
                  if (!(k2 & 0xffffu)) f[q * 8 + e * 2 + 0] = 0.f;
                  if (!(k2 >> 16)) f[q * 8 + e * 2 + 1] = 0.f;
                }
              }
            }
            if (row_ok) {
              float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) +
                                                      static_cast<size_t>(row) * p.Cout + col0);
#pragma unroll
              for (int q = 0; q < 8; ++q)
                dst[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
            }
          } else {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
            if (p.has_mask) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u =
                    *reinterpret_cast<const uint4*>(s_mask + soff + (((j0 + q) ^ swz) << 4));
                pk[q * 4 + 0] &= bf16x2_gt0_mask(u.x);
                pk[q * 4 + 1] &= bf16x2_gt0_mask(u.y);
                pk[q * 4 + 2] &= bf16x2_gt0_mask(u.z);
                pk[q * 4 + 3] &= bf16x2_gt0_mask(u.w);
              }
            }
            // stage in the TMA swizzle layout: 16-byte piece j of row r at piece j ^ swz
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(s_out + soff + (((j0 + q) ^ swz) << 4)) =
                  make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
        // half drained and staged: (after the last half) hand TMEM back, then store
        if (hf == kNHalf - 1 && mh == MT - 1) tc_fence_before();
        if (!p.out_f32) fence_proxy_async();             // generic smem writes -> async proxy
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        if (leader) {
          if (hf == kNHalf - 1 && mh == MT - 1) {
            if (CG2 && cta_rank != 0)     // the leader's MMA warp waits for both CTAs' epilogues
              mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
            else
              mbar_arrive(&tempty_bar[acc]);
          }
          if (!p.out_f32) {
#pragma unroll
            for (int sub = 0; sub < kNSub; ++sub)
              tma_store_2d(&tmC, s_out + sub * kSubBytes, nh + sub * kSubW, m0);
            tma_store_commit();
          }
          if (has_aux) {
            // the staging tiles were consumed by everyone before the barrier above: fetch the next
            // half's (same rows / next column half, else the next M tile of this CTA)
            if (hf + 1 < kNHalf) issue_aux(m0, nh + kHalfN);
            else if (mh + 1 < MT) issue_aux(m0 + kBM, n0);
            else if (it + 1 < my_tiles)
              issue_aux((m_first + (it + 1) * m_step) * kTileM + m_rank_off, n0);
          }
        }
        if (stats && st_on) {
          // column sums of the half tile as stored (bf16-rounded); rows past M were computed
          // from zero-filled operands and contribute zero.  Overlaps the TMA store (both read).
          const int sub = st_chunk / (kSubW / 8), jj = st_chunk % (kSubW / 8);
#pragma unroll 4
          for (int i = 0; i < kBM / kNRg; ++i) {
            const int rr = st_rg + i * kNRg;
            const int sw = (kRowBytes == 128) ? (rr & 7) : ((rr >> 1) & 3);
            const uint4 u = *reinterpret_cast<const uint4*>(s_out + sub * kSubBytes +
                                                            rr * kRowBytes + ((jj ^ sw) << 4));
            const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              bf16x2_sum_sq(acc_s[hf][2 * e], acc_q[hf][2 * e], acc_s[hf][2 * e + 1],
                            acc_q[hf][2 * e + 1], w4[e]);
          }
        }
        if (two_out) ebuf ^= 1;
      }
      }   // mh
    }
    if (leader && !p.out_f32) tma_store_wait_all();
    if (stats && my_tiles > 0) {
      // final cross-row-group reduction in the (now idle) output staging buffer
      float* red_sum = reinterpret_cast<float*>(s_out0);
      float* red_sq = red_sum + kNRg * BN;
      static_assert(2 * kNRg * BN * 4 <= Cfg::kTileBytes, "staging buffer too small for stats");
      asm volatile("bar.sync 1, 256;\n" ::: "memory");   // the last TMA store has drained
      if (st_on) {
#pragma unroll
        for (int hf = 0; hf < kNHalf; ++hf)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            red_sum[st_rg * BN + hf * kHalfN + st_chunk * 8 + e] = acc_s[hf][e];
            red_sq[st_rg * BN + hf * kHalfN + st_chunk * 8 + e] = acc_q[hf][e];
          }
      }
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      for (int col = st_t; col < BN; col += kEpiThreads) {
        float ss = 0.f, qq = 0.f;
        for (int g2 = 0; g2 < kNRg; ++g2) {
          ss += red_sum[g2 * BN + col];
          qq += red_sq[g2 * BN + col];
        }
        // one partial row per CTA of this N tile, plain stores: bn_finalize sums the rows in a
        // fixed order (deterministic; no pre-zeroed accumulator)
        float* row = p.ch_part +
                     static_cast<size_t>(CG2 ? m_first * 2 + static_cast<int>(cta_rank) : m_first) * 2 *
                         p.Cout;
        row[n0 + col] = ss;
        row[p.Cout + n0 + col] = qq;
      }
    }
    }
  epilogue_done:;
  }
  __syncwarp();
  tc_fence_before();
  if (CG2) cluster_sync();      // neither CTA may leave (or free TMEM) while the pair is in flight
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG2) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 convolution WITHOUT im2col copies ("halo" kernel).
//
// The im2col kernel above fetches every input pixel nine times (once per filter tap) from L2 into
// shared memory; for small-N layers (N <= 128: the stem, the 56x56 SK convs and their dgrads) that
// ingest -- not the tensor pipe, not HBM -- is the bound (147 KB of A per 128 x 64-channel tile).
// Here a CTA tile is a 16 x 8 PATCH of output pixels of one image, and ONE 18 x 10 pixel halo tile
// per 64-channel chunk is loaded by a 4-D tiled TMA box at (h0-1, w0-1) (out-of-image pixels are
// zero-filled: the padding).  The A operand of filter tap (r, s) is the same shared-memory tile read
// through a descriptor that starts (r*10 + s) pixel rows further on, with 10 pixel rows between 8-row
// groups (group g = output row h0+g, 8 pixels): the tensor core's swizzle is a function of the
// absolute shared-memory address, like TMA's, so any row may start a descriptor and the group stride
// need not be a multiple of the swizzle atom (probed on the B200: tools/microtests/halo_desc.cu,
// profiles/r02_microtest_halo_desc.txt).  23 KB of A per tile and chunk instead of 147 KB.
// Weight tiles [BN x CW] per (tap, chunk) stream through their own ring; when the whole N-tile slab
// fits the ring it is loaded once and stays (weights-stationary: the stem layers).
// Epilogue = the im2col kernel's (mask / add one half ahead, bf16 staging, TMA store, statistics),
// with 4-D boxes for the patch and the out-of-image rows excluded from the statistics.
// ------------------------------------------------------------------------------------------
constexpr int kHaloH = 18, kHaloW = 10, kPatchH = 16, kPatchW = 8;
constexpr int kHaloMaxB = 32;      // weight ring slots (barriers)

struct HaloParams {
  int H, W, B;            // output = input spatial size
  int Cin, Cout;
  int n_tiles, ph, pw;    // N tiles; patches per image
  int m_tiles;            // B * ph * pw
  int a_stages, b_slots, stationary;
  int out_bufs;           // 1 or 2 output staging tiles (2: the TMA store of tile i overlaps tile i+1)
  int split_epi;          // N <= 64: two independent 4-warp epilogue groups alternate tiles
  float* ch_part;
  int has_add, has_mask;
};

template <int BN, int CW>
struct HaloCfg {
  static constexpr int kRowB = CW * 2;
  static constexpr int kABytes = kHaloH * kHaloW * kRowB;
  static constexpr int kAStage = (kABytes + 1023) / 1024 * 1024;
  static constexpr int kBTile = BN * kRowB;
  static constexpr int kSubW = BN < 64 ? BN : 64;
  static constexpr int kTileBytes = kBM * BN * 2;
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
  static_assert(BN <= 128 && kBTile % 512 == 0, "halo kernel: N tile <= 128");
};

// Patch coordinates of the tiles first, first + step, ... advanced without divisions (three integer
// divisions per tile sat on the epilogue's critical path).
struct PatchIter {
  int img, row, col;          // image, patch row / column inside the image
  int d_img, d_row, d_col;
  int ph, pw;
  __device__ __forceinline__ PatchIter(int first, int step, int ph_, int pw_) : ph(ph_), pw(pw_) {
    const int tpi = ph * pw;
    img = first / tpi;
    int rem = first - img * tpi;
    row = rem / pw;
    col = rem - row * pw;
    d_img = step / tpi;
    rem = step - d_img * tpi;
    d_row = rem / pw;
    d_col = rem - d_row * pw;
  }
  __device__ __forceinline__ void next() {
    col += d_col;
    if (col >= pw) { col -= pw; ++row; }
    row += d_row;
    if (row >= ph) { row -= ph; ++img; }
    img += d_img;
  }
};

template <int BN, int CW>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmAdd,
                 const __grid_constant__ CUtensorMap tmMask, const HaloParams p) {
  using Cfg = HaloCfg<BN, CW>;
  constexpr int kRowB = Cfg::kRowB;
  constexpr int kKSteps = CW / 16;
  constexpr uint32_t kIdesc = make_idesc_bf16_m(128, BN, false, false);
  constexpr uint32_t kLayout = swizzle_layout_type(kRowB);
  constexpr int kSubW = Cfg::kSubW;
  constexpr int kSubRowB = kSubW * 2;
  constexpr int kSubBytes = kBM * kSubRowB;
  constexpr int kNSub = BN / kSubW;

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t a_full[4], a_empty[4];
  __shared__ uint64_t b_full[kHaloMaxB], b_empty[kHaloMaxB];
  __shared__ uint64_t tfull_bar[2], tempty_bar[2];
  __shared__ uint64_t aux_bars[2];
  __shared__ uint32_t tmem_base_smem;
  uint64_t& aux_bar = aux_bars[0];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int AS = p.a_stages, NB = p.b_slots;
  uint8_t* s_out0 = smem + AS * Cfg::kAStage + NB * Cfg::kBTile;
  uint8_t* s_add = s_out0 + p.out_bufs * Cfg::kTileBytes;
  uint8_t* s_mask = s_add + (p.has_add ? (p.split_epi ? 2 : 1) * Cfg::kTileBytes : 0);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nchunks = p.Cin / CW;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_first = blockIdx.x / p.n_tiles;
  const int m_step = gridDim.x / p.n_tiles;
  const int n0 = n_tile * BN;
  const int my_tiles = m_first < p.m_tiles ? (p.m_tiles - m_first + m_step - 1) / m_step : 0;
  const int tpi = p.ph * p.pw;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (p.has_add) tma_prefetch_desc(&tmAdd);
    if (p.has_mask) tma_prefetch_desc(&tmMask);
    for (int s = 0; s < AS; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 1);
    }
    mbar_init(&aux_bars[0], 1);
    mbar_init(&aux_bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_entry();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t a_base = smem_u32(smem), b_base = a_base + AS * Cfg::kAStage;
  const uint32_t afull0 = smem_u32(a_full), aempty0 = smem_u32(a_empty);
  const uint32_t bfull0 = smem_u32(b_full), bempty0 = smem_u32(b_empty);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // ring positions / phases advance incrementally (no runtime division in the issue loops)
    uint32_t sa = 0, a_ph = 0, sb = 0, b_ph = 0;
    if (p.stationary && my_tiles > 0) {
      // the whole weight slab of this N tile: loaded once, ONE barrier, stays for the CTA's lifetime
      if (elect_one()) {
        mbar_expect_tx_a(bfull0, static_cast<uint32_t>(nchunks) * 9 * Cfg::kBTile);
        for (int kc = 0; kc < nchunks; ++kc)
          for (int t = 0; t < 9; ++t)
            tma_load_2d_a(b_base + (kc * 9 + t) * Cfg::kBTile, &tmB, bfull0, t * p.Cin + kc * CW, n0);
      }
      __syncwarp();
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int tile = m_first + it * m_step;
      const int img = tile / tpi;
      const int rem = tile - img * tpi;
      const int h0 = (rem / p.pw) * kPatchH, w0 = (rem % p.pw) * kPatchW;
      for (int kc = 0; kc < nchunks; ++kc) {
        mbar_wait_a(aempty0 + sa * 8, a_ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx_a(afull0 + sa * 8, Cfg::kABytes);
          tma_load_4d_tile_a(a_base + sa * Cfg::kAStage, &tmA, afull0 + sa * 8, kc * CW, w0 - 1,
                             h0 - 1, img);
        }
        __syncwarp();
        if (++sa == static_cast<uint32_t>(AS)) { sa = 0; a_ph ^= 1; }
        if (!p.stationary) {
          for (int t = 0; t < 9; ++t) {
            mbar_wait_a(bempty0 + sb * 8, b_ph ^ 1);
            if (elect_one()) {
              mbar_expect_tx_a(bfull0 + sb * 8, Cfg::kBTile);
              tma_load_2d_a(b_base + sb * Cfg::kBTile, &tmB, bfull0 + sb * 8, t * p.Cin + kc * CW,
                            n0);
            }
            __syncwarp();
            if (++sb == static_cast<uint32_t>(NB)) { sb = 0; b_ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    uint32_t sa = 0, a_ph = 0, sb = 0, b_ph = 0;
    // descriptors are built once; every tap / k-step / ring slot is an add on the address field
    const uint64_t a_desc0 = make_smem_desc(a_base, 16, kHaloW * kRowB, kLayout);
    const uint64_t b_desc0 = make_smem_desc(b_base, 16, 8 * kRowB, kLayout);
    if (p.stationary && my_tiles > 0) {
      mbar_wait_a(bfull0, 0);
      tc_fence_after();
    }
    for (int it = 0; it < my_tiles; ++it) {
      const uint32_t acc = it & 1;
      mbar_wait_a(tempty0 + acc * 8, ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kc = 0; kc < nchunks; ++kc) {
        mbar_wait_a(afull0 + sa * 8, a_ph);
        tc_fence_after();
        const uint64_t a_st = a_desc0 + ((sa * Cfg::kAStage) >> 4);
        if (p.stationary) {
          if (elect_one()) {
            const uint64_t b_st = b_desc0 + ((static_cast<uint32_t>(kc) * 9 * Cfg::kBTile) >> 4);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              // tap (r, s): the halo tile read (r*10 + s) pixel rows further on; 8-row groups (one
              // output row each) 10 pixel rows apart
#pragma unroll
              for (int ks = 0; ks < kKSteps; ++ks)
                umma_bf16(tmem_d, a_st + ((((t / 3) * kHaloW + (t % 3)) * kRowB + ks * 32) >> 4),
                          b_st + ((t * Cfg::kBTile + ks * 32) >> 4), kIdesc,
                          (t | ks) ? 1u : static_cast<uint32_t>(kc != 0));
            }
            umma_commit_a(aempty0 + sa * 8);
            if (kc == nchunks - 1) umma_commit_a(tfull0 + acc * 8);
          }
          __syncwarp();
        } else {
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            mbar_wait_a(bfull0 + sb * 8, b_ph);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t a_tap = a_st + ((((t / 3) * kHaloW + (t % 3)) * kRowB) >> 4);
              const uint64_t b_tap = b_desc0 + ((sb * Cfg::kBTile) >> 4);
#pragma unroll
              for (int ks = 0; ks < kKSteps; ++ks)
                umma_bf16(tmem_d, a_tap + ((ks * 32) >> 4), b_tap + ((ks * 32) >> 4), kIdesc,
                          (kc | t | ks) ? 1u : 0u);
              umma_commit_a(bempty0 + sb * 8);
              if (t == 8) {
                umma_commit_a(aempty0 + sa * 8);
                if (kc == nchunks - 1) umma_commit_a(tfull0 + acc * 8);
              }
            }
            __syncwarp();
            if (++sb == static_cast<uint32_t>(NB)) { sb = 0; b_ph ^= 1; }
          }
        }
        if (++sa == static_cast<uint32_t>(AS)) { sa = 0; a_ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int quarter = warp & 3;
    const int egrp = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;                 // TMEM lane = patch pixel (r/8, r%8)
    const bool stats = p.ch_part != nullptr;
    const bool has_aux = p.has_add || p.has_mask;
    const int swz = (kSubRowB == 128) ? (r & 7) : ((r >> 1) & 3);
    constexpr int kNChunk = BN / 8;
    float acc_s[8], acc_q[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc_s[e] = acc_q[e] = 0.f;
    uint32_t aux_n = 0;
    // one 32-column chunk of the accumulator: TMEM -> (+add) (x mask) -> bf16 -> swizzled staging
    auto drain_chunk = [&](uint32_t taddr, int c, const uint8_t* sa_, const uint8_t* sm_,
                           uint8_t* so_) {
      uint32_t v[32];
      tmem_ld_32x32(taddr, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
      const int sub = (c * 32) / kSubW;
      const int j0 = ((c * 32) % kSubW) / 8;
      const int soff = sub * kSubBytes + r * kSubRowB;
      if (p.has_add) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 u = *reinterpret_cast<const uint4*>(sa_ + soff + (((j0 + q) ^ swz) << 4));
          const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            bf16x2_add(f[q * 8 + e * 2 + 0], f[q * 8 + e * 2 + 1], w4[e]);
        }
      }
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(f[2 * i], f[2 * i + 1]);
      if (p.has_mask) {        // ReLU mask of the destination tensor, applied to the packed halves
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 u = *reinterpret_cast<const uint4*>(sm_ + soff + (((j0 + q) ^ swz) << 4));
          pk[q * 4 + 0] &= bf16x2_gt0_mask(u.x);
          pk[q * 4 + 1] &= bf16x2_gt0_mask(u.y);
          pk[q * 4 + 2] &= bf16x2_gt0_mask(u.z);
          pk[q * 4 + 3] &= bf16x2_gt0_mask(u.w);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(so_ + soff + (((j0 + q) ^ swz) << 4)) =
            make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
    };
    // column sums of the tile as stored (bf16-rounded): this thread's 16-byte chunk `chunk` over the
    // rows rg, rg + nrg, ...; patch pixels outside the image excluded (they were computed from real
    // neighbours and are clipped by the store)
    auto stat_rows = [&](const uint8_t* so_, int chunk, int rg, int nrg, int h0, int w0) {
      const int sub = chunk / (kSubW / 8), jj = chunk % (kSubW / 8);
#pragma unroll 4
      for (int rr = rg; rr < kBM; rr += nrg) {
        if (h0 + (rr >> 3) >= p.H || w0 + (rr & 7) >= p.W) continue;
        const int sw = (kSubRowB == 128) ? (rr & 7) : ((rr >> 1) & 3);
        const uint4 u = *reinterpret_cast<const uint4*>(so_ + sub * kSubBytes + rr * kSubRowB +
                                                        ((jj ^ sw) << 4));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          bf16x2_sum_sq(acc_s[2 * e], acc_q[2 * e], acc_s[2 * e + 1], acc_q[2 * e + 1], w4[e]);
      }
    };
    auto issue_aux = [&](uint64_t* bar, uint8_t* sa_, uint8_t* sm_, int img, int h0, int w0) {
      mbar_expect_tx(bar, (p.has_add ? Cfg::kTileBytes : 0) + (p.has_mask ? Cfg::kTileBytes : 0));
#pragma unroll
      for (int sub = 0; sub < kNSub; ++sub) {
        if (p.has_add)
          tma_load_4d_tile_a(smem_u32(sa_ + sub * kSubBytes), &tmAdd, smem_u32(bar),
                             n0 + sub * kSubW, w0, h0, img);
        if (p.has_mask)
          tma_load_4d_tile_a(smem_u32(sm_ + sub * kSubBytes), &tmMask, smem_u32(bar),
                             n0 + sub * kSubW, w0, h0, img);
      }
    };

    if (BN <= 64 && p.split_epi) {
      // ---- two independent 4-warp groups: group g drains accumulator g of the tiles g, g + 2, ...
      // with its own staging / aux buffers, named barrier and TMA store queue, so the per-tile chain
      // (barrier -> TMEM load -> convert -> stage -> barrier -> store -> statistics) of one tile
      // overlaps the next tile's: the chain, not the tensor pipe, paced the N <= 64 tiles
      const int gt = threadIdx.x - 64 - egrp * 128;      // thread index inside the group
      const bool gleader = gt == 0;
      uint8_t* so_ = s_out0 + egrp * Cfg::kTileBytes;
      uint8_t* sa_ = s_add + egrp * Cfg::kTileBytes;
      uint8_t* sm_ = s_mask + egrp * Cfg::kTileBytes;
      uint64_t* abar = &aux_bars[egrp];
      constexpr int kNRgS = 128 / kNChunk;               // row groups inside one epilogue group
      const int st_chunk = gt % kNChunk, st_rg = gt / kNChunk;
      PatchIter pi(m_first + egrp * m_step, 2 * m_step, p.ph, p.pw);
      if (gleader && has_aux && egrp < my_tiles)
        issue_aux(abar, sa_, sm_, pi.img, pi.row * kPatchH, pi.col * kPatchW);
      for (int it = egrp; it < my_tiles; it += 2) {
        const int img = pi.img, h0 = pi.row * kPatchH, w0 = pi.col * kPatchW;
        pi.next();
        if (gleader) tma_store_wait_read();              // this group's previous store has read so_
        if (egrp == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
        else asm volatile("bar.sync 2, 128;\n" ::: "memory");
        mbar_wait(&tfull_bar[egrp], (it >> 1) & 1);
        tc_fence_after();
        if (has_aux) {
          mbar_wait(abar, aux_n & 1);
          ++aux_n;
        }
#pragma unroll
        for (int c = 0; c < BN / 32; ++c)
          drain_chunk(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + egrp * BN + c * 32, c,
                      sa_, sm_, so_);
        tc_fence_before();
        fence_proxy_async();
        if (egrp == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");
        else asm volatile("bar.sync 2, 128;\n" ::: "memory");
        if (gleader) {
          mbar_arrive(&tempty_bar[egrp]);
#pragma unroll
          for (int sub = 0; sub < kNSub; ++sub)
            tma_store_4d(&tmC, so_ + sub * kSubBytes, n0 + sub * kSubW, w0, h0, img);
          tma_store_commit();
          if (has_aux && it + 2 < my_tiles)
            issue_aux(abar, sa_, sm_, pi.img, pi.row * kPatchH, pi.col * kPatchW);
        }
        if (stats) stat_rows(so_, st_chunk, st_rg, kNRgS, h0, w0);
      }
      if (gleader) tma_store_wait_all();
      if (stats && my_tiles > 0) {
        // cross-group / cross-row-group reduction in the (now idle) staging buffers
        float* red_sum = reinterpret_cast<float*>(s_out0);
        float* red_sq = red_sum + 2 * kNRgS * BN;
        static_assert(BN > 64 || 2 * 2 * kNRgS * BN * 4 <= 2 * Cfg::kTileBytes,
                      "staging buffers too small for stats");
        asm volatile("bar.sync 3, 256;\n" ::: "memory");  // both groups' stores have drained
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          red_sum[(egrp * kNRgS + st_rg) * BN + st_chunk * 8 + e] = acc_s[e];
          red_sq[(egrp * kNRgS + st_rg) * BN + st_chunk * 8 + e] = acc_q[e];
        }
        asm volatile("bar.sync 3, 256;\n" ::: "memory");
        for (int col = threadIdx.x - 64; col < BN; col += kEpiThreads) {
          float ss = 0.f, qq = 0.f;
          for (int g2 = 0; g2 < 2 * kNRgS; ++g2) {
            ss += red_sum[g2 * BN + col];
            qq += red_sq[g2 * BN + col];
          }
          float* row = p.ch_part + static_cast<size_t>(m_first) * 2 * p.Cout;
          row[n0 + col] = ss;
          row[p.Cout + n0 + col] = qq;
        }
      }
    } else {
      // ---- all 8 warps on one tile at a time (two warps per TMEM lane quarter, alternating chunks)
      const bool leader = (warp == 2 && lane == 0);
      constexpr int kStatThreads = (BN == 32) ? 128 : kEpiThreads;
      constexpr int kNRg = kStatThreads / kNChunk;
      const int st_t = threadIdx.x - 64;
      const bool st_on = st_t < kStatThreads;
      const int st_chunk = st_t % kNChunk, st_rg = st_t / kNChunk;
      PatchIter pi(m_first, m_step, p.ph, p.pw);
      if (leader && has_aux && my_tiles > 0)
        issue_aux(&aux_bar, s_add, s_mask, pi.img, pi.row * kPatchH, pi.col * kPatchW);
      for (int it = 0; it < my_tiles; ++it) {
        const int acc = it & 1;
        const int img = pi.img, h0 = pi.row * kPatchH, w0 = pi.col * kPatchW;
        pi.next();
        // staging buffer free again: with two buffers only the store issued two tiles ago must
        // have finished reading (the previous tile's store drains under this tile's TMEM reads)
        uint8_t* s_out = s_out0 + ((p.out_bufs == 2 && (it & 1)) ? Cfg::kTileBytes : 0);
        if (leader) {
          if (p.out_bufs == 2) tma_store_wait_read1();
          else tma_store_wait_read();
        }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
        tc_fence_after();
        if (has_aux) {
          mbar_wait(&aux_bar, aux_n & 1);
          ++aux_n;
        }
#pragma unroll
        for (int c2 = 0; c2 < (BN / 32 + 1) / 2; ++c2) {
          const int c = c2 * 2 + egrp;
          if (c >= BN / 32) break;                       // warp-uniform
          drain_chunk(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + c * 32, c,
                      s_add, s_mask, s_out);
        }
        tc_fence_before();
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        if (leader) {
          mbar_arrive(&tempty_bar[acc]);
#pragma unroll
          for (int sub = 0; sub < kNSub; ++sub)
            tma_store_4d(&tmC, s_out + sub * kSubBytes, n0 + sub * kSubW, w0, h0, img);
          tma_store_commit();
          if (has_aux && it + 1 < my_tiles)
            issue_aux(&aux_bar, s_add, s_mask, pi.img, pi.row * kPatchH, pi.col * kPatchW);
        }
        if (stats && st_on) stat_rows(s_out, st_chunk, st_rg, kNRg, h0, w0);
      }
      if (leader) tma_store_wait_all();
      if (stats && my_tiles > 0) {
        float* red_sum = reinterpret_cast<float*>(s_out0);
        float* red_sq = red_sum + kNRg * BN;
        static_assert(2 * kNRg * BN * 4 <= Cfg::kTileBytes, "staging buffer too small for stats");
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        if (st_on) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            red_sum[st_rg * BN + st_chunk * 8 + e] = acc_s[e];
            red_sq[st_rg * BN + st_chunk * 8 + e] = acc_q[e];
          }
        }
        asm volatile("bar.sync 1, 256;\n" ::: "memory");
        for (int col = st_t; col < BN; col += kEpiThreads) {
          float ss = 0.f, qq = 0.f;
          for (int g2 = 0; g2 < kNRg; ++g2) {
            ss += red_sum[g2 * BN + col];
            qq += red_sq[g2 * BN + col];
          }
          float* row = p.ch_part + static_cast<size_t>(m_first) * 2 * p.Cout;
          row[n0 + col] = ss;
          row[p.Cout + n0 + col] = qq;
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// wgrad kernel:  D[(tap,ci) = M 128 rows][co = N] += sum over pixels of  x_im2col^T * dy
// ------------------------------------------------------------------------------------------
struct WgradParams {
  int P;           // pixels B*Ho*Wo (GEMM K)
  int Cout;        // GEMM N
  int Cin;
  int Ktot;        // kh*kw*Cin (GEMM M)
  int kw;
  int HoWo, Wo;
  int stride, pad_h_lo, pad_w_lo;
  int pix;               // pixels per stage (64 or 128)
  int stages_total;      // ceil(P / pix)
  int stages_per_split;
  float* dw;
};

constexpr int kWgPix = 64;   // default pixels per pipeline stage (4 UMMA K-steps); PIX = 128: 8 steps

// MT = 128-row blocks of (tap,ci) per CTA (1 or 2): with MT = 2 two accumulators share every dy
// stage, i.e. twice the tensor work per pipeline round-trip and half the dy traffic per FLOP.
// PIX = pixels (GEMM K) per pipeline stage: 64, or 128 (twice the MMAs per TMA / barrier round trip
// of the single-warp issue loops, which bound the small-N layers)
template <int BN, int MT, int NP = 1, int PIX = kWgPix>
struct WgradCfg {
  static constexpr int kAHalfBytes = PIX * 128 * 2;   // PIX pixels x 128 (tap,ci) columns
  static constexpr int kABytes = MT * kAHalfBytes;       // one plane
  static constexpr int kBBytes = PIX * BN * 2;    // PIX pixels x BN output channels, one plane
  static constexpr int kStageBytes = NP * (kABytes + kBBytes);
  static constexpr int kStagesFit = (kSmemBudget - 1024) / kStageBytes;
  static constexpr int kStages =
      PIX != kWgPix ? (kStagesFit > 4 ? 4 : kStagesFit)
      : NP == 3 ? (kStagesFit > 3 ? 3 : kStagesFit)
              : (MT == 1 ? ((BN >= 256) ? 4 : ((BN == 128) ? 3 : 4))
                         : (kStagesFit > 5 ? 5 : kStagesFit));
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024;
  // NP == 3: a second accumulator (columns BN..2BN) for the small cross terms, see FpropCfg
  static constexpr int kAccCols = (NP == 3 ? 2 : MT) * BN;
  static constexpr int kTmemCols = kAccCols < 32 ? 32 : kAccCols;
  static_assert(kAccCols <= 512 && (NP == 1 || MT == 1) && kStages >= (NP == 3 ? 2 : 3),
                "wgrad tile does not fit");
};

// CW: channel width of one im2col chunk of x (16/32/64), CWB: channel width of one dy chunk.
template <int BN, int CW, int CWB, bool IM2COL, int MT, int NP, int PIX>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                  const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                  const __grid_constant__ CUtensorMap tmDY1,
                  const __grid_constant__ CUtensorMap tmDY2, const WgradParams p) {
  using Cfg = WgradCfg<BN, MT, NP, PIX>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kAChunks = MT * 128 / CW;        // chunks of both 128-row blocks, consecutive
  constexpr int kBChunks = BN / CWB;
  constexpr int kAChunkBytes = PIX * CW * 2;
  constexpr int kBChunkBytes = PIX * CWB * 2;
  constexpr uint32_t kIdesc = make_idesc_bf16(BN, true, true);

  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[kStages];
  __shared__ uint64_t empty_bar[kStages];
  __shared__ uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * (MT * 128);   // first (tap,ci) column of dw handled here
  const bool second = MT == 2 && m0 + 128 < p.Ktot;   // the second 128-row block exists
  const int co0 = blockIdx.y * BN;
  const int ks_begin = blockIdx.z * p.stages_per_split;
  int ks_end = ks_begin + p.stages_per_split;
  if (ks_end > p.stages_total) ks_end = p.stages_total;
  const int num_ks = ks_end - ks_begin;   // >= 1 by construction of the grid

  int a_chunks = 0;
#pragma unroll
  for (int j = 0; j < kAChunks; ++j) a_chunks += (m0 + j * CW < p.Ktot) ? 1 : 0;
  int b_chunks = (p.Cout - co0 < BN ? p.Cout - co0 : BN) / CWB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the tail of
  // the preceding kernel; its outputs are read only after this point
  pdl_entry();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // whole warp converged, one elected lane issues (no election loops around TMA instructions)
    int stage = 0;
    uint32_t phase = 0;
    // (tap r, tap s, channel) of each A chunk of this CTA: fixed for the whole kernel
    int ch_r[kAChunks], ch_s[kAChunks], ch_c[kAChunks];
#pragma unroll
    for (int j = 0; j < kAChunks; ++j) {
      const int n = m0 + j * CW;
      const int tap = n / p.Cin;
      ch_c[j] = n - tap * p.Cin;
      ch_r[j] = tap / p.kw;
      ch_s[j] = tap - ch_r[j] * p.kw;
    }
    for (int it = 0; it < num_ks; ++it) {
      const int p0 = (ks_begin + it) * PIX;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + NP * Cfg::kABytes;
        mbar_expect_tx(&full_bar[stage], NP * (a_chunks * kAChunkBytes + b_chunks * kBChunkBytes));
        int img = 0, h0 = 0, w0 = 0;
        if (IM2COL) {
          img = p0 / p.HoWo;
          const int rem = p0 - img * p.HoWo;
          const int po = rem / p.Wo;
          const int qo = rem - po * p.Wo;
          h0 = po * p.stride - p.pad_h_lo;
          w0 = qo * p.stride - p.pad_w_lo;
        }
#pragma unroll
        for (int pl = 0; pl < NP; ++pl) {
          const CUtensorMap* mX = pl == 0 ? &tmX : (pl == 1 ? &tmX1 : &tmX2);
          const CUtensorMap* mDY = pl == 0 ? &tmDY : (pl == 1 ? &tmDY1 : &tmDY2);
#pragma unroll
          for (int j = 0; j < kAChunks; ++j) {
            if (j < a_chunks) {
              uint8_t* dst = sa + pl * Cfg::kABytes + j * kAChunkBytes;
              if (IM2COL) {
                tma_load_im2col_4d(dst, mX, &full_bar[stage], ch_c[j], w0, h0, img,
                                   (uint16_t)ch_s[j], (uint16_t)ch_r[j]);
              } else {
                tma_load_2d(dst, mX, &full_bar[stage], ch_c[j], p0);
              }
            }
          }
          for (int i = 0; i < b_chunks; ++i)
            tma_load_2d(sb + pl * Cfg::kBBytes + i * kBChunkBytes, mDY, &full_bar[stage],
                        co0 + i * CWB, p0);
        }
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    // MN-major operands: one pixel row = width*2 bytes, 8-row groups SBO apart, column blocks
    // (64/32/16 channels) LBO apart.  16 pixel rows per UMMA.
    const uint64_t a_desc0 = make_smem_desc(smem_u32(smem), kAChunkBytes, 8 * CW * 2,
                                            swizzle_layout_type(CW * 2));
    const uint64_t b_desc0 = make_smem_desc(smem_u32(smem) + NP * Cfg::kABytes, kBChunkBytes,
                                            8 * CWB * 2, swizzle_layout_type(CWB * 2));
    for (int it = 0; it < num_ks; ++it) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t da0 = a_desc0 + static_cast<uint64_t>(stage * (Cfg::kStageBytes >> 4));
        const uint64_t db0 = b_desc0 + static_cast<uint64_t>(stage * (Cfg::kStageBytes >> 4));
#pragma unroll
        for (int h = 0; h < MT; ++h) {
          if (h == 0 || second) {
#pragma unroll
            for (int ks = 0; ks < PIX / 16; ++ks) {
#pragma unroll
              for (int t = 0; t < (NP == 3 ? 6 : 1); ++t) {
                const int pa = NP == 3 ? plane_term_a(t) : 0;
                const int pb = NP == 3 ? plane_term_b(t) : 0;
                const bool small = NP == 3 && t < 5;
                umma_bf16(tmem_base + h * BN + (small ? BN : 0),
                          da0 + ((pa * Cfg::kABytes + h * Cfg::kAHalfBytes + ks * 16 * CW * 2) >> 4),
                          db0 + ((pb * Cfg::kBBytes + ks * 16 * CWB * 2) >> 4), kIdesc,
                          (it | ks | (small ? t : 0)) ? 1u : 0u);
              }
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        if (it == num_ks - 1) umma_commit(&tmem_full_bar);
      }
      __syncwarp();
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else {
    // TMEM lane = row of D = (tap,ci) column of dw; consecutive lanes -> consecutive addresses.
    const int quarter = warp & 3;
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int h = 0; h < MT; ++h) {
      if (h == 1 && !second) break;          // warp-uniform
      const int n = m0 + h * 128 + quarter * 32 + lane;
      const bool n_ok = n < p.Ktot;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        if (co0 + c * 32 >= p.Cout) break;   // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + h * BN + c * 32,
                      v);
        tmem_ld_wait();
        if (NP == 3) {   // + the small-term accumulator
          uint32_t v2[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + BN + c * 32, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
        }
        if (n_ok) {
          float* dst = p.dw + static_cast<size_t>(co0 + c * 32) * p.Ktot + n;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // (measured: replacing these atomics by plain stores changes the training step by
            // 0.13 ms of 25.5 -- profiles/r02_exp_knobs.txt -- the split-K reduction is not what
            // bounds wgrad)
            if (co0 + c * 32 + i < p.Cout)
              atomicAdd(dst + static_cast<size_t>(i) * p.Ktot, __uint_as_float(v[i]));
          }
        }
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

struct ConvMaps {
  CUtensorMap a[3], b[3], c, add, mask;
};

// output staging buffers of the conv GEMM epilogue: 0 = two where the ring stays deep, 1 = always
// one (default: the interleaved A/B of the c3 step shows no difference beyond noise -- 23.76 / 23.69
// / 23.86 ms for 1 / 0 / 2, profiles/r02_exp_ab.txt -- so the deeper ring is kept), 2 = two wherever
// shared memory allows (acnn_set_conv_out_bufs; ACNN_CONV_OUT_BUFS sets the initial value)
static int conv_out_bufs_default() {
  const char* e = getenv("ACNN_CONV_OUT_BUFS");
  return e ? (e[0] - '0') : 1;
}
static int g_conv_out_bufs = conv_out_bufs_default();

// split epilogue of the conv GEMM kernel: 0 = off, 1 (default) = where the epilogue chain paces the
// tile (K <= 256), 2 = wherever it fits (acnn_set_conv_split_epilogue; ACNN_CONV_SPLIT_EPI sets the
// initial value).  Interleaved A/B of the c3 step: 23.36 / 23.14 / 23.19 ms for 0 / 1 / 2
// (profiles/r02_exp_ab.txt); per layer profiles/r02_exp_split_epilogue_layers.txt
static int conv_split_epi_default() {
  const char* e = getenv("ACNN_CONV_SPLIT_EPI");
  return e ? (e[0] - '0') : 1;
}
static int g_conv_split_epi = conv_split_epi_default();
// the same for two-M-tile CTA tiles (the two groups take the two M tiles of every tile): 0 = off,
// 1 (default) = K <= 256, 2 = wherever it fits (acnn_set_conv_split_mt2; ACNN_CONV_SPLIT_MT2 sets
// the initial value).  Interleaved A/B of the c3 step: 23.28 / 23.11 / 23.17 ms for 0 / 1 / 2
static int conv_split_mt2_default() {
  const char* e = getenv("ACNN_CONV_SPLIT_MT2");
  return e ? (e[0] - '0') : 1;
}
static int g_conv_split_mt2 = conv_split_mt2_default();

template <int BN, int CW, bool IM2COL, int MT, int NP, bool CG2 = false>
static int launch_conv_gemm(const ConvMaps& tm, const ConvGemmParams& p, int per_n,
                            cudaStream_t stream) {
  using Cfg = FpropCfg<BN, MT, NP, CG2>;
  static bool attr_set = false;
  auto kern = conv_gemm_kernel<BN, CW, IM2COL, MT, NP, CG2>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget + 2048);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_gemm): %s", cudaGetErrorString(e));
      return ACNN_ERR_CUDA;
    }
    attr_set = true;
  }
  ConvGemmParams q = p;
  // a second output staging half tile where the ring stays deep enough without its bytes: all of K
  // in flight, or >= 4 stages (g_conv_out_bufs: 0 = this rule, 1 / 2 = force where it fits)
  q.out_bufs = 1;
  if (!p.out_f32 && g_conv_out_bufs != 1) {
    const int st2 = Cfg::stages_for(p.has_add, p.has_mask, p.out_f32, 2);
    const int num_kb = ceil_div(p.Ktot, kStageK);
    const int need = g_conv_out_bufs == 2 ? (MT == 2 ? 3 : 2)
                                          : (num_kb + 1 < 4 ? num_kb + 1 : 4);
    if (st2 >= need && st2 >= (MT == 2 ? 3 : 2) &&
        Cfg::smem_bytes(st2, p.has_add, p.has_mask, p.out_f32, 2) <= kSmemBudget + 1024)
      q.out_bufs = 2;
  }
  // split epilogue (two 4-warp groups alternate tiles): one M tile per CTA tile, bf16 output, no
  // bias; needs two output staging tiles and two sets of add / mask tiles.  Mode 1 = where the
  // epilogue chain, not the k-loop, paces a tile (K <= 256: at most four k-blocks) and the ring still
  // holds all of K or 3 stages; mode 2 = wherever it fits with >= 2 stages
  q.split_epi = 0;
  const int split_mode = MT == 2 ? g_conv_split_mt2 : g_conv_split_epi;
  if (NP == 1 && !CG2 && !p.out_f32 && !p.bias && split_mode > 0) {
    const int num_kb = ceil_div(p.Ktot, kStageK);
    const int st2 = Cfg::stages_for(p.has_add, p.has_mask, false, 2, 2);
    const int min_st = MT == 2 ? 3 : 2;
    const bool fits = st2 >= min_st && Cfg::smem_bytes(st2, p.has_add, p.has_mask, false, 2, 2) <=
                                           kSmemBudget + 1024;
    const int need = num_kb + 1 < 3 ? num_kb + 1 : 3;
    if (fits && (split_mode >= 2 || (num_kb <= 4 && st2 >= need))) {
      q.split_epi = 1;
      q.out_bufs = 2;
    }
  }
  const int aux_sets = q.split_epi ? 2 : 1;
  q.stages = Cfg::stages_for(p.has_add, p.has_mask, p.out_f32, q.out_bufs, aux_sets);
  q.m_tiles = ceil_div(p.M, (CG2 ? 2 : MT) * kBM);
  q.n_tiles = p.Cout / BN;
  const int smem =
      Cfg::smem_bytes(q.stages, p.has_add, p.has_mask, p.out_f32, q.out_bufs, aux_sets);
  if (CG2) {
    // per_n CTA PAIRS per N tile, launched as clusters of two (ranks 2i, 2i+1 share a TPC)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * per_n * q.n_tiles);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    (void)cudaLaunchKernelEx(&cfg, kern, tm.a[0], tm.b[0], tm.c, tm.add, tm.mask, tm.a[1], tm.a[2],
                             tm.b[1], tm.b[2], q);
  } else {
    launch_k(kern, dim3(per_n * q.n_tiles), dim3(kConvThreads), smem, stream, tm.a[0], tm.b[0], tm.c,
             tm.add, tm.mask, tm.a[1], tm.a[2], tm.b[1], tm.b[2], q);
  }
  count_launch();
  return check_launch("conv_gemm_kernel");
}

// -1: choose per problem; 1 / 2: force that many M tiles per CTA tile where the shape allows it
static int g_conv_mtiles_mode = -1;

// Tile shape and persistent grid of one conv GEMM: a pure function of the problem, shared by the
// launcher and by acnn_conv_stats_parts() (the number of partial statistics rows = per_n).
struct ConvTiling {
  int bn;      // N tile
  int mt;      // M tiles (128 pixels) per CTA tile
  int per_n;   // CTAs (pair = 1: CTA pairs) per N tile
  int pair;    // 1: CTA pairs (tcgen05 cta_group::2), one 256 x bn tile per pair
  int parts;   // rows of the partial statistics buffer (= CTAs per N tile)
};

// 1 (default): N = 256 tiles with K >= 512 run on CTA pairs; 0: off (acnn_set_conv_cta_pairs;
// ACNN_CONV_PAIRS=0|1 sets the initial value)
static int conv_pairs_default() {
  const char* e = getenv("ACNN_CONV_PAIRS");
  return e ? (e[0] == '1') : 1;
}
static int g_conv_pairs = conv_pairs_default();

static ConvTiling conv_tiling(int M, int Cout, int Ktot, int cw, bool has_add, bool has_mask,
                              bool out_f32, int np) {
  ConvTiling t;
  // N tile: 256 halves the A-operand traffic per FLOP (128 B/clk of smem reads at BN=128 is the
  // SM's whole shared-memory bandwidth); dense / tiny-M problems keep 128 for more CTAs
  t.bn = (Cout % 128 == 0) ? 128 : ((Cout % 64 == 0) ? 64 : 32);
  if (np == 1 && Cout % 256 == 0 && !out_f32 && (int64_t)ceil_div(M, kBM) * (Cout / 256) >= 96)
    t.bn = 256;
  // Two M tiles per CTA when TMEM has room (N <= 128), the smem ring stays >= 3 stages deep and
  // there are enough tiles to keep every SM busy for several rounds.
  // (at N = 128 that rules out only the add + mask epilogue: two 32 KiB aux staging tiles)
  t.mt = 1;
  if (np == 1 && t.bn <= 128 && !out_f32) {
    const bool fits = fprop_stages(t.bn, 2, 1, has_add, has_mask, out_f32) >= 3;
    const int64_t tiles = (int64_t)ceil_div(M, 2 * kBM) * (Cout / t.bn);
    if (fits && (g_conv_mtiles_mode == 2 ||
                 (g_conv_mtiles_mode == -1 && tiles >= 4 * (int64_t)num_sms())))
      t.mt = 2;
  }
  // persistent grid: a multiple of n_tiles so that every CTA keeps one N tile (its weights and
  // its per-channel statistics), at most one CTA per SM
  const int n_tiles = Cout / t.bn;
  // CTA pairs: measured per layer (profiles/r02_exp_cta_pairs.txt) -- K >= 512 gains 15-25 % (the
  // k-loop is shared-memory-ingest bound), small K loses 10-20 % (epilogue-bound tiles, and a pair
  // schedules half as many, twice as large units); 64-channel chunks only (instantiation count)
  t.pair = (g_conv_pairs && t.bn == 256 && np == 1 && t.mt == 1 && cw == 64 && Ktot >= 512 &&
            ceil_div(M, 2 * kBM) * n_tiles >= num_sms() / 2) ? 1 : 0;
  const int m_tiles = ceil_div(M, (t.pair ? 2 : t.mt) * kBM);
  t.per_n = (t.pair ? num_sms() / 2 : num_sms()) / n_tiles;
  if (t.per_n < 1) t.per_n = 1;
  if (t.per_n > m_tiles) t.per_n = m_tiles;
  t.parts = t.pair ? 2 * t.per_n : t.per_n;
  return t;
}

template <int BN, int MT, int NP>
static int dispatch_conv_cw(int cw, bool im2col, const ConvMaps& tm, const ConvGemmParams& p,
                            int per_n, cudaStream_t s) {
  if (im2col) {
    if (cw == 64) return launch_conv_gemm<BN, 64, true, MT, NP>(tm, p, per_n, s);
    if (cw == 32) return launch_conv_gemm<BN, 32, true, MT, NP>(tm, p, per_n, s);
    return launch_conv_gemm<BN, 16, true, MT, NP>(tm, p, per_n, s);
  }
  if (cw == 64) return launch_conv_gemm<BN, 64, false, MT, NP>(tm, p, per_n, s);
  if (cw == 32) return launch_conv_gemm<BN, 32, false, MT, NP>(tm, p, per_n, s);
  return launch_conv_gemm<BN, 16, false, MT, NP>(tm, p, per_n, s);
}

template <int BN>
static int dispatch_conv_gemm(const ConvTiling& t, int np, int cw, bool im2col, const ConvMaps& tm,
                              const ConvGemmParams& p, cudaStream_t s) {
  if constexpr (BN == 256) {
    if (t.pair) {     // full-width (64-channel) chunks only: bounds the instantiation count
      if (im2col) return launch_conv_gemm<256, 64, true, 1, 1, true>(tm, p, t.per_n, s);
      return launch_conv_gemm<256, 64, false, 1, 1, true>(tm, p, t.per_n, s);
    }
  }
  if constexpr (BN <= 128) {
    if (np == 3) return dispatch_conv_cw<BN, 1, 3>(cw, im2col, tm, p, t.per_n, s);
    if (t.mt == 2) return dispatch_conv_cw<BN, 2, 1>(cw, im2col, tm, p, t.per_n, s);
  }
  return dispatch_conv_cw<BN, 1, 1>(cw, im2col, tm, p, t.per_n, s);
}

static int chunk_width(int cin) { return cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 16); }

static int out_hw(const acnn_conv_geom& g, int* Ho, int* Wo) {
  *Ho = (g.H + g.pad_h_lo + g.pad_h_hi - g.kh) / g.stride + 1;
  *Wo = (g.W + g.pad_w_lo + g.pad_w_hi - g.kw) / g.stride + 1;
  return (*Ho > 0 && *Wo > 0) ? 1 : 0;
}

// elements of the input tensor (one operand plane)
static int64_t input_elems(const acnn_conv_geom& g) {
  int64_t pix, row, img;
  input_pitches(g, &pix, &row, &img);
  return (int64_t)g.B * img;
}

// ---- halo (im2col-free 3x3) kernel: eligibility, tiling, launch --------------------------------
// 0: off; 1 (default): where measured to pay (use_halo below); 2: wherever the kernel applies
// (acnn_set_conv_halo; ACNN_CONV_HALO=0|1|2 sets the initial value)
static int conv_halo_default() {
  const char* e = getenv("ACNN_CONV_HALO");
  return e ? (e[0] - '0') : 1;
}
static int g_conv_halo = conv_halo_default();

static int halo_bn(int Cout) { return Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32); }

// Shared-memory plan of the halo kernel for one problem (a pure function of the shape).
struct HaloPlan {
  int a_stages, b_slots, stationary, out_bufs, split_epi, smem;
};
// 1 (default): N <= 64 tiles use two independent epilogue groups; 0: all 8 warps on one tile
// (acnn_set_conv_halo_split; ACNN_CONV_HALO_SPLIT sets the initial value)
static int conv_halo_split_default() {
  const char* e = getenv("ACNN_CONV_HALO_SPLIT");
  return e ? (e[0] - '0') : 1;
}
static int g_conv_halo_split = conv_halo_split_default();

static HaloPlan halo_plan(int bn, int cw, int Cin, bool has_add, bool has_mask) {
  HaloPlan h;
  const int row_b = cw * 2;
  const int a_stage = (kHaloH * kHaloW * row_b + 1023) / 1024 * 1024;
  const int b_tile = bn * row_b;
  const int tile = kBM * bn * 2;
  const int nchunks = Cin / cw;
  // two output staging tiles where they are cheap (N <= 64: 8 / 16 KiB); the split epilogue also
  // doubles the add / mask staging (one set per group)
  h.out_bufs = bn <= 64 ? 2 : 1;
  h.split_epi = (bn <= 64 && g_conv_halo_split) ? 1 : 0;
  const int aux_sets = h.split_epi ? 2 : 1;
  const int fixed =
      1024 + tile * (h.out_bufs + (has_add ? aux_sets : 0) + (has_mask ? aux_sets : 0));
  h.a_stages = 3;
  h.b_slots = (kSmemBudget - fixed - h.a_stages * a_stage) / b_tile;
  if (h.b_slots < nchunks * 9) {         // one A stage fewer if that makes the slab stationary
    const int b2 = (kSmemBudget - fixed - 2 * a_stage) / b_tile;
    if (b2 >= nchunks * 9 || h.b_slots < 3) {
      h.a_stages = 2;
      h.b_slots = b2;
    }
  }
  if (h.b_slots > kHaloMaxB) h.b_slots = kHaloMaxB;
  h.stationary = (h.b_slots >= 2 && nchunks * 9 <= h.b_slots) ? 1 : 0;
  if (h.stationary) h.b_slots = nchunks * 9;
  h.smem = fixed + h.a_stages * a_stage + h.b_slots * b_tile;
  return h;
}

static bool use_halo(const acnn_conv_geom& g, int np, bool out_f32, bool has_bias, bool has_add,
                     bool has_mask) {
  if (g_conv_halo <= 0) return false;
  const bool applies = g.kh == 3 && g.kw == 3 && g.stride == 1 && g.pad_h_lo == 1 &&
                       g.pad_h_hi == 1 && g.pad_w_lo == 1 && g.pad_w_hi == 1 &&
                       g.x_pix_stride <= 0 && g.x_row_pitch <= 0 && g.x_img_pitch <= 0 && np == 1 &&
                       !out_f32 && !has_bias && (g.Cin % 64 == 0 || g.Cin == 32) &&
                       g.Cout % 32 == 0;
  if (!applies) return false;
  if (g_conv_halo >= 2) return true;
  // mode 1 = where measured to pay (profiles/r02_exp_halo_layers.txt): the weight slab of an N tile
  // stays in shared memory (otherwise the weight stream replaces the im2col re-reads as the ingest
  // bound: 56x56 128->64 dgrad 0.163 -> 0.243 ms) and the images have >= 56 rows (16 x 8 patches
  // waste <= 12.5 % of a 56 x 56 image, 27 % of 28 x 28)
  const HaloPlan h = halo_plan(halo_bn(g.Cout), g.Cin % 64 == 0 ? 64 : 32, g.Cin, has_add, has_mask);
  return h.stationary && g.H >= 56;
}

// CTAs per N tile of the halo kernel's persistent grid (= partial statistics rows)
static int halo_per_n(const acnn_conv_geom& g) {
  const int n_tiles = g.Cout / halo_bn(g.Cout);
  const int m_tiles = g.B * ceil_div(g.H, kPatchH) * ceil_div(g.W, kPatchW);
  int per_n = num_sms() / n_tiles;
  if (per_n < 1) per_n = 1;
  return per_n > m_tiles ? m_tiles : per_n;
}

static int make_map_4d(CUtensorMap* m, const void* base, int C, int W, int H, int B, int box_c,
                       int box_w, int box_h) {
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swizzle_enum(box_c * 2), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-d) failed (%d): C=%d W=%d H=%d B=%d box=%dx%dx%d", (int)r,
              C, W, H, B, box_c, box_w, box_h);
    return ACNN_ERR_CUDA;
  }
  return ACNN_OK;
}

template <int BN, int CW>
static int launch_conv_halo(const acnn_conv_geom& g, const void* x, const void* w, void* y,
                            float* ch_part, const void* add_src, const void* mask_src,
                            cudaStream_t stream) {
  using Cfg = HaloCfg<BN, CW>;
  static bool attr_set = false;
  auto kern = conv_halo_kernel<BN, CW>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget + 2048);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_halo): %s", cudaGetErrorString(e));
      return ACNN_ERR_CUDA;
    }
    attr_set = true;
  }
  HaloParams p;
  p.H = g.H; p.W = g.W; p.B = g.B; p.Cin = g.Cin; p.Cout = g.Cout;
  p.n_tiles = g.Cout / BN;
  p.ph = ceil_div(g.H, kPatchH);
  p.pw = ceil_div(g.W, kPatchW);
  p.m_tiles = g.B * p.ph * p.pw;
  p.ch_part = ch_part;
  p.has_add = add_src != nullptr;
  p.has_mask = mask_src != nullptr;
  const int nchunks = g.Cin / CW;
  const HaloPlan hp = halo_plan(BN, CW, g.Cin, p.has_add != 0, p.has_mask != 0);
  p.out_bufs = hp.out_bufs;
  p.split_epi = hp.split_epi;
  p.a_stages = hp.a_stages;
  p.b_slots = hp.b_slots;
  p.stationary = hp.stationary;
  ACNN_REQUIRE(p.b_slots >= 2, "conv (halo): shared memory does not fit");
  const int smem = hp.smem;
  CUtensorMap tmA, tmB, tmC, tmAdd, tmMask;
  int rc = make_map_4d(&tmA, x, g.Cin, g.W, g.H, g.B, CW, kHaloW, kHaloH);
  if (rc) return rc;
  if ((rc = make_map_2d(&tmB, w, g.Cout, 9 * g.Cin, 9 * g.Cin, BN, CW))) return rc;
  if ((rc = make_map_4d(&tmC, y, g.Cout, g.W, g.H, g.B, Cfg::kSubW, kPatchW, kPatchH))) return rc;
  tmAdd = tmMask = tmC;
  if (add_src && (rc = make_map_4d(&tmAdd, add_src, g.Cout, g.W, g.H, g.B, Cfg::kSubW, kPatchW,
                                   kPatchH)))
    return rc;
  if (mask_src && (rc = make_map_4d(&tmMask, mask_src, g.Cout, g.W, g.H, g.B, Cfg::kSubW, kPatchW,
                                    kPatchH)))
    return rc;
  launch_k(kern, dim3(halo_per_n(g) * p.n_tiles), dim3(kConvThreads), smem, stream, tmA, tmB, tmC,
           tmAdd, tmMask, p);
  count_launch();
  return check_launch("conv_halo_kernel");
}

static int conv_halo_host(const acnn_conv_geom& g, const void* x, const void* w, void* y,
                          float* ch_part, const void* add_src, const void* mask_src,
                          cudaStream_t stream) {
  const int bn = halo_bn(g.Cout);
  const bool c64 = g.Cin % 64 == 0;
  if (bn == 128) {
    return c64 ? launch_conv_halo<128, 64>(g, x, w, y, ch_part, add_src, mask_src, stream)
               : launch_conv_halo<128, 32>(g, x, w, y, ch_part, add_src, mask_src, stream);
  }
  if (bn == 64) {
    return c64 ? launch_conv_halo<64, 64>(g, x, w, y, ch_part, add_src, mask_src, stream)
               : launch_conv_halo<64, 32>(g, x, w, y, ch_part, add_src, mask_src, stream);
  }
  return c64 ? launch_conv_halo<32, 64>(g, x, w, y, ch_part, add_src, mask_src, stream)
             : launch_conv_halo<32, 32>(g, x, w, y, ch_part, add_src, mask_src, stream);
}

// precision 0: x / w are bf16.  precision 1 (fp32 parity mode): x and w each are THREE consecutive
// bf16 planes (acnn_split3 / acnn_prep_weights with planes = 3), plane p of x at x + p * numel(x),
// plane p of w at w + p * w_plane_stride elements; y must be fp32 (out_f32), no fused epilogue.
static int conv_gemm_host(const acnn_conv_geom& g, const void* x, const void* w, void* y,
                          float* ch_part, const void* add_src, const void* mask_src,
                          const float* bias, int out_f32, int precision, int64_t w_plane_stride,
                          cudaStream_t stream) {
  ACNN_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && g.Cin > 0 && g.Cout > 0, "conv: empty geometry");
  ACNN_REQUIRE(g.Cin % 16 == 0, "conv: Cin=%d must be a multiple of 16", g.Cin);
  ACNN_REQUIRE(g.Cout % 32 == 0, "conv: Cout=%d must be a multiple of 32", g.Cout);
  ACNN_REQUIRE(g.stride >= 1 && g.kh >= 1 && g.kw >= 1, "conv: bad kernel/stride");
  ACNN_REQUIRE(!(ch_part && out_f32), "conv: statistics only with bf16 output");
  ACNN_REQUIRE(precision == 0 || precision == 1, "conv: precision must be 0 (bf16) or 1 (fp32)");
  ACNN_REQUIRE(precision == 0 || (out_f32 && !add_src && !mask_src && !ch_part && w_plane_stride > 0),
               "conv: the fp32 (3-plane) mode needs fp32 output, a weight plane stride and no "
               "fused add / mask / statistics epilogue");
  int Ho, Wo;
  ACNN_REQUIRE(out_hw(g, &Ho, &Wo), "conv: empty output");
  ACNN_REQUIRE(g.pad_h_lo <= 128 && g.pad_w_lo <= 128 && g.kh <= 128 && g.kw <= 128,
               "conv: padding / filter exceed the TMA im2col corner range");
  int rc = load_driver_fns();
  if (rc) return rc;
  if (use_halo(g, precision ? 3 : 1, out_f32 != 0, bias != nullptr, add_src != nullptr,
               mask_src != nullptr))
    return conv_halo_host(g, x, w, y, ch_part, add_src, mask_src, stream);

  const bool plain = is_plain(g);
  const int cw = chunk_width(g.Cin);
  const int np = precision ? 3 : 1;
  ConvGemmParams p;
  p.M = g.B * Ho * Wo;
  p.Cout = g.Cout;
  p.Cin = g.Cin;
  p.Ktot = g.kh * g.kw * g.Cin;
  p.kw = g.kw;
  p.HoWo = Ho * Wo;
  p.Wo = Wo;
  p.stride = g.stride;
  p.pad_h_lo = g.pad_h_lo;
  p.pad_w_lo = g.pad_w_lo;
  p.b_sw_bytes = p.Ktot >= 64 ? 128 : p.Ktot * 2;
  p.y = y;
  p.ch_part = ch_part;
  p.bias = bias;
  p.has_add = add_src != nullptr;
  p.has_mask = mask_src != nullptr;
  p.out_f32 = out_f32;
  p.stages = p.m_tiles = p.n_tiles = 0;
  ACNN_REQUIRE(p.b_sw_bytes == 128 || p.b_sw_bytes == 64 || p.b_sw_bytes == 32,
               "conv: unsupported K=%d", p.Ktot);

  const ConvTiling t =
      conv_tiling(p.M, g.Cout, p.Ktot, cw, p.has_add, p.has_mask, out_f32 != 0, np);
  const int bn = t.bn;
  ConvMaps tm;
  const int64_t x_plane = input_elems(g);
  for (int pl = 0; pl < np; ++pl) {
    const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x) + pl * x_plane;
    const __nv_bfloat16* wp = static_cast<const __nv_bfloat16*>(w) + pl * w_plane_stride;
    if (plain) {
      rc = make_map_2d(&tm.a[pl], xp, p.M, g.Cin, g.Cin, kBM, cw);
    } else {
      rc = make_map_im2col(&tm.a[pl], xp, g, cw, kBM);
    }
    if (rc) return rc;
    rc = make_map_2d(&tm.b[pl], wp, g.Cout, p.Ktot, p.Ktot, t.pair ? bn / 2 : bn,
                     p.Ktot >= 64 ? 64 : p.Ktot);
    if (rc) return rc;
  }
  for (int pl = np; pl < 3; ++pl) {   // placeholders
    tm.a[pl] = tm.a[0];
    tm.b[pl] = tm.b[0];
  }
  tm.c = tm.add = tm.mask = tm.b[0];   // placeholders when unused
  const int subw = bn < 64 ? bn : 64;
  if (!out_f32 && (rc = make_map_2d(&tm.c, y, p.M, g.Cout, g.Cout, kBM, subw))) return rc;
  if (add_src && (rc = make_map_2d(&tm.add, add_src, p.M, g.Cout, g.Cout, kBM, subw))) return rc;
  if (mask_src && (rc = make_map_2d(&tm.mask, mask_src, p.M, g.Cout, g.Cout, kBM, subw))) return rc;
  if (bn == 256) return dispatch_conv_gemm<256>(t, np, cw, !plain, tm, p, stream);
  if (bn == 128) return dispatch_conv_gemm<128>(t, np, cw, !plain, tm, p, stream);
  if (bn == 64) return dispatch_conv_gemm<64>(t, np, cw, !plain, tm, p, stream);
  return dispatch_conv_gemm<32>(t, np, cw, !plain, tm, p, stream);
}

struct WgradMaps {
  CUtensorMap x[3], dy[3];
};

// fixed cost of one wgrad CTA (pipeline fill + atomic epilogue) in units of pipeline stages, for
// the split-K cost model (default 16: measured, profiles/r02_exp_wgrad_split.txt: -0.5 ms per c3 step
// against 0 = the round-1 "two waves of CTAs" rule; flat from 4 to 1000); acnn_set_wgrad_overhead_stages
static int g_wgrad_overhead_stages = 16;

template <int BN, int CW, int CWB, bool IM2COL, int MT, int NP, int PIX = kWgPix>
static int launch_wgrad(const WgradMaps& tm, WgradParams p, int n_tiles, int deterministic,
                        cudaStream_t stream) {
  using Cfg = WgradCfg<BN, MT, NP, PIX>;
  static bool attr_set = false;
  auto kern = wgrad_gemm_kernel<BN, CW, CWB, IM2COL, MT, NP, PIX>;
  const int m_tiles = ceil_div(p.Ktot, MT * 128);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(e));
      return ACNN_ERR_CUDA;
    }
    attr_set = true;
  }
  // split the pixel (K) range over CTAs; deterministic mode: no split -- every dw element receives
  // exactly one (atomic) add onto the zeroed buffer
  const int ctas_per_sm = Cfg::kSmemBytes <= 110 * 1024 ? 2 : 1;
  const int tiles = m_tiles * n_tiles;
  const int max_splits = p.stages_total >= 8 ? p.stages_total / 4 : 1;
  int splits;
  if (g_wgrad_overhead_stages <= 0) {
    // (round-1 rule: roughly two waves of CTAs)
    splits = ceil_div(2 * ctas_per_sm * num_sms(), tiles);
  } else {
    // cost model: an SM runs its CTAs' pipeline stages back to back (co-resident CTAs share the
    // tensor pipe) and pays a fixed pipeline-fill + epilogue cost, worth g_wgrad_overhead_stages
    // stages, once per round of resident CTAs; pick the split count with the least modelled time
    // (avoids e.g. 2.2 waves of CTAs where 0.97 waves do the same work in less time)
    splits = 1;
    int64_t best = -1;
    for (int s = 1; s <= max_splits; ++s) {
      const int per = ceil_div(p.stages_total, s);
      const int s_eff = ceil_div(p.stages_total, per);
      if (s_eff != s) continue;
      const int per_sm = ceil_div(tiles * s, num_sms());
      const int64_t t = (int64_t)per_sm * per +
                        (int64_t)ceil_div(per_sm, ctas_per_sm) * g_wgrad_overhead_stages;
      if (best < 0 || t < best) { best = t; splits = s; }
    }
  }
  if (splits > max_splits) splits = max_splits;
  if (splits < 1 || deterministic) splits = 1;
  p.stages_per_split = ceil_div(p.stages_total, splits);
  splits = ceil_div(p.stages_total, p.stages_per_split);
  dim3 grid(m_tiles, n_tiles, splits);
  launch_k(kern, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, tm.x[0], tm.dy[0], tm.x[1],
           tm.x[2], tm.dy[1], tm.dy[2], p);
  count_launch();
  return check_launch("wgrad_gemm_kernel");
}

template <int BN, int CW, bool IM2COL>
static int dispatch_wgrad_cwb(int cwb, int np, const WgradMaps& tm, const WgradParams& p, int nt,
                              int det, cudaStream_t s) {
  if constexpr (BN <= 128) {
    if (np == 3) {
      if constexpr (BN >= 64) {
        if (cwb == 64) return launch_wgrad<BN, CW, 64, IM2COL, 1, 3>(tm, p, nt, det, s);
      }
      return launch_wgrad<BN, CW, 32, IM2COL, 1, 3>(tm, p, nt, det, s);
    }
  }
  if constexpr (BN >= 64) {
    if (cwb == 64) {
      // two 128-row blocks per CTA (full-width chunks only, to bound the instantiation count)
      if constexpr (CW == 64) {
        // measured: pays when the (tap,ci) extent and the pixel count are both large (otherwise
        // halving the number of CTAs costs more than the shared dy stage saves)
        const bool big = (p.Ktot >= 1024 && p.P >= 50176) || p.Ktot >= 4096;
        if (p.pix == 64 && p.Ktot > 128 &&
            (g_conv_mtiles_mode == 2 || (g_conv_mtiles_mode == -1 && big)))
          return launch_wgrad<BN, CW, 64, IM2COL, 2, 1>(tm, p, nt, det, s);
      }
      if constexpr (BN <= 128) {
        if (p.pix == 128) return launch_wgrad<BN, CW, 64, IM2COL, 1, 1, 128>(tm, p, nt, det, s);
      }
      return launch_wgrad<BN, CW, 64, IM2COL, 1, 1>(tm, p, nt, det, s);
    }
  }
  if constexpr (BN <= 128) {
    if (p.pix == 128) return launch_wgrad<BN, CW, 32, IM2COL, 1, 1, 128>(tm, p, nt, det, s);
  }
  return launch_wgrad<BN, CW, 32, IM2COL, 1, 1>(tm, p, nt, det, s);
}

template <int BN, bool IM2COL>
static int dispatch_wgrad_cw(int cw, int cwb, int np, const WgradMaps& tm, const WgradParams& p,
                             int nt, int det, cudaStream_t s) {
  if (cw == 64) return dispatch_wgrad_cwb<BN, 64, IM2COL>(cwb, np, tm, p, nt, det, s);
  if (cw == 32) return dispatch_wgrad_cwb<BN, 32, IM2COL>(cwb, np, tm, p, nt, det, s);
  return dispatch_wgrad_cwb<BN, 16, IM2COL>(cwb, np, tm, p, nt, det, s);
}

template <bool IM2COL>
static int dispatch_wgrad(int bn, int cw, int cwb, int np, const WgradMaps& tm, const WgradParams& p,
                          int nt, int det, cudaStream_t s) {
  if (bn == 256) return dispatch_wgrad_cw<256, IM2COL>(cw, cwb, np, tm, p, nt, det, s);
  if (bn == 128) return dispatch_wgrad_cw<128, IM2COL>(cw, cwb, np, tm, p, nt, det, s);
  if (bn == 64) return dispatch_wgrad_cw<64, IM2COL>(cw, cwb, np, tm, p, nt, det, s);
  return dispatch_wgrad_cw<32, IM2COL>(cw, cwb, np, tm, p, nt, det, s);
}

// 0: choose per problem; 64 / 128: force the pixels per stage where the shape allows it
static int g_wgrad_pix = 0;
static bool wgrad_wants_pix128(const WgradParams&, int) {
  // measured (profiles/r02_exp_ab.txt, interleaved A/B under the one-round split rule): 128-pixel
  // stages wherever they fit (N tile <= 128) take 0.24 ms off the c3 step
  return true;
}

static int conv_wgrad_host(const acnn_conv_geom& g, const void* x, const void* dy, float* dw,
                           int precision, int deterministic, cudaStream_t stream) {
  ACNN_REQUIRE(g.Cin % 16 == 0 && g.Cout % 32 == 0, "wgrad: Cin %% 16 / Cout %% 32 required");
  ACNN_REQUIRE(precision == 0 || precision == 1, "wgrad: precision must be 0 (bf16) or 1 (fp32)");
  int Ho, Wo;
  ACNN_REQUIRE(out_hw(g, &Ho, &Wo), "wgrad: empty output");
  int rc = load_driver_fns();
  if (rc) return rc;
  const bool plain = is_plain(g);
  const int np = precision ? 3 : 1;
  WgradParams p;
  p.P = g.B * Ho * Wo;
  p.Cout = g.Cout;
  p.Cin = g.Cin;
  p.Ktot = g.kh * g.kw * g.Cin;
  p.kw = g.kw;
  p.HoWo = Ho * Wo;
  p.Wo = Wo;
  p.stride = g.stride;
  p.pad_h_lo = g.pad_h_lo;
  p.pad_w_lo = g.pad_w_lo;
  int bn = g.Cout >= 256 ? 256 : (g.Cout >= 128 ? 128 : (g.Cout >= 64 ? 64 : 32));
  if (np == 3 && bn > 128) bn = 128;   // three operand planes per stage: smem
  // 128-pixel stages: only the bf16 path, N tile <= 128 (smem), enough pixels to split
  p.pix = kWgPix;
  if (np == 1 && bn <= 128 && p.P >= 4096 &&
      (g_wgrad_pix == 128 || (g_wgrad_pix == 0 && wgrad_wants_pix128(p, bn))))
    p.pix = 128;
  p.stages_total = ceil_div(p.P, p.pix);
  p.stages_per_split = p.stages_total;
  p.dw = dw;
  const int cw = chunk_width(g.Cin);
  const int cwb = (g.Cout % 64 == 0) ? 64 : 32;
  ACNN_REQUIRE(g.Cout % bn == 0, "wgrad: Cout=%d not a multiple of its N tile %d", g.Cout, bn);
  WgradMaps tm;
  const int64_t x_plane = input_elems(g), dy_plane = (int64_t)p.P * g.Cout;
  for (int pl = 0; pl < np; ++pl) {
    const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x) + pl * x_plane;
    const __nv_bfloat16* dp = static_cast<const __nv_bfloat16*>(dy) + pl * dy_plane;
    rc = make_map_2d(&tm.dy[pl], dp, p.P, g.Cout, g.Cout, p.pix, cwb);
    if (rc) return rc;
    if (plain) {
      rc = make_map_2d(&tm.x[pl], xp, p.P, g.Cin, g.Cin, p.pix, cw);
    } else {
      rc = make_map_im2col(&tm.x[pl], xp, g, cw, p.pix);
    }
    if (rc) return rc;
  }
  for (int pl = np; pl < 3; ++pl) {
    tm.x[pl] = tm.x[0];
    tm.dy[pl] = tm.dy[0];
  }
  const int nt = g.Cout / bn;
  if (plain) return dispatch_wgrad<false>(bn, cw, cwb, np, tm, p, nt, deterministic, stream);
  return dispatch_wgrad<true>(bn, cw, cwb, np, tm, p, nt, deterministic, stream);
}

}  // namespace acnn

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int acnn_set_conv_mtiles(int mode) {
  const int prev = acnn::g_conv_mtiles_mode;
  acnn::g_conv_mtiles_mode = (mode == 1 || mode == 2) ? mode : -1;
  return prev;
}

int acnn_set_conv_halo(int mode) {
  const int prev = acnn::g_conv_halo;
  acnn::g_conv_halo = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return prev;
}

int acnn_set_conv_split_epilogue(int mode) {
  const int prev = acnn::g_conv_split_epi;
  acnn::g_conv_split_epi = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return prev;
}

int acnn_set_conv_split_mt2(int mode) {
  const int prev = acnn::g_conv_split_mt2;
  acnn::g_conv_split_mt2 = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return prev;
}

int acnn_set_conv_halo_split(int on) {
  const int prev = acnn::g_conv_halo_split;
  acnn::g_conv_halo_split = on ? 1 : 0;
  return prev;
}

int acnn_set_conv_out_bufs(int mode) {
  const int prev = acnn::g_conv_out_bufs;
  acnn::g_conv_out_bufs = mode < 0 ? 0 : (mode > 2 ? 2 : mode);
  return prev;
}

int acnn_set_wgrad_overhead_stages(int stages) {
  const int prev = acnn::g_wgrad_overhead_stages;
  acnn::g_wgrad_overhead_stages = stages < 0 ? 0 : stages;
  return prev;
}

int acnn_set_wgrad_pixels(int pix) {
  const int prev = acnn::g_wgrad_pix;
  acnn::g_wgrad_pix = (pix == 64 || pix == 128) ? pix : 0;
  return prev;
}

int acnn_conv_stats_parts(const acnn_conv_geom* g) {
  if (!g) return 0;
  int Ho, Wo;
  if (!acnn::out_hw(*g, &Ho, &Wo) || g->Cout % 32 != 0) return 0;
  if (acnn::use_halo(*g, 1, false, false, false, false)) return acnn::halo_per_n(*g);
  return acnn::conv_tiling(g->B * Ho * Wo, g->Cout, g->kh * g->kw * g->Cin,
                           acnn::chunk_width(g->Cin), false, false, false, 1).parts;
}

int acnn_set_conv_cta_pairs(int on) {
  const int prev = acnn::g_conv_pairs;
  acnn::g_conv_pairs = on ? 1 : 0;
  return prev;
}

int acnn_conv_fprop(const acnn_conv_geom* g, const void* x, const void* w, void* y,
                    float* ch_part, const void* add_src, const void* mask_src, const float* bias,
                    int out_f32, int precision, int64_t w_plane_stride, void* stream) {
  if (!g || !x || !w || !y) {
    acnn::set_error("acnn_conv_fprop: null argument");
    return ACNN_ERR_INVALID;
  }
  return acnn::conv_gemm_host(*g, x, w, y, ch_part, add_src, mask_src, bias, out_f32, precision,
                              w_plane_stride, static_cast<cudaStream_t>(stream));
}

int acnn_conv_dgrad(const acnn_conv_geom* g, const void* dy, const void* w_dgrad, void* dx,
                    const void* add_src, const void* mask_src, int precision,
                    int64_t w_plane_stride, void* stream) {
  if (!g || !dy || !w_dgrad || !dx) {
    acnn::set_error("acnn_conv_dgrad: null argument");
    return ACNN_ERR_INVALID;
  }
  if (g->stride != 1) {
    acnn::set_error("acnn_conv_dgrad: stride %d (zero-insert dy first, then call with stride 1)",
                    g->stride);
    return ACNN_ERR_UNSUPPORTED;
  }
  // dx = correlation of dy with the flipped, channel-transposed filter; padding k-1-pad.
  const int Ho = g->H + g->pad_h_lo + g->pad_h_hi - g->kh + 1;
  const int Wo = g->W + g->pad_w_lo + g->pad_w_hi - g->kw + 1;
  acnn_conv_geom t;
  t.B = g->B;
  t.H = Ho;
  t.W = Wo;
  t.Cin = g->Cout;
  t.Cout = g->Cin;
  t.kh = g->kh;
  t.kw = g->kw;
  t.stride = 1;
  t.pad_h_lo = g->kh - 1 - g->pad_h_lo;
  t.pad_h_hi = g->kh - 1 - g->pad_h_hi;
  t.pad_w_lo = g->kw - 1 - g->pad_w_lo;
  t.pad_w_hi = g->kw - 1 - g->pad_w_hi;
  t.x_pix_stride = t.x_row_pitch = t.x_img_pitch = t.reserved_ = 0;
  return acnn::conv_gemm_host(t, dy, w_dgrad, dx, nullptr, add_src, mask_src, nullptr,
                              precision ? 1 : 0, precision, w_plane_stride,
                              static_cast<cudaStream_t>(stream));
}

int acnn_conv_wgrad(const acnn_conv_geom* g, const void* x, const void* dy, float* dw,
                    int precision, int deterministic, void* stream) {
  if (!g || !x || !dy || !dw) {
    acnn::set_error("acnn_conv_wgrad: null argument");
    return ACNN_ERR_INVALID;
  }
  return acnn::conv_wgrad_host(*g, x, dy, dw, precision, deterministic,
                               static_cast<cudaStream_t>(stream));
}

}  // extern "C"
