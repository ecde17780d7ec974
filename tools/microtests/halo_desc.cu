// MICRO-TEST (sm_100a semantics probe, not part of the library): can a tcgen05 K-major SWIZZLE_128B
// shared-memory descriptor (a) start at an arbitrary 128-byte row of a TMA-written tile and (b) use a
// stride between 8-row groups that is not a multiple of 1024 bytes?  If so, the nine taps of a 3x3
// convolution can read shifted windows of ONE halo tile in shared memory instead of nine im2col copies.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o halo_desc halo_desc.cu -lcuda
//
// One CTA: TMA-loads an 18 x 10 pixel x 64 channel halo tile (4-D tiled map, OOB zero fill, start
// coordinate (-1,-1)) and nine 64 x 64 weight tiles, then runs experiments and prints the max error
// of each against a CPU reference.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../assembled_cnn_b200/csrc/ptx.cuh"

using namespace acnn;
using bf16 = __nv_bfloat16;

#ifndef CH
#define CH 64
#endif
constexpr int H = 20, W = 12, C = CH, NOUT = 64;
constexpr int ROWB = C * 2;                  // bytes of one pixel row (128: SWIZZLE_128B, 64: SWIZZLE_64B)
constexpr uint32_t LAYOUT = ROWB == 128 ? 2u : 4u;
constexpr int HH = 18, HW = 10;              // halo tile (16 x 8 outputs, 3x3, pad 1)
constexpr int kHaloBytes = HH * HW * C * 2;  // 23040
constexpr int kHaloPad = (kHaloBytes + 1023) / 1024 * 1024;
constexpr int kBTile = NOUT * C * 2;         // 8192 per tap
constexpr int NEXP = 16;

struct Exp { int kind; int shift; int bo_mode; };   // kind 0: linear rows, 1: 3x3 patch conv
__constant__ Exp c_exp[NEXP];

__device__ __forceinline__ void tma_load_4d_tile(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo, int bo_mode) {
  uint64_t d = make_smem_desc(addr, 16, sbo, LAYOUT);
  if (bo_mode == 1) d |= static_cast<uint64_t>((addr >> 7) & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* out,
      int nexp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar, mma_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const uint32_t s_halo = smem_u32(smem), s_b = s_halo + kHaloPad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<64>(&tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&full_bar, kHaloBytes + 9 * kBTile);
    tma_load_4d_tile(s_halo, &tmX, smem_u32(&full_bar), 0, -1, -1, 0);
    for (int t = 0; t < 9; ++t)
      tma_load_2d_a(s_b + t * kBTile, &tmW, smem_u32(&full_bar), t * C, 0);
  }
  mbar_wait(&full_bar, 0);
  tc_fence_after();
  constexpr uint32_t idesc = make_idesc_bf16(NOUT, false, false);
  for (int e = 0; e < nexp; ++e) {
    const Exp ex = c_exp[e];
    if (threadIdx.x == 0) {
      if (ex.kind == 0) {
        // rows = 128 consecutive pixels of the halo tile starting at pixel `shift`; tap-0 weights
        for (int ks = 0; ks < C / 16; ++ks)
          umma_bf16(tmem, desc_sw128(s_halo + ex.shift * ROWB + ks * 32, 8 * ROWB, ex.bo_mode),
                    desc_sw128(s_b + ks * 32, 8 * ROWB, 0), idesc, ks ? 1u : 0u);
      } else {
        // 16 x 8 output patch: row group g = output row, 8 pixels; groups 10 pixels (1280 B) apart
        for (int t = 0; t < 9; ++t)
          for (int ks = 0; ks < C / 16; ++ks)
            umma_bf16(tmem,
                      desc_sw128(s_halo + ((t / 3) * HW + (t % 3)) * ROWB + ks * 32, HW * ROWB,
                                 ex.bo_mode),
                      desc_sw128(s_b + t * kBTile + ks * 32, 8 * ROWB, 0), idesc, (t | ks) ? 1u : 0u);
      }
      umma_commit(&mma_bar);
    }
    mbar_wait(&mma_bar, e & 1);
    tc_fence_after();
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c * 32, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i)
        out[(static_cast<size_t>(e) * 128 + warp * 32 + lane) * NOUT + c * 32 + i] =
            __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc<64>(tmem);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float bfr(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fn);
  std::vector<float> X(H * W * C), Wt(NOUT * 9 * C);
  srand(1);
  for (auto& v : X) v = bfr((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : Wt) v = bfr((rand() % 2001 - 1000) / 4000.f);
  std::vector<bf16> Xh(X.size()), Wh(Wt.size());
  for (size_t i = 0; i < X.size(); ++i) Xh[i] = __float2bfloat16(X[i]);
  for (size_t i = 0; i < Wt.size(); ++i) Wh[i] = __float2bfloat16(Wt[i]);
  bf16 *dX, *dW;
  float* dOut;
  cudaMalloc(&dX, Xh.size() * 2);
  cudaMalloc(&dW, Wh.size() * 2);
  cudaMalloc(&dOut, NEXP * 128 * NOUT * 4);
  cudaMemcpy(dX, Xh.data(), Xh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, Wh.data(), Wh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dOut, 0, NEXP * 128 * NOUT * 4);
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[4] = {C, W, H, 1};
    cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {C, HW, HH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dX, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode X failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {9 * C, NOUT};
    cuuint64_t strides[1] = {9 * C * 2};
    cuuint32_t box[2] = {C, NOUT};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dW, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, ROWB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode W failed %d\n", (int)r); return 1; }
  }
  std::vector<Exp> exps;
  for (int s : {0, 8, 1, 3, 11, 20}) for (int bo : {0, 1}) exps.push_back({0, s, bo});
  exps.push_back({1, 0, 0});
  exps.push_back({1, 0, 1});
  const int nexp = (int)exps.size();
  cudaMemcpyToSymbol(c_exp, exps.data(), nexp * sizeof(Exp));
  const int smem = 1024 + kHaloPad + 9 * kBTile;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(tmX, tmW, dOut, nexp);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out((size_t)nexp * 128 * NOUT);
  cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
  // halo tile on the CPU: pixel (hh, hw) = input (hh - 1, hw - 1), zero outside
  auto halo = [&](int hh, int hw, int c) -> float {
    const int h = hh - 1, w = hw - 1;
    return (h < 0 || h >= H || w < 0 || w >= W) ? 0.f : X[(h * W + w) * C + c];
  };
  for (int ei = 0; ei < nexp; ++ei) {
    double worst = 0, ref_max = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < NOUT; ++n) {
        double ref = 0;
        if (exps[ei].kind == 0) {
          const int p = exps[ei].shift + m;
          for (int c = 0; c < C; ++c) ref += (double)halo(p / HW, p % HW, c) * Wt[(n * 9 + 0) * C + c];
        } else {
          const int r = m / 8, cc = m % 8;
          for (int t = 0; t < 9; ++t)
            for (int c = 0; c < C; ++c)
              ref += (double)halo(r + t / 3, cc + t % 3, c) * Wt[(n * 9 + t) * C + c];
        }
        const double err = fabs(ref - out[((size_t)ei * 128 + m) * NOUT + n]);
        if (err > worst) worst = err;
        if (fabs(ref) > ref_max) ref_max = fabs(ref);
      }
    printf("C=%d exp %2d kind=%s shift=%2d base_offset=%s : max err %.3e (ref max %.3f) %s\n", C, ei,
           exps[ei].kind ? "patch3x3" : "linear  ", exps[ei].shift, exps[ei].bo_mode ? "auto" : "0   ",
           worst, ref_max, worst < 1e-3 * ref_max ? "MATCH" : "MISMATCH");
  }
  return 0;
}
