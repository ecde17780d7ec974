// The tiny fully-connected layers of the SK / SE attention paths (<= 0.3 MMAC per image):
// CUDA-core fp32 kernels on [B, <=2048] descriptors.  nets/blocks.py:136-151 (SK), :171-182 (SE).
#include "common.h"
#include "vec.cuh"

namespace acnn {

constexpr int kTB = 4;       // batch rows per block
constexpr int kFT = 256;

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return 1.f / (1.f + __expf(-v));
  return v;
}

// out[b][j] = act(sum_k in[b][k] * W[j][k])          W row-major [J][K]
__global__ void __launch_bounds__(kFT)
fc_nt_kernel(const float* __restrict__ in, const float* __restrict__ W, float* __restrict__ out,
             int B, int K, int J, int act) {
  extern __shared__ float tile[];   // [kTB][K]
  const int b0 = blockIdx.x * kTB;
  for (int i = threadIdx.x; i < kTB * K; i += kFT) {
    const int t = i / K, k = i - t * K;
    tile[i] = (b0 + t < B) ? in[(size_t)(b0 + t) * K + k] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < J; j += kFT / 32) {
    float acc[kTB];
#pragma unroll
    for (int t = 0; t < kTB; ++t) acc[t] = 0.f;
    const float* wr = W + (size_t)j * K;
    for (int k = lane; k < K; k += 32) {
      const float w = __ldg(wr + k);
#pragma unroll
      for (int t = 0; t < kTB; ++t) acc[t] = fmaf(tile[t * K + k], w, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < kTB; ++t) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
    }
    if (lane < kTB && b0 + lane < B) out[(size_t)(b0 + lane) * J + j] = apply_act(acc[lane], act);
  }
}

// out[b][j] = post * sum_k in[b][k] * W[k][j]          W row-major [K][J]
__global__ void __launch_bounds__(kFT)
fc_nn_kernel(const float* __restrict__ in, const float* __restrict__ W, float* __restrict__ out,
             int B, int K, int J, float post) {
  extern __shared__ float tile[];   // [kTB][K]
  const int b0 = blockIdx.x * kTB;
  for (int i = threadIdx.x; i < kTB * K; i += kFT) {
    const int t = i / K, k = i - t * K;
    tile[i] = (b0 + t < B) ? in[(size_t)(b0 + t) * K + k] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < J; j += kFT) {
    float acc[kTB];
#pragma unroll
    for (int t = 0; t < kTB; ++t) acc[t] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(W + (size_t)k * J + j);
#pragma unroll
      for (int t = 0; t < kTB; ++t) acc[t] = fmaf(tile[t * K + k], w, acc[t]);
    }
#pragma unroll
    for (int t = 0; t < kTB; ++t)
      if (b0 + t < B) out[(size_t)(b0 + t) * J + j] = acc[t] * post;
  }
}

// dW[k][j] += sum_b A[b][k] * Bm[b][j]
__global__ void __launch_bounds__(kFT)
outer_acc_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* dW, int B, int K,
                 int J) {
  const int64_t idx = blockIdx.x * (int64_t)kFT + threadIdx.x;
  if (idx >= (int64_t)K * J) return;
  const int k = (int)(idx / J), j = (int)(idx - (int64_t)k * J);
  float acc = 0.f;
  for (int b = 0; b < B; ++b) acc = fmaf(__ldg(A + (size_t)b * K + k), __ldg(Bm + (size_t)b * J + j), acc);
  dW[idx] += acc;
}

// One warp per channel j: batch-norm over the batch dimension, then ReLU.
__global__ void __launch_bounds__(kFT)
bn_batch_relu_fwd_kernel(const float* __restrict__ zpre, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float* moving_mean, float* moving_var,
                         float momentum, float eps, int training, float* __restrict__ z,
                         float* __restrict__ bnstat, int B, int d) {
  const int j = blockIdx.x * (kFT / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= d) return;
  float mean, var;
  if (training) {
    float s = 0.f, q = 0.f;
    for (int b = lane; b < B; b += 32) {
      const float v = zpre[(size_t)b * d + j];
      s += v;
      q += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    mean = s / B;
    var = fmaxf(q / B - mean * mean, 0.f);
    if (lane == 0) {
      const float unbiased = var * ((float)B / fmaxf((float)B - 1.f, 1.f));
      moving_mean[j] = moving_mean[j] * momentum + mean * (1.f - momentum);
      moving_var[j] = moving_var[j] * momentum + unbiased * (1.f - momentum);
    }
  } else {
    mean = moving_mean[j];
    var = moving_var[j];
  }
  const float rstd = rsqrtf(var + eps);
  if (lane == 0) {
    bnstat[j] = mean;
    bnstat[d + j] = rstd;
  }
  const float sc = gamma[j] * rstd, sh = beta[j] - mean * sc;
  for (int b = lane; b < B; b += 32)
    z[(size_t)b * d + j] = fmaxf(fmaf(zpre[(size_t)b * d + j], sc, sh), 0.f);
}

// dz (in: grad wrt z; out: grad wrt zpre), one warp per channel.
__global__ void __launch_bounds__(kFT)
bn_batch_relu_bwd_kernel(float* dz, const float* __restrict__ z, const float* __restrict__ zpre,
                         const float* __restrict__ bnstat, const float* __restrict__ gamma,
                         float* dgamma, float* dbeta, int B, int d) {
  const int j = blockIdx.x * (kFT / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= d) return;
  const float mean = bnstat[j], rstd = bnstat[d + j];
  float s1 = 0.f, s2 = 0.f;
  for (int b = lane; b < B; b += 32) {
    const size_t i = (size_t)b * d + j;
    const float g = z[i] > 0.f ? dz[i] : 0.f;
    s1 += g;
    s2 += g * ((zpre[i] - mean) * rstd);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float k1 = gamma[j] * rstd;
  for (int b = lane; b < B; b += 32) {
    const size_t i = (size_t)b * d + j;
    const float g = z[i] > 0.f ? dz[i] : 0.f;
    const float xh = (zpre[i] - mean) * rstd;
    dz[i] = k1 * (g - s1 / B - xh * s2 / B);
  }
  if (lane == 0) {
    dgamma[j] += s2;
    dbeta[j] += s1;
  }
}

__global__ void sk_gate_fwd_kernel(const float* __restrict__ a, float* __restrict__ att, int B,
                                   int f) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * f) return;
  const int b = (int)(i / f), c = (int)(i - (int64_t)b * f);
  const float d = a[(size_t)b * 2 * f + c] - a[(size_t)b * 2 * f + f + c];
  att[i] = 1.f / (1.f + expf(-d));
}

__global__ void sk_gate_bwd_kernel(const float* __restrict__ dA, const float* __restrict__ att,
                                   float* __restrict__ da, int B, int f) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * f) return;
  const int b = (int)(i / f), c = (int)(i - (int64_t)b * f);
  const float t = att[i] * (1.f - att[i]) * dA[i];
  da[(size_t)b * 2 * f + c] = t;
  da[(size_t)b * 2 * f + f + c] = -t;
}

// out = in * e * (1 - e)   (sigmoid backward)   /   out = in * (h > 0)   (relu backward)
__global__ void ew_bwd_kernel(const float* in, const float* __restrict__ act,
                              float* out, int64_t n, int mode) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (mode == 2) ? in[i] * act[i] * (1.f - act[i]) : (act[i] > 0.f ? in[i] : 0.f);
}

static int fc_nt(const float* in, const float* W, float* out, int B, int K, int J, int act,
                 cudaStream_t s) {
  ACNN_REQUIRE(kTB * K * 4 <= 48 * 1024, "fc: K=%d too large", K);
  fc_nt_kernel<<<ceil_div(B, kTB), kFT, kTB * K * 4, s>>>(in, W, out, B, K, J, act);
  count_launch();
  return check_launch("fc_nt");
}
static int fc_nn(const float* in, const float* W, float* out, int B, int K, int J, float post,
                 cudaStream_t s) {
  ACNN_REQUIRE(kTB * K * 4 <= 48 * 1024, "fc: K=%d too large", K);
  fc_nn_kernel<<<ceil_div(B, kTB), kFT, kTB * K * 4, s>>>(in, W, out, B, K, J, post);
  count_launch();
  return check_launch("fc_nn");
}
static int outer_acc(const float* A, const float* Bm, float* dW, int B, int K, int J,
                     cudaStream_t s) {
  outer_acc_kernel<<<(int)ceil_div64((int64_t)K * J, kFT), kFT, 0, s>>>(A, Bm, dW, B, K, J);
  count_launch();
  return check_launch("outer_acc");
}

}  // namespace acnn

using namespace acnn;

extern "C" {

int acnn_sk_fc_fwd(const float* s, const float* w1, const float* gamma, const float* beta,
                   float* moving_mean, float* moving_var, float momentum, float eps, int training,
                   const float* w2, float* zpre, float* bnstat, float* z, float* att,
                   float* scratch, int B, int f, int d, void* stream) {
  ACNN_REQUIRE(s && w1 && gamma && beta && moving_mean && moving_var && w2 && zpre && bnstat && z &&
                   att && scratch, "sk_fc_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = fc_nt(s, w1, zpre, B, f, d, 0, st);
  if (rc) return rc;
  bn_batch_relu_fwd_kernel<<<ceil_div(d, kFT / 32), kFT, 0, st>>>(
      zpre, gamma, beta, moving_mean, moving_var, momentum, eps, training, z, bnstat, B, d);
  count_launch();
  if ((rc = check_launch("sk bn_batch_relu_fwd"))) return rc;
  if ((rc = fc_nt(z, w2, scratch, B, d, 2 * f, 0, st))) return rc;
  sk_gate_fwd_kernel<<<(int)ceil_div64((int64_t)B * f, 256), 256, 0, st>>>(scratch, att, B, f);
  count_launch();
  return check_launch("sk_gate_fwd");
}

int acnn_sk_fc_bwd(const float* dA, const float* att, const float* z, const float* zpre,
                   const float* bnstat, const float* gamma, const float* s, const float* w1,
                   const float* w2, float* dw1, float* dw2, float* dgamma, float* dbeta, float* ds,
                   float* scratch, int B, int f, int d, void* stream) {
  ACNN_REQUIRE(dA && att && z && zpre && bnstat && gamma && s && w1 && w2 && dw1 && dw2 && dgamma &&
                   dbeta && ds && scratch, "sk_fc_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* da = scratch;                       // [B][2f]
  float* dz = scratch + (size_t)B * 2 * f;   // [B][d]
  sk_gate_bwd_kernel<<<(int)ceil_div64((int64_t)B * f, 256), 256, 0, st>>>(dA, att, da, B, f);
  count_launch();
  int rc = check_launch("sk_gate_bwd");
  if (rc) return rc;
  if ((rc = outer_acc(da, z, dw2, B, 2 * f, d, st))) return rc;        // dW2[2f][d]
  if ((rc = fc_nn(da, w2, dz, B, 2 * f, d, 1.f, st))) return rc;        // dz = da * W2
  bn_batch_relu_bwd_kernel<<<ceil_div(d, kFT / 32), kFT, 0, st>>>(dz, z, zpre, bnstat, gamma,
                                                                  dgamma, dbeta, B, d);
  count_launch();
  if ((rc = check_launch("sk bn_batch_relu_bwd"))) return rc;
  if ((rc = outer_acc(dz, s, dw1, B, d, f, st))) return rc;             // dW1[d][f]
  return fc_nn(dz, w1, ds, B, d, f, 1.f, st);                           // ds = dzpre * W1
}

int acnn_se_fc_fwd(const float* q, const float* w1, const float* w2, float* h, float* e, int B,
                   int C, int r, void* stream) {
  ACNN_REQUIRE(q && w1 && w2 && h && e, "se_fc_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = fc_nt(q, w1, h, B, C, r, 1, st);
  if (rc) return rc;
  return fc_nt(h, w2, e, B, r, C, 2, st);
}

int acnn_se_fc_bwd(const float* de, const float* e, const float* h, const float* q,
                   const float* w1, const float* w2, float* dw1, float* dw2, float* dq,
                   float* scratch, int B, int C, int r, int HW, void* stream) {
  ACNN_REQUIRE(de && e && h && q && w1 && w2 && dw1 && dw2 && dq && scratch,
               "se_fc_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* da2 = scratch;                     // [B][C]
  float* dh = scratch + (size_t)B * C;      // [B][r]
  const int64_t n2 = (int64_t)B * C, n1 = (int64_t)B * r;
  ew_bwd_kernel<<<(int)ceil_div64(n2, 256), 256, 0, st>>>(de, e, da2, n2, 2);
  count_launch();
  int rc = check_launch("se sigmoid bwd");
  if (rc) return rc;
  if ((rc = outer_acc(da2, h, dw2, B, C, r, st))) return rc;            // dW2[C][r]
  if ((rc = fc_nn(da2, w2, dh, B, C, r, 1.f, st))) return rc;           // dh = da2 * W2
  ew_bwd_kernel<<<(int)ceil_div64(n1, 256), 256, 0, st>>>(dh, h, dh, n1, 1);
  count_launch();
  if ((rc = check_launch("se relu bwd"))) return rc;
  if ((rc = outer_acc(dh, q, dw1, B, r, C, st))) return rc;             // dW1[r][C]
  return fc_nn(dh, w1, dq, B, r, C, 1.f / HW, st);                      // dq = da1 * W1 / HW
}

}  // extern "C"
