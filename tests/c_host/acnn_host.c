/* A host written in plain C against include/acnn_model.h: nothing but pointers and sizes cross the
 * boundary.  `acnn_host plan` needs no GPU (acnn_create is host logic); `acnn_host step` allocates the
 * caller-owned buffers with cudaMalloc, draws the reference's initializers on the host, feeds host
 * arrays and runs three Assemble-ResNet-50 training steps (mixup type 1) through acnn_step, printing the loss.
 * Built and run by tests/test_native_plan_cpu.py (plan) and tests/test_native_model_gpu.py (step):
 *   gcc -std=c99 -O1 -I include -I /usr/local/cuda/include tests/c_host/acnn_host.c \
 *       -o assembled_cnn_b200/build/acnn_host -L assembled_cnn_b200 -l:libacnn.so \
 *       -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/assembled_cnn_b200 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "acnn_model.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != 0) {                                                              \
      fprintf(stderr, "%s failed (rc=%d): %s\n", #call, rc_, acnn_last_error()); \
      return 1;                                                                  \
    }                                                                            \
  } while (0)
#define CUDA(call)                                                           \
  do {                                                                       \
    cudaError_t e_ = (call);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      fprintf(stderr, "%s failed: %s\n", #call, cudaGetErrorString(e_));     \
      return 1;                                                              \
    }                                                                        \
  } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static double uniform01(void) { /* xorshift64* */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 0x2545F4914F6CDD1Dull) >> 11) / 9007199254740992.0;
}
static double normal01(void) {
  double u = uniform01(), v = uniform01();
  return sqrt(-2.0 * log(u + 1e-300)) * cos(6.283185307179586 * v);
}

int main(int argc, char** argv) {
  const int do_step = argc > 1 && strcmp(argv[1], "step") == 0;
  acnn_model_config cfg;
  acnn_model_config_init(&cfg);
  /* scripts/train_assemble_from_scratch.sh: Assemble-ResNet-50 = resnet_version 2 + SK + sconv/3 */
  cfg.resnet_version = 2;
  cfg.use_sk_block = 1;
  strcpy(cfg.anti_alias_type, "sconv");
  cfg.anti_alias_filter_size = 3;
  cfg.batch = 8;
  cfg.height = cfg.width = 64;
  cfg.mixup_type = 1;
  cfg.label_smoothing = 0.1;
  acnn_model* m = NULL;
  CHECK(acnn_create(&cfg, &m));
  acnn_model_sizes s;
  CHECK(acnn_model_get_sizes(m, &s));
  acnn_variable_info vi;
  CHECK(acnn_variable_info_get(m, 0, &vi));
  printf("plan: %d variables (first %s), %lld parameters, %d tensors, %d+%d+%d ops, workspace %.1f MB\n",
         s.n_variables, vi.name, (long long)s.param_elems, s.n_tensors, s.n_forward, s.n_backward,
         s.n_update, s.workspace_bytes / 1048576.0);
  if (!do_step) {
    acnn_destroy(m);
    return 0;
  }

  /* caller-owned device buffers */
  float *params, *grads, *mom, *state;
  void *wf, *wd, *ws;
  CUDA(cudaMalloc((void**)&params, s.param_elems * 4));
  CUDA(cudaMalloc((void**)&grads, s.param_elems * 4));
  CUDA(cudaMalloc((void**)&mom, s.param_elems * 4));
  CUDA(cudaMalloc((void**)&state, s.state_elems * 4));
  CUDA(cudaMalloc(&wf, s.w_fprop_elems * 2));
  CUDA(cudaMalloc(&wd, s.w_dgrad_elems * 2));
  CUDA(cudaMalloc(&ws, s.workspace_bytes));
  CUDA(cudaMemset(mom, 0, s.param_elems * 4));
  cudaStream_t st;
  CUDA(cudaStreamCreate(&st));
  CHECK(acnn_bind(m, params, grads, mom, state, wf, wd, ws, st));

  /* the reference's initializers (nets/model_helper.py:77 variance_scaling, gamma 1, beta 0,
   * moving_mean 0 / moving_variance 1), drawn on the host in the flat layouts acnn_variable_info gives */
  float* hp_ = (float*)calloc(s.param_elems, 4);
  float* hs_ = (float*)calloc(s.state_elems, 4);
  for (int i = 0; i < s.n_variables; ++i) {
    CHECK(acnn_variable_info_get(m, i, &vi));
    float* dst = (vi.buffer == ACNN_BUF_PARAMS ? hp_ : hs_) + vi.offset;
    if (!strcmp(vi.kind, "conv_kernel") || !strcmp(vi.kind, "dense_kernel")) {
      double fan_in = 1;
      for (int d = 0; d + 1 < vi.tf_rank; ++d) fan_in *= (double)vi.tf_shape[d];
      /* dense rows beyond num_classes (padding to ld_logits) stay zero */
      const long long live = !strcmp(vi.kind, "dense_kernel") ? vi.tf_shape[0] * vi.tf_shape[1] : vi.size;
      for (long long k = 0; k < live; ++k) dst[k] = (float)(normal01() / sqrt(fan_in));
    } else if (!strcmp(vi.kind, "gamma") || !strcmp(vi.kind, "moving_variance")) {
      for (long long k = 0; k < vi.size; ++k) dst[k] = 1.0f;
    }
  }
  CUDA(cudaMemcpy(params, hp_, s.param_elems * 4, cudaMemcpyHostToDevice));
  CUDA(cudaMemcpy(state, hs_, s.state_elems * 4, cudaMemcpyHostToDevice));

  /* host inputs: 2*batch images (mixup type 1), labels, one lambda per mixed example */
  const int Bin = s.input_batch, H = cfg.height, W = cfg.width;
  float* x = (float*)malloc((size_t)Bin * H * W * 3 * 4);
  int32_t* y = (int32_t*)malloc((size_t)Bin * 4);
  float* lam = (float*)malloc((size_t)cfg.batch * 4);
  for (long long k = 0; k < (long long)Bin * H * W * 3; ++k) x[k] = (float)(64.0 * normal01());
  for (int k = 0; k < Bin; ++k) y[k] = 1 + (int32_t)(uniform01() * 1000);
  for (int k = 0; k < cfg.batch; ++k) lam[k] = (float)uniform01();
  float hyper[8] = {0.05f, 0.9f, 1e-4f, 1.0f, 1.0f, 0, 0, 0}; /* lr, momentum, wd, grad_scale, keep_prob */
  float loss[4], l2_first = 0, l2_last = 0;
  float* logits = (float*)malloc((size_t)cfg.batch * cfg.num_classes * 4);
  for (int step = 0; step < 3; ++step) {
    CHECK(acnn_set_inputs(m, x, y, lam, NULL, NULL, st));
    CHECK(acnn_set_hparams(m, hyper, st));
    CHECK(acnn_step(m, st));
    CHECK(acnn_get_loss(m, loss, st));
    CHECK(acnn_get_logits(m, logits, st));
    CUDA(cudaStreamSynchronize(st));
    printf("step %d: cross_entropy %.5f l2_loss %.5f logits[0][1] %.5f\n", step, loss[0], loss[1], logits[1]);
    /* 1001 classes: ln(1001) = 6.9 at initialisation; anything non-finite or far off is a failure
     * (values are checked against the oracle through the same call sequence in
     * tests/test_native_model_gpu.py::test_c_abi_call_sequence_with_host_arrays_against_oracle) */
    if (!(loss[0] > 3.0f && loss[0] < 15.0f) || !(loss[1] > 0) || !(logits[1] == logits[1])) return 2;
    if (step == 0) l2_first = loss[1];
    l2_last = loss[1];
  }
  /* the optimizer ran: the L2 term of the weights moved between the first and the last step */
  if (l2_last == l2_first) {
    fprintf(stderr, "weights did not change: l2_loss %.6f -> %.6f\n", l2_first, l2_last);
    return 3;
  }
  acnn_destroy(m);
  printf("ok\n");
  return 0;
}
