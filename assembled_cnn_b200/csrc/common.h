// Host-side helpers shared by every translation unit of libacnn.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/acnn.h"

namespace acnn {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Returns ACNN_OK or records the CUDA error (launch-configuration errors surface here).
int check_launch(const char* what);

#define ACNN_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ::acnn::set_error(__VA_ARGS__);  \
      return ACNN_ERR_INVALID;         \
    }                                  \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace acnn
