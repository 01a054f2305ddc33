// Shared helpers for the caesar-mrcnn B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define MRCNN_OK 0
#define MRCNN_ERR_INVALID (-1)
#define MRCNN_ERR_CUDA (-2)
#define MRCNN_ERR_UNSUPPORTED (-3)
#define MRCNN_ERR_NOTFOUND (-4)

// last-error plumbing (thread-local string, read through mrcnn_last_error())
void mrcnn_set_error(const char* fmt, ...);

#define MRCNN_CHECK_CUDA(expr)                                                        \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      mrcnn_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,           \
                      cudaGetErrorString(_e));                                        \
      return MRCNN_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define MRCNN_REQUIRE(cond, ...)                                                      \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      mrcnn_set_error(__VA_ARGS__);                                                   \
      return MRCNN_ERR_INVALID;                                                       \
    }                                                                                 \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
