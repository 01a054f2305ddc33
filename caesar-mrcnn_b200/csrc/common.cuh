// Shared helpers for the caesar-mrcnn B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

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

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL): every kernel of the detect plan is launched with the
// programmatic-stream-serialization attribute and starts with pdl_prologue(): the next kernel's CTAs are
// placed on SMs as soon as this grid's CTAs free them and run their own prologue (barrier init, TMEM
// allocation, tensor-map prefetch, index arithmetic) while the tail of this grid is still running;
// griddepcontrol.wait then holds them until this grid has completed and its memory is visible.
// Both instructions are no-ops in a kernel launched without the attribute.  MRCNN_B200_PDL=0 turns the
// attribute off.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_launch_dependents();
  pdl_wait();
}
#endif

bool mrcnn_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t mrcnn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mrcnn_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
