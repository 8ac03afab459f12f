// Shared host-side helpers for libtod_b200.so: error reporting and CUDA checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <string>
#include "../../include/tod.h"

namespace tod {

std::string& last_error();  // thread-local (api.cu)

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

#define TOD_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::tod::fail(TOD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define TOD_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r < 0) return _r;     \
  } while (0)

// Programmatic dependent launch (PDL).  While the graph executor sets pdl_next(), the launch helpers below attach
// cudaLaunchAttributeProgrammaticStreamSerialization, so the kernel may start (and run its constant-only prologue:
// barrier init, TMEM allocation, table / filter loads) while its stream predecessor drains.  Such kernels execute
// griddepcontrol.wait before they touch any activation, and every kernel fires griddepcontrol.launch_dependents first.
bool& pdl_next();  // thread-local (api.cu)

template <class... KArgs, class... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  if (pdl_next()) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// selects the device and verifies it is a Blackwell (sm_100) part: there is no fallback path.
int select_device(int device);

}  // namespace tod
