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

// selects the device and verifies it is a Blackwell (sm_100) part: there is no fallback path.
int select_device(int device);

}  // namespace tod
