// Library-wide entry points of libtod_b200.so: error string, ABI version, device selection.
#include "common.h"

namespace tod {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

bool& pdl_next() {
  static thread_local bool v = false;
  return v;
}

int select_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(TOD_ERR_NO_DEVICE, "no CUDA device is visible (%s); this library has no CPU path",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= count) return fail(TOD_ERR_INVALID_ARG, "device %d out of range (0..%d)", device, count - 1);
  cudaDeviceProp prop{};
  TOD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(TOD_ERR_NO_DEVICE, "device %d (%s, sm_%d%d) is not a Blackwell sm_100 part; the kernels are built for sm_100a only",
                device, prop.name, prop.major, prop.minor);
  TOD_CUDA(cudaSetDevice(device));
  return TOD_OK;
}

}  // namespace tod

extern "C" {

const char* tod_last_error(void) { return tod::last_error().c_str(); }

int tod_abi_version(void) { return TOD_ABI_VERSION; }

int tod_device_count(int* count) {
  if (!count) return tod::fail(TOD_ERR_INVALID_ARG, "tod_device_count: null argument");
  *count = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return tod::fail(TOD_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = n;
  return TOD_OK;
}

}  // extern "C"
