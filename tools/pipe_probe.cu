// Throughput of IDP.4A / IMAD / PRMT / LOP3 per SM on sm_100a: sizing input for the depthwise kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipe_probe tools/pipe_probe.cu && gpurun_out/pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(int* out, int n, int seed) {
  int a0 = threadIdx.x + seed, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
  const int b = seed * 0x01010101 + 7;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) { a0 = __dp4a(a0, b, a0); a1 = __dp4a(a1, b, a1); a2 = __dp4a(a2, b, a2); a3 = __dp4a(a3, b, a3); a4 = __dp4a(a4, b, a4); a5 = __dp4a(a5, b, a5); a6 = __dp4a(a6, b, a6); a7 = __dp4a(a7, b, a7); }
      if (OP == 1) { a0 = a0 * b + a1; a1 = a1 * b + a2; a2 = a2 * b + a3; a3 = a3 * b + a4; a4 = a4 * b + a5; a5 = a5 * b + a6; a6 = a6 * b + a7; a7 = a7 * b + a0; }
      if (OP == 2) { a0 = __byte_perm(a0, a1, b); a1 = __byte_perm(a1, a2, b); a2 = __byte_perm(a2, a3, b); a3 = __byte_perm(a3, a4, b); a4 = __byte_perm(a4, a5, b); a5 = __byte_perm(a5, a6, b); a6 = __byte_perm(a6, a7, b); a7 = __byte_perm(a7, a0, b); }
      if (OP == 3) { a0 = (a0 & b) ^ a1; a1 = (a1 & b) ^ a2; a2 = (a2 & b) ^ a3; a3 = (a3 & b) ^ a4; a4 = (a4 & b) ^ a5; a5 = (a5 & b) ^ a6; a6 = (a6 & b) ^ a7; a7 = (a7 & b) ^ a0; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int OP>
void run(const char* name) {
  int* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int n = 4096;
  k<OP><<<148 * 8, 256>>>(out, 16, 1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<148 * 8, 256>>>(out, n, 1);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = double(148) * 8 * 256 * n * 64;  // thread-level instructions
  int mhz = 0;
  cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
  std::printf("%-6s %8.3f ms  %.1f thread-instr / clk / SM (at %d MHz nominal)\n", name, ms, ops / (ms * 1e-3) / (mhz * 1e3) / 148, mhz / 1000);
  cudaFree(out);
}

int main() {
  run<0>("IDP4A");
  run<1>("IMAD");
  run<2>("PRMT");
  run<3>("LOP3");
  return 0;
}
