// Prints (%smid, %warpid) of every warp of the CTAs that land on SM 0 and 1 (2 CTAs of 4 warps per SM).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint32_t* out) {
  extern __shared__ uint32_t pad[];
  uint32_t smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2] = smid; out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + 1] = wid; }
  // keep the CTA alive for a while so that 2 CTAs per SM are co-resident
  long long t0 = clock64(); while (clock64() - t0 < 2000000) {}
}
int main(int argc, char** argv) {
  const int warps = argc > 1 ? atoi(argv[1]) : 4, per_sm = argc > 2 ? atoi(argv[2]) : 2;
  const int blocks = 148 * per_sm;
  uint32_t* out; cudaMalloc(&out, blocks * 8 * 2 * 4);
  cudaMemset(out, 0xff, blocks * 8 * 2 * 4);
  const int smem = 200 * 1024 / per_sm;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<blocks, warps * 32, smem>>>(out); cudaDeviceSynchronize();
  static uint32_t h[148 * 8 * 8 * 2];
  cudaMemcpy(h, out, blocks * 8 * 2 * 4, cudaMemcpyDeviceToHost);
  for (int b = 0; b < blocks; b++) if (h[b * 16] < 2) { printf("cta %3d sm %u warpids:", b, h[b * 16]); for (int w = 0; w < warps; w++) printf(" %u", h[(b * 8 + w) * 2 + 1]); printf("\n"); }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
