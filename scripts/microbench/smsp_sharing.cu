// Does a latency-bound chain warp slow down when it shares its SM sub-partition with throughput warps,
// and do warps w of a 4-warp CTA land on sub-partition w % 4?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smsp_sharing smsp_sharing.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr uint32_t M = 0x00ff00ffu;
constexpr int N = 1 << 14;

__device__ __forceinline__ void chain(const uint32_t* in, uint32_t* out, long long* cyc) {
  uint32_t c[16], n[16];
  for (int i = 0; i < 16; i++) { c[i] = (in[(threadIdx.x + i) & 63] & M) + 0x01000100u; n[i] = in[64 + ((threadIdx.x + 3 * i) & 63)] & M; }
  uint32_t w = in[threadIdx.x & 63] & M, nw = n[15];
  long long t0 = clock64();
  for (int it = 0; it < N / 16; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      w = (c[i] + __vimin3_u16x2(n[i], w, nw) + __vimax3_u16x2(n[i], w, nw)) & M;
      nw = n[i];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = w;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 4 + (threadIdx.x >> 5)] = t1 - t0;
}
__device__ __forceinline__ void heavy(const uint32_t* in, uint32_t* out, int iters) {
  uint32_t a[8];
  for (int i = 0; i < 8; i++) a[i] = in[(threadIdx.x + 7 * i) & 63];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = __byte_perm(a[i], a[(i + 1) & 7], 0x5140) ^ (a[(i + 3) & 7] & 0x7f7f7f7fu);
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mode 0: warps [c,c,h,h]   mode 1: even CTAs [c,h,c,h], odd CTAs [h,c,h,c]   mode 2: chain warps only (w 0,1)
// mode 3: [c,h,h,h]
__global__ void k(const uint32_t* in, uint32_t* out, long long* cyc, int mode, int heavy_iters) {
  extern __shared__ uint32_t pad[];
  const int w = threadIdx.x >> 5;
  bool is_chain;
  if (mode == 0 || mode == 2) is_chain = w < 2;
  else if (mode == 1) is_chain = ((w + blockIdx.x) & 1) == 0;
  else is_chain = w == 0;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 4 + w] = 0;
  if (is_chain) chain(in, out, cyc);
  else if (mode != 2) heavy(in, out, heavy_iters);
}
int main() {
  uint32_t *in, *out; long long* cyc;
  const int blocks = 148 * 2;
  cudaMalloc(&in, 4096); cudaMalloc(&out, blocks * 128 * 4); cudaMalloc(&cyc, blocks * 4 * 8);
  uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = 2654435761u * (i + 1);
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[] = {"[c,c,h,h] all CTAs (chains together)", "[c,h,c,h]/[h,c,h,c] (chain next to heavy)", "chains only", "[c,h,h,h]"};
  static long long hc[blocks * 4];
  for (int mode = 0; mode < 4; mode++) {
    for (int rep = 0; rep < 2; rep++) { k<<<blocks, 128, 100 * 1024>>>(in, out, cyc, mode, 6000); cudaDeviceSynchronize(); }
    cudaMemcpy(hc, cyc, sizeof hc, cudaMemcpyDeviceToHost);
    double s = 0; int cnt = 0; long long mx = 0;
    for (int i = 0; i < blocks * 4; i++) if (hc[i]) { s += hc[i]; cnt++; if (hc[i] > mx) mx = hc[i]; }
    printf("%-45s chain warps %4d  avg %.2f  max %.2f cycles/step\n", names[mode], cnt, s / cnt / N, (double)mx / N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
