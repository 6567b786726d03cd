// Dependent-issue latency of the inverse-ClampedGradient chain step variants on one warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_latency chain_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int N = 4096;
constexpr uint32_t M = 0x00ff00ffu;

template <int VAR>
__global__ void k(const uint32_t* in, uint32_t* out, long long* cyc) {
  __shared__ uint32_t cs[64], ns[64];
  if (threadIdx.x < 64) { cs[threadIdx.x] = in[threadIdx.x] & M; ns[threadIdx.x] = in[64 + threadIdx.x] & M; }
  __syncthreads();
  uint32_t c[16], n[16];
  for (int i = 0; i < 16; i++) { c[i] = cs[(threadIdx.x + i) & 63] + 0x01000100u; n[i] = ns[(threadIdx.x + 3 * i) & 63]; }
  uint32_t w = in[threadIdx.x] & M, nw = n[15], w2 = (in[threadIdx.x] >> 8) & M, nw2 = n[14];
  long long t0 = clock64();
  for (int it = 0; it < N / 16; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (VAR == 0) {          // packed u16x2: min3, max3, add, mask
        w = (c[i] + __vimin3_u16x2(n[i], w, nw) + __vimax3_u16x2(n[i], w, nw)) & M;
        nw = n[i];
      } else if (VAR == 1) {   // two interleaved chains
        w = (c[i] + __vimin3_u16x2(n[i], w, nw) + __vimax3_u16x2(n[i], w, nw)) & M;
        w2 = (c[15 - i] + __vimin3_u16x2(n[15 - i], w2, nw2) + __vimax3_u16x2(n[15 - i], w2, nw2)) & M;
        nw = n[i]; nw2 = n[15 - i];
      } else if (VAR == 2) {   // 2-input min / max with precomputed lo, hi (c holds lo, n holds hi here)
        w = (c[i] + __vminu2(w, n[i]) + __vmaxu2(w, c[(i + 1) & 15])) & M;
      } else if (VAR == 3) {   // 32-bit top-byte form: min3, max3, add (no mask)
        w = c[i] + __vimin3_u32(n[i], w, nw) + __vimax3_u32(n[i], w, nw);
        nw = n[i];
      } else if (VAR == 4) {   // plain 32-bit min/max (2-input)
        w = c[i] + min(w, n[i]) + max(w, c[(i + 1) & 15]);
      } else if (VAR == 5) {   // add + mask only
        w = (c[i] + w + nw) & M;
      } else if (VAR == 6) {   // single LOP3 chain
        w = (w ^ c[i]) & n[i];
      } else if (VAR == 7) {   // single VIMNMX3 chain
        w = __vimin3_u16x2(n[i], w, c[i]);
      } else if (VAR == 8) {   // single IADD3 chain
        w = w + c[i] + n[i];
      }
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = w + w2;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  uint32_t *in, *out; long long* cyc;
  cudaMalloc(&in, 4096); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  uint32_t h[1024]; for (int i = 0; i < 1024; i++) h[i] = 2654435761u * (i + 1);
  cudaMemcpy(in, h, 4096, cudaMemcpyHostToDevice);
  const char* names[] = {"u16x2 min3+max3+iadd3+lop", "two interleaved chains (per step of both)", "u16x2 2-input min+max+iadd3+lop",
                         "u32 min3+max3+iadd3", "u32 2-input min+max+iadd3", "iadd3+lop", "lop3", "vimnmx3.u16x2", "iadd3"};
  for (int v = 0; v < 9; v++) {
    for (int rep = 0; rep < 2; rep++) {
      switch (v) {
        case 0: k<0><<<1, 32>>>(in, out, cyc); break; case 1: k<1><<<1, 32>>>(in, out, cyc); break;
        case 2: k<2><<<1, 32>>>(in, out, cyc); break; case 3: k<3><<<1, 32>>>(in, out, cyc); break;
        case 4: k<4><<<1, 32>>>(in, out, cyc); break; case 5: k<5><<<1, 32>>>(in, out, cyc); break;
        case 6: k<6><<<1, 32>>>(in, out, cyc); break; case 7: k<7><<<1, 32>>>(in, out, cyc); break;
        default: k<8><<<1, 32>>>(in, out, cyc); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-45s %6.2f cycles/step\n", names[v], (double)c / N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
