// Which hardware warp slots (%warpid) share an SM sub-partition (scheduler)?  One block of 32 warps on one SM; warps 0 and j
// run an issue-bound loop of independent FMAs together: if they share a scheduler the pair takes ~2x the solo time.
// Also: how fast does ONE warp run a dependent FMA chain / an LDS->FMA chain (cycles per instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/smsp_map tools/microbench/smsp_map.cu && /tmp/smsp_map
#include <cstdio>
#include <cuda_runtime.h>

__global__ void pair_kernel(int j, long long* out, unsigned* wid_out, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned wid;
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if (lane == 0) wid_out[warp] = wid;
  __syncthreads();
  if (warp != 0 && warp != j) return;
  float a0 = lane, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
  const float m = 1.0001f, c = 0.5f;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
      a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && warp == 0) out[j] = t1 - t0;
  sink[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void chain_kernel(long long* out, float* sink) {
  __shared__ float sm[1024];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  float a = lane;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a = fmaf(a, 1.0001f, 0.5f);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;  // 65536 dependent FMAs
  // dependent LDS -> FMA -> address chain
  int idx = lane;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v = sm[idx & 1023];
      a = fmaf(a, v, 0.5f);
      idx = idx + 33 + (a > 1e30f ? 1 : 0);
    }
  }
  t1 = clock64();
  if (threadIdx.x == 0) out[1] = t1 - t0;  // 16384 x (LDS + FMA + ...)
  sink[threadIdx.x] = a + idx;
}

int main() {
  long long* out; unsigned* wid; float* sink;
  cudaMallocManaged(&out, 64 * sizeof(long long));
  cudaMallocManaged(&wid, 64 * sizeof(unsigned));
  cudaMalloc(&sink, 4096);
  for (int j = 0; j < 32; ++j) {
    pair_kernel<<<1, 1024>>>(j, out, wid, sink);
    cudaDeviceSynchronize();
  }
  printf("warpid of warps: ");
  for (int w = 0; w < 32; ++w) printf("%u ", wid[w]);
  printf("\nsolo (j=0): %lld cycles for 131072 FMAs per lane-warp\n", out[0]);
  for (int j = 1; j < 32; ++j) printf("pair (0,%2d): %.2fx solo\n", j, (double)out[j] / out[0]);
  chain_kernel<<<1, 32>>>(out + 40, sink);
  cudaDeviceSynchronize();
  printf("dependent FMA chain: %.2f cycles/FMA\n", out[40] / 65536.0);
  printf("dependent LDS->FMA->addr chain: %.2f cycles/step\n", out[41] / 16384.0);
  printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
