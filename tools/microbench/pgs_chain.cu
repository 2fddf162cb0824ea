// Latency / issue microbenchmark for the building blocks of the constraint-space PGS loop (csrc/physics.cu substep_solve):
// dependent SHFL.IDX, dependent LDS, the motor-event chain (FFMA -> FADD -> FMNMX x2 -> FADD -> SHFL -> FFMA with one LDS),
// measured for 1..8 warps per SM sub-partition.  One block of 32 * W warps per SM; every warp runs the same chain N times.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/pgs_chain.bin tools/microbench/pgs_chain.cu
//   tools/microbench/pgs_chain.bin            (prints cycles per chain element for each warp count)
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

template <int KIND>
__global__ void chain_kernel(float* out, long long* cycles, int iters) {
  __shared__ float tab[32 * 64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) tab[i] = 1.0f / (1 + (i % 97));
  __syncthreads();
  float v = 0.001f * lane, lam = 0.f, acc = 0.f;
  const float inv = 0.5f, rhs = 0.25f;
  unsigned addr = (unsigned)__cvta_generic_to_shared(tab) + 4u * lane;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 4
    for (int j = 0; j < 8; ++j) {
      if (KIND == 0) {            // dependent shuffle
        v = __shfl_sync(0xffffffffu, v, (lane + j) & 31) + 1.0f;
      } else if (KIND == 1) {     // dependent shared-memory load (address depends on the previous value)
        float x;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr + ((__float_as_uint(v) & 7u) << 7)));
        v = x + 1.0e-6f * j;
      } else if (KIND == 2) {     // dependent FFMA / FMNMX chain of one row update, no communication
        const float cand = fminf(fmaxf(lam + fmaf(-v, inv, rhs), -2.f), 2.f);
        const float d = cand - lam;
        lam = cand;
        v = fmaf(0.3f, d, v);
      } else {                    // the motor event: candidate, shuffle broadcast from the owner lane, coefficient load, update
        float c;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c) : "r"(addr + 128u * j));
        const float cand = fminf(fmaxf(lam + fmaf(-v, inv, rhs), -2.f), 2.f);
        const float d = cand - lam;
        const float dj = __shfl_sync(0xffffffffu, d, j);
        if (lane == j) { lam = cand; acc = d; }
        v = fmaf(c, dj, v);
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x * (blockDim.x / 32) + warp] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + lam + acc;
}

template <int KIND>
static void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * 32 * sizeof(long long));
  printf("%-34s", name);
  for (int wps = 1; wps <= 8; ++wps) {        // warps per sub-partition
    const int warps = 4 * wps;
    chain_kernel<KIND><<<148, 32 * warps>>>(out, cyc, 16);          // warm-up (I-cache)
    chain_kernel<KIND><<<148, 32 * warps>>>(out, cyc, N / 8);
    cudaDeviceSynchronize();
    long long h[32];
    cudaMemcpy(h, cyc, warps * sizeof(long long), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int w = 0; w < warps; ++w) s += (double)h[w];
    printf(" %6.1f", s / warps / N);
  }
  printf("   cycles per element, 1..8 warps per sub-partition\n");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("dependent SHFL.IDX (+FADD)");
  run<1>("dependent LDS (+FADD)");
  run<2>("row update chain, no communication");
  run<3>("motor event (cand + SHFL + LDS + FFMA)");
  return 0;
}
