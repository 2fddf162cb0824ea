// Shared helpers for libbmi_b200: error plumbing, launch accounting, Philox4x32-10.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/bmi.h"

namespace bmi {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline cudaStream_t as_stream(bmi_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define BMI_CUDA_CHECK(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::bmi::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return BMI_ERR_CUDA;                                                            \
    }                                                                                 \
  } while (0)

#define BMI_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::bmi::set_error(__VA_ARGS__);      \
      return BMI_ERR_ARG;                 \
    }                                     \
  } while (0)

// call after every <<<>>> launch: counts it and surfaces launch-configuration errors
#define BMI_LAUNCHED()                                   \
  do {                                                   \
    ::bmi::g_launches.fetch_add(1, std::memory_order_relaxed); \
    BMI_CUDA_CHECK(cudaGetLastError());                  \
  } while (0)

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter-based: output block = f(key, counter).
// oracle/philox.py restates exactly this function in numpy.
// ---------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t v[4];
};

__host__ __device__ inline Philox4 philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32);
  uint32_t c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c0;
    uint64_t p1 = (uint64_t)M1 * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// 53-bit uniform in [0,1) from two 32-bit words, numpy's random_sample construction:
// (a >> 5) * 2^26 + (b >> 6), divided by 2^53.
__host__ __device__ inline double u53(uint32_t a, uint32_t b) {
  return (double)(((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// 24-bit uniform in [0,1) as float
__host__ __device__ inline float u24(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }

// distinct Philox "hi" counter words per consumer so streams never overlap
enum : uint64_t { kStreamHer = 1, kStreamExplore = 2, kStreamReset = 3 };

}  // namespace bmi
