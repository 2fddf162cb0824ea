// Running mean/std normaliser kernels (normalizer.py:5-70, ddpg_agent.py:163-171).
//
// numpy semantics that are reproduced bit-for-bit:
//   update():   local_sum (f32) += v.sum(axis=0) (f64, rows added sequentially)  — the f32
//               accumulator is promoted to f64, added, rounded once back to f32
//               (ufunc add, casting='same_kind'); same for np.square(v).sum(axis=0);
//   recompute_stats(): all float32 arithmetic (numpy 1.19 value-based casting keeps
//               np.maximum(np.square(eps), x) in float32 — the reference pins numpy 1.19.2);
//   normalize(): float64 arithmetic on (v - mean)/std, clip.
#include "common.cuh"

namespace bmi {

// one thread per column; rows are consumed in order so the sum matches numpy exactly.
template <typename T>
__global__ void norm_update_kernel(const T* __restrict__ vin, int64_t n_rows, int size, double clip,
                                   float* __restrict__ lsum, float* __restrict__ lsumsq,
                                   float* __restrict__ lcount) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  struct Clipped {  // np.clip(v, -clip, clip) on load
    const T* p; double c;
    __device__ double operator[](int64_t i) const { return fmin(fmax((double)p[i], -c), c); }
  } v{vin, clip};
  if (j < size && n_rows > 0) {
    double s = (double)v[j];
    double q = __dmul_rn(s, s);
    int64_t i = 1;
    for (; i + 4 <= n_rows; i += 4) {  // 4 loads in flight, adds stay sequential
      double a0 = (double)v[(i + 0) * size + j], a1 = (double)v[(i + 1) * size + j];
      double a2 = (double)v[(i + 2) * size + j], a3 = (double)v[(i + 3) * size + j];
      s = __dadd_rn(s, a0); q = __dadd_rn(q, __dmul_rn(a0, a0));
      s = __dadd_rn(s, a1); q = __dadd_rn(q, __dmul_rn(a1, a1));
      s = __dadd_rn(s, a2); q = __dadd_rn(q, __dmul_rn(a2, a2));
      s = __dadd_rn(s, a3); q = __dadd_rn(q, __dmul_rn(a3, a3));
    }
    for (; i < n_rows; ++i) {
      double a = (double)v[i * size + j];
      s = __dadd_rn(s, a);
      q = __dadd_rn(q, __dmul_rn(a, a));
    }
    lsum[j] = (float)__dadd_rn((double)lsum[j], s);
    lsumsq[j] = (float)__dadd_rn((double)lsumsq[j], q);
  }
  if (j == 0) lcount[0] = __fadd_rn(lcount[0], (float)n_rows);  // f32 += python int
}

// Large inputs (the vectorised agent feeds 2e5 rows per cycle; the reference feeds 100): rows are split into
// chunks of kNormChunk rows, each chunk is summed sequentially, and the chunk sums are added sequentially in
// chunk order — a fixed, documented order that oracle/learner_oracle.py restates.  n_rows <= kNormChunk keeps
// numpy's exact order (one chunk).
constexpr int kNormChunk = 1024;

template <typename T>
__global__ void norm_chunk_kernel(const T* __restrict__ vin, int64_t n_rows, int size, double clip,
                                  double* __restrict__ part) {  // part[chunk][2][size]
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t c = blockIdx.y;
  if (j >= size) return;
  const int64_t r0 = c * kNormChunk, r1 = min(n_rows, r0 + (int64_t)kNormChunk);
  auto ld = [&](int64_t i) { return fmin(fmax((double)vin[i * size + j], -clip), clip); };
  double s = ld(r0);
  double q = __dmul_rn(s, s);
  for (int64_t i = r0 + 1; i < r1; ++i) {
    const double a = ld(i);
    s = __dadd_rn(s, a);
    q = __dadd_rn(q, __dmul_rn(a, a));
  }
  part[(c * 2 + 0) * size + j] = s;
  part[(c * 2 + 1) * size + j] = q;
}

__global__ void norm_chunk_finish_kernel(const double* __restrict__ part, int64_t n_chunks, int64_t n_rows, int size,
                                         float* __restrict__ lsum, float* __restrict__ lsumsq,
                                         float* __restrict__ lcount) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < size) {
    double s = part[j], q = part[size + j];
    for (int64_t c = 1; c < n_chunks; ++c) {
      s = __dadd_rn(s, part[(c * 2 + 0) * size + j]);
      q = __dadd_rn(q, part[(c * 2 + 1) * size + j]);
    }
    lsum[j] = (float)__dadd_rn((double)lsum[j], s);
    lsumsq[j] = (float)__dadd_rn((double)lsumsq[j], q);
  }
  if (j == 0) lcount[0] = __fadd_rn(lcount[0], (float)n_rows);
}

static double* g_norm_scratch = nullptr;
static size_t g_norm_scratch_bytes = 0;

__global__ void norm_recompute_kernel(float* lsum, float* lsumsq, float* lcount, float* tsum,
                                      float* tsumsq, float* tcount, float* mean, float* stdv,
                                      int size, float eps, float world) {
  __shared__ float cnt;
  if (threadIdx.x == 0) {
    // _mpi_average: buf /= size (normalizer.py:63), a float32 true divide
    float c = __fdiv_rn(lcount[0], world);
    tcount[0] = __fadd_rn(tcount[0], c);
    lcount[0] = 0.0f;
    cnt = tcount[0];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < size; j += blockDim.x) {
    float s = lsum[j], q = lsumsq[j];
    s = __fdiv_rn(s, world);
    q = __fdiv_rn(q, world);
    float ts = __fadd_rn(tsum[j], s);
    float tq = __fadd_rn(tsumsq[j], q);
    tsum[j] = ts;
    tsumsq[j] = tq;
    lsum[j] = 0.0f;
    lsumsq[j] = 0.0f;
    float m = __fdiv_rn(ts, cnt);
    mean[j] = m;
    float var = __fsub_rn(__fdiv_rn(tq, cnt), __fmul_rn(m, m));
    float e2 = (float)((double)eps * (double)eps);  // np.square(python float) -> f64 -> f32
    stdv[j] = __fsqrt_rn(fmaxf(e2, var));
  }
}

template <typename T, typename TO>
__global__ void norm_normalize_kernel(const T* __restrict__ v, int64_t n, int size,
                                      const float* __restrict__ mean,
                                      const float* __restrict__ stdv, double clip,
                                      TO* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * size) return;
  int j = (int)(i % size);
  double z = __ddiv_rn(__dsub_rn((double)v[i], (double)mean[j]), (double)stdv[j]);
  out[i] = (TO)fmin(fmax(z, -clip), clip);
}

template <typename T>
__global__ void preproc_inputs_kernel(const T* __restrict__ obs, const T* __restrict__ g, int64_t n,
                                      int Do, int Dg, const float* __restrict__ om,
                                      const float* __restrict__ os, const float* __restrict__ gm,
                                      const float* __restrict__ gs, double clip,
                                      float* __restrict__ x) {
  const int Dx = Do + Dg;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * Dx) return;
  int64_t row = i / Dx;
  int j = (int)(i % Dx);
  double v, m, s;
  if (j < Do) {
    v = (double)obs[row * Do + j]; m = (double)om[j]; s = (double)os[j];
  } else {
    v = (double)g[row * Dg + (j - Do)]; m = (double)gm[j - Do]; s = (double)gs[j - Do];
  }
  double z = __ddiv_rn(__dsub_rn(v, m), s);
  x[i] = (float)fmin(fmax(z, -clip), clip);
}

}  // namespace bmi

using namespace bmi;

extern "C" int bmi_norm_update(const void* v, int64_t n_rows, int32_t size, int32_t dtype, double pre_clip,
                               float* lsum, float* lsumsq, float* lcount, bmi_stream_t stream) {
  const double clip = pre_clip > 0.0 ? pre_clip : INFINITY;
  BMI_REQUIRE(size > 0 && n_rows >= 0, "bmi_norm_update: bad sizes");
  BMI_REQUIRE(dtype == BMI_F32 || dtype == BMI_F64, "bmi_norm_update: bad dtype %d", dtype);
  BMI_REQUIRE(lsum && lsumsq && lcount && (v || n_rows == 0), "bmi_norm_update: null pointer");
  unsigned grid = (unsigned)((size + 31) / 32);
  if (n_rows > kNormChunk) {  // chunked deterministic order (see norm_chunk_kernel)
    const int64_t n_chunks = (n_rows + kNormChunk - 1) / kNormChunk;
    const size_t need = (size_t)n_chunks * 2 * size * sizeof(double);
    if (need > g_norm_scratch_bytes) {  // grows rarely; not legal inside a stream capture (update() is never captured)
      if (g_norm_scratch) BMI_CUDA_CHECK(cudaFree(g_norm_scratch));
      g_norm_scratch = nullptr;
      g_norm_scratch_bytes = 0;
      BMI_CUDA_CHECK(cudaMalloc(&g_norm_scratch, need));
      g_norm_scratch_bytes = need;
    }
    dim3 g2(grid, (unsigned)n_chunks);
    if (dtype == BMI_F64)
      norm_chunk_kernel<double><<<g2, 32, 0, as_stream(stream)>>>((const double*)v, n_rows, size, clip, g_norm_scratch);
    else
      norm_chunk_kernel<float><<<g2, 32, 0, as_stream(stream)>>>((const float*)v, n_rows, size, clip, g_norm_scratch);
    BMI_LAUNCHED();
    norm_chunk_finish_kernel<<<grid, 32, 0, as_stream(stream)>>>(g_norm_scratch, n_chunks, n_rows, size, lsum, lsumsq, lcount);
    BMI_LAUNCHED();
    return BMI_OK;
  }
  if (dtype == BMI_F64)
    norm_update_kernel<double><<<grid, 32, 0, as_stream(stream)>>>((const double*)v, n_rows, size, clip,
                                                                   lsum, lsumsq, lcount);
  else
    norm_update_kernel<float><<<grid, 32, 0, as_stream(stream)>>>((const float*)v, n_rows, size, clip,
                                                                  lsum, lsumsq, lcount);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_norm_recompute(float* lsum, float* lsumsq, float* lcount, float* tsum,
                                  float* tsumsq, float* tcount, float* mean, float* stdv,
                                  int32_t size, float eps, float world, bmi_stream_t stream) {
  BMI_REQUIRE(size > 0 && world >= 1.0f, "bmi_norm_recompute: bad size/world");
  BMI_REQUIRE(lsum && lsumsq && lcount && tsum && tsumsq && tcount && mean && stdv,
              "bmi_norm_recompute: null pointer");
  norm_recompute_kernel<<<1, 64, 0, as_stream(stream)>>>(lsum, lsumsq, lcount, tsum, tsumsq, tcount,
                                                         mean, stdv, size, eps, world);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_norm_normalize(const void* v, int64_t n_rows, int32_t size, int32_t dtype,
                                  const float* mean, const float* stdv, double clip, void* out,
                                  int32_t out_dtype, bmi_stream_t stream) {
  BMI_REQUIRE(size > 0 && n_rows >= 0, "bmi_norm_normalize: bad sizes");
  BMI_REQUIRE((dtype == BMI_F32 || dtype == BMI_F64) && (out_dtype == BMI_F32 || out_dtype == BMI_F64),
              "bmi_norm_normalize: bad dtype");
  if (n_rows == 0) return BMI_OK;
  BMI_REQUIRE(v && mean && stdv && out, "bmi_norm_normalize: null pointer");
  unsigned grid = (unsigned)((n_rows * size + 255) / 256);
  cudaStream_t st = as_stream(stream);
  if (dtype == BMI_F64 && out_dtype == BMI_F64)
    norm_normalize_kernel<double, double><<<grid, 256, 0, st>>>((const double*)v, n_rows, size, mean, stdv, clip, (double*)out);
  else if (dtype == BMI_F64)
    norm_normalize_kernel<double, float><<<grid, 256, 0, st>>>((const double*)v, n_rows, size, mean, stdv, clip, (float*)out);
  else if (out_dtype == BMI_F64)
    norm_normalize_kernel<float, double><<<grid, 256, 0, st>>>((const float*)v, n_rows, size, mean, stdv, clip, (double*)out);
  else
    norm_normalize_kernel<float, float><<<grid, 256, 0, st>>>((const float*)v, n_rows, size, mean, stdv, clip, (float*)out);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_preproc_inputs(const void* obs, const void* g, int64_t n, int32_t Do, int32_t Dg,
                                  int32_t dtype, const float* om, const float* os, const float* gm,
                                  const float* gs, double clip, float* x, bmi_stream_t stream) {
  BMI_REQUIRE(n >= 0 && Do > 0 && Dg > 0, "bmi_preproc_inputs: bad sizes");
  BMI_REQUIRE(dtype == BMI_F32 || dtype == BMI_F64, "bmi_preproc_inputs: bad dtype %d", dtype);
  if (n == 0) return BMI_OK;
  BMI_REQUIRE(obs && g && om && os && gm && gs && x, "bmi_preproc_inputs: null pointer");
  unsigned grid = (unsigned)((n * (Do + Dg) + 255) / 256);
  if (dtype == BMI_F64)
    preproc_inputs_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)obs, (const double*)g, n, Do, Dg, om, os, gm, gs, clip, x);
  else
    preproc_inputs_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)obs, (const float*)g, n, Do, Dg, om, os, gm, gs, clip, x);
  BMI_LAUNCHED();
  return BMI_OK;
}
