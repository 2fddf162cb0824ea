// DDPG learner: the update as two hand-written kernels (ddpg_fused.cuh: the default for the reference's shapes) or, for
// other hidden widths / BMI_DDPG_CUBLAS=1, as a chain of cuBLASLt fp32 GEMMs (bias+ReLU fused in the GEMM epilogue) with fused
// loss / head / dReLU+bias-grad kernels; the policy forward (bmi_ddpg_act); multi-tensor Adam on flat parameter buffers, alone
// or fused with the gradient sum over ranks through NVLink peer memory; Polyak averaging.
//
// Reference behaviour restated (paths relative to the reference tree):
//   models.py:11-44            actor / critic
//   ddpg_agent.py:250-277      _update_network (targets, losses, two backward passes, Adam)
//   ddpg_agent.py:220-222      _soft_update_target_network
//   ddpg_agent.py:174-184      _select_actions
//   utils.py:18-27             flat parameter order (named_parameters)
//
// Row-major convention: every activation is [rows][features]; torch Linear weights are
// [out][in].  gemm_rm() maps C[M][N] = op(A) op(B) onto column-major cuBLASLt by swapping the
// operands (C^T = op(B)^T op(A)^T), so a per-feature bias is a cuBLASLt row-bias epilogue.
#include <cublasLt.h>

#include <map>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "ddpg_fused.cuh"

namespace bmi {

#define BMI_CUBLAS_CHECK(expr)                                                         \
  do {                                                                                 \
    cublasStatus_t _s = (expr);                                                        \
    if (_s != CUBLAS_STATUS_SUCCESS) {                                                 \
      ::bmi::set_error("%s:%d %s -> cublas status %d", __FILE__, __LINE__, #expr, (int)_s); \
      return BMI_ERR_CUBLAS;                                                           \
    }                                                                                  \
  } while (0)

struct GemmKey {
  int opA, opB, M, N, K, lda, ldb, ldc, epi, alA, alB, alC;
  bool operator<(const GemmKey& o) const {
    return std::tie(opA, opB, M, N, K, lda, ldb, ldc, epi, alA, alB, alC) <
           std::tie(o.opA, o.opB, o.M, o.N, o.K, o.lda, o.ldb, o.ldc, o.epi, o.alA, o.alB, o.alC);
  }
};

// largest power of two (<= 256) dividing the base address (cuBLASLt sees the pitch itself)
static inline int ptr_align(const void* p) {
  uintptr_t v = (uintptr_t)p | 256u;
  return (int)(v & (~v + 1));
}

struct GemmPlan {
  cublasLtMatmulDesc_t desc = nullptr;
  cublasLtMatrixLayout_t la = nullptr, lb = nullptr, lc = nullptr;
  cublasLtMatmulAlgo_t algo;
};

enum { EPI_NONE = 0, EPI_BIAS = 1, EPI_RELU_BIAS = 2 };

// layer parameter offsets inside a flat buffer (named_parameters order)
struct NetLayout {
  int in_dim, hidden, out_dim;
  int64_t w[4], b[4];
  int64_t count;
  void init(int in_d, int hid, int out_d) {
    in_dim = in_d; hidden = hid; out_dim = out_d;
    int ins[4] = {in_d, hid, hid, hid};
    int outs[4] = {hid, hid, hid, out_d};
    int64_t off = 0;
    for (int l = 0; l < 4; ++l) {
      w[l] = off; off += (int64_t)outs[l] * ins[l];
      b[l] = off; off += outs[l];
    }
    count = off;
  }
  int fan_in(int l) const { return l == 0 ? in_dim : hidden; }
  int fan_out(int l) const { return l == 3 ? out_dim : hidden; }
};

}  // namespace bmi

using namespace bmi;

struct bmi_ddpg {
  bmi_ddpg_config cfg;
  NetLayout la, lc;  // actor, critic
  float *actor, *critic, *actor_t, *critic_t;  // caller-owned flat parameters
  float* grads = nullptr;                       // [actor | pad to 64 floats | critic]
  int64_t na_pad = 0, n_grads = 0;
  float *adam_m = nullptr, *adam_v = nullptr;   // same layout as grads
  int* adam_step = nullptr;                     // device step counter
  float* adam_scal = nullptr;                   // two slots of [step_size_a, step_size_c, sqrt(bc2), -]: slot (t & 1) for step t
  cublasLtHandle_t lt = nullptr;
  void* workspace = nullptr;
  size_t workspace_bytes = 8u << 20;
  std::map<GemmKey, GemmPlan> plans;
  // training activations (batch rows)
  float *xc = nullptr, *h1 = nullptr, *h2 = nullptr, *h3 = nullptr;        // scratch chain
  float *ch1 = nullptr, *ch2 = nullptr, *ch3 = nullptr, *q = nullptr;      // critic(x, A)
  float *ah1 = nullptr, *ah2 = nullptr, *ah3 = nullptr, *az = nullptr, *aa = nullptr;  // actor(x)
  float *qh1 = nullptr, *qh2 = nullptr, *qh3 = nullptr, *qa = nullptr, *xca = nullptr; // critic(x, pi(x))
  float *a_next = nullptr, *q_next = nullptr, *y = nullptr;
  float *d1 = nullptr, *d2 = nullptr, *dq = nullptr, *dq_const = nullptr, *dxc = nullptr, *dz = nullptr;    // backward scratch
  // hand-written two-launch update (ddpg_fused.cuh): chosen at creation when the shapes fit (hidden 256, ...)
  bool fused = false;
  float *cd1 = nullptr, *cd2 = nullptr, *cd3 = nullptr, *fd1 = nullptr, *fd2 = nullptr, *fd3 = nullptr, *loss_part = nullptr;
  float* bg_part = nullptr;        // drelu_bgrad: partial column sums [row chunks][hidden]
  unsigned* bg_ticket = nullptr;   // ... and one ticket counter per 32-column group
  unsigned* adam_ticket = nullptr; // adam_kernel: blocks finished (the last one advances the step counter)
  // policy (act) activations (max_act_rows)
  float *ph1 = nullptr, *ph2 = nullptr, *pz = nullptr;
  std::vector<void*> owned;
  // ---- peer-memory gradient exchange (fused sum-over-ranks + Adam over NVLink) ----
  int p2p_rank = 0, p2p_world = 1;
  int* p2p_sync = nullptr;                 // [0..7] ready flags, [8..15] done flags, [16] epoch, [17] block counter,
                                           // [18] local go flag, [19] timeout flag   (this rank's copy, IPC-exported)
  float* peer_grads[8] = {nullptr};        // peer_grads[q] = rank q's gradient buffer mapped here (own buffer for q == rank)
  int* peer_sync[8] = {nullptr};
  std::vector<void*> ipc_opened;
};

namespace bmi {

static int dalloc(bmi_ddpg* h, float** p, size_t n) {
  void* q = nullptr;
  BMI_CUDA_CHECK(cudaMalloc(&q, n * sizeof(float)));
  BMI_CUDA_CHECK(cudaMemset(q, 0, n * sizeof(float)));
  h->owned.push_back(q);
  *p = (float*)q;
  return BMI_OK;
}

// C[M][N] (row-major, ldc) = op(A)[M][K] * op(B)[K][N]  (+ bias[N], ReLU)
static int gemm_rm(bmi_ddpg* h, cudaStream_t st, int opA, int opB, int M, int N, int K,
                   const float* A, int lda, const float* B, int ldb, float* C, int ldc, int epi,
                   const float* bias) {
  GemmKey key{opA, opB, M, N, K, lda, ldb, ldc, epi, ptr_align(A), ptr_align(B), ptr_align(C)};
  auto it = h->plans.find(key);
  if (it == h->plans.end()) {
    GemmPlan p;
    BMI_CUBLAS_CHECK(cublasLtMatmulDescCreate(&p.desc, CUBLAS_COMPUTE_32F, CUDA_R_32F));
    // column-major call: m = N, n = M, first operand = B, second = A
    cublasOperation_t ta = opB ? CUBLAS_OP_T : CUBLAS_OP_N;
    cublasOperation_t tb = opA ? CUBLAS_OP_T : CUBLAS_OP_N;
    BMI_CUBLAS_CHECK(cublasLtMatmulDescSetAttribute(p.desc, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof(ta)));
    BMI_CUBLAS_CHECK(cublasLtMatmulDescSetAttribute(p.desc, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof(tb)));
    cublasLtEpilogue_t e = epi == EPI_RELU_BIAS ? CUBLASLT_EPILOGUE_RELU_BIAS
                           : epi == EPI_BIAS    ? CUBLASLT_EPILOGUE_BIAS
                                                : CUBLASLT_EPILOGUE_DEFAULT;
    BMI_CUBLAS_CHECK(cublasLtMatmulDescSetAttribute(p.desc, CUBLASLT_MATMUL_DESC_EPILOGUE, &e, sizeof(e)));
    // stored (pre-op) shapes, column-major view of the row-major arrays
    // first operand (row-major B): opB==N -> stored [K][N] row-major == col-major N x K
    BMI_CUBLAS_CHECK(cublasLtMatrixLayoutCreate(&p.la, CUDA_R_32F, opB ? K : N, opB ? N : K, ldb));
    BMI_CUBLAS_CHECK(cublasLtMatrixLayoutCreate(&p.lb, CUDA_R_32F, opA ? M : K, opA ? K : M, lda));
    BMI_CUBLAS_CHECK(cublasLtMatrixLayoutCreate(&p.lc, CUDA_R_32F, N, M, ldc));
    // the bias pointer is part of the descriptor, but only its presence matters for the
    // heuristic; set a placeholder now and the real pointer on every call
    if (epi != EPI_NONE) {
      const void* bp = bias;
      BMI_CUBLAS_CHECK(cublasLtMatmulDescSetAttribute(p.desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bp, sizeof(bp)));
    }
    cublasLtMatmulPreference_t pref;
    BMI_CUBLAS_CHECK(cublasLtMatmulPreferenceCreate(&pref));
    BMI_CUBLAS_CHECK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES,
                                                          &h->workspace_bytes, sizeof(h->workspace_bytes)));
    // tell the heuristic how well the operands are really aligned (default assumes 256 B)
    uint32_t al_first = (uint32_t)key.alB, al_second = (uint32_t)key.alA, al_c = (uint32_t)key.alC;
    BMI_CUBLAS_CHECK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_A_BYTES, &al_first, sizeof(al_first)));
    BMI_CUBLAS_CHECK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_B_BYTES, &al_second, sizeof(al_second)));
    BMI_CUBLAS_CHECK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_C_BYTES, &al_c, sizeof(al_c)));
    BMI_CUBLAS_CHECK(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_D_BYTES, &al_c, sizeof(al_c)));
    cublasLtMatmulHeuristicResult_t res;
    int found = 0;
    cublasStatus_t s = cublasLtMatmulAlgoGetHeuristic(h->lt, p.desc, p.la, p.lb, p.lc, p.lc, pref, 1, &res, &found);
    cublasLtMatmulPreferenceDestroy(pref);
    if (s != CUBLAS_STATUS_SUCCESS || found == 0) {
      set_error("cublasLt heuristic found no algorithm for gemm opA=%d opB=%d M=%d N=%d K=%d epi=%d (status %d)",
                opA, opB, M, N, K, epi, (int)s);
      return BMI_ERR_CUBLAS;
    }
    p.algo = res.algo;
    it = h->plans.emplace(key, p).first;
  }
  GemmPlan& p = it->second;
  if (epi != EPI_NONE) {
    const void* bp = bias;
    BMI_CUBLAS_CHECK(cublasLtMatmulDescSetAttribute(p.desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bp, sizeof(bp)));
  }
  const float one = 1.0f, zero = 0.0f;
  BMI_CUBLAS_CHECK(cublasLtMatmul(h->lt, p.desc, &one, B, p.la, A, p.lb, &zero, C, p.lc, C, p.lc, &p.algo,
                                  h->workspace, h->workspace_bytes, st));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return BMI_OK;
}

// ---- elementwise / reduction kernels ------------------------------------------------------
// xc[r] = [x[r], a[r] / amax]
__global__ void concat_scale_kernel(const float* __restrict__ x, const float* __restrict__ a, int n,
                                    int Dx, int Da, float amax, float* __restrict__ xc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int D = Dx + Da;
  if (i >= n * D) return;
  int r = i / D, j = i % D;
  xc[i] = j < Dx ? x[r * Dx + j] : __fdiv_rn(a[r * Da + (j - Dx)], amax);
}

// a = amax tanh(z) and xc[r] = [x[r], a[r] / amax] in one pass (the actor head feeding a critic input)
__global__ void tanh_concat_kernel(const float* __restrict__ x, const float* __restrict__ z, int n, int Dx, int Da,
                                   float amax, float* __restrict__ a_out, float* __restrict__ xc) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int D = Dx + Da;
  if (i >= n * D) return;
  int r = i / D, j = i % D;
  if (j < Dx) { xc[i] = x[r * Dx + j]; return; }
  const float a = amax * tanhf(z[r * Da + (j - Dx)]);
  a_out[r * Da + (j - Dx)] = a;
  xc[i] = __fdiv_rn(a, amax);
}

__global__ void tanh_scale_kernel(const float* __restrict__ z, int n, float amax,
                                  float* __restrict__ a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = amax * tanhf(z[i]);
}

// y = clamp(r + gamma * q_next, -1/(1-gamma), 0)   (ddpg_agent.py:255-260)
__global__ void target_kernel(const float* __restrict__ r, const float* __restrict__ qn, int n,
                              float gamma, float clip_ret, float* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fminf(fmaxf(__fadd_rn(r[i], __fmul_rn(gamma, qn[i])), -clip_ret), 0.0f);
}

__device__ __forceinline__ float block_sum(float v, float* sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = l < (int)(blockDim.x >> 5) ? sm[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in warp 0
}

// target y = clamp(r + gamma q', -1/(1-gamma), 0) (ddpg_agent.py:255-260) fused with
// critic loss = mean((y - q)^2); dq = 2 (q - y) / B          single block
__global__ void critic_loss_kernel(const float* __restrict__ r, const float* __restrict__ qn, float gamma, float clip_ret,
                                   const float* __restrict__ q, int B, float* __restrict__ y_out,
                                   float* __restrict__ dq, float* __restrict__ loss) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float yi = fminf(fmaxf(__fadd_rn(r[i], __fmul_rn(gamma, qn[i])), -clip_ret), 0.0f);
    y_out[i] = yi;
    float d = yi - q[i];
    acc += d * d;
    dq[i] = -2.0f * d / (float)B;
  }
  float t = block_sum(acc, sm);
  if (threadIdx.x == 0) *loss = t / (float)B;
}

// actor loss = -mean(qa) + l2 * mean((a/amax)^2);  dqa = -1/B;
// dz = (da_q/amax + 2 l2 a / (amax^2 B Da)) * amax (1 - tanh^2),  tanh = a/amax
__global__ void actor_loss_kernel(const float* __restrict__ qa, const float* __restrict__ a,
                                  const float* __restrict__ dxc, int B, int Dx, int Da, float amax,
                                  float l2, float* __restrict__ dz, float* __restrict__ loss) {
  __shared__ float sm[32];
  float accq = 0.f, acca = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) accq += qa[i];
  const int n = B * Da;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int r = i / Da, j = i % Da;
    float th = a[i] / amax;
    acca += th * th;
    float da = dxc[r * (Dx + Da) + Dx + j] / amax + l2 * 2.0f * th / (amax * (float)n);
    dz[i] = da * amax * (1.0f - th * th);
  }
  float tq = block_sum(accq, sm);
  float ta = block_sum(acca, sm);
  if (threadIdx.x == 0) *loss = -tq / (float)B + l2 * ta / (float)n;
}

__global__ void fill_kernel(float* p, int n, float v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// dZ = dH * (H > 0) in place, db[j] = sum_r dZ[r][j].  Grid (cols / 32, rows / 32), block (32, 8): every thread owns four
// rows of one column (all four loads in flight before the first use); the block's column sums go to part[blockIdx.y][c]
// and the LAST block of a column group to finish (ticket counter) adds the partial sums in row-chunk order, so the result is
// deterministic (no float atomics).  Round 1: one block per 32 columns looping over all rows = 8 blocks in total and
// 21.8 us per launch in the ncu capture -- 9 launches = more than half of an update.
__global__ void drelu_bgrad_kernel(float* __restrict__ dH, const float* __restrict__ H, int rows,
                                   int cols, float* __restrict__ db, float* __restrict__ part, unsigned* __restrict__ ticket) {
  __shared__ float sm[8][33];
  __shared__ bool last;
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * 32 + threadIdx.y;
  float acc = 0.f;
  if (c < cols) {
    float g[4], h[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + 8 * k;
      g[k] = r < rows ? dH[(size_t)r * cols + c] : 0.f;
      h[k] = r < rows ? H[(size_t)r * cols + c] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + 8 * k;
      const float v = h[k] > 0.f ? g[k] : 0.f;
      if (r < rows) dH[(size_t)r * cols + c] = v;
      acc += v;
    }
  }
  if (db == nullptr) return;   // whole grid: no parameter gradients wanted
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    if (c < cols) part[(size_t)blockIdx.y * cols + c] = t;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned n = atomicAdd(ticket + blockIdx.x, 1u);
    last = n == gridDim.y - 1;
    if (last) ticket[blockIdx.x] = 0u;   // ready for the next launch (stream ordered)
  }
  __syncthreads();
  if (last && threadIdx.y == 0 && c < cols) {
    __threadfence();
    float t = 0.f;
    for (unsigned k = 0; k < gridDim.y; ++k) t += __ldcg(part + (size_t)k * cols + c);
    db[c] = t;
  }
}

// db[j] = sum_r dZ[r][j] (no mask) for the output layer
__global__ void bgrad_kernel(const float* __restrict__ dZ, int rows, int cols, float* __restrict__ db) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < cols) {
#pragma unroll 8
    for (int r = threadIdx.y; r < rows; r += 8) acc += dZ[(size_t)r * cols + c];   // unrolled: the loads are independent
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][threadIdx.x];
    db[c] = t;
  }
}

// bias-corrected step sizes of optimiser step t, as torch computes them (python doubles, then float32): scal[0..2].
// No separate one-thread launch: the scalars live in a two-slot ring (slot t & 1 for step t).  The launch of step t reads its
// slot -- written by the launch of step t - 1, or by the host for t = 1 -- while ONE thread of the grid evaluates the two
// pow() calls for step t + 1 off the critical path (about 4 us, as long as the whole Adam pass).  The step counter is advanced
// by the LAST block to finish (adam_kernel: ticket; adam_p2p_kernel: its "done" section), after every block has read it.
__host__ __device__ inline void adam_scalars(int t, float lr_a, float lr_c, float b1, float b2, float* scal) {
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  scal[0] = (float)((double)lr_a / bc1);
  scal[1] = (float)((double)lr_c / bc1);
  scal[2] = (float)sqrt(bc2);
}

// torch.optim.Adam (_single_tensor_adam) on the concatenated [actor | critic] buffers
__global__ void adam_kernel(float* __restrict__ pa, float* __restrict__ pc, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, int64_t na, int64_t na_pad,
                            int64_t n, int* step_ctr, unsigned* ticket, float* scal_ring, float lr_a, float lr_c, float b1,
                            float b2, float eps) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = *step_ctr + 1;                            // this launch is optimiser step t
  const float* scal = scal_ring + 4 * (t & 1);
  const int64_t writer = na < na_pad ? na : ((n & 255) ? n : 0);   // an idle thread of the grid if there is one
  if (i == writer) adam_scalars(t + 1, lr_a, lr_c, b1, b2, scal_ring + 4 * ((t + 1) & 1));
  const bool live = i < n && !(i >= na && i < na_pad);
  if (live) {
    const float gi = g[i], m0 = m[i], v0 = v[i];
    float mi = m0 + (gi - m0) * (1.0f - b1);            // exp_avg.lerp_(grad, 1-beta1)
    float vi = v0 * b2 + (1.0f - b2) * gi * gi;         // mul_(beta2).addcmul_(g, g, 1-beta2)
    m[i] = mi;
    v[i] = vi;
    float denom = sqrtf(vi) / scal[2] + eps;
    float step = i < na ? scal[0] : scal[1];
    float* p = i < na ? pa + i : pc + (i - na_pad);
    *p = *p - step * (mi / denom);
  }
  __syncthreads();                                        // the ring writer, if it is in this block
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {       // every block has read the counter: advance it
      *ticket = 0u;
      *step_ctr = *step_ctr + 1;
    }
  }
}

// ---- fused gradient sum over ranks + Adam through NVLink peer memory -----------------------------------------
// Replaces [ncclAllReduce(sum) -> adam_kernel] (utils.py:43-48 + ddpg_agent.py:272,277) by ONE kernel: every rank
// reads the gradient buffers of all ranks directly (peer loads), adds them in rank order (so every rank computes
// the identical sum) and applies Adam to its own replica.  Cross-GPU ordering uses two flag barriers in peer
// memory: "ready" (all backward passes have finished) before the reads, "done" (all ranks have finished reading)
// before anyone may overwrite its gradients again.  Spins are bounded (about 30 s of SM clocks: host skew between ranks
// -- first-call module loads, checkpoint I/O on rank 0 -- is seconds at most) so a dead peer cannot hang the GPU.  A
// time-out is FATAL and sticky: the kernel returns WITHOUT touching the parameters (a partial gradient sum would make
// the replicas diverge silently), every later launch returns immediately, and the host raises at the next check
// (ddpg_agent.update_many / learn -> p2p_timed_out()).
struct P2PArgs {
  const float* grads[8];
  int* sync[8];
  int rank, world;
};
enum { PS_READY = 0, PS_DONE = 8, PS_EPOCH = 16, PS_COUNT = 17, PS_GO = 18, PS_TIMEOUT = 19, PS_WORDS = 32 };

__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool spin_until_ge(const int* p, int target, int* timeout_flag) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < target) {
    if (clock64() - t0 > 60000000000ll || ld_acquire_sys(timeout_flag) != 0) {
      *timeout_flag = 1;
      return false;
    }
  }
  return true;
}

__global__ void __launch_bounds__(256) adam_p2p_kernel(P2PArgs pa, float* __restrict__ p_actor, float* __restrict__ p_critic,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t na,
                                                       int64_t na_pad, int64_t n, int* step_ctr, float* scal_ring,
                                                       float lr_a, float lr_c, float b1, float b2, float eps) {
  int* mine = pa.sync[pa.rank];
  if (ld_acquire_sys(mine + PS_TIMEOUT) != 0) return;   // sticky: a previous launch timed out, the replica is frozen
  const int epoch = mine[PS_EPOCH] + 1;   // written only by the last block of the previous launch (stream ordered)
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = 1;
  const int t_opt = *step_ctr + 1;                        // this launch is optimiser step t_opt
  const float* scal = scal_ring + 4 * (t_opt & 1);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 32)   // next step's scalars, off the critical path
    adam_scalars(t_opt + 1, lr_a, lr_c, b1, b2, scal_ring + 4 * ((t_opt + 1) & 1));
  __syncthreads();
  if (blockIdx.x == 0) {
    // "ready": tell every rank that this rank's gradients are complete, then wait for everybody
    if (threadIdx.x < pa.world) {
      __threadfence_system();
      st_release_sys(pa.sync[threadIdx.x] + PS_READY + pa.rank, epoch);
      if (!spin_until_ge(mine + PS_READY + threadIdx.x, epoch, mine + PS_TIMEOUT)) ok_s = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(mine + PS_GO, ok_s ? epoch : 0x7fffffff);   // 0x7fffffff releases the waiters into the abort path
  } else {
    if (threadIdx.x == 0 && !spin_until_ge(mine + PS_GO, epoch, mine + PS_TIMEOUT)) ok_s = 0;
    __syncthreads();
  }
  if (!ok_s || ld_acquire_sys(mine + PS_TIMEOUT) != 0) return;   // whole block: no Adam on a partial sum
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= na && i < na_pad) continue;
    float gi = 0.f;
    for (int q = 0; q < pa.world; ++q) gi += __ldcv(pa.grads[q] + i);   // rank order: identical sum on every rank
    float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    float vi = v[i] * b2 + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / scal[2] + eps;
    const float step = i < na ? scal[0] : scal[1];
    float* p = i < na ? p_actor + i : p_critic + (i - na_pad);
    *p = *p - step * (mi / denom);
  }
  // "done": the last block to finish tells every rank that this rank no longer reads their gradients
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int prev = atomicAdd(mine + PS_COUNT, 1);
    if (prev == (int)gridDim.x - 1) {
      mine[PS_COUNT] = 0;
      *step_ctr = *step_ctr + 1;           // every block has read it
      for (int q = 0; q < pa.world; ++q) st_release_sys(pa.sync[q] + PS_DONE + pa.rank, epoch);
      for (int q = 0; q < pa.world; ++q) spin_until_ge(mine + PS_DONE + q, epoch, mine + PS_TIMEOUT);
      st_release_sys(mine + PS_EPOCH, epoch);
    }
  }
}

// target = (1 - polyak) * param + polyak * target, separately rounded like torch
__global__ void polyak_kernel(float* __restrict__ tgt, const float* __restrict__ src, int64_t n,
                              float c_src, float c_tgt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tgt[i] = __fadd_rn(__fmul_rn(c_src, src[i]), __fmul_rn(c_tgt, tgt[i]));
}

__global__ void select_actions_kernel(const float* __restrict__ pi, int64_t n, int Da, float amax,
                                      float noise_eps, float random_eps, float late_clip,
                                      uint64_t seed, const uint64_t* __restrict__ counter,
                                      float* __restrict__ out) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint64_t c = *counter + (uint64_t)r;
  // block 0: two Box-Muller pairs -> up to 4 gaussians; block 1: 4 uniforms; block 2: bernoulli
  Philox4 pg = philox4x32_10(seed, 3 * c, kStreamExplore);
  Philox4 pu = philox4x32_10(seed, 3 * c + 1, kStreamExplore);
  Philox4 pb = philox4x32_10(seed, 3 * c + 2, kStreamExplore);
  const bool take_random = u24(pb.v[0]) < random_eps;
  for (int j = 0; j < Da; ++j) {
    int pair = (j >> 1) & 1;
    float u1 = 1.0f - u24(pg.v[2 * pair]);  // (0,1]
    float u2 = u24(pg.v[2 * pair + 1]);
    float rad = sqrtf(-2.0f * logf(u1));
    float gz = (j & 1) ? rad * sinf(6.28318530717958647692f * u2) : rad * cosf(6.28318530717958647692f * u2);
    float a = pi[r * Da + j] + noise_eps * amax * gz;
    a = fminf(fmaxf(a, -amax), amax);
    float ra = -amax + 2.0f * amax * u24(pu.v[j & 3]);
    if (take_random) a = ra;  // a += 1 * (ra - a)
    if (late_clip > 0.f) a = fminf(fmaxf(a, -late_clip), late_clip);
    out[r * Da + j] = a;
  }
}

__global__ void advance_counter_kernel2(uint64_t* counter, uint64_t by) { *counter += by; }

// forward through the three hidden layers (+ optional output layer) of a net
static int mlp_hidden(bmi_ddpg* h, cudaStream_t st, const NetLayout& L, const float* P, const float* x,
                      int rows, float* o1, float* o2, float* o3) {
  int rc;
  const int H = L.hidden;
  if ((rc = gemm_rm(h, st, 0, 1, rows, H, L.in_dim, x, L.in_dim, P + L.w[0], L.in_dim, o1, H, EPI_RELU_BIAS, P + L.b[0]))) return rc;
  if ((rc = gemm_rm(h, st, 0, 1, rows, H, H, o1, H, P + L.w[1], H, o2, H, EPI_RELU_BIAS, P + L.b[1]))) return rc;
  if ((rc = gemm_rm(h, st, 0, 1, rows, H, H, o2, H, P + L.w[2], H, o3, H, EPI_RELU_BIAS, P + L.b[2]))) return rc;
  return BMI_OK;
}

static int mlp_out(bmi_ddpg* h, cudaStream_t st, const NetLayout& L, const float* P, const float* h3,
                   int rows, float* z) {
  return gemm_rm(h, st, 0, 1, rows, L.out_dim, L.hidden, h3, L.hidden, P + L.w[3], L.hidden, z,
                 L.out_dim, EPI_BIAS, P + L.b[3]);
}

// backward through a 4-layer MLP.  dz: [rows][out] grad wrt output pre-activation (clobbered
// scratch d1/d2 hold hidden grads).  If G != nullptr parameter grads are written at G + offsets.
// If dx != nullptr the input gradient [rows][in] is written.
static int mlp_backward(bmi_ddpg* h, cudaStream_t st, const NetLayout& L, const float* P, float* G,
                        const float* x, const float* a1, const float* a2, const float* a3,
                        const float* dz, int rows, float* dx) {
  int rc;
  const int H = L.hidden, O = L.out_dim, I = L.in_dim;
  dim3 blk(32, 8);
  if (G) {
    // dW4[O][H] = dz^T a3 ; db4 = colsum(dz)
    if ((rc = gemm_rm(h, st, 1, 0, O, H, rows, dz, O, a3, H, G + L.w[3], H, EPI_NONE, nullptr))) return rc;
    bgrad_kernel<<<(O + 31) / 32, blk, 0, st>>>(dz, rows, O, G + L.b[3]);
    BMI_LAUNCHED();
  }
  // d3 = (dz W4) * relu'(a3)
  if ((rc = gemm_rm(h, st, 0, 0, rows, H, O, dz, O, P + L.w[3], H, h->d1, H, EPI_NONE, nullptr))) return rc;
  float* dump = nullptr;  // no parameter gradients wanted: the mask kernel skips its column sums
  const dim3 grd((H + 31) / 32, (rows + 31) / 32);
  drelu_bgrad_kernel<<<grd, blk, 0, st>>>(h->d1, a3, rows, H, G ? G + L.b[2] : dump, h->bg_part, h->bg_ticket);
  BMI_LAUNCHED();
  if (G)
    if ((rc = gemm_rm(h, st, 1, 0, H, H, rows, h->d1, H, a2, H, G + L.w[2], H, EPI_NONE, nullptr))) return rc;
  // d2 = (d3 W3) * relu'(a2)
  if ((rc = gemm_rm(h, st, 0, 0, rows, H, H, h->d1, H, P + L.w[2], H, h->d2, H, EPI_NONE, nullptr))) return rc;
  drelu_bgrad_kernel<<<grd, blk, 0, st>>>(h->d2, a2, rows, H, G ? G + L.b[1] : dump, h->bg_part, h->bg_ticket);
  BMI_LAUNCHED();
  if (G)
    if ((rc = gemm_rm(h, st, 1, 0, H, H, rows, h->d2, H, a1, H, G + L.w[1], H, EPI_NONE, nullptr))) return rc;
  // d1 = (d2 W2) * relu'(a1)
  if ((rc = gemm_rm(h, st, 0, 0, rows, H, H, h->d2, H, P + L.w[1], H, h->d1, H, EPI_NONE, nullptr))) return rc;
  drelu_bgrad_kernel<<<grd, blk, 0, st>>>(h->d1, a1, rows, H, G ? G + L.b[0] : dump, h->bg_part, h->bg_ticket);
  BMI_LAUNCHED();
  if (G)
    if ((rc = gemm_rm(h, st, 1, 0, H, I, rows, h->d1, H, x, I, G + L.w[0], I, EPI_NONE, nullptr))) return rc;
  if (dx)
    if ((rc = gemm_rm(h, st, 0, 0, rows, I, H, h->d1, H, P + L.w[0], I, dx, I, EPI_NONE, nullptr))) return rc;
  return BMI_OK;
}

}  // namespace bmi

extern "C" int64_t bmi_ddpg_actor_param_count(const bmi_ddpg_config* c) {
  if (!c) return -1;
  NetLayout L;
  L.init(c->obs_dim + c->goal_dim, c->hidden, c->act_dim);
  return L.count;
}
extern "C" int64_t bmi_ddpg_critic_param_count(const bmi_ddpg_config* c) {
  if (!c) return -1;
  NetLayout L;
  L.init(c->obs_dim + c->goal_dim + c->act_dim, c->hidden, 1);
  return L.count;
}

extern "C" int bmi_ddpg_create(bmi_ddpg** out, const bmi_ddpg_config* cfg, float* actor, float* critic,
                               float* actor_t, float* critic_t) {
  BMI_REQUIRE(out && cfg && actor && critic && actor_t && critic_t, "bmi_ddpg_create: null pointer");
  BMI_REQUIRE(cfg->obs_dim > 0 && cfg->goal_dim > 0 && cfg->act_dim > 0 && cfg->hidden > 0 &&
                  cfg->batch > 0 && cfg->max_act_rows > 0,
              "bmi_ddpg_create: bad dims");
  BMI_REQUIRE(cfg->gamma >= 0.f && cfg->gamma < 1.f && cfg->clip_return > 0.f,
              "bmi_ddpg_create: gamma must be in [0,1) and clip_return > 0");
  bmi_ddpg* h = new bmi_ddpg();
  h->cfg = *cfg;
  const int Dx = cfg->obs_dim + cfg->goal_dim, Da = cfg->act_dim, H = cfg->hidden, B = cfg->batch;
  h->la.init(Dx, H, Da);
  h->lc.init(Dx + Da, H, 1);
  h->actor = actor; h->critic = critic; h->actor_t = actor_t; h->critic_t = critic_t;
  cublasStatus_t cs = cublasLtCreate(&h->lt);
  if (cs != CUBLAS_STATUS_SUCCESS) {
    set_error("cublasLtCreate failed with status %d", (int)cs);
    delete h;
    return BMI_ERR_CUBLAS;
  }
  int rc = 0;
  h->na_pad = (h->la.count + 63) / 64 * 64;
  h->n_grads = h->na_pad + h->lc.count;
  const int64_t np = h->n_grads;
#define A_(p, n) if (!rc) rc = dalloc(h, &h->p, (size_t)(n))
  A_(grads, np); A_(adam_m, np); A_(adam_v, np); A_(adam_scal, 8);
  A_(xc, B * (Dx + Da)); A_(h1, B * H); A_(h2, B * H); A_(h3, B * H);
  A_(ch1, B * H); A_(ch2, B * H); A_(ch3, B * H); A_(q, B);
  A_(ah1, B * H); A_(ah2, B * H); A_(ah3, B * H); A_(az, B * Da); A_(aa, B * Da);
  A_(qh1, B * H); A_(qh2, B * H); A_(qh3, B * H); A_(qa, B); A_(xca, B * (Dx + Da));
  A_(a_next, B * Da); A_(q_next, B); A_(y, B);
  A_(d1, B * H); A_(d2, B * H); A_(dq, B > H ? B : H); A_(dq_const, B); A_(bg_part, (size_t)((B + 31) / 32) * H); A_(dxc, B * (Dx + Da)); A_(dz, B * Da > H ? B * Da : H);
  A_(ph1, (size_t)cfg->max_act_rows * H); A_(ph2, (size_t)cfg->max_act_rows * H);
  A_(pz, (size_t)cfg->max_act_rows * Da);
  // BMI_DDPG_CUBLAS=1 keeps the cuBLASLt chain (A/B measurements, tests of both paths)
  const char* force_lt = getenv("BMI_DDPG_CUBLAS");
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15u) == 0; };
  h->fused = !(force_lt && force_lt[0] == '1') && H == FH && B % 32 == 0 && Dx + Da <= FIN && Da <= FOUT && al16(actor) &&
             al16(critic) && al16(actor_t) && al16(critic_t);
  if (h->fused) {
    A_(cd1, B * H); A_(cd2, B * H); A_(cd3, B * H); A_(fd1, B * H); A_(fd2, B * H); A_(fd3, B * H);
    A_(loss_part, (size_t)(B / FR) * 4);
  }
#undef A_
  if (!rc) {
    void* p = nullptr;
    const size_t nt = (size_t)(H + 31) / 32 + 1;
    if (cudaMalloc(&p, sizeof(unsigned) * nt) != cudaSuccess || cudaMemset(p, 0, sizeof(unsigned) * nt) != cudaSuccess) {
      set_error("bmi_ddpg_create: cudaMalloc(tickets) failed");
      rc = BMI_ERR_CUDA;
    } else {
      h->owned.push_back(p);
      h->bg_ticket = (unsigned*)p;
      h->adam_ticket = (unsigned*)p + (nt - 1);
    }
  }
  if (!rc) {   // Adam step sizes of step 1 (slot 1 of the ring); every launch prepares its successor's
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
    adam_scalars(1, cfg->lr_actor, cfg->lr_critic, cfg->adam_beta1, cfg->adam_beta2, sc);
    if (cudaMemcpy(h->adam_scal + 4, sc, sizeof(sc), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("bmi_ddpg_create: cudaMemcpy(adam scalars) failed");
      rc = BMI_ERR_CUDA;
    }
  }
  if (!rc) {   // d(-mean Q)/dQ of the actor loss: constant, written once
    fill_kernel<<<(B + 255) / 256, 256>>>(h->dq_const, B, -1.0f / (float)B);
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("bmi_ddpg_create: fill failed"); rc = BMI_ERR_CUDA; }
  }
  if (!rc) {
    void* p = nullptr;
    if (cudaMalloc(&p, sizeof(int)) != cudaSuccess || cudaMemset(p, 0, sizeof(int)) != cudaSuccess) {
      set_error("bmi_ddpg_create: cudaMalloc(step) failed");
      rc = BMI_ERR_CUDA;
    } else {
      h->owned.push_back(p);
      h->adam_step = (int*)p;
    }
  }
  if (!rc) {
    if (cudaMalloc(&h->workspace, h->workspace_bytes) != cudaSuccess) {
      set_error("bmi_ddpg_create: cudaMalloc(workspace) failed");
      rc = BMI_ERR_CUDA;
    }
  }
  if (rc) {
    bmi_ddpg_destroy(h);
    return rc;
  }
  *out = h;
  return BMI_OK;
}

extern "C" int bmi_ddpg_destroy(bmi_ddpg* h) {
  if (!h) return BMI_OK;
  for (auto& kv : h->plans) {
    cublasLtMatmulDescDestroy(kv.second.desc);
    cublasLtMatrixLayoutDestroy(kv.second.la);
    cublasLtMatrixLayoutDestroy(kv.second.lb);
    cublasLtMatrixLayoutDestroy(kv.second.lc);
  }
  for (void* p : h->ipc_opened) cudaIpcCloseMemHandle(p);
  for (void* p : h->owned) cudaFree(p);
  if (h->workspace) cudaFree(h->workspace);
  if (h->lt) cublasLtDestroy(h->lt);
  delete h;
  return BMI_OK;
}

extern "C" int bmi_ddpg_act(bmi_ddpg* h, const float* x, int64_t n, int32_t use_target, float* actions,
                            bmi_stream_t stream) {
  BMI_REQUIRE(h && x && actions, "bmi_ddpg_act: null pointer");
  BMI_REQUIRE(n >= 0 && n <= h->cfg.max_act_rows, "bmi_ddpg_act: n=%lld exceeds max_act_rows=%d",
              (long long)n, h->cfg.max_act_rows);
  if (n == 0) return BMI_OK;
  cudaStream_t st = as_stream(stream);
  const float* P = use_target ? h->actor_t : h->actor;
  const int H = h->cfg.hidden, Da = h->cfg.act_dim;
  int rc;
  // ping-pong two hidden buffers: ph1 -> ph2 -> ph1
  const NetLayout& L = h->la;
  if ((rc = gemm_rm(h, st, 0, 1, (int)n, H, L.in_dim, x, L.in_dim, P + L.w[0], L.in_dim, h->ph1, H, EPI_RELU_BIAS, P + L.b[0]))) return rc;
  if ((rc = gemm_rm(h, st, 0, 1, (int)n, H, H, h->ph1, H, P + L.w[1], H, h->ph2, H, EPI_RELU_BIAS, P + L.b[1]))) return rc;
  if ((rc = gemm_rm(h, st, 0, 1, (int)n, H, H, h->ph2, H, P + L.w[2], H, h->ph1, H, EPI_RELU_BIAS, P + L.b[2]))) return rc;
  if ((rc = mlp_out(h, st, L, P, h->ph1, (int)n, h->pz))) return rc;
  int tot = (int)n * Da;
  tanh_scale_kernel<<<(tot + 255) / 256, 256, 0, st>>>(h->pz, tot, h->cfg.action_max, actions);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_ddpg_backward(bmi_ddpg* h, const float* x, const float* xn, const float* actions,
                                 const float* r, float* losses, bmi_stream_t stream) {
  BMI_REQUIRE(h && x && xn && actions && r && losses, "bmi_ddpg_backward: null pointer");
  cudaStream_t st = as_stream(stream);
  const bmi_ddpg_config& c = h->cfg;
  const int B = c.batch, Dx = c.obs_dim + c.goal_dim, Da = c.act_dim, Dc = Dx + Da;
  const float amax = c.action_max;
  float* Ga = h->grads;
  float* Gc = h->grads + h->na_pad;
  int rc;
  if (h->fused) {
    FusedArgs fa;
    fa.P[0] = h->actor; fa.P[1] = h->critic; fa.P[2] = h->actor_t; fa.P[3] = h->critic_t;
    for (int l = 0; l < 4; ++l) {
      fa.wa[l] = (int)h->la.w[l]; fa.ba[l] = (int)h->la.b[l];
      fa.wc[l] = (int)h->lc.w[l]; fa.bc[l] = (int)h->lc.b[l];
    }
    fa.x = x; fa.xn = xn; fa.act = actions; fa.r = r;
    fa.xc = h->xc; fa.ch1 = h->ch1; fa.ch2 = h->ch2; fa.ch3 = h->ch3;
    fa.cd1 = h->cd1; fa.cd2 = h->cd2; fa.cd3 = h->cd3; fa.dq = h->dq;
    fa.ah1 = h->ah1; fa.ah2 = h->ah2; fa.ah3 = h->ah3;
    fa.fd1 = h->fd1; fa.fd2 = h->fd2; fa.fd3 = h->fd3; fa.dz = h->dz;
    fa.loss_part = h->loss_part;
    fa.B = B; fa.Dx = Dx; fa.Da = Da;
    fa.amax = amax; fa.gamma = c.gamma; fa.clip_ret = c.clip_return; fa.l2 = c.action_l2;
    ddpg_rows_kernel<<<B / FR, FT, 0, st>>>(fa);
    BMI_LAUNCHED();
    WgradArgs wg;
    const int H = c.hidden;
    //            critic W4      W3       W2       W1       actor W4   W3       W2       W1
    const float* Dp[WG_P] = {h->dq, h->cd3, h->cd2, h->cd1, h->dz, h->fd3, h->fd2, h->fd1};
    const float* Ap[WG_P] = {h->ch3, h->ch2, h->ch1, h->xc, h->ah3, h->ah2, h->ah1, x};
    const int ldD[WG_P] = {1, H, H, H, Da, H, H, H}, ldA[WG_P] = {H, H, H, Dc, H, H, H, Dx};
    const int Nj[WG_P] = {1, H, H, H, Da, H, H, H}, Nk[WG_P] = {H, H, H, Dc, H, H, H, Dx};
    float* gWp[WG_P] = {Gc + h->lc.w[3], Gc + h->lc.w[2], Gc + h->lc.w[1], Gc + h->lc.w[0],
                        Ga + h->la.w[3], Ga + h->la.w[2], Ga + h->la.w[1], Ga + h->la.w[0]};
    float* gbp[WG_P] = {Gc + h->lc.b[3], Gc + h->lc.b[2], Gc + h->lc.b[1], Gc + h->lc.b[0],
                        Ga + h->la.b[3], Ga + h->la.b[2], Ga + h->la.b[1], Ga + h->la.b[0]};
    int tiles = 0;
    for (int p = 0; p < WG_P; ++p) {
      wg.D[p] = Dp[p]; wg.Ac[p] = Ap[p]; wg.gW[p] = gWp[p]; wg.gb[p] = gbp[p];
      wg.ldD[p] = ldD[p]; wg.ldA[p] = ldA[p]; wg.Nj[p] = Nj[p]; wg.Nk[p] = Nk[p];
      wg.tile0[p] = tiles;
      tiles += Nj[p] <= FOUT ? 1 : ((Nj[p] + WG_TJ - 1) / WG_TJ) * ((Nk[p] + WG_TK - 1) / WG_TK);   // output layers: one CTA
    }
    wg.tile0[WG_P] = tiles;
    wg.loss_part = h->loss_part; wg.losses = losses; wg.n_part = B / FR; wg.B = B; wg.Da = Da; wg.l2 = c.action_l2;
    ddpg_wgrad_kernel<<<tiles, WG_T, 0, st>>>(wg);
    BMI_LAUNCHED();
    return BMI_OK;
  }
  // ---- target: y = clamp(r + gamma * Q'(x', pi'(x')), -1/(1-gamma), 0) -------------------
  if ((rc = mlp_hidden(h, st, h->la, h->actor_t, xn, B, h->h1, h->h2, h->h3))) return rc;
  if ((rc = mlp_out(h, st, h->la, h->actor_t, h->h3, B, h->az))) return rc;
  tanh_concat_kernel<<<(B * Dc + 255) / 256, 256, 0, st>>>(xn, h->az, B, Dx, Da, amax, h->a_next, h->xc);
  BMI_LAUNCHED();
  if ((rc = mlp_hidden(h, st, h->lc, h->critic_t, h->xc, B, h->h1, h->h2, h->h3))) return rc;
  if ((rc = mlp_out(h, st, h->lc, h->critic_t, h->h3, B, h->q_next))) return rc;
  // ---- critic loss + backward -----------------------------------------------------------
  concat_scale_kernel<<<(B * Dc + 255) / 256, 256, 0, st>>>(x, actions, B, Dx, Da, amax, h->xc);
  BMI_LAUNCHED();
  if ((rc = mlp_hidden(h, st, h->lc, h->critic, h->xc, B, h->ch1, h->ch2, h->ch3))) return rc;
  if ((rc = mlp_out(h, st, h->lc, h->critic, h->ch3, B, h->q))) return rc;
  critic_loss_kernel<<<1, 256, 0, st>>>(r, h->q_next, c.gamma, c.clip_return, h->q, B, h->y, h->dq, losses + 1);
  BMI_LAUNCHED();
  if ((rc = mlp_backward(h, st, h->lc, h->critic, Gc, h->xc, h->ch1, h->ch2, h->ch3, h->dq, B, nullptr))) return rc;
  // ---- actor loss + backward ------------------------------------------------------------
  if ((rc = mlp_hidden(h, st, h->la, h->actor, x, B, h->ah1, h->ah2, h->ah3))) return rc;
  if ((rc = mlp_out(h, st, h->la, h->actor, h->ah3, B, h->az))) return rc;
  tanh_concat_kernel<<<(B * Dc + 255) / 256, 256, 0, st>>>(x, h->az, B, Dx, Da, amax, h->aa, h->xca);
  BMI_LAUNCHED();
  if ((rc = mlp_hidden(h, st, h->lc, h->critic, h->xca, B, h->qh1, h->qh2, h->qh3))) return rc;
  if ((rc = mlp_out(h, st, h->lc, h->critic, h->qh3, B, h->qa))) return rc;
  // d(-mean Q)/dQ = -1/B: a constant vector, filled once at creation (dq_const)
  if ((rc = mlp_backward(h, st, h->lc, h->critic, nullptr, h->xca, h->qh1, h->qh2, h->qh3, h->dq_const, B, h->dxc))) return rc;
  actor_loss_kernel<<<1, 256, 0, st>>>(h->qa, h->aa, h->dxc, B, Dx, Da, amax, c.action_l2, h->dz, losses);
  BMI_LAUNCHED();
  // dz now holds the gradient wrt the actor's output pre-activation; mlp_backward with G
  // only uses d1/d2 as scratch so dz stays intact while it is consumed
  if ((rc = mlp_backward(h, st, h->la, h->actor, Ga, x, h->ah1, h->ah2, h->ah3, h->dz, B, nullptr))) return rc;
  return BMI_OK;
}

extern "C" int bmi_ddpg_grad_buffer(bmi_ddpg* h, float** grads, int64_t* n) {
  BMI_REQUIRE(h && grads && n, "bmi_ddpg_grad_buffer: null pointer");
  *grads = h->grads;
  *n = h->n_grads;
  return BMI_OK;
}

extern "C" int bmi_ddpg_adam_step(bmi_ddpg* h, bmi_stream_t stream) {
  BMI_REQUIRE(h, "bmi_ddpg_adam_step: null handle");
  cudaStream_t st = as_stream(stream);
  const bmi_ddpg_config& c = h->cfg;
  const int64_t n = h->n_grads;
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->actor, h->critic, h->grads, h->adam_m, h->adam_v,
                                                            h->la.count, h->na_pad, n, h->adam_step, h->adam_ticket,
                                                            h->adam_scal, c.lr_actor, c.lr_critic, c.adam_beta1, c.adam_beta2, c.adam_eps);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_ddpg_p2p_export(bmi_ddpg* h, void* handles128) {
  BMI_REQUIRE(h && handles128, "bmi_ddpg_p2p_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  if (!h->p2p_sync) {
    BMI_CUDA_CHECK(cudaMalloc(&h->p2p_sync, PS_WORDS * sizeof(int)));
    BMI_CUDA_CHECK(cudaMemset(h->p2p_sync, 0, PS_WORDS * sizeof(int)));
    h->owned.push_back(h->p2p_sync);
  }
  cudaIpcMemHandle_t hg, hs;
  BMI_CUDA_CHECK(cudaIpcGetMemHandle(&hg, h->grads));
  BMI_CUDA_CHECK(cudaIpcGetMemHandle(&hs, h->p2p_sync));
  memcpy(handles128, &hg, 64);
  memcpy((char*)handles128 + 64, &hs, 64);
  return BMI_OK;
}

extern "C" int bmi_ddpg_p2p_attach(bmi_ddpg* h, int32_t rank, int32_t world, const void* all_handles) {
  BMI_REQUIRE(h && all_handles, "bmi_ddpg_p2p_attach: null pointer");
  BMI_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world, "bmi_ddpg_p2p_attach: bad rank %d / world %d (max 8)", rank, world);
  BMI_REQUIRE(h->p2p_sync, "bmi_ddpg_p2p_attach: call bmi_ddpg_p2p_export first");
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      h->peer_grads[q] = h->grads;
      h->peer_sync[q] = h->p2p_sync;
      continue;
    }
    cudaIpcMemHandle_t hg, hs;
    memcpy(&hg, (const char*)all_handles + (size_t)q * 128, 64);
    memcpy(&hs, (const char*)all_handles + (size_t)q * 128 + 64, 64);
    void *pg = nullptr, *ps = nullptr;
    BMI_CUDA_CHECK(cudaIpcOpenMemHandle(&pg, hg, cudaIpcMemLazyEnablePeerAccess));
    BMI_CUDA_CHECK(cudaIpcOpenMemHandle(&ps, hs, cudaIpcMemLazyEnablePeerAccess));
    h->ipc_opened.push_back(pg);
    h->ipc_opened.push_back(ps);
    h->peer_grads[q] = (float*)pg;
    h->peer_sync[q] = (int*)ps;
  }
  h->p2p_rank = rank;
  h->p2p_world = world;
  return BMI_OK;
}

extern "C" int bmi_ddpg_adam_step_p2p(bmi_ddpg* h, bmi_stream_t stream) {
  BMI_REQUIRE(h, "bmi_ddpg_adam_step_p2p: null handle");
  BMI_REQUIRE(h->p2p_world >= 1 && h->peer_grads[h->p2p_rank] != nullptr, "bmi_ddpg_adam_step_p2p: peers not attached");
  cudaStream_t st = as_stream(stream);
  const bmi_ddpg_config& c = h->cfg;
  P2PArgs pa;
  for (int q = 0; q < 8; ++q) { pa.grads[q] = h->peer_grads[q]; pa.sync[q] = h->peer_sync[q]; }
  pa.rank = h->p2p_rank;
  pa.world = h->p2p_world;
  adam_p2p_kernel<<<64, 256, 0, st>>>(pa, h->actor, h->critic, h->adam_m, h->adam_v, h->la.count, h->na_pad, h->n_grads,
                                      h->adam_step, h->adam_scal, c.lr_actor, c.lr_critic, c.adam_beta1, c.adam_beta2, c.adam_eps);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_ddpg_p2p_status(bmi_ddpg* h, int32_t* timed_out) {
  BMI_REQUIRE(h && timed_out, "bmi_ddpg_p2p_status: null pointer");
  *timed_out = 0;
  if (h->p2p_sync) BMI_CUDA_CHECK(cudaMemcpy(timed_out, h->p2p_sync + PS_TIMEOUT, sizeof(int), cudaMemcpyDeviceToHost));
  return BMI_OK;
}

extern "C" int bmi_ddpg_soft_update(bmi_ddpg* h, bmi_stream_t stream) {
  BMI_REQUIRE(h, "bmi_ddpg_soft_update: null handle");
  cudaStream_t st = as_stream(stream);
  // (1 - polyak) is evaluated in double by python, then rounded to f32 when it meets the tensor
  const float c_src = h->cfg.one_minus_polyak, c_tgt = h->cfg.polyak;
  polyak_kernel<<<(unsigned)((h->la.count + 255) / 256), 256, 0, st>>>(h->actor_t, h->actor, h->la.count, c_src, c_tgt);
  BMI_LAUNCHED();
  polyak_kernel<<<(unsigned)((h->lc.count + 255) / 256), 256, 0, st>>>(h->critic_t, h->critic, h->lc.count, c_src, c_tgt);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_select_actions(const float* pi, int64_t n, int32_t Da, float amax, float noise_eps,
                                  float random_eps, float late_clip, uint64_t seed, uint64_t* counter,
                                  float* out, bmi_stream_t stream) {
  BMI_REQUIRE(n >= 0 && Da > 0 && Da <= 4, "bmi_select_actions: act_dim must be in [1,4]");
  if (n == 0) return BMI_OK;
  BMI_REQUIRE(pi && counter && out, "bmi_select_actions: null pointer");
  cudaStream_t st = as_stream(stream);
  select_actions_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(pi, n, Da, amax, noise_eps, random_eps,
                                                                     late_clip, seed, counter, out);
  BMI_LAUNCHED();
  advance_counter_kernel2<<<1, 1, 0, st>>>(counter, (uint64_t)n);
  BMI_LAUNCHED();
  return BMI_OK;
}
