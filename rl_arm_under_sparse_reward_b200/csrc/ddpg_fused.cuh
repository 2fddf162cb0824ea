// Hand-written DDPG update for the reference's network shape (hidden = 256): TWO launches instead of the ~54 of the
// cuBLASLt chain (ddpg.cu: bmi_ddpg_backward).  Restates ddpg_agent.py:250-270,274-275 + models.py:11-44.
//
//   ddpg_rows_kernel   every CTA owns FR = 2 rows of the batch and runs the WHOLE per-row computation on them:
//                      target actor -> target critic -> y; critic forward + loss + delta chain; actor forward; critic(x, pi(x))
//                      forward + delta chain down to dQ/da; actor delta chain.  A row never needs another row, so there is no
//                      grid-wide synchronisation: the only shared data are the weights, streamed from L2 by every CTA
//                      (16 hidden x hidden matrices = 4.2 MB per CTA; 128 CTAs at batch 256).  Layer products are fp32 FFMA:
//                      a warp owns 32 output neurons, every lane a 8-wide slice of the input vector (two coalesced 512-byte
//                      weight-row loads per neuron), and the 32 x FR partial sums per lane are reduced with a transposing
//                      butterfly (31 shuffles per row instead of 160).  The delta (dgrad) products read the SAME row-major
//                      weights with the lane owning a slice of the layer's INPUT, so they need no lane reduction at all, only a
//                      sum over the 8 warps through shared memory.  Activations and deltas go to global memory for ...
//   ddpg_wgrad_kernel  ... the weight / bias gradients: 8 products  dW = delta^T . activation  (sum over the batch rows) as
//                      32 x 64 output tiles (144 CTAs, one wave), classic shared-memory fp32 GEMM with register prefetch; the
//                      CTAs of the first tile column also produce the bias gradients, CTA 0 the two scalar losses.
//
// The layer routines are deliberately NOT inlined: the kernel runs 33 layer steps back to back without a loop, and inlined it
// was 290 KB of straight-line code, i.e. an instruction-cache miss per instruction.
//
// Results: same quantities as the cuBLASLt chain up to fp32 summation order (tests/test_gpu_ddpg.py runs both paths
// against the torch oracle and against each other).
#pragma once

namespace bmi {

constexpr int FH = 256;        // hidden width this path is built for
constexpr int FR = 2;          // batch rows per CTA
constexpr int FT = 256;        // threads per CTA
constexpr int FW = FT / 32;    // warps per CTA
constexpr int FIN = 64;        // padded width of the input vectors (obs + goal + action <= 64)
constexpr int FOUT = 8;        // largest output layer (act_dim <= 8)

struct FusedArgs {
  const float* P[4];           // flat parameters: 0 actor, 1 critic, 2 actor target, 3 critic target
  int wa[4], ba[4];            // actor layer offsets (floats) inside its flat buffer
  int wc[4], bc[4];            // critic layer offsets
  const float *x, *xn, *act, *r;
  float *xc, *ch1, *ch2, *ch3, *cd1, *cd2, *cd3, *dq;   // critic(x, a): input rows, activations, deltas
  float *ah1, *ah2, *ah3, *fd1, *fd2, *fd3, *dz;        // actor(x): activations, deltas (its input rows are x)
  float* loss_part;            // [gridDim.x][4]: sum qa, sum th^2, sum (y - q)^2 over the CTA's rows
  int B, Dx, Da;
  float amax, gamma, clip_ret, l2;
};

__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

// v[i] (i < 32) summed over the 32 lanes; lane l returns the total of element l
__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = lane & s;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// out[r][j] = act(b[j] + sum_k W[j][k] in[r][k]),  W [FH][FH] row-major.  in / out: shared [FR][FH]; gout: global rows or null
template <bool RELU>
__device__ __noinline__ void f_fwd_hidden(const float* __restrict__ W, const float* __restrict__ b, const float* in,
                                             float* out, float* gout) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float4 i0[FR], i1[FR];
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    i0[r] = *reinterpret_cast<const float4*>(in + r * FH + l * 4);
    i1[r] = *reinterpret_cast<const float4*>(in + r * FH + 128 + l * 4);
  }
  float acc[FR][32];
  const float4* Wp = reinterpret_cast<const float4*>(W + (size_t)(w * 32) * FH) + l;   // row jj: + jj * 64 float4
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 8) {
    float4 a[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a[u] = __ldg(Wp + (j0 + u) * 64);
      c[u] = __ldg(Wp + (j0 + u) * 64 + 32);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int r = 0; r < FR; ++r) acc[r][j0 + u] = dot4(a[u], i0[r]) + dot4(c[u], i1[r]);
  }
  const float bj = __ldg(b + threadIdx.x);
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    float v = butterfly32(acc[r], l) + bj;          // lane l: neuron w * 32 + l == threadIdx.x
    if (RELU) v = fmaxf(v, 0.f);
    out[r * FH + threadIdx.x] = v;
    if (gout) gout[r * FH + threadIdx.x] = v;
  }
}

// first layer: out[r][j] = relu(b[j] + sum_{k < K} W[j][k] in[r][k]),  W [FH][K] row-major, K <= 64; in: shared [FR][FIN]
__device__ __noinline__ void f_fwd_first(const float* __restrict__ W, const float* __restrict__ b, const float* in, int K,
                                            float* out, float* gout) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float i0[FR], i1[FR];
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    i0[r] = l < K ? in[r * FIN + l] : 0.f;
    i1[r] = l + 32 < K ? in[r * FIN + 32 + l] : 0.f;
  }
  float acc[FR][32];
  const float* Wp = W + (size_t)(w * 32) * K;
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 8) {
    float a[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a[u] = l < K ? __ldg(Wp + (j0 + u) * K + l) : 0.f;
      c[u] = l + 32 < K ? __ldg(Wp + (j0 + u) * K + 32 + l) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int r = 0; r < FR; ++r) acc[r][j0 + u] = fmaf(c[u], i1[r], a[u] * i0[r]);
  }
  const float bj = __ldg(b + threadIdx.x);
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    const float v = fmaxf(butterfly32(acc[r], l) + bj, 0.f);
    out[r * FH + threadIdx.x] = v;
    if (gout) gout[r * FH + threadIdx.x] = v;
  }
}

// output layer: z[r][j] = b[j] + sum_k W[j][k] in[r][k],  j < N <= FOUT (one warp per output); z: shared [FR][FOUT]
__device__ __noinline__ void f_fwd_out(const float* __restrict__ W, const float* __restrict__ b, const float* in, int N,
                                          float* z) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (w >= N) return;
  const float4* Wp = reinterpret_cast<const float4*>(W + (size_t)w * FH) + l;
  const float4 a = __ldg(Wp), c = __ldg(Wp + 32);
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    float v = dot4(a, *reinterpret_cast<const float4*>(in + r * FH + l * 4)) +
              dot4(c, *reinterpret_cast<const float4*>(in + r * FH + 128 + l * 4));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) z[r * FOUT + w] = v + __ldg(b + w);
  }
}

// delta below an output layer: dx[r][k] = [h[r][k] > 0] sum_{j < N} dz[r][j] W[j][k];  dz: shared [FR][FOUT]
__device__ __noinline__ void f_bwd_out(const float* __restrict__ W, const float* dz, int N, const float* h, float* dx,
                                          float* gdx) {
  const int k = threadIdx.x;
  float acc[FR];
#pragma unroll
  for (int r = 0; r < FR; ++r) acc[r] = 0.f;
  for (int j = 0; j < N; ++j) {
    const float wv = __ldg(W + j * FH + k);
#pragma unroll
    for (int r = 0; r < FR; ++r) acc[r] = fmaf(dz[r * FOUT + j], wv, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    const float v = h[r * FH + k] > 0.f ? acc[r] : 0.f;
    dx[r * FH + k] = v;
    if (gdx) gdx[r * FH + k] = v;
  }
}

// delta below a hidden layer: dx[r][k] = [h[r][k] > 0] sum_j dy[r][j] W[j][k].  Contains a __syncthreads (partial sums of
// the 8 warps in part [FW][FR][FH]); dy / h / dx: shared [FR][FH], dx must not alias dy
__device__ __noinline__ void f_bwd_hidden(const float* __restrict__ W, const float* dy, const float* h, float* dx,
                                             float* gdx, float* part) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float4 a0[FR], a1[FR];
#pragma unroll
  for (int r = 0; r < FR; ++r) a0[r] = a1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* Wp = reinterpret_cast<const float4*>(W + (size_t)(w * 32) * FH) + l;
#pragma unroll
  for (int j0 = 0; j0 < 32; j0 += 8) {
    float4 a[8], c[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a[u] = __ldg(Wp + (j0 + u) * 64);
      c[u] = __ldg(Wp + (j0 + u) * 64 + 32);
    }
#pragma unroll
    for (int r = 0; r < FR; ++r) {
      const float4 d0 = *reinterpret_cast<const float4*>(dy + r * FH + w * 32 + j0);
      const float4 d1 = *reinterpret_cast<const float4*>(dy + r * FH + w * 32 + j0 + 4);
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a0[r].x = fmaf(d[u], a[u].x, a0[r].x); a0[r].y = fmaf(d[u], a[u].y, a0[r].y);
        a0[r].z = fmaf(d[u], a[u].z, a0[r].z); a0[r].w = fmaf(d[u], a[u].w, a0[r].w);
        a1[r].x = fmaf(d[u], c[u].x, a1[r].x); a1[r].y = fmaf(d[u], c[u].y, a1[r].y);
        a1[r].z = fmaf(d[u], c[u].z, a1[r].z); a1[r].w = fmaf(d[u], c[u].w, a1[r].w);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    *reinterpret_cast<float4*>(part + (w * FR + r) * FH + l * 4) = a0[r];
    *reinterpret_cast<float4*>(part + (w * FR + r) * FH + 128 + l * 4) = a1[r];
  }
  __syncthreads();
  const int k = threadIdx.x;
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < FW; ++ww) s += part[(ww * FR + r) * FH + k];
    const float v = h[r * FH + k] > 0.f ? s : 0.f;
    dx[r * FH + k] = v;
    if (gdx) gdx[r * FH + k] = v;
  }
}

// gradient of the first layer's pre-activations wrt input columns [c0, c0 + n):  out[r][c] = sum_j d[r][j] W[j][c0 + c],
// W [FH][K] row-major, n <= FOUT.  Contains a __syncthreads; red: shared [FW][FR * FOUT]; out: shared [FR][FOUT]
__device__ __forceinline__ void f_bwd_first_cols(const float* __restrict__ W, int K, int c0, int n, const float* d, float* out,
                                                 float* red) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, j = threadIdx.x;
  float v[FR][FOUT];
#pragma unroll
  for (int c = 0; c < FOUT; ++c) {
    const float wv = c < n ? __ldg(W + (size_t)j * K + c0 + c) : 0.f;
#pragma unroll
    for (int r = 0; r < FR; ++r) v[r][c] = d[r * FH + j] * wv;
  }
#pragma unroll
  for (int r = 0; r < FR; ++r)
#pragma unroll
    for (int c = 0; c < FOUT; ++c) {
      float s = v[r][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (l == 0) red[w * (FR * FOUT) + r * FOUT + c] = s;
    }
  __syncthreads();
  if (threadIdx.x < FR * FOUT) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < FW; ++ww) s += red[ww * (FR * FOUT) + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(FT, 1) ddpg_rows_kernel(const FusedArgs A) {
  __shared__ __align__(16) float s_vec[14][FR * FH];     // 0-2 scratch, 3-5 critic h, 6-8 actor h, 9-11 critic(x, pi) h, 12-13 deltas
  __shared__ __align__(16) float s_part[FW * FR * FH];
  __shared__ __align__(16) float s_in[3][FR * FIN];      // 0: [xn, pi'(xn)/amax]  1: [x, a/amax]  2: [x, pi(x)/amax]
  __shared__ float s_z[FR * FOUT], s_dz[FR * FOUT], s_th[FR * FOUT], s_da[FR * FOUT], s_red[FW * FR * FOUT];
  __shared__ float s_qn[FR * FOUT], s_q[FR * FOUT], s_qa[FR * FOUT];
  const int t = threadIdx.x;
  const int row0 = blockIdx.x * FR;
  const int Dx = A.Dx, Da = A.Da, Dc = Dx + Da;
  const float *Pa = A.P[0], *Pc = A.P[1], *Pat = A.P[2], *Pct = A.P[3];
  float *sA = s_vec[0], *sB = s_vec[1];
  float *ch1 = s_vec[3], *ch2 = s_vec[4], *ch3 = s_vec[5], *ah1 = s_vec[6], *ah2 = s_vec[7], *ah3 = s_vec[8];
  float *qh1 = s_vec[9], *qh2 = s_vec[10], *qh3 = s_vec[11], *sD0 = s_vec[12], *sD1 = s_vec[13];
  const size_t g0 = (size_t)row0 * FH;      // this CTA's rows in the [B][FH] global arrays

  // ---- inputs -------------------------------------------------------------------------------------------------------------
  for (int i = t; i < FR * FIN; i += FT) {
    const int r = i / FIN, k = i - r * FIN;
    const size_t row = row0 + r;
    float vn = 0.f, vc = 0.f, va = 0.f;
    if (k < Dx) {
      vn = A.xn[row * Dx + k];
      vc = va = A.x[row * Dx + k];
    } else if (k < Dc) {
      vc = __fdiv_rn(A.act[row * Da + (k - Dx)], A.amax);
    }
    s_in[0][i] = vn; s_in[1][i] = vc; s_in[2][i] = va;
    if (k < Dc) A.xc[row * Dc + k] = vc;
  }
  __syncthreads();

  // ---- target: y = clamp(r + gamma Q'(x', pi'(x')), -clip, 0)  (ddpg_agent.py:252-260) ---------------------------------------
  f_fwd_first(Pat + A.wa[0], Pat + A.ba[0], s_in[0], Dx, sA, nullptr);
  __syncthreads();
  f_fwd_hidden<true>(Pat + A.wa[1], Pat + A.ba[1], sA, sB, nullptr);
  __syncthreads();
  f_fwd_hidden<true>(Pat + A.wa[2], Pat + A.ba[2], sB, sA, nullptr);
  __syncthreads();
  f_fwd_out(Pat + A.wa[3], Pat + A.ba[3], sA, Da, s_z);
  __syncthreads();
  if (t < FR * Da) {
    const int r = t / Da, c = t - r * Da;
    const float a = A.amax * tanhf(s_z[r * FOUT + c]);
    s_in[0][r * FIN + Dx + c] = __fdiv_rn(a, A.amax);
  }
  __syncthreads();
  f_fwd_first(Pct + A.wc[0], Pct + A.bc[0], s_in[0], Dc, sA, nullptr);
  __syncthreads();
  f_fwd_hidden<true>(Pct + A.wc[1], Pct + A.bc[1], sA, sB, nullptr);
  __syncthreads();
  f_fwd_hidden<true>(Pct + A.wc[2], Pct + A.bc[2], sB, sA, nullptr);
  __syncthreads();
  f_fwd_out(Pct + A.wc[3], Pct + A.bc[3], sA, 1, s_qn);
  __syncthreads();

  // ---- critic(x, a): forward, loss, delta chain (ddpg_agent.py:262-263,274-275) ---------------------------------------------
  f_fwd_first(Pc + A.wc[0], Pc + A.bc[0], s_in[1], Dc, ch1, A.ch1 + g0);
  __syncthreads();
  f_fwd_hidden<true>(Pc + A.wc[1], Pc + A.bc[1], ch1, ch2, A.ch2 + g0);
  __syncthreads();
  f_fwd_hidden<true>(Pc + A.wc[2], Pc + A.bc[2], ch2, ch3, A.ch3 + g0);
  __syncthreads();
  f_fwd_out(Pc + A.wc[3], Pc + A.bc[3], ch3, 1, s_q);
  __syncthreads();
  float l_c = 0.f, l_q = 0.f, l_t = 0.f;     // thread 0: loss partial sums of this CTA
  if (t == 0) {
#pragma unroll
    for (int r = 0; r < FR; ++r) {
      const float y = fminf(fmaxf(__fadd_rn(A.r[row0 + r], __fmul_rn(A.gamma, s_qn[r * FOUT])), -A.clip_ret), 0.0f);
      const float d = y - s_q[r * FOUT];
      l_c += d * d;
      const float dq = -2.0f * d / (float)A.B;
      s_dz[r * FOUT] = dq;
      A.dq[row0 + r] = dq;
    }
  }
  __syncthreads();
  f_bwd_out(Pc + A.wc[3], s_dz, 1, ch3, sD0, A.cd3 + g0);
  __syncthreads();
  f_bwd_hidden(Pc + A.wc[2], sD0, ch2, sD1, A.cd2 + g0, s_part);
  __syncthreads();
  f_bwd_hidden(Pc + A.wc[1], sD1, ch1, sD0, A.cd1 + g0, s_part);
  __syncthreads();

  // ---- actor(x) and critic(x, pi(x)) forward (ddpg_agent.py:265-267) -------------------------------------------------------
  f_fwd_first(Pa + A.wa[0], Pa + A.ba[0], s_in[2], Dx, ah1, A.ah1 + g0);
  __syncthreads();
  f_fwd_hidden<true>(Pa + A.wa[1], Pa + A.ba[1], ah1, ah2, A.ah2 + g0);
  __syncthreads();
  f_fwd_hidden<true>(Pa + A.wa[2], Pa + A.ba[2], ah2, ah3, A.ah3 + g0);
  __syncthreads();
  f_fwd_out(Pa + A.wa[3], Pa + A.ba[3], ah3, Da, s_z);
  __syncthreads();
  if (t < FR * Da) {
    const int r = t / Da, c = t - r * Da;
    const float a = A.amax * tanhf(s_z[r * FOUT + c]);
    s_th[r * FOUT + c] = a / A.amax;
    s_in[2][r * FIN + Dx + c] = __fdiv_rn(a, A.amax);
  }
  __syncthreads();
  f_fwd_first(Pc + A.wc[0], Pc + A.bc[0], s_in[2], Dc, qh1, nullptr);
  __syncthreads();
  f_fwd_hidden<true>(Pc + A.wc[1], Pc + A.bc[1], qh1, qh2, nullptr);
  __syncthreads();
  f_fwd_hidden<true>(Pc + A.wc[2], Pc + A.bc[2], qh2, qh3, nullptr);
  __syncthreads();
  f_fwd_out(Pc + A.wc[3], Pc + A.bc[3], qh3, 1, s_qa);
  if (t < FR * FOUT) s_dz[t] = -1.0f / (float)A.B;          // d(-mean Q)/dQ
  __syncthreads();

  // ---- dQ/da through the critic, then the actor's delta chain (ddpg_agent.py:266-270) -------------------------------------
  f_bwd_out(Pc + A.wc[3], s_dz, 1, qh3, sD0, nullptr);
  __syncthreads();
  f_bwd_hidden(Pc + A.wc[2], sD0, qh2, sD1, nullptr, s_part);
  __syncthreads();
  f_bwd_hidden(Pc + A.wc[1], sD1, qh1, sD0, nullptr, s_part);
  __syncthreads();
  f_bwd_first_cols(Pc + A.wc[0], Dc, Dx, Da, sD0, s_da, s_red);
  __syncthreads();
  if (t == 0)
#pragma unroll
    for (int r = 0; r < FR; ++r) l_q += s_qa[r * FOUT];
  if (t < FR * Da) {
    const int r = t / Da, c = t - r * Da;
    const float th = s_th[r * FOUT + c];
    const float n = (float)(A.B * Da);
    const float da = s_da[r * FOUT + c] / A.amax + A.l2 * 2.0f * th / (A.amax * n);
    const float dz = da * A.amax * (1.0f - th * th);
    s_dz[r * FOUT + c] = dz;
    A.dz[(size_t)(row0 + r) * Da + c] = dz;
  }
  __syncthreads();
  if (t == 0) {
    for (int r = 0; r < FR; ++r)
      for (int c = 0; c < Da; ++c) l_t += s_th[r * FOUT + c] * s_th[r * FOUT + c];
    float* lp = A.loss_part + (size_t)blockIdx.x * 4;
    lp[0] = l_q; lp[1] = l_t; lp[2] = l_c; lp[3] = 0.f;
  }
  f_bwd_out(Pa + A.wa[3], s_dz, Da, ah3, sD0, A.fd3 + g0);
  __syncthreads();
  f_bwd_hidden(Pa + A.wa[2], sD0, ah2, sD1, A.fd2 + g0, s_part);
  __syncthreads();
  f_bwd_hidden(Pa + A.wa[1], sD1, ah1, sD0, A.fd1 + g0, s_part);
}

// ---- weight / bias gradients ---------------------------------------------------------------------------------------------------
// problem p:  gW[j][k] = sum_r D[r][j] A[r][k]  (j < Nj, k < Nk),  gb[j] = sum_r D[r][j];  D [B][ldD], A [B][ldA] row-major
constexpr int WG_P = 8;         // problems
constexpr int WG_TJ = 32, WG_TK = 64, WG_RC = 32;
struct WgradArgs {
  const float* D[WG_P];
  const float* Ac[WG_P];
  float* gW[WG_P];
  float* gb[WG_P];
  int ldD[WG_P], ldA[WG_P], Nj[WG_P], Nk[WG_P], tile0[WG_P + 1];   // tile0: first CTA of the problem
  const float* loss_part;
  float* losses;
  int n_part, B, Da;
  float l2;
};

__global__ void __launch_bounds__(256) ddpg_wgrad_kernel(const WgradArgs G) {
  __shared__ __align__(16) float sD[WG_RC][WG_TJ];
  __shared__ __align__(16) float sA[WG_RC][WG_TK];
  const int t = threadIdx.x;
  int p = 0;
#pragma unroll
  for (int q = 1; q < WG_P; ++q) p += (int)blockIdx.x >= G.tile0[q];
  const int tile = blockIdx.x - G.tile0[p];
  const int Nj = G.Nj[p], Nk = G.Nk[p], ldD = G.ldD[p], ldA = G.ldA[p];
  const int tk_n = (Nk + WG_TK - 1) / WG_TK;
  const int j0 = (tile / tk_n) * WG_TJ, k0 = (tile % tk_n) * WG_TK;
  const float* __restrict__ D = G.D[p];
  const float* __restrict__ Am = G.Ac[p];
  // staging roles: D chunk 32 x 32 = 4 per thread (row t / 32 + 8 i, col t % 32); A chunk 32 x 64 = 8 per thread
  const int dc = t & 31, dr = t >> 5, ac = t & 63, ar = t >> 6;
  const bool d_ok = j0 + dc < Nj, a_ok = k0 + ac < Nk;
  float pd[4], pa[8];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) pd[i] = d_ok ? D[(size_t)(r0 + dr + 8 * i) * ldD + j0 + dc] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) pa[i] = a_ok ? Am[(size_t)(r0 + ar + 4 * i) * ldA + k0 + ac] : 0.f;
  };
  // compute roles: thread owns outputs j = jp * 2 + {0, 1}, k = kq * 4 + {0..3}
  const int kq = t & 15, jp = t >> 4;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float bsum = 0.f;                                         // thread t < 32: column sum of D (bias gradient)
  fetch(0);
  for (int r0 = 0; r0 < G.B; r0 += WG_RC) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) sD[dr + 8 * i][dc] = pd[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) sA[ar + 4 * i][ac] = pa[i];
    __syncthreads();
    if (r0 + WG_RC < G.B) fetch(r0 + WG_RC);
#pragma unroll
    for (int r = 0; r < WG_RC; ++r) {
      const float2 d = *reinterpret_cast<const float2*>(&sD[r][jp * 2]);
      const float4 a = *reinterpret_cast<const float4*>(&sA[r][kq * 4]);
      acc[0][0] = fmaf(d.x, a.x, acc[0][0]); acc[0][1] = fmaf(d.x, a.y, acc[0][1]);
      acc[0][2] = fmaf(d.x, a.z, acc[0][2]); acc[0][3] = fmaf(d.x, a.w, acc[0][3]);
      acc[1][0] = fmaf(d.y, a.x, acc[1][0]); acc[1][1] = fmaf(d.y, a.y, acc[1][1]);
      acc[1][2] = fmaf(d.y, a.z, acc[1][2]); acc[1][3] = fmaf(d.y, a.w, acc[1][3]);
    }
    if (k0 == 0 && t < WG_TJ) {
#pragma unroll
      for (int r = 0; r < WG_RC; ++r) bsum += sD[r][t];
    }
  }
  float* __restrict__ gW = G.gW[p];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int j = j0 + jp * 2 + a;
    if (j >= Nj) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = k0 + kq * 4 + b;
      if (k < Nk) gW[(size_t)j * Nk + k] = acc[a][b];
    }
  }
  if (k0 == 0 && t < WG_TJ && j0 + t < Nj) G.gb[p][j0 + t] = bsum;
  if (blockIdx.x == 0 && t == 0) {
    float sq = 0.f, st = 0.f, sc = 0.f;
    for (int i = 0; i < G.n_part; ++i) {
      sq += G.loss_part[i * 4 + 0];
      st += G.loss_part[i * 4 + 1];
      sc += G.loss_part[i * 4 + 2];
    }
    G.losses[0] = -sq / (float)G.B + G.l2 * st / (float)(G.B * G.Da);
    G.losses[1] = sc / (float)G.B;
  }
}

}  // namespace bmi
