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
//                      32 x 64 output tiles (152 CTAs of 128 threads, 4 x 4 outputs per thread), shared-memory fp32 GEMM with register prefetch; the
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

// out[r][j] = act(b[j] + sum_k W[j][k] in[r][k]),  W [FH][FH] row-major.  in / out: shared [FR][FH]; gout: global rows or null.
// STARTS with the barrier that makes `in` visible -- after the first batch of weight loads has been issued, so that their L2
// latency overlaps the wait; inside, the loads of batch b + 1 are in flight while batch b is multiplied.
template <bool RELU>
__device__ __noinline__ void f_fwd_hidden(const float* __restrict__ W, const float* __restrict__ b, const float* in,
                                          float* out, float* gout) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const float4* Wp = reinterpret_cast<const float4*>(W + (size_t)(w * 32) * FH) + l;   // row jj: + jj * 64 float4
  float4 a[2][8], c[2][8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    a[0][u] = __ldg(Wp + u * 64);
    c[0][u] = __ldg(Wp + u * 64 + 32);
  }
  const float bj = __ldg(b + threadIdx.x);
  __syncthreads();
  float4 i0[FR], i1[FR];
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    i0[r] = *reinterpret_cast<const float4*>(in + r * FH + l * 4);
    i1[r] = *reinterpret_cast<const float4*>(in + r * FH + 128 + l * 4);
  }
  float acc[FR][32];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g < 3) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a[(g + 1) & 1][u] = __ldg(Wp + ((g + 1) * 8 + u) * 64);
        c[(g + 1) & 1][u] = __ldg(Wp + ((g + 1) * 8 + u) * 64 + 32);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int r = 0; r < FR; ++r) acc[r][g * 8 + u] = dot4(a[g & 1][u], i0[r]) + dot4(c[g & 1][u], i1[r]);
  }
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    float v = butterfly32(acc[r], l) + bj;          // lane l: neuron w * 32 + l == threadIdx.x
    if (RELU) v = fmaxf(v, 0.f);
    out[r * FH + threadIdx.x] = v;
    if (gout) gout[r * FH + threadIdx.x] = v;
  }
}

// first layer: out[r][j] = relu(b[j] + sum_{k < K} W[j][k] in[r][k]),  W [FH][K] row-major, K <= 64; in: shared [FR][FIN].
// Starts with the barrier that makes `in` visible (after its weight loads have been issued)
__device__ __noinline__ void f_fwd_first(const float* __restrict__ W, const float* __restrict__ b, const float* in, int K,
                                         float* out, float* gout) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const float* Wp = W + (size_t)(w * 32) * K;
  float a[32], c[32];                           // the whole 32 x K slab of this warp: at most 64 loads per lane, all in flight
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    a[u] = l < K ? __ldg(Wp + u * K + l) : 0.f;
    c[u] = l + 32 < K ? __ldg(Wp + u * K + 32 + l) : 0.f;
  }
  const float bj = __ldg(b + threadIdx.x);
  __syncthreads();
  float i0[FR], i1[FR];
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    i0[r] = l < K ? in[r * FIN + l] : 0.f;
    i1[r] = l + 32 < K ? in[r * FIN + 32 + l] : 0.f;
  }
  float acc[FR][32];
#pragma unroll
  for (int u = 0; u < 32; ++u)
#pragma unroll
    for (int r = 0; r < FR; ++r) acc[r][u] = fmaf(c[u], i1[r], a[u] * i0[r]);
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    const float v = fmaxf(butterfly32(acc[r], l) + bj, 0.f);
    out[r * FH + threadIdx.x] = v;
    if (gout) gout[r * FH + threadIdx.x] = v;
  }
}

// output layer: z[r][j] = b[j] + sum_k W[j][k] in[r][k],  j < N <= FOUT (one warp per output); z: shared [FR][FOUT].
// Starts with the barrier that makes `in` visible
__device__ __noinline__ void f_fwd_out(const float* __restrict__ W, const float* __restrict__ b, const float* in, int N,
                                       float* z) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
  float bw = 0.f;
  if (w < N) {
    const float4* Wp = reinterpret_cast<const float4*>(W + (size_t)w * FH) + l;
    a = __ldg(Wp);
    c = __ldg(Wp + 32);
    bw = __ldg(b + w);
  }
  __syncthreads();
  if (w >= N) return;
#pragma unroll
  for (int r = 0; r < FR; ++r) {
    float v = dot4(a, *reinterpret_cast<const float4*>(in + r * FH + l * 4)) +
              dot4(c, *reinterpret_cast<const float4*>(in + r * FH + 128 + l * 4));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) z[r * FOUT + w] = v + bw;
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

// delta below a hidden layer: dx[r][k] = [h[r][k] > 0] sum_j dy[r][j] W[j][k].  Starts with the barrier that makes dy visible
// (after the first weight loads) and contains a second one (partial sums of the 8 warps in part [FW][FR][FH]);
// dy / h / dx: shared [FR][FH], dx must not alias dy
__device__ __noinline__ void f_bwd_hidden(const float* __restrict__ W, const float* dy, const float* h, float* dx,
                                          float* gdx, float* part) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const float4* Wp = reinterpret_cast<const float4*>(W + (size_t)(w * 32) * FH) + l;
  float4 a[2][8], c[2][8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    a[0][u] = __ldg(Wp + u * 64);
    c[0][u] = __ldg(Wp + u * 64 + 32);
  }
  __syncthreads();                               // dy (and h) visible; part free again
  float4 a0[FR], a1[FR];
#pragma unroll
  for (int r = 0; r < FR; ++r) a0[r] = a1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g < 3) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        a[(g + 1) & 1][u] = __ldg(Wp + ((g + 1) * 8 + u) * 64);
        c[(g + 1) & 1][u] = __ldg(Wp + ((g + 1) * 8 + u) * 64 + 32);
      }
    }
#pragma unroll
    for (int r = 0; r < FR; ++r) {
      const float4 d0 = *reinterpret_cast<const float4*>(dy + r * FH + w * 32 + g * 8);
      const float4 d1 = *reinterpret_cast<const float4*>(dy + r * FH + w * 32 + g * 8 + 4);
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 wa = a[g & 1][u], wc = c[g & 1][u];
        a0[r].x = fmaf(d[u], wa.x, a0[r].x); a0[r].y = fmaf(d[u], wa.y, a0[r].y);
        a0[r].z = fmaf(d[u], wa.z, a0[r].z); a0[r].w = fmaf(d[u], wa.w, a0[r].w);
        a1[r].x = fmaf(d[u], wc.x, a1[r].x); a1[r].y = fmaf(d[u], wc.y, a1[r].y);
        a1[r].z = fmaf(d[u], wc.z, a1[r].z); a1[r].w = fmaf(d[u], wc.w, a1[r].w);
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
  // (f_fwd_first / f_fwd_hidden / f_fwd_out / f_bwd_hidden START with the barrier that publishes what was written before them)
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

  // ---- target: y = clamp(r + gamma Q'(x', pi'(x')), -clip, 0)  (ddpg_agent.py:252-260) ---------------------------------------
  f_fwd_first(Pat + A.wa[0], Pat + A.ba[0], s_in[0], Dx, sA, nullptr);
  f_fwd_hidden<true>(Pat + A.wa[1], Pat + A.ba[1], sA, sB, nullptr);
  f_fwd_hidden<true>(Pat + A.wa[2], Pat + A.ba[2], sB, sA, nullptr);
  f_fwd_out(Pat + A.wa[3], Pat + A.ba[3], sA, Da, s_z);
  __syncthreads();
  if (t < FR * Da) {
    const int r = t / Da, c = t - r * Da;
    const float a = A.amax * tanhf(s_z[r * FOUT + c]);
    s_in[0][r * FIN + Dx + c] = __fdiv_rn(a, A.amax);
  }
  f_fwd_first(Pct + A.wc[0], Pct + A.bc[0], s_in[0], Dc, sA, nullptr);
  f_fwd_hidden<true>(Pct + A.wc[1], Pct + A.bc[1], sA, sB, nullptr);
  f_fwd_hidden<true>(Pct + A.wc[2], Pct + A.bc[2], sB, sA, nullptr);
  f_fwd_out(Pct + A.wc[3], Pct + A.bc[3], sA, 1, s_qn);

  // ---- critic(x, a): forward, loss, delta chain (ddpg_agent.py:262-263,274-275) ---------------------------------------------
  f_fwd_first(Pc + A.wc[0], Pc + A.bc[0], s_in[1], Dc, ch1, A.ch1 + g0);
  f_fwd_hidden<true>(Pc + A.wc[1], Pc + A.bc[1], ch1, ch2, A.ch2 + g0);
  f_fwd_hidden<true>(Pc + A.wc[2], Pc + A.bc[2], ch2, ch3, A.ch3 + g0);
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
  f_bwd_hidden(Pc + A.wc[2], sD0, ch2, sD1, A.cd2 + g0, s_part);
  f_bwd_hidden(Pc + A.wc[1], sD1, ch1, sD0, A.cd1 + g0, s_part);

  // ---- actor(x) and critic(x, pi(x)) forward (ddpg_agent.py:265-267) -------------------------------------------------------
  f_fwd_first(Pa + A.wa[0], Pa + A.ba[0], s_in[2], Dx, ah1, A.ah1 + g0);
  f_fwd_hidden<true>(Pa + A.wa[1], Pa + A.ba[1], ah1, ah2, A.ah2 + g0);
  f_fwd_hidden<true>(Pa + A.wa[2], Pa + A.ba[2], ah2, ah3, A.ah3 + g0);
  f_fwd_out(Pa + A.wa[3], Pa + A.ba[3], ah3, Da, s_z);
  __syncthreads();
  if (t < FR * Da) {
    const int r = t / Da, c = t - r * Da;
    const float a = A.amax * tanhf(s_z[r * FOUT + c]);
    s_th[r * FOUT + c] = a / A.amax;
    s_in[2][r * FIN + Dx + c] = __fdiv_rn(a, A.amax);
  }
  f_fwd_first(Pc + A.wc[0], Pc + A.bc[0], s_in[2], Dc, qh1, nullptr);
  f_fwd_hidden<true>(Pc + A.wc[1], Pc + A.bc[1], qh1, qh2, nullptr);
  f_fwd_hidden<true>(Pc + A.wc[2], Pc + A.bc[2], qh2, qh3, nullptr);
  f_fwd_out(Pc + A.wc[3], Pc + A.bc[3], qh3, 1, s_qa);
  if (t < FR * FOUT) s_dz[t] = -1.0f / (float)A.B;          // d(-mean Q)/dQ   (s_dz: last read two barriers ago)
  __syncthreads();

  // ---- dQ/da through the critic, then the actor's delta chain (ddpg_agent.py:266-270) -------------------------------------
  f_bwd_out(Pc + A.wc[3], s_dz, 1, qh3, sD0, nullptr);
  f_bwd_hidden(Pc + A.wc[2], sD0, qh2, sD1, nullptr, s_part);
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
  f_bwd_hidden(Pa + A.wa[2], sD0, ah2, sD1, A.fd2 + g0, s_part);
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

constexpr int WG_T = 128;       // threads: 4 (j) x 4 (k) outputs each
__global__ void __launch_bounds__(WG_T) ddpg_wgrad_kernel(const WgradArgs G) {
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
  if (blockIdx.x == gridDim.x - 1 && t == WG_T - 1) {     // the two scalar losses (an idle thread of the last tile's epilogue)
    float sq = 0.f, st = 0.f, sc = 0.f;
    for (int i = 0; i < G.n_part; ++i) {
      sq += G.loss_part[i * 4 + 0];
      st += G.loss_part[i * 4 + 1];
      sc += G.loss_part[i * 4 + 2];
    }
    G.losses[0] = -sq / (float)G.B + G.l2 * st / (float)(G.B * G.Da);
    G.losses[1] = sc / (float)G.B;
  }
  if (Nj <= FOUT) {
    // output layers (1 or act_dim rows): ONE CTA for the whole product, so that the grid stays below one CTA per SM
    // (152 tiles on 148 SMs doubled the kernel's duration: two CTAs shared an SM).  Thread t owns columns t and t + 128.
    float acc[FOUT][2], bs[FOUT];
#pragma unroll
    for (int j = 0; j < FOUT; ++j) acc[j][0] = acc[j][1] = bs[j] = 0.f;
    const bool k0v = t < Nk, k1v = t + WG_T < Nk;
    float* sDj = &sA[0][0];                                // [rows of a 256-row block][FOUT], 8 KB of the staging buffer
    for (int rb = 0; rb < G.B; rb += 256) {
      const int nr = min(256, G.B - rb);
      __syncthreads();
      for (int i = t; i < nr * FOUT; i += WG_T) {
        const int r = i / FOUT, j = i - r * FOUT;
        sDj[i] = j < Nj ? D[(size_t)(rb + r) * ldD + j] : 0.f;
      }
      __syncthreads();
      for (int r0 = 0; r0 < nr; r0 += 8) {                 // nr is a multiple of 32 (checked on the host)
        float a0[8], a1[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          a0[u] = k0v ? Am[(size_t)(rb + r0 + u) * ldA + t] : 0.f;
          a1[u] = k1v ? Am[(size_t)(rb + r0 + u) * ldA + t + WG_T] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int j = 0; j < FOUT; ++j) {
            const float d = sDj[(r0 + u) * FOUT + j];
            acc[j][0] = fmaf(d, a0[u], acc[j][0]);
            acc[j][1] = fmaf(d, a1[u], acc[j][1]);
            bs[j] += d;
          }
      }
    }
#pragma unroll
    for (int j = 0; j < FOUT; ++j) {
      if (j >= Nj) break;
      if (k0v) G.gW[p][(size_t)j * Nk + t] = acc[j][0];
      if (k1v) G.gW[p][(size_t)j * Nk + t + WG_T] = acc[j][1];
      if (t == 0) G.gb[p][j] = bs[j];
    }
    return;
  }
  // staging roles: D chunk 32 x 32 = 8 per thread (row t / 32 + 4 i, col t % 32); A chunk 32 x 64 = 16 per thread
  const int dc = t & 31, dr = t >> 5, ac = t & 63, ar = t >> 6;
  const bool d_ok = j0 + dc < Nj, a_ok = k0 + ac < Nk;
  float pd[8], pa[16];
  auto fetch = [&](int r0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) pd[i] = d_ok ? D[(size_t)(r0 + dr + 4 * i) * ldD + j0 + dc] : 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) pa[i] = a_ok ? Am[(size_t)(r0 + ar + 2 * i) * ldA + k0 + ac] : 0.f;
  };
  // compute roles: thread owns outputs j = jq * 4 + {0..3}, k = kq * 4 + {0..3}: two 16-byte shared loads per 16 FFMA
  const int kq = t & 15, jq = t >> 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float bsum = 0.f;                                         // thread t < 32: column sum of D (bias gradient)
  fetch(0);
  for (int r0 = 0; r0 < G.B; r0 += WG_RC) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) sD[dr + 4 * i][dc] = pd[i];
#pragma unroll
    for (int i = 0; i < 16; ++i) sA[ar + 2 * i][ac] = pa[i];
    __syncthreads();
    if (r0 + WG_RC < G.B) fetch(r0 + WG_RC);
#pragma unroll
    for (int r = 0; r < WG_RC; ++r) {
      const float4 d = *reinterpret_cast<const float4*>(&sD[r][jq * 4]);
      const float4 a = *reinterpret_cast<const float4*>(&sA[r][kq * 4]);
      const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(dv[i], a.x, acc[i][0]); acc[i][1] = fmaf(dv[i], a.y, acc[i][1]);
        acc[i][2] = fmaf(dv[i], a.z, acc[i][2]); acc[i][3] = fmaf(dv[i], a.w, acc[i][3]);
      }
    }
    if (k0 == 0 && t < WG_TJ) {
#pragma unroll
      for (int r = 0; r < WG_RC; ++r) bsum += sD[r][t];
    }
  }
  float* __restrict__ gW = G.gW[p];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = j0 + jq * 4 + a;
    if (j >= Nj) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = k0 + kq * 4 + b;
      if (k < Nk) gW[(size_t)j * Nk + k] = acc[a][b];
    }
  }
  if (k0 == 0 && t < WG_TJ && j0 + t < Nj) G.gb[p][j0 + t] = bsum;
}

}  // namespace bmi
