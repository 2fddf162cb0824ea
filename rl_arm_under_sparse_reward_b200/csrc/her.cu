// Episode store, sparse reward and HER "future" relabelling kernels.
//
// Reference behaviour restated here (paths relative to the reference tree):
//   replay_buffer.py:32-43   store_episode        -> store_kernel
//   bmirobot_env_push_F.py:20-23,84-90  reward    -> reward_kernel / inside her kernels
//   her.py:13-41             sample_her_transitions -> her_gather_kernel / her_inputs_lane_kernel (wide layouts: her_inputs_kernel)
//   ddpg_agent.py:229-248    clip + normalise + concat + f32 cast -> her_inputs_lane_kernel
//
// Bit-exactness: every float64 operation that numpy performs is issued with the
// round-to-nearest intrinsics (__dadd_rn/__dmul_rn/...) so nvcc cannot contract them
// into FMAs; numpy's reduction order for a 3-vector norm is ((s0+s1)+s2) (verified
// against the unmodified reference: tests/test_oracle_learner.py::test_her_bit_exact_vs_reference and
// tests/test_gpu_her.py::test_reward_kernel_matches_numpy).
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace bmi {

template <typename T>
__device__ __forceinline__ double ld_as_f64(const T* p) {
  return (double)(*p);
}

// ---- store ---------------------------------------------------------------------------
template <typename TD, typename TS>
__global__ void store_kernel(TD* __restrict__ dst, const TS* __restrict__ src,
                             const int64_t* __restrict__ slots, int64_t row_elems) {
  const int64_t e = blockIdx.x;
  const int64_t slot = slots[e];
  if (slot < 0) return;  // superseded duplicate (numpy fancy assignment: the last write wins)
  const TS* s = src + e * row_elems;
  TD* d = dst + slot * row_elems;
  for (int64_t i = blockIdx.y * blockDim.x + threadIdx.x; i < row_elems;
       i += (int64_t)gridDim.y * blockDim.x)
    d[i] = (TD)s[i];
}

template <typename TD, typename TS>
static int store_key(void* dst, const void* src, const int64_t* slots, int64_t n_ep,
                     int64_t row_elems, cudaStream_t st) {
  if (n_ep == 0) return BMI_OK;
  int gy = (int)((row_elems + 255) / 256);
  if (gy > 16) gy = 16;
  dim3 grid((unsigned)n_ep, gy);
  store_kernel<TD, TS><<<grid, 256, 0, st>>>((TD*)dst, (const TS*)src, slots, row_elems);
  BMI_LAUNCHED();
  return BMI_OK;
}

// ---- reward --------------------------------------------------------------------------
template <typename T>
__global__ void reward_kernel(const T* __restrict__ ag, const T* __restrict__ g, int64_t n,
                              int gd, double thr, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < gd; ++k) {
    double d = __dsub_rn((double)ag[i * gd + k], (double)g[i * gd + k]);
    double sq = __dmul_rn(d, d);
    s = (k == 0) ? sq : __dadd_rn(s, sq);
  }
  double dist = __dsqrt_rn(s);
  out[i] = (dist > thr) ? -1.0f : -0.0f;  // -(bool).astype(f32): False -> -0.0
}

// ---- HER gather ------------------------------------------------------------------------
struct HerRow {
  int64_t ep, t, ft;
  bool her;
};

__device__ __forceinline__ HerRow her_row(const int64_t* ep_idx, const int64_t* t_idx,
                                          const double* u_her, const double* u_off, int64_t b,
                                          int T, double future_p) {
  HerRow r;
  r.ep = ep_idx[b];
  r.t = t_idx[b];
  r.her = u_her[b] < future_p;                                   // her.py:28
  double off = __dmul_rn(u_off[b], (double)((int64_t)T - r.t));  // her.py:30
  r.ft = r.t + 1 + (int64_t)off;                                 // her.py:31-32 (trunc)
  return r;
}

// distance of two goal vectors held one component per lane (lanes [0,gd)), numpy order
__device__ __forceinline__ double warp_goal_dist(double a, double g, int gd, int lane) {
  double d = __dsub_rn(a, g);
  double sq = __dmul_rn(d, d);
  double s = __shfl_sync(0xffffffffu, sq, 0);
  for (int k = 1; k < gd; ++k) s = __dadd_rn(s, __shfl_sync(0xffffffffu, sq, k));
  return __dsqrt_rn(s);
}

// one warp per sampled transition
template <typename T>
__global__ void her_gather_kernel(bmi_episodes buf, const int64_t* __restrict__ ep_idx,
                                  const int64_t* __restrict__ t_idx,
                                  const double* __restrict__ u_her,
                                  const double* __restrict__ u_off, int64_t B, double future_p,
                                  double thr, bmi_transitions out) {
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int Tn = buf.T, Do = buf.obs_dim, Dg = buf.goal_dim, Da = buf.act_dim;
  const HerRow r = her_row(ep_idx, t_idx, u_her, u_off, b, Tn, future_p);
  const T* obs = (const T*)buf.obs + (r.ep * (Tn + 1) + r.t) * Do;  // row t, row t+1 follows
  const T* ag = (const T*)buf.ag + (r.ep * (Tn + 1) + r.t) * Dg;
  const T* gsrc = r.her ? (const T*)buf.ag + (r.ep * (Tn + 1) + r.ft) * Dg
                        : (const T*)buf.g + (r.ep * Tn + r.t) * Dg;
  const T* act = (const T*)buf.actions + (r.ep * Tn + r.t) * Da;
  for (int i = lane; i < Do; i += 32) {
    T o = obs[i], on = obs[Do + i];
    if (out.obs) ((T*)out.obs)[b * Do + i] = o;
    if (out.obs_next) ((T*)out.obs_next)[b * Do + i] = on;
  }
  for (int i = lane; i < Da; i += 32)
    if (out.actions) ((T*)out.actions)[b * Da + i] = act[i];
  T a0 = 0, a1 = 0, gv = 0;
  if (lane < Dg) {
    a0 = ag[lane];
    a1 = ag[Dg + lane];
    gv = gsrc[lane];
    if (out.ag) ((T*)out.ag)[b * Dg + lane] = a0;
    if (out.ag_next) ((T*)out.ag_next)[b * Dg + lane] = a1;
    if (out.g) ((T*)out.g)[b * Dg + lane] = gv;
  }
  double dist = warp_goal_dist((double)a1, (double)gv, Dg, lane);
  if (lane == 0 && out.r) out.r[b] = (dist > thr) ? -1.0f : -0.0f;
}

__device__ __forceinline__ double clipd(double v, double lim) {
  // np.clip(v, -lim, lim) == minimum(maximum(v, -lim), lim)
  return fmin(fmax(v, -lim), lim);
}

__device__ __forceinline__ float norm_clip_f32(double v, double clip_obs, float mean, float stdv,
                                               double clip_range) {
  double c = clipd(v, clip_obs);                                        // _preproc_og
  double z = __ddiv_rn(__dsub_rn(c, (double)mean), (double)stdv);       // normalizer.normalize
  return (float)clipd(z, clip_range);                                   // torch.tensor(f32)
}

// Fused sampler for the learner: one block gathers a TILE of 64 sampled transitions element-parallel.  64 threads
// resolve the tile's (episode, t, future t) rows into shared memory; then every thread owns several (sample, element)
// items and issues all its loads before the first use, so a block keeps ~3.5 K independent 4-byte loads in flight
// instead of one 108-byte row per warp (the one-warp-per-sample version was latency bound at 8 % of the HBM roofline).
// Arithmetic per element is unchanged (float64 clip / normalise / clip, then the float32 cast of ddpg_agent.py:229-248).
// Small batches (one DDPG update: 256 samples) use 8-sample tiles so that the batch still spreads over 32 SMs.
constexpr int HER_THREADS = 256;
template <typename T, int HER_TILE>
__global__ void __launch_bounds__(HER_THREADS)
her_inputs_kernel(bmi_episodes buf, const int64_t* __restrict__ ep_idx,
                  const int64_t* __restrict__ t_idx, const double* __restrict__ u_her,
                  const double* __restrict__ u_off, int64_t B, double future_p, double thr,
                  double clip_obs, double clip_range, const float* __restrict__ o_mean,
                  const float* __restrict__ o_std, const float* __restrict__ g_mean,
                  const float* __restrict__ g_std, float* __restrict__ x, float* __restrict__ xn,
                  float* __restrict__ actions, float* __restrict__ rew) {
  __shared__ int64_t s_row[HER_TILE];    // (ep * (T + 1) + t): row index into obs / ag
  __shared__ int64_t s_goal[HER_TILE];   // row index of the goal source (>= 0: ag row, < 0: ~row of g)
  __shared__ int64_t s_tr[HER_TILE];     // (ep * T + t): row index into g / actions
  const int64_t b0 = (int64_t)blockIdx.x * HER_TILE;
  const int nb = (int)((B - b0) < HER_TILE ? (B - b0) : HER_TILE);
  const int Tn = buf.T, Do = buf.obs_dim, Dg = buf.goal_dim, Da = buf.act_dim;
  const int Dx = Do + Dg;
  if ((int)threadIdx.x < nb) {
    const HerRow r = her_row(ep_idx, t_idx, u_her, u_off, b0 + threadIdx.x, Tn, future_p);
    s_row[threadIdx.x] = r.ep * (Tn + 1) + r.t;
    s_tr[threadIdx.x] = r.ep * Tn + r.t;
    s_goal[threadIdx.x] = r.her ? (r.ep * (Tn + 1) + r.ft) : ~(r.ep * Tn + r.t);
  }
  __syncthreads();
  const T* obs = (const T*)buf.obs;
  const T* agb = (const T*)buf.ag;
  const T* gb = (const T*)buf.g;
  const T* actb = (const T*)buf.actions;
  // Warp roles, so that every kind of item is ONE memory round trip running concurrently with the others:
  // warps 0..5 observation pairs, warp 6 goals + rewards, warp 7 actions.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int OBS_WARPS = HER_THREADS / 32 - 2, OBS_THREADS = OBS_WARPS * 32;
  if (warp < OBS_WARPS) {
    // ---- observation pairs (row t and row t + 1 are adjacent in the episode-major store) ----------------------------
    constexpr int UNR = (HER_TILE * 27 + OBS_THREADS - 1) / OBS_THREADS;   // one batch of loads covers a 27-wide tile
    const int n_obs = nb * Do;
    for (int base = threadIdx.x; base < n_obs; base += OBS_THREADS * UNR) {
      T o[UNR], on[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int idx = base + u * OBS_THREADS;
        o[u] = 0; on[u] = 0;
        if (idx < n_obs) {
          const int sI = idx / Do, i = idx - sI * Do;
          const T* p = obs + s_row[sI] * Do + i;
          o[u] = p[0];
          on[u] = p[Do];
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int idx = base + u * OBS_THREADS;
        if (idx < n_obs) {
          const int sI = idx / Do, i = idx - sI * Do;
          const float m = o_mean[i], sd = o_std[i];
          const int64_t b = b0 + sI;
          x[b * Dx + i] = norm_clip_f32((double)o[u], clip_obs, m, sd, clip_range);
          xn[b * Dx + i] = norm_clip_f32((double)on[u], clip_obs, m, sd, clip_range);
        }
      }
    }
  } else if (warp == OBS_WARPS) {
    // ---- goals (relabelled or stored) and rewards: lane = sample (two passes for a 64-sample tile) --------------------
    for (int sI = lane; sI < nb; sI += 32) {
      const int64_t gr = s_goal[sI];
      const T* an = agb + (s_row[sI] + 1) * Dg;
      const T* gs = gr >= 0 ? agb + gr * Dg : gb + (~gr) * Dg;
      const int64_t b = b0 + sI;
      double acc = 0.0;   // numpy order: ((d0^2 + d1^2) + d2^2), then sqrt
      for (int k = 0; k < Dg; ++k) {
        const T gv = gs[k];
        const float gn = norm_clip_f32((double)gv, clip_obs, g_mean[k], g_std[k], clip_range);
        x[b * Dx + Do + k] = gn;
        xn[b * Dx + Do + k] = gn;  // g_next is the same relabelled goal (ddpg_agent.py:231)
        const double d = __dsub_rn((double)an[k], (double)gv);
        const double sq = __dmul_rn(d, d);
        acc = k == 0 ? sq : __dadd_rn(acc, sq);
      }
      rew[b] = (__dsqrt_rn(acc) > thr) ? -1.0f : -0.0f;
    }
  } else {
    // ---- actions -------------------------------------------------------------------------------------------------------
    for (int idx = lane; idx < nb * Da; idx += 32) {
      const int sI = idx / Da, k = idx - sI * Da;
      actions[(b0 + sI) * Da + k] = (float)actb[s_tr[sI] * Da + k];
    }
  }
}

// (c - m) / sd rounded to float64 exactly like __ddiv_rn, at the price of one multiplication: q~ = d * (1 / sd) is
// within 3 units of the correctly rounded quotient, so after the float32 cast the two can only differ when q~ lies
// within a few units of a float32 rounding boundary (low 29 mantissa bits == 0x10000000), of the float32 subnormal
// range or of a non-finite value -- those rare lanes take the exact division.
__device__ __noinline__ double ddiv_exact(double d, double sd) { return __ddiv_rn(d, sd); }
__device__ __forceinline__ double div_as_f32_exact(double d, double sd, double rsd) {
  double q = __dmul_rn(d, rsd);
  const unsigned lo = (unsigned)__double2loint(q) & 0x1FFFFFFFu;
  const unsigned hi = (unsigned)__double2hiint(q) & 0x7FF00000u;
  const bool risky = (lo - 0x0FFFFFF8u) <= 16u || ((hi - 0x38400000u) >= 0x44C00000u && d != 0.0);   // 2^-123 <= |q| < 2^977
  if (risky) q = ddiv_exact(d, sd);
  return q;
}

// clip -> normalise -> clip -> float32 of one stored value (ddpg_agent.py:229-248 with normalizer.py:66-70).  F32CLIP: both
// clip limits are float32-representable, so the outer clip commutes with the (monotonic) float32 rounding and, for float32
// storage, the inner one can run before the widening -- single FMNMX instructions instead of float64 compare / select pairs.
template <typename T, bool F32CLIP>
__device__ __forceinline__ float norm_clip_lane(T v, double clip_obs, double m, double sd, double rsd, double clip_range) {
  double c;
  if (F32CLIP && sizeof(T) == 4) {
    const float lim = (float)clip_obs;
    c = (double)fminf(fmaxf((float)v, -lim), lim);
  } else {
    c = clipd((double)v, clip_obs);
  }
  const double q = div_as_f32_exact(__dsub_rn(c, m), sd, rsd);
  if (F32CLIP) {
    const float lim = (float)clip_range;
    return fminf(fmaxf((float)q, -lim), lim);
  }
  return (float)clipd(q, clip_range);
}

// Lane-per-element sampler for the common narrow layouts (obs_dim + goal_dim <= 32, act_dim <= 32, < 2^32 stored rows): a
// warp handles HER_S sampled transitions at a time and lane i owns column i of the network input -- lanes [0, Do) the
// observation pair (two loads, rows t and t + 1), lanes [Do, Do + Dg) the (relabelled) goal and ag_next, lanes [0, Da)
// also the action.  The sampled rows of a 16-sample chunk are resolved one per lane (coalesced index loads) and broadcast
// with shuffles; all loads of HER_S samples are issued before the first use; both input rows leave as one coalesced
// 4 * Dx-byte store each.  A persistent grid (8 blocks per SM) amortises the per-lane constants (mean, std, 1 / std).
// sq_thr is the largest float64 whose correctly rounded square root is <= thr, so "sqrt(acc) > thr" is "acc > sq_thr".
// Results identical bit for bit to the float64-division chain.
// HER_CH: samples per warp chunk, their (episode, t, future t) rows resolved one per lane; HER_S: samples whose loads are in
// flight together.  16 / 8 for bandwidth-bound batches; 4 / 4 for launch-bound ones (a warp's chunk is then ONE memory round
// trip and a batch of 256 still spreads over 64 warps).  DG: goal_dim known at compile time (0: use buf.goal_dim)
template <typename T, bool F32CLIP, int DG, int HER_CH, int HER_S>
__global__ void __launch_bounds__(128, 8)
her_inputs_lane_kernel(bmi_episodes buf, const int64_t* __restrict__ ep_idx, const int64_t* __restrict__ t_idx,
                       const double* __restrict__ u_her, const double* __restrict__ u_off, int64_t B, double future_p,
                       double sq_thr, double clip_obs, double clip_range, const float* __restrict__ o_mean,
                       const float* __restrict__ o_std, const float* __restrict__ g_mean,
                       const float* __restrict__ g_std, float* __restrict__ x, float* __restrict__ xn,
                       float* __restrict__ actions, float* __restrict__ rew) {
  const int lane = threadIdx.x & 31;
  const unsigned Tn = buf.T, Do = buf.obs_dim, Dg = DG ? DG : buf.goal_dim, Da = buf.act_dim, Dx = Do + Dg;
  const bool is_obs = lane < Do, is_goal = !is_obs && lane < Dx;
  const int kg = lane - Do;
  double m = 0.0, sd = 1.0;
  if (is_obs) { m = (double)o_mean[lane]; sd = (double)o_std[lane]; }
  if (is_goal) { m = (double)g_mean[kg]; sd = (double)g_std[kg]; }
  const double rsd = __ddiv_rn(1.0, sd);
  // per-lane address recipe: value A = baseA[idxA * strideA], value B = baseB[(row + 1) * strideA]
  const T* const agk = (const T*)buf.ag + (is_goal ? kg : 0);
  const T* const gk = (const T*)buf.g + (is_goal ? kg : 0);
  const T* const obl = (const T*)buf.obs + (is_obs ? lane : 0);
  const T* const baseB = is_goal ? agk : obl;
  const unsigned strideA = is_goal ? Dg : Do;
  const unsigned Bu = (unsigned)B;                        // the host guarantees B * max(Dx, Da) < 2^32
  const unsigned n_chunks = (Bu + HER_CH - 1) / HER_CH;
  const unsigned warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned n_warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned chunk = warp0; chunk < n_chunks; chunk += n_warps) {
    const unsigned c0 = chunk * HER_CH;
    // ---- lane l resolves sample c0 + (l mod HER_CH); the tail repeats the last sample (not stored) --------------------
    unsigned l_row, l_tr, l_goal, her_mask;
    {
      const unsigned b = min(c0 + (lane & (HER_CH - 1)), Bu - 1u);
      const unsigned ep = (unsigned)ep_idx[b], t = (unsigned)t_idx[b];
      const bool her = u_her[b] < future_p;                                                                  // her.py:28
      const unsigned ft = t + 1u + (unsigned)__double2int_rz(__dmul_rn(u_off[b], (double)(int)(Tn - t)));    // her.py:30-32
      l_row = ep * (Tn + 1u) + t;
      l_tr = ep * Tn + t;
      l_goal = her ? l_row - t + ft : l_tr;     // row of ag (relabelled) or of g (stored goal)
      her_mask = __ballot_sync(0xffffffffu, her);
    }
    // ---- actions of the whole chunk: element j = (sample, component) -----------------------------------------------------
    for (unsigned j0 = 0; j0 < HER_CH * Da; j0 += 32) {      // uniform trip count: the shuffle needs every lane
      const unsigned j = j0 + lane;
      const bool valid = j < HER_CH * Da;
      const unsigned sI = valid ? j / Da : 0u, k = j - sI * Da;
      const unsigned tr = __shfl_sync(0xffffffffu, l_tr, sI);
      if (valid && c0 + sI < Bu) actions[(c0 + sI) * Da + k] = (float)((const T*)buf.actions)[(uint64_t)tr * Da + k];
    }
#pragma unroll
    for (int h = 0; h < HER_CH / HER_S; ++h) {
      T va[HER_S], vb[HER_S];
#pragma unroll
      for (int s = 0; s < HER_S; ++s) {
        const int src = h * HER_S + s;
        const unsigned row = __shfl_sync(0xffffffffu, l_row, src);
        const unsigned gl = __shfl_sync(0xffffffffu, l_goal, src);
        const bool her = (her_mask >> src) & 1u;
        const T* baseA = is_goal ? (her ? agk : gk) : obl;
        va[s] = baseA[(uint64_t)(is_goal ? gl : row) * strideA];
        vb[s] = baseB[(uint64_t)(row + 1u) * strideA];
      }
#pragma unroll
      for (int s = 0; s < HER_S; ++s) {
        const unsigned b = c0 + h * HER_S + s;
        const float xa = norm_clip_lane<T, F32CLIP>(va[s], clip_obs, m, sd, rsd, clip_range);
        float xb = xa;                                    // g_next is the same relabelled goal (ddpg_agent.py:231)
        if (is_obs) xb = norm_clip_lane<T, F32CLIP>(vb[s], clip_obs, m, sd, rsd, clip_range);
        // reward: numpy order ((d0^2 + d1^2) + d2^2) (bmirobot_env_push_F.py:84-90)
        const double d = __dsub_rn((double)vb[s], (double)va[s]);
        const double sq = __dmul_rn(d, d);
        double acc = __shfl_sync(0xffffffffu, sq, Do);
#pragma unroll
        for (unsigned k = 1; k < Dg; ++k) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, sq, Do + k));
        if (b < Bu) {
          if (lane < Dx) {
            x[b * Dx + lane] = xa;
            xn[b * Dx + lane] = xb;
          }
          if (lane == 0) rew[b] = (acc > sq_thr) ? -1.0f : -0.0f;
        }
      }
    }
  }
}

// largest float64 a with sqrt_rn(a) <= thr (sqrt_rn is monotonic, the host's sqrt is correctly rounded)
static double sqrt_threshold(double thr) {
  if (!(thr >= 0.0)) return -1.0;                 // sqrt(acc) >= 0 > thr for every acc (NaN thr: comparison false anyway)
  double a = thr * thr;
  while (std::sqrt(a) <= thr) a = std::nextafter(a, INFINITY);
  while (std::sqrt(a) > thr) a = std::nextafter(a, -INFINITY);
  return a;
}

// ---- device-side draws -------------------------------------------------------------------
__global__ void her_draw_kernel(uint64_t seed, const uint64_t* __restrict__ counter, int64_t B,
                                const int64_t* __restrict__ n_valid_p, int T,
                                int64_t* __restrict__ ep_idx, int64_t* __restrict__ t_idx,
                                double* __restrict__ u_her, double* __restrict__ u_off) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint64_t c = *counter + (uint64_t)b;
  const int64_t n_valid = *n_valid_p;
  Philox4 p0 = philox4x32_10(seed, 2 * c, kStreamHer);
  Philox4 p1 = philox4x32_10(seed, 2 * c + 1, kStreamHer);
  int64_t ep = (int64_t)__dmul_rn(u53(p0.v[0], p0.v[1]), (double)n_valid);
  if (ep >= n_valid) ep = n_valid - 1;
  ep_idx[b] = ep;
  t_idx[b] = (int64_t)(((uint64_t)p0.v[2] * (uint64_t)T) >> 32);
  u_her[b] = u53(p1.v[0], p1.v[1]);
  u_off[b] = u53(p1.v[2], p1.v[3]);
}

__global__ void advance_counter_kernel(uint64_t* counter, uint64_t by) { *counter += by; }

// ---- rollout record ----------------------------------------------------------------------
template <typename T>
__global__ void rollout_record_kernel(bmi_episodes ep, int t, const float* __restrict__ obs,
                                      const float* __restrict__ ag, const float* __restrict__ g,
                                      const float* __restrict__ act) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= ep.n_episodes) return;
  const int Tn = ep.T, Do = ep.obs_dim, Dg = ep.goal_dim, Da = ep.act_dim;
  T* o = (T*)ep.obs + (e * (Tn + 1) + t) * Do;
  for (int i = lane; i < Do; i += 32) o[i] = (T)obs[e * Do + i];
  if (lane < Dg) ((T*)ep.ag)[(e * (Tn + 1) + t) * Dg + lane] = (T)ag[e * Dg + lane];
  if (t < Tn && g != nullptr && act != nullptr) {
    if (lane < Dg) ((T*)ep.g)[(e * Tn + t) * Dg + lane] = (T)g[e * Dg + lane];
    if (lane < Da) ((T*)ep.actions)[(e * Tn + t) * Da + lane] = (T)act[e * Da + lane];
  }
}

static int check_eps(const bmi_episodes* b, const char* what) {
  BMI_REQUIRE(b != nullptr, "%s: null episodes", what);
  BMI_REQUIRE(b->dtype == BMI_F32 || b->dtype == BMI_F64, "%s: bad dtype %d", what, b->dtype);
  BMI_REQUIRE(b->T > 0 && b->obs_dim > 0 && b->goal_dim > 0 && b->goal_dim <= 32 &&
                  b->act_dim > 0 && b->act_dim <= 32,
              "%s: bad dims T=%d obs=%d goal=%d act=%d", what, b->T, b->obs_dim, b->goal_dim,
              b->act_dim);
  return BMI_OK;
}

}  // namespace bmi

using namespace bmi;

extern "C" int bmi_buffer_store(const bmi_episodes* dst, const bmi_episodes* src,
                                const int64_t* slots, bmi_stream_t stream) {
  int rc;
  if ((rc = check_eps(dst, "bmi_buffer_store(dst)"))) return rc;
  if ((rc = check_eps(src, "bmi_buffer_store(src)"))) return rc;
  BMI_REQUIRE(dst->T == src->T && dst->obs_dim == src->obs_dim && dst->goal_dim == src->goal_dim &&
                  dst->act_dim == src->act_dim,
              "bmi_buffer_store: src/dst dims differ");
  BMI_REQUIRE(slots != nullptr || src->n_episodes == 0, "bmi_buffer_store: null slots");
  cudaStream_t st = as_stream(stream);
  const int64_t n = src->n_episodes;
  const int64_t T = src->T;
  const int64_t rows[4] = {(T + 1) * src->obs_dim, (T + 1) * src->goal_dim, T * src->goal_dim,
                           T * src->act_dim};
  void* d[4] = {dst->obs, dst->ag, dst->g, dst->actions};
  const void* s[4] = {src->obs, src->ag, src->g, src->actions};
  for (int k = 0; k < 4; ++k) {
    if (dst->dtype == BMI_F64 && src->dtype == BMI_F64)
      rc = store_key<double, double>(d[k], s[k], slots, n, rows[k], st);
    else if (dst->dtype == BMI_F64)
      rc = store_key<double, float>(d[k], s[k], slots, n, rows[k], st);
    else if (src->dtype == BMI_F64)
      rc = store_key<float, double>(d[k], s[k], slots, n, rows[k], st);
    else
      rc = store_key<float, float>(d[k], s[k], slots, n, rows[k], st);
    if (rc) return rc;
  }
  return BMI_OK;
}

extern "C" int bmi_compute_reward(const void* ag, const void* g, int64_t n, int32_t goal_dim,
                                  int32_t dtype, double thr, float* out, bmi_stream_t stream) {
  BMI_REQUIRE(n >= 0 && goal_dim > 0, "bmi_compute_reward: bad sizes n=%lld gd=%d", (long long)n,
              goal_dim);
  BMI_REQUIRE(dtype == BMI_F32 || dtype == BMI_F64, "bmi_compute_reward: bad dtype %d", dtype);
  if (n == 0) return BMI_OK;
  BMI_REQUIRE(ag && g && out, "bmi_compute_reward: null pointer");
  unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == BMI_F64)
    reward_kernel<double><<<grid, 256, 0, as_stream(stream)>>>((const double*)ag, (const double*)g,
                                                               n, goal_dim, thr, out);
  else
    reward_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)ag, (const float*)g, n,
                                                              goal_dim, thr, out);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_her_sample(const bmi_episodes* buf, int64_t n_valid, const int64_t* ep_idx,
                              const int64_t* t_idx, const double* u_her, const double* u_off,
                              int64_t B, double future_p, double thr, const bmi_transitions* out,
                              bmi_stream_t stream) {
  int rc;
  if ((rc = check_eps(buf, "bmi_her_sample"))) return rc;
  BMI_REQUIRE(B >= 0, "bmi_her_sample: negative batch");
  if (B == 0) return BMI_OK;
  BMI_REQUIRE(n_valid > 0 && n_valid <= buf->n_episodes,
              "bmi_her_sample: n_valid=%lld outside (0, %lld] (sampling an empty buffer)",
              (long long)n_valid, (long long)buf->n_episodes);
  BMI_REQUIRE(ep_idx && t_idx && u_her && u_off && out, "bmi_her_sample: null pointer");
  unsigned grid = (unsigned)((B + 3) / 4);
  if (buf->dtype == BMI_F64)
    her_gather_kernel<double><<<grid, 128, 0, as_stream(stream)>>>(*buf, ep_idx, t_idx, u_her,
                                                                   u_off, B, future_p, thr, *out);
  else
    her_gather_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(*buf, ep_idx, t_idx, u_her,
                                                                  u_off, B, future_p, thr, *out);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_her_sample_inputs(const bmi_episodes* buf, int64_t n_valid,
                                     const int64_t* ep_idx, const int64_t* t_idx,
                                     const double* u_her, const double* u_off, int64_t B,
                                     double future_p, double thr, double clip_obs,
                                     double clip_range, const float* o_mean, const float* o_std,
                                     const float* g_mean, const float* g_std, float* x, float* xn,
                                     float* actions, float* r, bmi_stream_t stream) {
  int rc;
  if ((rc = check_eps(buf, "bmi_her_sample_inputs"))) return rc;
  BMI_REQUIRE(B >= 0, "bmi_her_sample_inputs: negative batch");
  if (B == 0) return BMI_OK;
  BMI_REQUIRE(n_valid != 0, "bmi_her_sample_inputs: sampling an empty buffer");
  BMI_REQUIRE(ep_idx && t_idx && u_her && u_off && o_mean && o_std && g_mean && g_std && x && xn &&
                  actions && r,
              "bmi_her_sample_inputs: null pointer");
#define BMI_HER_LAUNCH(TT, TILE)                                                                             \
  her_inputs_kernel<TT, TILE><<<(unsigned)((B + (TILE) - 1) / (TILE)), HER_THREADS, 0, as_stream(stream)>>>(        \
      *buf, ep_idx, t_idx, u_her, u_off, B, future_p, thr, clip_obs, clip_range, o_mean, o_std, g_mean, g_std, x, \
      xn, actions, r)
#define BMI_HER_LANE(TT, FC)                                                                                  \
  { if (buf->goal_dim == 3) BMI_HER_LANE_(TT, FC, 3) else BMI_HER_LANE_(TT, FC, 0) }
#define BMI_HER_LANE_(TT, FC, DG)                                                                             \
  { if (B > 16384) { BMI_HER_LANE__(TT, FC, DG, 16, 8); } else { BMI_HER_LANE__(TT, FC, DG, 4, 4); } }
#define BMI_HER_LANE__(TT, FC, DG, CH, S)                                                                     \
  her_inputs_lane_kernel<TT, FC, DG, CH, S><<<lane_grid, 128, 0, as_stream(stream)>>>(                                    \
      *buf, ep_idx, t_idx, u_her, u_off, B, future_p, sq_thr, clip_obs, clip_range, o_mean, o_std, g_mean, g_std, x, \
      xn, actions, r)
  const bool narrow = buf->obs_dim + buf->goal_dim <= 32 &&     // act_dim <= 32 is checked by check_eps
                      (double)buf->n_episodes * (buf->T + 1) < 4294967296.0 && (double)B * 32.0 < 4294967296.0 &&
                      std::isfinite(thr);
  if (narrow) {
    const int64_t her_ch = B > 16384 ? 16 : 4;
    const int64_t n_chunks = (B + her_ch - 1) / her_ch;
    const unsigned lane_grid = (unsigned)std::min<int64_t>((n_chunks + 3) / 4, 148 * 8);
    const double sq_thr = sqrt_threshold(thr);
    const bool fc = (double)(float)clip_obs == clip_obs && (double)(float)clip_range == clip_range;
    if (buf->dtype == BMI_F64) { if (fc) { BMI_HER_LANE(double, true); } else { BMI_HER_LANE(double, false); } }
    else { if (fc) { BMI_HER_LANE(float, true); } else { BMI_HER_LANE(float, false); } }
  } else if (buf->dtype == BMI_F64) {
    if (B <= 8192) BMI_HER_LAUNCH(double, 8); else BMI_HER_LAUNCH(double, 64);
  } else {
    if (B <= 8192) BMI_HER_LAUNCH(float, 8); else BMI_HER_LAUNCH(float, 64);
  }
#undef BMI_HER_LANE
#undef BMI_HER_LANE_
#undef BMI_HER_LANE__
#undef BMI_HER_LAUNCH
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_her_draw(uint64_t seed, uint64_t* counter, int64_t B, const int64_t* n_valid,
                            int32_t T, int64_t* ep_idx, int64_t* t_idx, double* u_her,
                            double* u_off, bmi_stream_t stream) {
  BMI_REQUIRE(B >= 0 && T > 0, "bmi_her_draw: bad sizes");
  if (B == 0) return BMI_OK;
  BMI_REQUIRE(counter && n_valid && ep_idx && t_idx && u_her && u_off, "bmi_her_draw: null pointer");
  her_draw_kernel<<<(unsigned)((B + 127) / 128), 128, 0, as_stream(stream)>>>(
      seed, counter, B, n_valid, T, ep_idx, t_idx, u_her, u_off);
  BMI_LAUNCHED();
  advance_counter_kernel<<<1, 1, 0, as_stream(stream)>>>(counter, (uint64_t)B);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_rollout_record(const bmi_episodes* ep, int32_t t, const float* obs,
                                  const float* ag, const float* g, const float* act,
                                  bmi_stream_t stream) {
  int rc;
  if ((rc = check_eps(ep, "bmi_rollout_record"))) return rc;
  BMI_REQUIRE(t >= 0 && t <= ep->T, "bmi_rollout_record: t=%d outside [0,%d]", t, ep->T);
  BMI_REQUIRE(obs && ag, "bmi_rollout_record: null obs/ag");
  BMI_REQUIRE(t == ep->T || (g && act), "bmi_rollout_record: null g/actions for t<T");
  if (ep->n_episodes == 0) return BMI_OK;
  unsigned grid = (unsigned)((ep->n_episodes + 3) / 4);
  if (ep->dtype == BMI_F64)
    rollout_record_kernel<double><<<grid, 128, 0, as_stream(stream)>>>(*ep, t, obs, ag, g, act);
  else
    rollout_record_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(*ep, t, obs, ag, g, act);
  BMI_LAUNCHED();
  return BMI_OK;
}
