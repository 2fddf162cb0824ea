// Vectorised bmirobot environment: one CUDA thread block (one warp) per env instance.
//
// Reference behaviour restated (paths relative to the reference tree; the arithmetic itself
// lives in PyBullet, so the algorithm follows oracle/bmi_physics_oracle.c line for line):
//   bmirobot_env/bmirobot_env_push_F.py:92-108   step: clip, action[3]=0, IK + motors, 20 sub-steps
//   bmirobot_env/bmirobot_env_push_F.py:110-165  reset (block / goal placement ranges)
//   bmirobot_env/bmirobot_env_push_F.py:169-237  27-float observation
//   bmirobot_env/bmirobot.py:129-191             applyAction / sent_hand_moving
//   bmirobot_env/bmirobot_inverse_kinematics.py:28-33  position-only DLS IK of link 11
//   bmirobot_env/bmirobot_env_pickandplace_v2.py:92-95,116-131  pick task deltas
//
// Per sub-step: forward kinematics -> mass matrix + bias by 10 lane-parallel recursive
// Newton-Euler sweeps (lane j < 9: unit acceleration e_j, lane 9: velocity/gravity/damping bias)
// -> Cholesky -> unconstrained velocities -> contact generation (lane = vertex) -> constraint
// rows (lane = row) -> projected Gauss-Seidel with warp-shuffle dot products (lane = generalized
// velocity) -> semi-implicit Euler.  fp32 throughout, no tensor cores.
// The kinematic tree + solver constants (header + 9 link records of the model blob, 1408 B) are
// staged into shared memory by one TMA bulk copy per block; convex-polytope vertex/plane pools
// are read through the read-only path (they are shared by every block and stay in L1/L2).
#include "common.cuh"
#include "../../include/bmi_model.h"

namespace bmi {

constexpr int NL = 9;          // links / joints of the right arm
constexpr int NU = 15;         // generalized velocities: 9 joints + block linear 3 + angular 3
constexpr int EE = 8;          // right_hand2
constexpr int MAXC = 10;       // contacts per sub-step
constexpr int MAXA = 6;        // ... of which at most 6 involve an arm link
constexpr int MAXNC = 17;      // non-contact rows: 9 motors + up to 8 limit rows
constexpr int STAGED = BMI_MODEL_HDR + BMI_MAX_LINKS * BMI_LINK_STRIDE;  // floats staged by TMA
constexpr int HID = 256;         // hidden width of the actor (models.py:15-17)
constexpr unsigned FULL = 0xffffffffu;
#ifndef BMI_BLOCKS_PER_SM
#define BMI_BLOCKS_PER_SM 4
#endif
#ifndef BMI_ENVS_PER_BLOCK
#define BMI_ENVS_PER_BLOCK 7
#endif
constexpr int BLOCKS_PER_SM = BMI_BLOCKS_PER_SM;  // x ENVW envs: 28 envs per SM -> 4096 envs resident in one wave

// topology of the right arm: chain 0..6, two fingers on link 6 (checked against the blob)
__host__ __device__ constexpr int parent_of(int i) { return i == 0 ? -1 : (i <= 6 ? i - 1 : 6); }
__host__ __device__ constexpr bool is_ancestor_or_self(int a, int l) {
  return a == l || (a <= 6 && l >= a);  // every chain link j<=6 is an ancestor of all l>=j
}

// Env instances per thread block: ENVW env warps (one env each) + ONE solver warp that runs the constraint solver of
// all the block's envs with one LANE per env (solver_loop).  The envs share one staged model copy.  ENVW <= 8 keeps the
// solver's per-lane shared-memory accesses (env stride = 16 B x odd) free of bank conflicts.
constexpr int ENVW = BMI_ENVS_PER_BLOCK;
constexpr int WARPS = ENVW + 1;
static_assert(ENVW >= 1 && ENVW <= 8, "one solver lane per env, conflict-free for <= 8 envs per block");
#ifndef BMI_SMEM_PAD
#define BMI_SMEM_PAD 0
#endif
constexpr int MP12 = 12;         // padded row length of the float4-readable 9-vectors (M^-1 columns, arm Jacobian rows)

struct __align__(16) Smem {      // per-env (per-warp) working set
  const float* model;             // block-shared header params + link records (the TMA destination)
  int req;                        // request sequence number posted by the env warp (-1: the env is finished)
  int nc, na, n_nc, n_bt;         // contacts, contacts on arm links, non-contact rows, block-on-table contacts
  float R[NL][9], p[NL][3], z[NL][3], c[NL][3], Rl[NL][9];
  union {
    float L[NL * NL];             // mass matrix / its Cholesky factor (dead once M^-1 is known)
    float4 MinvR4[NL * MP12 / 4]; // motor-row update vectors, rotated: row r holds M^-1[(r+k) % 9][r] at k = 0..8
  };
  float4 MinvP4[NL * MP12 / 4];   // M^-1, column r at floats [r*12, r*12+9): one solver row update = 3 float4 loads
  float q[NL], qd[NL], qt[NL], bias[NL], acc[NL];
  float u[16];
  float dvout[16];                // solver result: velocity deltas of the 15 generalized velocities
  float tauw[NL + 1][NL];             // per-lane RNEA output rows
  float bp[3], bq[4], bv[3], bw[3], goal[3];
  float Rb[9], Ibinv[9], bvert[8][3];
  // contacts
  float cx[MAXC][3], cn[MAXC][3], cdist[MAXC], cmu[MAXC];
  int clink[MAXC], chasb[MAXC];
  // rows
  union {
    struct {
      union {
        float4 rd[3 * MAXC][4];            // per contact row: Jb[6] Wb[6] | invd rhs diag mu
        struct { float A[NL * NL], b[NL]; } ik;   // IK scratch (the IK runs before the sub-steps)
      };
      float4 Ja4[3 * MAXA * MP12 / 4], Wa4[3 * MAXA * MP12 / 4];  // arm parts (only contacts that touch an arm link)
    };
    struct { float x[32], hA[HID], hB[HID]; } pol;  // policy activations (fused rollout; between env steps)
  };
  float obs[BMI_OBS_DIM + BMI_GOAL_DIM];   // last observation + achieved goal (fused rollout)
  int carm[MAXC];                      // arm slot of a contact or -1
  float lam[3 * MAXC];
  float invd[MAXNC], rhs[MAXNC], lo[MAXNC], hi[MAXNC], lamn[MAXNC];  // non-contact rows
  int ncj[MAXNC];                 // joint index (+1, sign = direction) of each non-contact row
  float mdiag[NL];                // diagonal of M^-1
  float qik[NL];
#if BMI_SMEM_PAD > 0
  float pad_[BMI_SMEM_PAD];
#endif
};
// the solver warp reads env `lane`'s rows: an env stride of 16 B x odd keeps 8 lanes on distinct banks for both
// 32-bit and 128-bit shared loads
static_assert(sizeof(Smem) % 16 == 0 && (sizeof(Smem) / 16) % 2 == 1, "adjust BMI_SMEM_PAD: Smem stride must be 16 B x odd");

struct EnvParams {
  int task;
  float bh[3], bmass, binertia[3], bmu;
};

__device__ __forceinline__ float P(const Smem& s, int i) { return s.model[i]; }
__device__ __forceinline__ const float* LK(const Smem& s, int i) { return s.model + BMI_MODEL_HDR + i * BMI_LINK_STRIDE; }

__device__ __forceinline__ float& MINV(Smem& s, int i, int j) { return reinterpret_cast<float*>(s.MinvP4)[j * MP12 + i]; }
__device__ __forceinline__ float* JA(Smem& s, int row) { return reinterpret_cast<float*>(s.Ja4) + row * MP12; }
__device__ __forceinline__ float* WA(Smem& s, int row) { return reinterpret_cast<float*>(s.Wa4) + row * MP12; }

__device__ __forceinline__ void cross3(float* o, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void mat_vec(float* o, const float* A, const float* v) {
  float x = A[0] * v[0] + A[1] * v[1] + A[2] * v[2], y = A[3] * v[0] + A[4] * v[1] + A[5] * v[2],
        z = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void matT_vec(float* o, const float* A, const float* v) {
  float x = A[0] * v[0] + A[3] * v[1] + A[6] * v[2], y = A[1] * v[0] + A[4] * v[1] + A[7] * v[2],
        z = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ float warp_sum16(float v) {  // sum over lanes 0..15 (butterfly), result in all 16
  v += __shfl_xor_sync(FULL, v, 8);
  v += __shfl_xor_sync(FULL, v, 4);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 1);
  return v;
}

// sin and cos for |x| up to a few turns (joint angles, half-angles): two-constant Cody-Waite reduction to [-pi/4, pi/4]
// + the cephes single-precision minimax polynomials (<= 2 ulp).  Compact on purpose: sincosf() drags a 2 KB slow path
// into every caller and this kernel is instruction-fetch bound.
__device__ __forceinline__ void sincos_compact(float x, float* sn, float* cs) {
  const float kf = rintf(x * 0.63661977236758134f);
  const int k = (int)kf;
  float r = fmaf(-kf, 1.5707962513e+00f, x);
  r = fmaf(-kf, 7.5497894159e-08f, r);
  const float r2 = r * r;
  const float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  const float pc = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2,
                        fmaf(-0.5f, r2, 1.0f));
  const float a = (k & 1) ? pc : ps, b = (k & 1) ? ps : pc;
  *sn = (k & 2) ? -a : a;
  *cs = ((k + 1) & 2) ? -b : b;
}

// ---- TMA staging of the joint tree ---------------------------------------------------------
__device__ __forceinline__ void stage_model(float* model_s, unsigned long long* mbar_s,
                                            const float* __restrict__ model_g, int tid) {
  const int lane = tid;  // thread 0 of the block issues the copy; every thread of every warp waits on the barrier
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_s);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(model_s);
  constexpr unsigned bytes = STAGED * sizeof(float);
  static_assert(bytes % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(model_g), "r"(bytes), "r"(mbar)
                 : "memory");
  }
  __syncthreads();  // barrier initialised and armed before anyone polls it
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mbar)
        : "memory");
  }
}

// ---- forward kinematics ----------------------------------------------------------------------
__device__ __noinline__ void fk(Smem& s, const float* q, int lane) {
  if (lane < NL) {  // local rotation jrot * Rodrigues(axis, q)
    const float* lk = LK(s, lane);
    const float ux = lk[ML_AXIS], uy = lk[ML_AXIS + 1], uz = lk[ML_AXIS + 2];
    float sn, cs;
    sincos_compact(q[lane], &sn, &cs);
    const float C = 1.f - cs;
    float Rq[9] = {cs + ux * ux * C,      ux * uy * C - uz * sn, ux * uz * C + uy * sn,
                   uy * ux * C + uz * sn, cs + uy * uy * C,      uy * uz * C - ux * sn,
                   uz * ux * C - uy * sn, uz * uy * C + ux * sn, cs + uz * uz * C};
    const float* Jr = lk + ML_JROT;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int cc = 0; cc < 3; ++cc)
        s.Rl[lane][3 * r + cc] = Jr[3 * r] * Rq[cc] + Jr[3 * r + 1] * Rq[3 + cc] + Jr[3 * r + 2] * Rq[6 + cc];
  }
  __syncwarp();
#pragma unroll 1
  for (int i = 0; i < NL; ++i) {
    const int pa = parent_of(i);
    if (lane < 9) {
      const int r = lane / 3, cc = lane % 3;
      float v;
      if (pa < 0) v = s.Rl[i][lane];
      else v = s.R[pa][3 * r] * s.Rl[i][cc] + s.R[pa][3 * r + 1] * s.Rl[i][3 + cc] + s.R[pa][3 * r + 2] * s.Rl[i][6 + cc];
      s.R[i][lane] = v;
    } else if (lane < 12) {
      const int a = lane - 9;
      const float* jp = LK(s, i) + ML_JPOS;
      float v;
      if (pa < 0) v = P(s, MP_BASE_PX + a) + jp[a];
      else v = s.p[pa][a] + s.R[pa][3 * a] * jp[0] + s.R[pa][3 * a + 1] * jp[1] + s.R[pa][3 * a + 2] * jp[2];
      s.p[i][a] = v;
    }
    __syncwarp();
  }
  if (lane < NL) {
    const float* lk = LK(s, lane);
    float t[3];
    mat_vec(t, s.R[lane], lk + ML_AXIS);
    s.z[lane][0] = t[0]; s.z[lane][1] = t[1]; s.z[lane][2] = t[2];
    mat_vec(t, s.R[lane], lk + ML_COM);
    s.c[lane][0] = s.p[lane][0] + t[0]; s.c[lane][1] = s.p[lane][1] + t[1]; s.c[lane][2] = s.p[lane][2] + t[2];
  }
  __syncwarp();
}

// ---- recursive Newton-Euler, one independent sweep per lane ------------------------------------
// tau_j = z_j . sum_{k in subtree(j)} [ N_k + (c_k - p_j) x F_k ]   (accumulated pairwise so that no
// per-link force arrays are needed; the tree topology is a compile-time constant).
// Kept as a compact runtime loop (not unrolled): the kernel is instruction-fetch sensitive (profiles/r01_*).
__device__ __noinline__ void rnea_lane(const Smem& s, bool is_bias, int unit, float gz, float kl, float ka,
                                       float* tau) {  // tau: 9 floats in shared memory, private to this lane
  constexpr bool kBias = true;  // the velocity terms are evaluated by every sweep (zeros for the unit sweeps)
  float w[3] = {0, 0, 0}, al[3] = {0, 0, 0}, a[3] = {0, 0, -gz}, vo[3] = {0, 0, 0};
  float w6[3] = {0, 0, 0}, al6[3] = {0, 0, 0}, a6[3] = {0, 0, 0}, vo6[3] = {0, 0, 0};
  for (int j = 0; j < NL; ++j) tau[j] = 0.f;
#pragma unroll 1
  for (int i = 0; i < NL; ++i) {
    const int pa = parent_of(i);
    if (i == 7 || i == 8) {  // fingers hang off link 6
#pragma unroll
      for (int k = 0; k < 3; ++k) { w[k] = w6[k]; al[k] = al6[k]; a[k] = a6[k]; vo[k] = vo6[k]; }
    }
    float t[3], t2[3];
    if (pa >= 0) {
      float r[3] = {s.p[i][0] - s.p[pa][0], s.p[i][1] - s.p[pa][1], s.p[i][2] - s.p[pa][2]};
      if (kBias) { cross3(t, w, r); vo[0] += t[0]; vo[1] += t[1]; vo[2] += t[2]; }
      cross3(t, al, r);
      a[0] += t[0]; a[1] += t[1]; a[2] += t[2];
      if (kBias) { cross3(t, w, r); cross3(t2, w, t); a[0] += t2[0]; a[1] += t2[1]; a[2] += t2[2]; }
    }
    const float qdi = is_bias ? s.qd[i] : 0.f;
    const float qddi = (!is_bias && unit == i) ? 1.f : 0.f;
    const float* zi = s.z[i];
    if (kBias) { cross3(t, w, zi); al[0] += qdi * t[0]; al[1] += qdi * t[1]; al[2] += qdi * t[2]; }
    al[0] += qddi * zi[0]; al[1] += qddi * zi[1]; al[2] += qddi * zi[2];
    if (kBias) { w[0] += qdi * zi[0]; w[1] += qdi * zi[1]; w[2] += qdi * zi[2]; }
    if (i == 6) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { w6[k] = w[k]; al6[k] = al[k]; a6[k] = a[k]; vo6[k] = vo[k]; }
    }
    const float* lk = LK(s, i);
    const float mass = lk[ML_MASS];
    float rc[3] = {s.c[i][0] - s.p[i][0], s.c[i][1] - s.p[i][1], s.c[i][2] - s.p[i][2]};
    float ac[3];
    cross3(t, al, rc);
    ac[0] = a[0] + t[0]; ac[1] = a[1] + t[1]; ac[2] = a[2] + t[2];
    float F[3], N[3];
    if (kBias) {
      cross3(t, w, rc); cross3(t2, w, t);
      ac[0] += t2[0]; ac[1] += t2[1]; ac[2] += t2[2];
      float vc[3] = {vo[0] + t[0], vo[1] + t[1], vo[2] + t[2]};
      const float vn = sqrtf(dot3(vc, vc));
      const float kd = mass * (kl + kl * vn);
      F[0] = mass * ac[0] + kd * vc[0]; F[1] = mass * ac[1] + kd * vc[1]; F[2] = mass * ac[2] + kd * vc[2];
    } else {
      F[0] = mass * ac[0]; F[1] = mass * ac[1]; F[2] = mass * ac[2];
    }
    {
      float ll[3], tl[3];
      matT_vec(ll, s.R[i], al);
      tl[0] = lk[ML_INERTIA] * ll[0]; tl[1] = lk[ML_INERTIA + 1] * ll[1]; tl[2] = lk[ML_INERTIA + 2] * ll[2];
      mat_vec(N, s.R[i], tl);
      if (kBias) {
        float wl[3], Iw[3];
        matT_vec(wl, s.R[i], w);
        tl[0] = lk[ML_INERTIA] * wl[0]; tl[1] = lk[ML_INERTIA + 1] * wl[1]; tl[2] = lk[ML_INERTIA + 2] * wl[2];
        mat_vec(Iw, s.R[i], tl);
        cross3(t, w, Iw);
        const float wn = sqrtf(dot3(w, w));
        const float kk = ka + ka * wn;
        N[0] += t[0] + kk * Iw[0]; N[1] += t[1] + kk * Iw[1]; N[2] += t[2] + kk * Iw[2];
      }
    }
#pragma unroll 1
    for (int j = i; j >= 0; j = parent_of(j)) {  // every joint on the path base -> link i feels link i's wrench
      float r[3] = {s.c[i][0] - s.p[j][0], s.c[i][1] - s.p[j][1], s.c[i][2] - s.p[j][2]};
      cross3(t, r, F);
      tau[j] += s.z[j][0] * (N[0] + t[0]) + s.z[j][1] * (N[1] + t[1]) + s.z[j][2] * (N[2] + t[2]);
    }
  }
}

// Cholesky of the 9x9 SPD matrix in s.L (lower, in place, reciprocal diagonal); lanes cooperate per column.
__device__ __noinline__ void chol9(float* Lm, int lane) {
#pragma unroll 1
  for (int j = 0; j < NL; ++j) {
    float d = 0.f;
    if (lane == 0) {
      d = Lm[j * NL + j];
      for (int k = 0; k < j; ++k) d -= Lm[j * NL + k] * Lm[j * NL + k];
      d = rsqrtf(fmaxf(d, 1e-20f));   // the diagonal stores 1 / L_jj: the solves multiply instead of divide
      Lm[j * NL + j] = d;
    }
    d = __shfl_sync(FULL, d, 0);
    if (lane > j && lane < NL) {
      float sacc = Lm[lane * NL + j];
      for (int k = 0; k < j; ++k) sacc -= Lm[lane * NL + k] * Lm[j * NL + k];
      Lm[lane * NL + j] = sacc * d;
    }
    __syncwarp();
  }
}
// per-lane triangular solves  L L^T x = b   (b, x: 9 registers)
__device__ __forceinline__ void chol9_solve(const float* Lm, const float* b, float* x) {
  float y[NL];
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    float sacc = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) sacc -= Lm[i * NL + k] * y[k];
    y[i] = sacc * Lm[i * NL + i];
  }
#pragma unroll
  for (int i = NL - 1; i >= 0; --i) {
    float sacc = y[i];
#pragma unroll
    for (int k = i + 1; k < NL; ++k) sacc -= Lm[k * NL + i] * x[k];
    x[i] = sacc * Lm[i * NL + i];
  }
}

// ---- inverse kinematics (BussIK DLS restated, see oracle solve_ik) ------------------------------
__device__ __noinline__ void solve_ik(Smem& s, const float* target, int lane) {
  if (lane < NL) s.qik[lane] = s.q[lane];
  __syncwarp();
  const int iters = (int)P(s, MP_IK_ITERS);
  const float damp = P(s, MP_IK_DAMPING), thr = P(s, MP_IK_THRESH), maxang = P(s, MP_IK_MAX_ANGLE);
  for (int it = 0; it < iters; ++it) {
    fk(s, s.qik, lane);
    float e[3] = {target[0] - s.p[EE][0], target[1] - s.p[EE][1], target[2] - s.p[EE][2]};
    if (sqrtf(dot3(e, e)) <= thr) break;  // uniform across the warp
    // Jacobian column of joint `lane` (joint 7 is not on the path to the EE)
    float Jc[3] = {0, 0, 0};
    if (lane < NL && lane != 7) {
      float r[3] = {s.p[EE][0] - s.p[lane][0], s.p[EE][1] - s.p[lane][1], s.p[EE][2] - s.p[lane][2]};
      cross3(Jc, s.z[lane], r);
    }
    // A = J^T J + damp I, b = J^T e ; lane i owns row i
    float Arow[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      float jx = __shfl_sync(FULL, Jc[0], j), jy = __shfl_sync(FULL, Jc[1], j), jz = __shfl_sync(FULL, Jc[2], j);
      Arow[j] = Jc[0] * jx + Jc[1] * jy + Jc[2] * jz + ((j == lane) ? damp : 0.f);
    }
    if (lane < NL) {
#pragma unroll
      for (int j = 0; j < NL; ++j) s.ik.A[lane * NL + j] = Arow[j];
      s.ik.b[lane] = Jc[0] * e[0] + Jc[1] * e[1] + Jc[2] * e[2];
    }
    __syncwarp();
    chol9(s.ik.A, lane);
    float bb[NL], x[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) bb[j] = s.ik.b[j];
    chol9_solve(s.ik.A, bb, x);  // every lane solves the same system (cheap, avoids a broadcast)
    float mx = 0.f;
#pragma unroll
    for (int j = 0; j < NL; ++j) mx = fmaxf(mx, fabsf(x[j]));
    const float sc = mx > maxang ? maxang / mx : 1.f;
    __syncwarp();
    if (lane < NL) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < NL; ++j) if (j == lane) v = x[j];
      s.qik[lane] += sc * v;
    }
    __syncwarp();
  }
}

// ---- contact generation ---------------------------------------------------------------------------
// keep the `cap` lanes with the smallest d (< margin); ties resolved towards the lower lane; returns the
// ballot mask of the selected lanes
__device__ __noinline__ unsigned select_deepest(float d, bool valid, float margin, int cap, int lane) {
  // order-preserving map float -> uint, then one REDUX.MIN + one ballot per round (compact: this used to be 25 KB of
  // unrolled shuffle butterflies, and the kernel is instruction-fetch bound)
  const unsigned b = __float_as_uint(d);
  unsigned key = (valid && d < margin) ? (b ^ ((b & 0x80000000u) ? 0xffffffffu : 0x80000000u)) : 0xffffffffu;
  unsigned picked = 0;
#pragma unroll 1
  for (int r = 0; r < cap; ++r) {
    const unsigned best = __reduce_min_sync(FULL, key);
    if (best == 0xffffffffu) break;
    const int idx = __ffs(__ballot_sync(FULL, key == best)) - 1;  // ties: lower lane
    picked |= 1u << idx;
    if (lane == idx) key = 0xffffffffu;
  }
  return picked;
}

__device__ __noinline__ void push_contacts(Smem& s, unsigned mask, int lane, int link, int hasb, const float* x,
                                              const float* n, float dist, float mu) {
  if (mask == 0) return;
  const int base = s.nc, abase = s.na;
  // capacity: MAXC contacts in total, MAXA of them on arm links; later candidates (lane order) are dropped
  int allowed = MAXC - base;
  if (link >= 0) allowed = min(allowed, MAXA - abase);
  const int rank = __popc(mask & ((1u << lane) - 1));
  const int n_add = min(__popc(mask), max(allowed, 0));
  if (((mask >> lane) & 1u) && rank < n_add) {
    const int slot = base + rank;
    s.cx[slot][0] = x[0]; s.cx[slot][1] = x[1]; s.cx[slot][2] = x[2];
    s.cn[slot][0] = n[0]; s.cn[slot][1] = n[1]; s.cn[slot][2] = n[2];
    s.cdist[slot] = dist; s.cmu[slot] = mu; s.clink[slot] = link; s.chasb[slot] = hasb;
    s.carm[slot] = link >= 0 ? abase + rank : -1;
  }
  __syncwarp();
  if (lane == 0) {
    s.nc = base + n_add;
    if (link >= 0) s.na = abase + n_add;
  }
  __syncwarp();
}

__device__ __noinline__ void find_contacts(Smem& s, const EnvParams& ep, const float* __restrict__ model_g,
                                           float block_margin, int lane) {
  if (lane == 0) { s.nc = 0; s.na = 0; }
  // block frame
  if (lane == 0) {
    const float x = s.bq[0], y = s.bq[1], z = s.bq[2], w = s.bq[3];
    s.Rb[0] = 1 - 2 * (y * y + z * z); s.Rb[1] = 2 * (x * y - z * w);     s.Rb[2] = 2 * (x * z + y * w);
    s.Rb[3] = 2 * (x * y + z * w);     s.Rb[4] = 1 - 2 * (x * x + z * z); s.Rb[5] = 2 * (y * z - x * w);
    s.Rb[6] = 2 * (x * z - y * w);     s.Rb[7] = 2 * (y * z + x * w);     s.Rb[8] = 1 - 2 * (x * x + y * y);
  }
  __syncwarp();
  const float tz = P(s, MP_TABLE_Z);
  float bvx[3] = {0, 0, 0};
  if (lane < 8) {
    float l[3] = {(lane & 1 ? 1.f : -1.f) * ep.bh[0], (lane & 2 ? 1.f : -1.f) * ep.bh[1], (lane & 4 ? 1.f : -1.f) * ep.bh[2]};
    mat_vec(bvx, s.Rb, l);
    bvx[0] += s.bp[0]; bvx[1] += s.bp[1]; bvx[2] += s.bp[2];
    s.bvert[lane][0] = bvx[0]; s.bvert[lane][1] = bvx[1]; s.bvert[lane][2] = bvx[2];
  }
  __syncwarp();
  const float up[3] = {0.f, 0.f, 1.f};
  {  // block vertices vs table plane, up to 4 deepest
    const float d = bvx[2] - tz;
    unsigned m = select_deepest(d, lane < 8, P(s, MP_TABLE_MARGIN), 4, lane);
    push_contacts(s, m, lane, -1, 1, bvx, up, d, ep.bmu * P(s, MP_MU_TABLE));
    if (lane == 0) s.n_bt = s.nc;  // contacts [0, n_bt) touch no arm link; every later one does
  }
  const int ns = (int)P(s, MP_N_SHAPES);
  const float* shapes = model_g + (int)P(s, MP_SHAPES_OFF);
  const float* pool = model_g + (int)P(s, MP_POOL_OFF);
  const float brad = sqrtf(ep.bh[0] * ep.bh[0] + ep.bh[1] * ep.bh[1] + ep.bh[2] * ep.bh[2]);
  for (int si = 0; si < ns; ++si) {
    const float* sh = shapes + si * BMI_SHAPE_STRIDE;
    const int l = (int)__ldg(sh + MS_LINK), nv = (int)__ldg(sh + MS_NVERTS), np = (int)__ldg(sh + MS_NPLANES);
    const float* verts = pool + (int)__ldg(sh + MS_VERT_OFF);
    const float* planes = pool + (int)__ldg(sh + MS_PLANE_OFF);
    const float smu = __ldg(sh + MS_MU);
    float wv[3] = {0, 0, 0};
    if (lane < nv) {
      float lv[3] = {__ldg(verts + 3 * lane), __ldg(verts + 3 * lane + 1), __ldg(verts + 3 * lane + 2)};
      mat_vec(wv, s.R[l], lv);
      wv[0] += s.p[l][0]; wv[1] += s.p[l][1]; wv[2] += s.p[l][2];
    }
    {  // hull vertices vs table plane, up to 2 deepest
      const float d = wv[2] - tz;
      unsigned m = select_deepest(d, lane < nv, P(s, MP_CONTACT_MARGIN), 2, lane);
      push_contacts(s, m, lane, l, 0, wv, up, d, smu * P(s, MP_MU_TABLE));
    }
    // broadphase: bounding spheres
    float lc[3] = {__ldg(sh + MS_SPHERE_C), __ldg(sh + MS_SPHERE_C + 1), __ldg(sh + MS_SPHERE_C + 2)}, cw[3];
    mat_vec(cw, s.R[l], lc);
    float dd[3] = {s.bp[0] - cw[0] - s.p[l][0], s.bp[1] - cw[1] - s.p[l][1], s.bp[2] - cw[2] - s.p[l][2]};
    if (sqrtf(dot3(dd, dd)) > __ldg(sh + MS_SPHERE_R) + brad + block_margin) continue;  // uniform
    // candidates: lanes 0..7 = block vertex vs hull planes, lanes 8..8+nv-1 = hull vertex vs block box
    float d = 3.0e38f, nrm[3] = {0, 0, 0}, x[3] = {0, 0, 0};
    bool valid = false;
    if (lane < 8) {
      float r[3] = {bvx[0] - s.p[l][0], bvx[1] - s.p[l][1], bvx[2] - s.p[l][2]}, xl[3];
      matT_vec(xl, s.R[l], r);
      float best = -1e30f;
      int bpi = 0;
#pragma unroll 2
      for (int pi = 0; pi < np; ++pi) {
        const float4 pl = __ldg(reinterpret_cast<const float4*>(planes) + pi);
        const float sd = pl.x * xl[0] + pl.y * xl[1] + pl.z * xl[2] + pl.w;
        if (sd > best) { best = sd; bpi = pi; }
      }
      const float4 pl = __ldg(reinterpret_cast<const float4*>(planes) + bpi);
      float ln[3] = {pl.x, pl.y, pl.z};
      mat_vec(nrm, s.R[l], ln);
      d = best; valid = true;
      x[0] = bvx[0]; x[1] = bvx[1]; x[2] = bvx[2];
    }
    // hull vertices are owned by lanes 0..nv-1 but candidate slots are 8..8+nv-1: shift by 8 lanes
    {
      const int src = lane - 8;
      float hx = __shfl_sync(FULL, wv[0], src & 31), hy = __shfl_sync(FULL, wv[1], src & 31), hz = __shfl_sync(FULL, wv[2], src & 31);
      if (lane >= 8 && src < nv) {
        float r[3] = {hx - s.bp[0], hy - s.bp[1], hz - s.bp[2]}, xb[3];
        matT_vec(xb, s.Rb, r);
        float best = -1e30f, sg = 1.f;
        int ba = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float sd = fabsf(xb[a]) - ep.bh[a];
          if (sd > best) { best = sd; ba = a; sg = xb[a] >= 0.f ? 1.f : -1.f; }
        }
        nrm[0] = -sg * s.Rb[ba]; nrm[1] = -sg * s.Rb[3 + ba]; nrm[2] = -sg * s.Rb[6 + ba];
        d = best; valid = true;
        x[0] = hx; x[1] = hy; x[2] = hz;
      }
    }
    // (hull polytopes are baked with <= 24 vertices, so 8 + nv <= 32 candidates fit one warp)
    unsigned m = select_deepest(d, valid, block_margin, 3, lane);
    push_contacts(s, m, lane, l, 1, x, nrm, d, ep.bmu * smu);
  }
  __syncwarp();
}

__device__ __forceinline__ void plane_space(const float* n, float* p, float* q) {
  if (fabsf(n[2]) > 0.70710678f) {
    const float a = n[1] * n[1] + n[2] * n[2], k = rsqrtf(a);
    p[0] = 0.f; p[1] = -n[2] * k; p[2] = n[1] * k;
    q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
  } else {
    const float a = n[0] * n[0] + n[1] * n[1], k = rsqrtf(a);
    p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0.f;
    q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
  }
}

// ---- one simulation sub-step ------------------------------------------------------------------------
// substep_pre (one warp per env): dynamics terms, contacts and constraint rows; pgs_thread (one LANE per env) solves;
// substep_post (one warp per env) integrates.
__device__ __noinline__ void substep_pre(Smem& s, const EnvParams& ep, const float* __restrict__ model_g, int lane) {
  const float dt = P(s, MP_DT), gz = P(s, MP_GRAVITY), kl = P(s, MP_LIN_DAMP), ka = P(s, MP_ANG_DAMP);
  fk(s, s.q, lane);
  {  // mass matrix columns (lanes 0..8) and bias (lane 9)
    if (lane <= NL) {  // one code path for all ten sweeps (no divergence): unit lanes see zero velocity / gravity
      const bool is_bias = lane == NL;
      float* tau = s.tauw[lane];
      rnea_lane(s, is_bias, lane, is_bias ? gz : 0.f, is_bias ? kl : 0.f, is_bias ? ka : 0.f, tau);
      for (int i = 0; i < NL; ++i) {
        if (is_bias) s.bias[i] = tau[i];
        else s.L[i * NL + lane] = tau[i];
      }
    }
  }
  __syncwarp();
  if (lane < NL) {  // symmetrise (lower triangle is what Cholesky reads)
    float v[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) v[j] = 0.5f * (s.L[lane * NL + j] + s.L[j * NL + lane]);
    __syncwarp(0x1ff);
#pragma unroll
    for (int j = 0; j < NL; ++j) s.L[lane * NL + j] = v[j];
  }
  __syncwarp();
  chol9(s.L, lane);
  // M^-1 columns (lanes 0..8) and unconstrained acceleration (lane 9)
  if (lane <= NL) {
    float b[NL], x[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) b[i] = lane < NL ? (i == lane ? 1.f : 0.f) : (-LK(s, i)[ML_DAMPING] * s.qd[i] - s.bias[i]);
    chol9_solve(s.L, b, x);
    if (lane < NL) {
#pragma unroll
      for (int i = 0; i < NL; ++i) MINV(s, i, lane) = x[i];
    } else {
#pragma unroll
      for (int i = 0; i < NL; ++i) s.acc[i] = x[i];
    }
  }
  __syncwarp();
  // rotated copy for the solver's rolled motor loop (its 9 arm deltas rotate through registers, see solver_loop)
  for (int idx = lane; idx < NL * NL; idx += 32) {
    const int r = idx / NL, k = idx - r * NL;
    int i = r + k;
    if (i >= NL) i -= NL;
    reinterpret_cast<float*>(s.MinvR4)[r * MP12 + k] = MINV(s, i, r);
  }
  // predicted (unconstrained) velocities
  if (lane < NL) s.u[lane] = s.qd[lane] + dt * s.acc[lane];
  else if (lane < 12) {
    const int a = lane - 9;
    const float vn = sqrtf(dot3(s.bv, s.bv));
    s.u[lane] = s.bv[a] + dt * (-(kl + kl * vn) * s.bv[a]) + (a == 2 ? dt * gz : 0.f);
  } else if (lane < 15) {
    const int a = lane - 12;
    const float wn = sqrtf(dot3(s.bw, s.bw));
    s.u[lane] = s.bw[a] + dt * (-(ka + ka * wn) * s.bw[a]);
  } else if (lane == 15) s.u[15] = 0.f;
  find_contacts(s, ep, model_g, P(s, MP_BLOCK_MARGIN), lane);
  if (lane < 9) {  // world-frame inverse inertia of the block: R diag(1/I) R^T
    const int r = lane / 3, cc = lane % 3;
    s.Ibinv[lane] = s.Rb[3 * r] * s.Rb[3 * cc] / ep.binertia[0] + s.Rb[3 * r + 1] * s.Rb[3 * cc + 1] / ep.binertia[1] +
                    s.Rb[3 * r + 2] * s.Rb[3 * cc + 2] / ep.binertia[2];
  }
  __syncwarp();
  // ---- non-contact rows: motors (always) then violated joint limits --------------------------------
  const float max_imp = P(s, MP_MOTOR_FORCE) * dt;
  int n_nc = NL;
  if (lane < NL) {
    const float w = MINV(s, lane, lane);
    const float target = P(s, MP_MOTOR_KP) * (s.qt[lane] - s.q[lane]) / dt + (1.f - P(s, MP_MOTOR_KD)) * s.qd[lane];
    s.invd[lane] = 1.f / w;
    s.mdiag[lane] = w;
    s.rhs[lane] = (target - s.u[lane]) / w;
    s.lo[lane] = -max_imp; s.hi[lane] = max_imp; s.lamn[lane] = 0.f;
    s.ncj[lane] = lane + 1;
  }
  {
    // limit candidates: lane = 2*j + side
    bool viol = false;
    float pen = 0.f;
    const int j = lane >> 1, side = lane & 1;
    if (lane < 2 * NL) {
      pen = side == 0 ? s.q[j] - LK(s, j)[ML_LO] : LK(s, j)[ML_HI] - s.q[j];
      viol = !(pen > 0.f);
    }
    unsigned m = __ballot_sync(FULL, viol);
    const int slot = NL + __popc(m & ((1u << lane) - 1));
    if (viol && slot < MAXNC) {
      const float sgn = side == 0 ? 1.f : -1.f;
      const float w = MINV(s, j, j);
      s.invd[slot] = 1.f / w;
      s.rhs[slot] = (-pen * P(s, MP_ERP_JOINT) / dt - sgn * s.u[j]) / w;
      s.lo[slot] = 0.f; s.hi[slot] = P(s, MP_JOINT_LIMIT_IMPULSE); s.lamn[slot] = 0.f;
      s.ncj[slot] = side == 0 ? (j + 1) : -(j + 1);
    }
    n_nc = min(MAXNC, NL + __popc(m));
  }
  __syncwarp();
  // ---- contact rows: lane = row (3 rows per contact: normal, tangent 1, tangent 2) ------------------
  // Row storage: the block part of every row (J and M^-1 J^T over the block's 6 velocities) plus its scalars is
  // one 64-byte record read with broadcast LDS.128; the arm part (9 + 9 floats) exists only for contacts that
  // touch an arm link.
  const int nc = s.nc;
  const int n_rows_c = 3 * nc;
  for (int base = 0; base < n_rows_c; base += 32) {
    const int ri = base + lane;
    if (ri < n_rows_c) {
      const int ci = ri < nc ? ri : (ri - nc) / 2;
      const int kind = ri < nc ? 0 : 1 + ((ri - nc) & 1);
      float n[3] = {s.cn[ci][0], s.cn[ci][1], s.cn[ci][2]}, dir[3], t1[3], t2[3];
      plane_space(n, t1, t2);
#pragma unroll
      for (int a = 0; a < 3; ++a) dir[a] = kind == 0 ? n[a] : (kind == 1 ? t1[a] : t2[a]);
      const float x[3] = {s.cx[ci][0], s.cx[ci][1], s.cx[ci][2]};
      const int link = s.clink[ci], hasb = s.chasb[ci];
      float J[NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) J[a] = 0.f;
      if (hasb) {
        float r[3] = {x[0] - s.bp[0], x[1] - s.bp[1], x[2] - s.bp[2]}, t[3];
        cross3(t, r, dir);
#pragma unroll
        for (int a = 0; a < 3; ++a) { J[9 + a] = dir[a]; J[12 + a] = t[a]; }
      }
      float Wv[NU];
      float diag = 0.f, rel = 0.f;
      if (link >= 0) {  // arm part, compact runtime loops through this row's shared-memory slot
        const float sgn = hasb ? -1.f : 1.f;
        const int as = s.carm[ci] * 3 + kind;
        float* Jr = JA(s, as);
        float* Wr = WA(s, as);
        for (int j = 0; j < NL; ++j) Jr[j] = 0.f;
#pragma unroll 1
        for (int j = link; j >= 0; j = parent_of(j)) {  // joints on the path base -> link
          float r[3] = {x[0] - s.p[j][0], x[1] - s.p[j][1], x[2] - s.p[j][2]}, cr[3];
          cross3(cr, s.z[j], r);
          Jr[j] = sgn * dot3(dir, cr);
        }
#pragma unroll 1
        for (int i = 0; i < NL; ++i) {
          float acc = 0.f;
          for (int j = 0; j < NL; ++j) acc += MINV(s, i, j) * Jr[j];
          Wr[i] = acc;
          diag += Jr[i] * acc;
          rel += Jr[i] * s.u[i];
        }
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) Wv[9 + a] = J[9 + a] / ep.bmass;
      mat_vec(Wv + 12, s.Ibinv, J + 12);
#pragma unroll
      for (int a = 9; a < NU; ++a) { diag += J[a] * Wv[a]; rel += J[a] * s.u[a]; }
      const float invd = 1.f / diag;
      float rhs;
      if (kind == 0) {
        const float pen = s.cdist[ci] + P(s, MP_LINEAR_SLOP);
        float pos_err = 0.f, vel_err = -rel;
        if (pen > 0.f) vel_err -= pen / dt; else pos_err = -pen * P(s, MP_ERP_CONTACT) / dt;
        rhs = (pos_err + vel_err) * invd;
      } else {
        rhs = -rel * invd;
      }
      s.rd[ri][0] = make_float4(J[9], J[10], J[11], J[12]);
      s.rd[ri][1] = make_float4(J[13], J[14], Wv[9], Wv[10]);
      s.rd[ri][2] = make_float4(Wv[11], Wv[12], Wv[13], Wv[14]);
      s.rd[ri][3] = make_float4(invd, rhs, diag, s.cmu[ci]);
      s.lam[ri] = 0.f;
    }
  }
  __syncwarp();
  if (lane == 0) s.n_nc = n_nc;
  __syncwarp();
}

// ---- projected Gauss-Seidel, ONE THREAD per env, served by the block's solver warp ------------------------------
// The solve is a strictly sequential chain (row r needs row r-1's update), so a warp per env leaves 31 lanes idle for
// ~90 % of the sub-step's instructions.  Here lane l of the solver warp owns env slot l of the block: all 15 velocity
// deltas live in its registers, row records come from the env's shared-memory slot (lane-strided, conflict-free), no
// shuffles.  Row order, clamps and the residual exit are the oracle's (pgs_solve in oracle/bmi_physics_oracle.c).
//
// The solver warp is a SERVER: every trip of its loop advances each busy lane by ONE Gauss-Seidel iteration, lanes are
// at different iteration numbers of different requests.  An env that converges after 30 iterations gets its answer
// then, integrates and prepares its next sub-step on its own warp while a neighbour with arm contacts is still
// grinding through its 150 — no barrier couples the envs (iteration counts: median 32, 12 % of the sub-steps hit 150).
// Protocol per env slot: the env warp writes its rows, then `req = seq` (fence + volatile store); the solver lane polls
// `req`, solves and writes dvout, then the solver warp arrives on the slot's named barrier where the env warp is parked.

// Shared-memory loads the compiler must not hoist out of the iteration loop: the M^-1 columns and motor-row scalars are
// loop invariant, and hoisting 140 floats into registers spills them to LOCAL memory (seen in the SASS).
__device__ __forceinline__ float4 lds_v4(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds_f(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds_volatile_i(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_volatile_i(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
// Named hardware barrier (ids 1..ENVW, id 0 is __syncthreads): the env warp parks in bar.sync — no issue slots burnt,
// unlike an mbarrier.try_wait loop, whose time-out is short enough that 24 waiting warps took 43 % of the SM's issued
// instructions — and the solver WARP arrives on it (bar.arrive counts whole warps) when the slot's lane has answered.
__device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

// env warp side: post request `seq` (>0) for the rows just written, sleep until the solver lane has answered
__device__ __forceinline__ void solver_request(Smem& s, int seq, int slot, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    sts_volatile_i(&s.req, seq);
  }
  named_bar_sync(slot + 1);
}
__device__ __forceinline__ void solver_release(Smem& s, int lane) {  // the env is finished: its solver lane retires
  __syncwarp();
  if (lane == 0) sts_volatile_i(&s.req, -1);
}

#ifdef BMI_PROF
// Debug build only (tools/prof_rollout_phases.py): per-env cycle counters of the fused rollout.
__device__ unsigned long long g_prof[8192 * 8];
#define PROF_T0() long long prof_t0 = clock64()
#define PROF_ADD(e, k) do { const long long t1_ = clock64(); if (lane == 0 && (e) < 8192) g_prof[(e) * 8 + (k)] += (unsigned long long)(t1_ - prof_t0); prof_t0 = t1_; } while (0)
#define PROF_CNT(e, k, v) do { if (lane == 0 && (e) < 8192) g_prof[(e) * 8 + (k)] += (unsigned long long)(v); } while (0)
#else
#define PROF_T0()
#define PROF_ADD(e, k)
#define PROF_CNT(e, k, v)
#endif

__device__ __noinline__ void solver_loop(Smem* sw, int lane, unsigned live_mask) {
  const bool mine = lane < ENVW && ((live_mask >> lane) & 1u);
  Smem& s = sw[mine ? lane : 0];
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(&s);
  const unsigned a_minvr = sbase + (unsigned)offsetof(Smem, MinvR4), a_minv = sbase + (unsigned)offsetof(Smem, MinvP4),
                 a_rhs = sbase + (unsigned)offsetof(Smem, rhs), a_invd = sbase + (unsigned)offsetof(Smem, invd),
                 a_mdiag = sbase + (unsigned)offsetof(Smem, mdiag), a_lamn = sbase + (unsigned)offsetof(Smem, lamn);
  // arm deltas dv0..dv8 (named registers: the motor loop ROTATES them so that a rolled loop can index "the current
  // joint" statically; after 9 rows they are back in place) and the block's six deltas
  float dv0 = 0.f, dv1 = 0.f, dv2 = 0.f, dv3 = 0.f, dv4 = 0.f, dv5 = 0.f, dv6 = 0.f, dv7 = 0.f, dv8 = 0.f;
  float dvb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float max_imp = P(s, MP_MOTOR_FORCE) * P(s, MP_DT);
  const int max_it = (int)P(s, MP_SOLVER_ITERS);
  const float thresh = P(s, MP_RESIDUAL_THRESH);
  bool busy = false, finished = !mine;
  int seen = 0, it = 0, n_nc = 0, nc = 0, n_bt = 0;
#define BMI_ARM_APPLY(w0, w1, w2, d)                                                                        \
  do {                                                                                                      \
    dv0 = fmaf((w0).x, (d), dv0); dv1 = fmaf((w0).y, (d), dv1); dv2 = fmaf((w0).z, (d), dv2);               \
    dv3 = fmaf((w0).w, (d), dv3); dv4 = fmaf((w1).x, (d), dv4); dv5 = fmaf((w1).y, (d), dv5);               \
    dv6 = fmaf((w1).z, (d), dv6); dv7 = fmaf((w1).w, (d), dv7); dv8 = fmaf((w2), (d), dv8);                 \
  } while (0)
  auto arm_dot = [&](const float4* J4) -> float {
    const float4 j0 = J4[0], j1 = J4[1];
    const float j2 = reinterpret_cast<const float*>(J4)[8];
    const float a = j0.x * dv0 + j0.y * dv1 + j0.z * dv2;
    const float b = j0.w * dv3 + j1.x * dv4 + j1.y * dv5;
    const float c = j1.z * dv6 + j1.w * dv7 + j2 * dv8;
    return (a + b) + c;
  };
  auto blk_dot = [&](const float4& r0, const float4& r1) -> float {
    const float d0 = r0.x * dvb[0] + r0.y * dvb[1] + r0.z * dvb[2];
    const float d1 = r0.w * dvb[3] + r1.x * dvb[4] + r1.y * dvb[5];
    return d0 + d1;
  };
  auto blk_apply = [&](const float4& r1, const float4& r2, float d) {
    dvb[0] = fmaf(r1.z, d, dvb[0]); dvb[1] = fmaf(r1.w, d, dvb[1]); dvb[2] = fmaf(r2.x, d, dvb[2]);
    dvb[3] = fmaf(r2.y, d, dvb[3]); dvb[4] = fmaf(r2.z, d, dvb[4]); dvb[5] = fmaf(r2.w, d, dvb[5]);
  };
#pragma unroll 1
  while (true) {
    if (!busy && !finished) {  // poll this slot's request word
      const int r = lds_volatile_i(&s.req);
      if (r != seen) {
        seen = r;
        if (r < 0) finished = true;
        else {
          __threadfence_block();
          busy = true; it = 0;
          n_nc = s.n_nc; nc = s.nc; n_bt = s.n_bt;
          dv0 = dv1 = dv2 = dv3 = dv4 = dv5 = dv6 = dv7 = dv8 = 0.f;
#pragma unroll
          for (int i = 0; i < 6; ++i) dvb[i] = 0.f;
        }
      }
    }
    if (__ballot_sync(FULL, busy) == 0u) {
      if (__all_sync(FULL, finished)) break;
      __nanosleep(100);
      continue;
    }
    bool answered = false;
    if (busy) {  // ONE Gauss-Seidel iteration of this lane's request
      float resid = 0.f;
      // ---- motors: J = e_r, bounds +-max_imp.  Three rows per trip of a rolled loop, rotating the arm registers by
      // three so that the row's own delta is always dv0 / dv1 / dv2 (M^-1 columns are stored rotated to match).
#pragma unroll 1
      for (int r3 = 0; r3 < NL; r3 += 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const unsigned o = (unsigned)(r3 + k) * 4u, oc = (unsigned)(r3 + k) * (MP12 * 4u);
          const float dvr = k == 0 ? dv0 : (k == 1 ? dv1 : dv2);
          float d = lds_f(a_rhs + o) - dvr * lds_f(a_invd + o);
          const float old = lds_f(a_lamn + o);
          const float sum = fminf(fmaxf(old + d, -max_imp), max_imp);
          d = sum - old;
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_lamn + o), "f"(sum) : "memory");
          // rotated column: entry j multiplies the register that currently holds joint (r + j) % 9, i.e. register
          // (k + j) % 9 of this unrolled group
          const float4 w0 = lds_v4(a_minvr + oc), w1 = lds_v4(a_minvr + oc + 16);
          const float w2 = lds_f(a_minvr + oc + 32);
          if (k == 0) { BMI_ARM_APPLY(w0, w1, w2, d); }
          else if (k == 1) {
            dv1 = fmaf(w0.x, d, dv1); dv2 = fmaf(w0.y, d, dv2); dv3 = fmaf(w0.z, d, dv3); dv4 = fmaf(w0.w, d, dv4);
            dv5 = fmaf(w1.x, d, dv5); dv6 = fmaf(w1.y, d, dv6); dv7 = fmaf(w1.z, d, dv7); dv8 = fmaf(w1.w, d, dv8);
            dv0 = fmaf(w2, d, dv0);
          } else {
            dv2 = fmaf(w0.x, d, dv2); dv3 = fmaf(w0.y, d, dv3); dv4 = fmaf(w0.z, d, dv4); dv5 = fmaf(w0.w, d, dv5);
            dv6 = fmaf(w1.x, d, dv6); dv7 = fmaf(w1.y, d, dv7); dv8 = fmaf(w1.z, d, dv8); dv0 = fmaf(w1.w, d, dv0);
            dv1 = fmaf(w2, d, dv1);
          }
          const float res = d * lds_f(a_mdiag + o);
          resid = fmaxf(resid, res * res);
        }
        // rotate by three: register j now holds what register (j + 3) % 9 held
        const float t0 = dv0, t1 = dv1, t2 = dv2;
        dv0 = dv3; dv1 = dv4; dv2 = dv5; dv3 = dv6; dv4 = dv7; dv5 = dv8; dv6 = t0; dv7 = t1; dv8 = t2;
      }
      // ---- violated joint limits: J = +-e_j (rare; the registers are back in joint order here)
#pragma unroll 1
      for (int r = NL; r < n_nc; ++r) {
        const int jj = s.ncj[r];
        const int j = abs(jj) - 1;
        const float sgn = jj > 0 ? 1.f : -1.f;
        float dvj = dv0;
        dvj = j == 1 ? dv1 : dvj; dvj = j == 2 ? dv2 : dvj; dvj = j == 3 ? dv3 : dvj; dvj = j == 4 ? dv4 : dvj;
        dvj = j == 5 ? dv5 : dvj; dvj = j == 6 ? dv6 : dvj; dvj = j == 7 ? dv7 : dvj; dvj = j == 8 ? dv8 : dvj;
        float d = s.rhs[r] - sgn * dvj * s.invd[r];
        const float old = s.lamn[r];
        const float sum = fminf(fmaxf(old + d, s.lo[r]), s.hi[r]);
        d = sum - old;
        s.lamn[r] = sum;
        const unsigned a = a_minv + (unsigned)j * (MP12 * 4u);
        const float4 w0 = lds_v4(a), w1 = lds_v4(a + 16);
        const float w2 = lds_f(a + 32);
        const float sd = sgn * d;
        BMI_ARM_APPLY(w0, w1, w2, sd);
        const float res = d * s.mdiag[j];
        resid = fmaxf(resid, res * res);
      }
      // ---- contact normals.  Contacts [0, n_bt) are block-on-table (no arm part, all lanes alike), the rest touch an
      // arm link: two loops instead of a per-row branch.
#pragma unroll 1
      for (int c = 0; c < n_bt; ++c) {
        const float4 r0 = s.rd[c][0], r1 = s.rd[c][1], r2 = s.rd[c][2], r3 = s.rd[c][3];
        float d = r3.y - blk_dot(r0, r1) * r3.x;
        const float old = s.lam[c];
        const float sum = fmaxf(old + d, 0.f);
        d = sum - old;
        s.lam[c] = sum;
        blk_apply(r1, r2, d);
        const float res = d * r3.z;
        resid = fmaxf(resid, res * res);
      }
#pragma unroll 1
      for (int c = n_bt; c < nc; ++c) {
        const float4 r0 = s.rd[c][0], r1 = s.rd[c][1], r2 = s.rd[c][2], r3 = s.rd[c][3];
        const int asr = s.carm[c] * 3;
        float d = r3.y - (blk_dot(r0, r1) + arm_dot(s.Ja4 + asr * (MP12 / 4))) * r3.x;
        const float old = s.lam[c];
        const float sum = fmaxf(old + d, 0.f);
        d = sum - old;
        s.lam[c] = sum;
        blk_apply(r1, r2, d);
        {
          const float4* W4 = s.Wa4 + asr * (MP12 / 4);
          const float4 w0 = W4[0], w1 = W4[1];
          const float w2 = reinterpret_cast<const float*>(W4)[8];
          BMI_ARM_APPLY(w0, w1, w2, d);
        }
        const float res = d * r3.z;
        resid = fmaxf(resid, res * res);
      }
      // ---- friction cones (same split)
#pragma unroll 1
      for (int c = 0; c < nc; ++c) {
        const int ra = nc + 2 * c, rb = ra + 1;
        const float4 a0 = s.rd[ra][0], a1 = s.rd[ra][1], a2 = s.rd[ra][2], a3 = s.rd[ra][3];
        const float4 b0 = s.rd[rb][0], b1 = s.rd[rb][1], b2 = s.rd[rb][2], b3 = s.rd[rb][3];
        const float lim = a3.w * s.lam[c];
        float ja = blk_dot(a0, a1), jb = blk_dot(b0, b1);
        const bool arm = c >= n_bt;
        const int asr = arm ? s.carm[c] * 3 : 0;
        if (arm) {
          ja += arm_dot(s.Ja4 + (asr + 1) * (MP12 / 4));
          jb += arm_dot(s.Ja4 + (asr + 2) * (MP12 / 4));
        }
        const float oa = s.lam[ra], ob = s.lam[rb];
        float sa = oa + (a3.y - ja * a3.x), sb = ob + (b3.y - jb * b3.x);
        const float n2 = sa * sa + sb * sb;
        const float sc = n2 > lim * lim ? lim * rsqrtf(n2) : 1.f;  // branch-free cone projection (x * 1 is exact)
        sa *= sc; sb *= sc;
        const float da = sa - oa, db = sb - ob;
        s.lam[ra] = sa; s.lam[rb] = sb;
        blk_apply(a1, a2, da);
        blk_apply(b1, b2, db);
        if (arm) {
          const float4* Wa = s.Wa4 + (asr + 1) * (MP12 / 4);
          const float4* Wb = s.Wa4 + (asr + 2) * (MP12 / 4);
          const float4 u0 = Wa[0], u1 = Wa[1], v0 = Wb[0], v1 = Wb[1];
          const float u2 = reinterpret_cast<const float*>(Wa)[8], v2 = reinterpret_cast<const float*>(Wb)[8];
          BMI_ARM_APPLY(u0, u1, u2, da);
          BMI_ARM_APPLY(v0, v1, v2, db);
        }
        const float r1_ = da * a3.z, r2_ = db * b3.z;
        resid = fmaxf(resid, fmaxf(r1_ * r1_, r2_ * r2_));
      }
      ++it;
      if (resid <= thresh || it >= max_it) {  // answer: velocity deltas, then release the env warp
        s.dvout[0] = dv0; s.dvout[1] = dv1; s.dvout[2] = dv2; s.dvout[3] = dv3; s.dvout[4] = dv4;
        s.dvout[5] = dv5; s.dvout[6] = dv6; s.dvout[7] = dv7; s.dvout[8] = dv8;
#pragma unroll
        for (int i = 0; i < 6; ++i) s.dvout[NL + i] = dvb[i];
        s.dvout[15] = 0.f;
#ifdef BMI_PROF
        s.dvout[15] = (float)it;
#endif
        __threadfence_block();
        busy = false;
        answered = true;
      }
    }
    // wake the env warps whose lanes answered in this trip (warp-uniform loop: bar.arrive is a whole-warp operation)
    unsigned fin = __ballot_sync(FULL, answered);
    while (fin) {
      const int b = __ffs(fin) - 1;
      fin &= fin - 1u;
      named_bar_arrive(b + 1);
    }
  }
#undef BMI_ARM_APPLY
}

// ---- integrate ----------------------------------------------------------------------------------------
__device__ __noinline__ void substep_post(Smem& s, int lane) {
  const float dt = P(s, MP_DT);
  const float unew = lane < 16 ? s.u[lane] + s.dvout[lane] : 0.f;
  if (lane < NL) {
    s.qd[lane] = unew;
    s.q[lane] += dt * unew;
  } else if (lane < 12) {
    s.bv[lane - 9] = unew;
    s.bp[lane - 9] += dt * unew;
  } else if (lane < 15) {
    s.bw[lane - 12] = unew;
  }
  __syncwarp();
  if (lane == 0) {  // quaternion exponential map
    const float wn = sqrtf(dot3(s.bw, s.bw)), th = wn * dt;
    float ax[3];
    float sh, ch;
    sincos_compact(0.5f * th, &sh, &ch);
    if (wn < 1e-12f) { ax[0] = s.bw[0] * 0.5f * dt; ax[1] = s.bw[1] * 0.5f * dt; ax[2] = s.bw[2] * 0.5f * dt; }
    else { const float sc = sh / wn; ax[0] = s.bw[0] * sc; ax[1] = s.bw[1] * sc; ax[2] = s.bw[2] * sc; }
    const float dq[4] = {ax[0], ax[1], ax[2], ch}, q0[4] = {s.bq[0], s.bq[1], s.bq[2], s.bq[3]};
    float r[4];
    r[3] = dq[3] * q0[3] - dq[0] * q0[0] - dq[1] * q0[1] - dq[2] * q0[2];
    r[0] = dq[3] * q0[0] + dq[0] * q0[3] + dq[1] * q0[2] - dq[2] * q0[1];
    r[1] = dq[3] * q0[1] - dq[0] * q0[2] + dq[1] * q0[3] + dq[2] * q0[0];
    r[2] = dq[3] * q0[2] + dq[0] * q0[1] - dq[1] * q0[0] + dq[2] * q0[3];
    const float inv = rsqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    s.bq[0] = r[0] * inv; s.bq[1] = r[1] * inv; s.bq[2] = r[2] * inv; s.bq[3] = r[3] * inv;
  }
  __syncwarp();
}

// ---- observation -------------------------------------------------------------------------------------------
__device__ __noinline__ void observe(Smem& s, int lane, float* __restrict__ obs, float* __restrict__ ag) {
  fk(s, s.q, lane);
  if (lane == 0) {
    float w[3] = {0, 0, 0}, vo[3] = {0, 0, 0};
    int prev = -1;
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      if (i == 7) continue;  // hand1 is not on the path to the EE
      if (prev >= 0) {
        float r[3] = {s.p[i][0] - s.p[prev][0], s.p[i][1] - s.p[prev][1], s.p[i][2] - s.p[prev][2]}, t[3];
        cross3(t, w, r);
        vo[0] += t[0]; vo[1] += t[1]; vo[2] += t[2];
      }
      w[0] += s.qd[i] * s.z[i][0]; w[1] += s.qd[i] * s.z[i][1]; w[2] += s.qd[i] * s.z[i][2];
      prev = i;
    }
    float rc[3] = {s.c[EE][0] - s.p[EE][0], s.c[EE][1] - s.p[EE][1], s.c[EE][2] - s.p[EE][2]}, t[3];
    cross3(t, w, rc);
    const float* R = s.R[EE];
    float eul[3];
    const float sarg = -R[6];
    if (sarg <= -0.99999f) { eul[0] = 0.f; eul[1] = -1.57079632679f; eul[2] = atan2f(-R[1], -R[2]); }
    else if (sarg >= 0.99999f) { eul[0] = 0.f; eul[1] = 1.57079632679f; eul[2] = atan2f(-R[1], R[2]); }
    else { eul[0] = atan2f(R[7], R[8]); eul[1] = asinf(sarg); eul[2] = atan2f(R[3], R[0]); }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      obs[a] = s.p[EE][a];
      obs[3 + a] = eul[a];
      obs[6 + a] = vo[a] + t[a];   // link velocity reported at the COM
      obs[9 + a] = w[a];
      obs[12 + a] = s.bp[a];
      obs[15 + a] = eul[a];        // reference bug kept: blockOrn slot repeats the gripper euler
      obs[18 + a] = s.bp[a] - s.p[EE][a];
      obs[21 + a] = s.bv[a];
      obs[24 + a] = s.bw[a];
      ag[a] = s.bp[a];
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void load_state(Smem& s, const float* __restrict__ st, int lane) {
  for (int i = lane; i < BMI_ENV_STATE_DIM; i += 32) {
    const float v = st[i];
    if (i < ST_QD) s.q[i - ST_Q] = v;
    else if (i < ST_QT) s.qd[i - ST_QD] = v;
    else if (i < ST_BPOS) s.qt[i - ST_QT] = v;
    else if (i < ST_BQUAT) s.bp[i - ST_BPOS] = v;
    else if (i < ST_BVEL) s.bq[i - ST_BQUAT] = v;
    else if (i < ST_BANG) s.bv[i - ST_BVEL] = v;
    else if (i < ST_GOAL) s.bw[i - ST_BANG] = v;
    else if (i < ST_PAD) s.goal[i - ST_GOAL] = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void store_state(const Smem& s, float* __restrict__ st, int lane) {
  for (int i = lane; i < BMI_ENV_STATE_DIM; i += 32) {
    float v = 0.f;
    if (i < ST_QD) v = s.q[i - ST_Q];
    else if (i < ST_QT) v = s.qd[i - ST_QD];
    else if (i < ST_BPOS) v = s.qt[i - ST_QT];
    else if (i < ST_BQUAT) v = s.bp[i - ST_BPOS];
    else if (i < ST_BVEL) v = s.bq[i - ST_BQUAT];
    else if (i < ST_BANG) v = s.bv[i - ST_BVEL];
    else if (i < ST_GOAL) v = s.bw[i - ST_BANG];
    else if (i < ST_PAD) v = s.goal[i - ST_GOAL];
    st[i] = v;
  }
}

__device__ __forceinline__ float goal_dist(const Smem& s) {
  const float dx = s.bp[0] - s.goal[0], dy = s.bp[1] - s.goal[1], dz = s.bp[2] - s.goal[2];
  return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}
// ---- block layout ---------------------------------------------------------------------------------------------
struct BlockSmem {
  float model_s[STAGED];
  unsigned long long mbar;
  int smsp_of[WARPS];   // SM sub-partition (scheduler) each warp of the block sits on
  int target_smsp;      // sub-partition this block's solver warp should sit on
  int pad_[3 + (WARPS % 2 ? 1 : 0)];
  Smem sw[ENVW];
};
static_assert(offsetof(BlockSmem, sw) % 16 == 0, "Smem slots must be 16-byte aligned");
static_assert((sizeof(BlockSmem) + 1024) * BLOCKS_PER_SM <= 228 * 1024, "BLOCKS_PER_SM blocks (+1 KB reserved each) must fit the SM's 228 KB");
template <int N> struct PrintSize;
#ifdef BMI_PRINT_SIZES
PrintSize<sizeof(Smem)> print_smem_size;
#endif
extern __shared__ __align__(16) unsigned char bmi_dyn_smem[];

// Block prologue: stage the model, initialise the per-env mbarriers, elect the solver warp.
// The four blocks that share an SM each run one latency-bound solver warp; those must sit on DIFFERENT sub-partitions
// (an early version had all four on scheduler 0: 72 % busy there, 8 % on the other three).  The k-th block to arrive on
// an SM (atomic counter per SM, never reset: only k mod 4 matters) takes sub-partition k mod 4 and elects its first warp
// whose hardware slot (%warpid mod 4) lives there.  Returns the solver warp's index in the block.
__device__ __forceinline__ int block_begin(BlockSmem& bs, const float* __restrict__ model_g, int* __restrict__ sm_arrivals,
                                           int warp, int lane) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < ENVW; ++i) {
      bs.sw[i].req = 0;
      bs.sw[i].model = bs.model_s;
    }
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    bs.target_smsp = atomicAdd(sm_arrivals + (smid & 1023u), 1) & 3;
  }
  if (lane == 0) {
    unsigned wid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    bs.smsp_of[warp] = (int)(wid & 3u);
  }
  stage_model(bs.model_s, &bs.mbar, model_g, threadIdx.x);  // mbarrier-init fence + __syncthreads inside
  int solver = WARPS - 1;
#pragma unroll
  for (int w = WARPS - 1; w >= 0; --w) if (bs.smsp_of[w] == bs.target_smsp) solver = w;
  return solver;
}

// clip, (pick: auto-grip), IK, motor set-points  (bmirobot_env_push_F.py:92-101) — one warp per env
__device__ __noinline__ void env_step_begin(Smem& s, const EnvParams& ep, const float* __restrict__ model_g,
                                            const float* a_in, int lane) {
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = fminf(fmaxf(a_in[i], -0.5f), 0.5f);
  if (ep.task == BMI_TASK_PUSH) a[3] = 0.f;  // bmirobot_env_push_F.py:94
  fk(s, s.q, lane);
  if (ep.task == BMI_TASK_PICK) {  // auto-grip (bmirobot_env_pickandplace_v2.py:94-95)
    find_contacts(s, ep, model_g, 1e-4f, lane);
    bool touch = false;
    for (int c = 0; c < s.nc; ++c) touch |= (s.chasb[c] && s.clink[c] >= 0 && s.cdist[c] < 1e-4f);
    if (touch) a[3] = -1.f;
  }
  // applyAction (bmirobot.py:129-162)
  float target[3] = {fminf(fmaxf(s.p[EE][0] + a[0], -1.f), 1.f), fminf(fmaxf(s.p[EE][1] + a[1], -1.f), 1.f),
                     fminf(fmaxf(s.p[EE][2] + a[2], 0.f), 1.f)};
  __syncwarp();
  solve_ik(s, target, lane);
  if (lane < 7) s.qt[lane] = s.qik[lane];
  else if (lane == 7) s.qt[7] = s.q[7] + a[3];  // sent_hand_moving (bmirobot.py:163-191)
  else if (lane == 8) s.qt[8] = s.q[8] - a[3];
  __syncwarp();
}

// One env step of this warp's env: n_substeps x [set-up | solve on the solver warp | integrate].  `seq` is
// the warp's running request number.
__device__ __forceinline__ void env_step_warp(Smem& s, const EnvParams& ep, const float* __restrict__ model_g,
                                              const float* a_in, int lane, int& seq, int slot, int e) {
  PROF_T0();
  env_step_begin(s, ep, model_g, a_in, lane);
  PROF_ADD(e, 1);
  const int nsub = (int)P(s, MP_N_SUBSTEPS);
  for (int i = 0; i < nsub; ++i) {
    substep_pre(s, ep, model_g, lane);
    PROF_ADD(e, 2);
    solver_request(s, ++seq, slot, lane);
    PROF_ADD(e, 3);
#ifdef BMI_PROF
    PROF_CNT(e, 5, s.dvout[15]);
    PROF_CNT(e, 6, s.dvout[15] >= 150.f ? 1 : 0);
    PROF_CNT(e, 7, s.nc);
    __syncwarp();
    if (lane == 0) s.dvout[15] = 0.f;
    __syncwarp();
#endif
    substep_post(s, lane);
    PROF_ADD(e, 4);
  }
}

__device__ __forceinline__ unsigned live_env_mask(int n_envs) {
  const int left = n_envs - (int)blockIdx.x * ENVW;
  return left >= ENVW ? ((1u << ENVW) - 1u) : ((1u << max(left, 0)) - 1u);
}

// ---- kernels ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WARPS, BLOCKS_PER_SM)
env_step_kernel(const float* __restrict__ model_g, EnvParams ep, int n_envs, int* __restrict__ sm_arrivals,
                float* __restrict__ state, const float* __restrict__ actions, float* __restrict__ obs,
                float* __restrict__ ag, float* __restrict__ reward, float* __restrict__ success) {
  BlockSmem& bs = *reinterpret_cast<BlockSmem*>(bmi_dyn_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int solver_warp = block_begin(bs, model_g, sm_arrivals, warp, lane);
  if (warp == solver_warp) { solver_loop(bs.sw, lane, live_env_mask(n_envs)); return; }
  const int slot = warp < solver_warp ? warp : warp - 1;
  const int e = blockIdx.x * ENVW + slot;
  if (e >= n_envs) return;  // whole warp; its slot is not in the live mask
  Smem& s = bs.sw[slot];
  int seq = 0;
  load_state(s, state + (size_t)e * BMI_ENV_STATE_DIM, lane);
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = actions[e * 4 + i];
  env_step_warp(s, ep, model_g, a, lane, seq, slot, e);
  solver_release(s, lane);
  observe(s, lane, obs + (size_t)e * BMI_OBS_DIM, ag + (size_t)e * BMI_GOAL_DIM);
  if (lane == 0) {
    const float dist = goal_dist(s);
    const float thr = P(s, MP_DIST_THRESHOLD);
    if (success) success[e] = dist < thr ? 1.f : 0.f;
    if (reward) reward[e] = dist > thr ? -1.f : -0.f;
  }
  store_state(s, state + (size_t)e * BMI_ENV_STATE_DIM, lane);
}

// ---- fused rollout: policy MLP + exploration noise + episode record + env step, T steps per launch ---------------
struct RolloutArgs {
  int T, explore;
  const float* actor_t;      // transposed actor weights: Wt1[Dx][HID] b1 Wt2[HID][HID] b2 Wt3[HID][HID] b3 Wt4[HID][4] b4
  const float *o_mean, *o_std, *g_mean, *g_std;
  float clip_range, action_max, noise_eps, random_eps, late_clip;
  unsigned long long seed;
  const unsigned long long* counter;
  float *ep_obs, *ep_ag, *ep_g, *ep_act;   // [n][T+1][27] [n][T+1][3] [n][T][3] [n][T][4] or null
  const float* init;                       // [n][8] placements: the episode starts with a reset
  float *obs, *ag, *g, *success;           // final observation / flags
};

// one hidden layer: out[HID] = relu(Wt[n_in][HID]^T x + b); lane owns outputs 8*lane .. 8*lane+7
__device__ __noinline__ void policy_layer(const float* __restrict__ Wt, const float* __restrict__ b, const float* x,
                                             int n_in, float* out, int lane) {
  float acc[8];
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * lane), b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * lane + 1);
  acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
  const float4* W4 = reinterpret_cast<const float4*>(Wt) + 2 * lane;
#pragma unroll 4
  for (int k = 0; k < n_in; ++k) {
    const float xk = x[k];
    const float4 w0 = __ldg(W4 + (size_t)k * (HID / 4)), w1 = __ldg(W4 + (size_t)k * (HID / 4) + 1);
    acc[0] = fmaf(xk, w0.x, acc[0]); acc[1] = fmaf(xk, w0.y, acc[1]); acc[2] = fmaf(xk, w0.z, acc[2]); acc[3] = fmaf(xk, w0.w, acc[3]);
    acc[4] = fmaf(xk, w1.x, acc[4]); acc[5] = fmaf(xk, w1.y, acc[5]); acc[6] = fmaf(xk, w1.z, acc[6]); acc[7] = fmaf(xk, w1.w, acc[7]);
  }
  float4* o4 = reinterpret_cast<float4*>(out) + 2 * lane;
  o4[0] = make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f));
  o4[1] = make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f));
  __syncwarp();
}

__device__ __noinline__ float norm_clip(float v, float m, float sd, float clip) {
  // ddpg_agent._preproc_inputs: float64 (v - mean) / std, clip, then float32 (same as bmi_preproc_inputs)
  const double z = __ddiv_rn(__dsub_rn((double)v, (double)m), (double)sd);
  return (float)fmin(fmax(z, -(double)clip), (double)clip);
}

__device__ __noinline__ Philox4 philox_explore(unsigned long long seed, unsigned long long ctr) {
  return philox4x32_10(seed, ctr, kStreamExplore);
}

__global__ void __launch_bounds__(32 * WARPS, BLOCKS_PER_SM)
rollout_kernel(const float* __restrict__ model_g, EnvParams ep, int n_envs, int* __restrict__ sm_arrivals,
               float* __restrict__ state, RolloutArgs ra) {
  BlockSmem& bs = *reinterpret_cast<BlockSmem*>(bmi_dyn_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int solver_warp = block_begin(bs, model_g, sm_arrivals, warp, lane);
  if (warp == solver_warp) { solver_loop(bs.sw, lane, live_env_mask(n_envs)); return; }
  const int slot = warp < solver_warp ? warp : warp - 1;
  const int e = blockIdx.x * ENVW + slot;
  if (e >= n_envs) return;  // whole warp; its slot is not in the live mask
  Smem& s = bs.sw[slot];
  int seq = 0;
  float* st = state + (size_t)e * BMI_ENV_STATE_DIM;
  if (ra.init != nullptr) {  // reset (bmirobot_env_push_F.py:110-165)
    const float* in = ra.init + (size_t)e * 8;
    for (int i = lane; i < BMI_ENV_STATE_DIM; i += 32) {
      float v = 0.f;
      if (i >= ST_BPOS && i < ST_BPOS + 3) v = in[i - ST_BPOS];
      else if (i == ST_BQUAT + 2 || i == ST_BQUAT + 3) {
        float sy, cy;
        sincos_compact(0.5f * in[3], &sy, &cy);
        v = i == ST_BQUAT + 2 ? sy : cy;
      }
      else if (i >= ST_GOAL && i < ST_GOAL + 3) v = in[4 + i - ST_GOAL];
      st[i] = v;
    }
    __syncwarp();
  }
  load_state(s, st, lane);
  observe(s, lane, s.obs, s.obs + BMI_OBS_DIM);
  constexpr int Do = BMI_OBS_DIM, Dg = BMI_GOAL_DIM, Da = BMI_ACT_DIM, Dx = Do + Dg;
  const float* Wt1 = ra.actor_t;
  const float* b1 = Wt1 + Dx * HID;
  const float* Wt2 = b1 + HID;
  const float* b2 = Wt2 + HID * HID;
  const float* Wt3 = b2 + HID;
  const float* b3 = Wt3 + HID * HID;
  const float* Wt4 = b3 + HID;
  const float* b4 = Wt4 + HID * Da;
  const unsigned long long ctr0 = ra.explore ? *ra.counter : 0ull;
  for (int t = 0; t < ra.T; ++t) {
    PROF_T0();
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    // ---- record obs / ag / g of step t -----------------------------------------------------------------
    if (ra.ep_obs) {
      if (lane < Do) ra.ep_obs[((size_t)e * (ra.T + 1) + t) * Do + lane] = s.obs[lane];
      if (lane < Dg) {
        ra.ep_ag[((size_t)e * (ra.T + 1) + t) * Dg + lane] = s.obs[Do + lane];
        ra.ep_g[((size_t)e * ra.T + t) * Dg + lane] = s.goal[lane];
      }
    }
    // ---- policy: normalise -> 3 hidden layers -> tanh head (ddpg_agent.py:113-116) ------------------------
    if (lane < Do) s.pol.x[lane] = norm_clip(s.obs[lane], ra.o_mean[lane], ra.o_std[lane], ra.clip_range);
    else if (lane < Dx) s.pol.x[lane] = norm_clip(s.goal[lane - Do], ra.g_mean[lane - Do], ra.g_std[lane - Do], ra.clip_range);
    __syncwarp();
    policy_layer(Wt1, b1, s.pol.x, Dx, s.pol.hA, lane);
    policy_layer(Wt2, b2, s.pol.hA, HID, s.pol.hB, lane);
    policy_layer(Wt3, b3, s.pol.hB, HID, s.pol.hA, lane);
    float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = lane * 8 + kk;
      const float hk = s.pol.hA[k];
      const float4 w = __ldg(reinterpret_cast<const float4*>(Wt4) + k);
      z[0] = fmaf(hk, w.x, z[0]); z[1] = fmaf(hk, w.y, z[1]); z[2] = fmaf(hk, w.z, z[2]); z[3] = fmaf(hk, w.w, z[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = z[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      a[j] = ra.action_max * tanhf(v + __ldg(b4 + j));
    }
    if (ra.explore) {  // _select_actions (ddpg_agent.py:174-184): same Philox stream as bmi_select_actions
      const unsigned long long c = ctr0 + (unsigned long long)t * (unsigned long long)n_envs + (unsigned long long)e;
      const Philox4 pg = philox_explore(ra.seed, 3 * c);
      const Philox4 pu = philox_explore(ra.seed, 3 * c + 1);
      const Philox4 pb = philox_explore(ra.seed, 3 * c + 2);
      const bool take_random = u24(pb.v[0]) < ra.random_eps;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int pair = (j >> 1) & 1;
        const float u1 = 1.0f - u24(pg.v[2 * pair]);
        const float u2 = u24(pg.v[2 * pair + 1]);
        const float rad = sqrtf(-2.0f * logf(u1));
        const float gz = (j & 1) ? rad * sinf(6.28318530717958647692f * u2) : rad * cosf(6.28318530717958647692f * u2);
        float v = a[j] + ra.noise_eps * ra.action_max * gz;
        v = fminf(fmaxf(v, -ra.action_max), ra.action_max);
        const float rv = -ra.action_max + 2.0f * ra.action_max * u24(pu.v[j & 3]);
        if (take_random) v = rv;
        if (ra.late_clip > 0.f) v = fminf(fmaxf(v, -ra.late_clip), ra.late_clip);
        a[j] = v;
      }
    }
    if (ra.ep_act && lane < Da) {
      float v = a[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) if (lane == j) v = a[j];
      ra.ep_act[((size_t)e * ra.T + t) * Da + lane] = v;
    }
    __syncwarp();
    // ---- env step (the sub-step solves run on the block's solver warp) -------------------------------------------
    PROF_ADD(e, 0);
    env_step_warp(s, ep, model_g, a, lane, seq, slot, e);
    observe(s, lane, s.obs, s.obs + Do);
  }
  solver_release(s, lane);
  if (ra.ep_obs) {
    if (lane < Do) ra.ep_obs[((size_t)e * (ra.T + 1) + ra.T) * Do + lane] = s.obs[lane];
    if (lane < Dg) ra.ep_ag[((size_t)e * (ra.T + 1) + ra.T) * Dg + lane] = s.obs[Do + lane];
  }
  if (ra.obs && lane < Do) ra.obs[(size_t)e * Do + lane] = s.obs[lane];
  if (ra.ag && lane < Dg) ra.ag[(size_t)e * Dg + lane] = s.obs[Do + lane];
  if (ra.g && lane < Dg) ra.g[(size_t)e * Dg + lane] = s.goal[lane];
  if (ra.success && lane == 0) ra.success[e] = goal_dist(s) < P(s, MP_DIST_THRESHOLD) ? 1.f : 0.f;
  store_state(s, st, lane);
}

// W[out][in] (torch layout) -> Wt[in][out]
__global__ void transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt, int n_out, int n_in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_out * n_in) {
    const int o = i / n_in, k = i % n_in;
    Wt[(size_t)k * n_out + o] = W[i];
  }
}

constexpr int RESET_WARPS = 4;
__global__ void __launch_bounds__(32 * RESET_WARPS)
env_reset_kernel(const float* __restrict__ model_g, int n_envs, float* __restrict__ state,
                 const unsigned char* __restrict__ mask, const float* __restrict__ init, float* __restrict__ obs,
                 float* __restrict__ ag, float* __restrict__ g) {
  __shared__ __align__(16) float model_s[STAGED];
  __shared__ unsigned long long mbar_s;
  __shared__ Smem sw[RESET_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.x * RESET_WARPS + warp;
  stage_model(model_s, &mbar_s, model_g, threadIdx.x);
  if (e >= n_envs) return;
  Smem& s = sw[warp];
  if (lane == 0) s.model = model_s;
  __syncwarp();
  float* st = state + (size_t)e * BMI_ENV_STATE_DIM;
  if (mask == nullptr || mask[e]) {
    const float* in = init + (size_t)e * 8;
    for (int i = lane; i < BMI_ENV_STATE_DIM; i += 32) {
      float v = 0.f;
      if (i >= ST_BPOS && i < ST_BPOS + 3) v = in[i - ST_BPOS];
      else if (i == ST_BQUAT + 2 || i == ST_BQUAT + 3) {
        float sy, cy;
        sincos_compact(0.5f * in[3], &sy, &cy);
        v = i == ST_BQUAT + 2 ? sy : cy;
      }
      else if (i >= ST_GOAL && i < ST_GOAL + 3) v = in[4 + i - ST_GOAL];
      st[i] = v;
    }
    __syncwarp();
  }
  load_state(s, st, lane);
  observe(s, lane, obs + (size_t)e * BMI_OBS_DIM, ag + (size_t)e * BMI_GOAL_DIM);
  if (lane < 3) g[e * 3 + lane] = s.goal[lane];
}

// rejection-sampled block / goal placement (bmirobot_env_push_F.py:117-132; pick: pickandplace_v2.py:116-131)
__global__ void env_sample_init_kernel(int n, int task, uint64_t seed, const uint64_t* __restrict__ counter,
                                       float* __restrict__ init) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const uint64_t c = *counter + (uint64_t)e;
  float x = 0, y = 0, ang = 0, xt = 0, yt = 0, zt = 0.2f;
  for (int attempt = 0; attempt < 100; ++attempt) {
    Philox4 p0 = philox4x32_10(seed, c * 128 + 2 * attempt, kStreamReset);
    Philox4 p1 = philox4x32_10(seed, c * 128 + 2 * attempt + 1, kStreamReset);
    x = 0.15f + 0.2f * u24(p0.v[0]);
    y = u24(p0.v[1]) * 0.3f + 0.2f;
    ang = 3.14f * 0.5f + 3.1415925438f * u24(p0.v[2]);
    xt = 0.35f * u24(p0.v[3]);
    if (task == BMI_TASK_PUSH) { yt = u24(p1.v[0]) * 0.3f + 0.2f; zt = 0.2f; }
    else { yt = u24(p1.v[0]) * 0.25f + 0.3f; zt = 0.3f + 0.2f * u24(p1.v[1]); }
    const float dx = x - xt, dy = y - yt, dz = 0.2f - zt;
    if (sqrtf(dx * dx + dy * dy + dz * dz) >= 0.15f) break;
  }
  float* o = init + (size_t)e * 8;
  o[0] = x; o[1] = y; o[2] = 0.2f; o[3] = ang; o[4] = xt; o[5] = yt; o[6] = zt; o[7] = 0.f;
}
__global__ void advance_counter_kernel3(uint64_t* counter, uint64_t by) { *counter += by; }

__global__ void copy_state_kernel(float* __restrict__ dst, const float* __restrict__ src, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

}  // namespace bmi

using namespace bmi;

struct bmi_env {
  int n_envs = 0;
  EnvParams ep;
  float* model_dev = nullptr;
  float* state_dev = nullptr;
  int64_t model_floats = 0;
  int* sm_arrivals = nullptr;   // per-SM block arrival counters (solver-warp placement), 1024 ints
};

extern "C" int bmi_env_create(bmi_env** out, int32_t n_envs, int32_t task, const void* blob, int64_t bytes) {
  BMI_REQUIRE(out && blob, "bmi_env_create: null pointer");
  BMI_REQUIRE(n_envs > 0, "bmi_env_create: n_envs must be positive");
  BMI_REQUIRE(task == BMI_TASK_PUSH || task == BMI_TASK_PICK, "bmi_env_create: unknown task %d", task);
  BMI_REQUIRE(bytes >= (int64_t)(BMI_MODEL_HDR * sizeof(float)) && bytes % 4 == 0, "bmi_env_create: bad model blob size");
  const float* b = (const float*)blob;
  const int64_t n = bytes / 4;
  BMI_REQUIRE(b[MP_MAGIC] == BMI_MODEL_MAGIC && (int64_t)b[MP_TOTAL] == n && n <= BMI_MODEL_MAX_FLOATS,
              "bmi_env_create: model blob magic/size mismatch");
  BMI_REQUIRE((int)b[MP_N_LINKS] == NL && (int)b[MP_EE_LINK] == EE && (int)b[MP_N_SHAPES] <= BMI_MAX_SHAPES &&
                  (int)b[MP_LINKS_OFF] == BMI_MODEL_HDR,
              "bmi_env_create: model does not match the compiled arm topology");
  for (int i = 0; i < NL; ++i) {
    const float* lk = b + BMI_MODEL_HDR + i * BMI_LINK_STRIDE;
    BMI_REQUIRE((int)lk[ML_PARENT] == parent_of(i), "bmi_env_create: link %d has parent %d, kernel expects %d", i,
                (int)lk[ML_PARENT], parent_of(i));
  }
  for (int si = 0; si < (int)b[MP_N_SHAPES]; ++si) {
    const float* sh = b + (int)b[MP_SHAPES_OFF] + si * BMI_SHAPE_STRIDE;
    BMI_REQUIRE((int)sh[MS_NVERTS] <= 24 && ((int)b[MP_POOL_OFF] + (int)sh[MS_PLANE_OFF]) % 4 == 0,
                "bmi_env_create: shape %d needs <= 24 vertices and 16-byte aligned planes", si);
  }
  {  // the env kernels keep ENVW env working sets per block in dynamic shared memory (> 48 KB: opt-in)
    BMI_CUDA_CHECK(cudaFuncSetAttribute(env_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlockSmem)));
    BMI_CUDA_CHECK(cudaFuncSetAttribute(rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BlockSmem)));
  }
  bmi_env* h = new bmi_env();
  h->n_envs = n_envs;

  h->ep.task = task;
  const int o = task == BMI_TASK_PUSH ? MP_PUSH_HX : MP_PICK_HX;
  for (int a = 0; a < 3; ++a) h->ep.bh[a] = b[o + a];
  h->ep.bmass = b[o + 3];
  h->ep.bmu = b[o + 4];
  const float lx = 2 * h->ep.bh[0], ly = 2 * h->ep.bh[1], lz = 2 * h->ep.bh[2], mm = h->ep.bmass / 12.f;
  h->ep.binertia[0] = mm * (ly * ly + lz * lz);
  h->ep.binertia[1] = mm * (lx * lx + lz * lz);
  h->ep.binertia[2] = mm * (lx * lx + ly * ly);
  h->model_floats = n;
  if (cudaMalloc(&h->model_dev, n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&h->state_dev, (size_t)n_envs * BMI_ENV_STATE_DIM * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&h->sm_arrivals, 1024 * sizeof(int)) != cudaSuccess) {
    set_error("bmi_env_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    bmi_env_destroy(h);
    return BMI_ERR_CUDA;
  }
  BMI_CUDA_CHECK(cudaMemcpy(h->model_dev, blob, n * sizeof(float), cudaMemcpyHostToDevice));
  BMI_CUDA_CHECK(cudaMemset(h->state_dev, 0, (size_t)n_envs * BMI_ENV_STATE_DIM * sizeof(float)));
  BMI_CUDA_CHECK(cudaMemset(h->sm_arrivals, 0, 1024 * sizeof(int)));
  *out = h;
  return BMI_OK;
}

extern "C" int bmi_env_destroy(bmi_env* h) {
  if (!h) return BMI_OK;
  if (h->model_dev) cudaFree(h->model_dev);
  if (h->state_dev) cudaFree(h->state_dev);
  if (h->sm_arrivals) cudaFree(h->sm_arrivals);
  delete h;
  return BMI_OK;
}

extern "C" int32_t bmi_env_num_envs(const bmi_env* h) { return h ? h->n_envs : -1; }

extern "C" int bmi_env_reset(bmi_env* h, const uint8_t* mask, const float* init, float* obs, float* ag, float* g,
                             bmi_stream_t stream) {
  BMI_REQUIRE(h && init && obs && ag && g, "bmi_env_reset: null pointer");
  env_reset_kernel<<<(h->n_envs + RESET_WARPS - 1) / RESET_WARPS, 32 * RESET_WARPS, 0, as_stream(stream)>>>(h->model_dev, h->n_envs, h->state_dev, mask,
                                                                                           init, obs, ag, g);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_env_sample_init(bmi_env* h, uint64_t seed, uint64_t* counter, float* init, bmi_stream_t stream) {
  BMI_REQUIRE(h && counter && init, "bmi_env_sample_init: null pointer");
  env_sample_init_kernel<<<(h->n_envs + 127) / 128, 128, 0, as_stream(stream)>>>(h->n_envs, h->ep.task, seed, counter, init);
  BMI_LAUNCHED();
  advance_counter_kernel3<<<1, 1, 0, as_stream(stream)>>>(counter, (uint64_t)h->n_envs);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_env_step(bmi_env* h, const float* actions, float* obs, float* ag, float* reward, float* success,
                            bmi_stream_t stream) {
  BMI_REQUIRE(h && actions && obs && ag, "bmi_env_step: null pointer");
  env_step_kernel<<<(h->n_envs + ENVW - 1) / ENVW, 32 * WARPS, sizeof(BlockSmem), as_stream(stream)>>>(h->model_dev, h->ep, h->n_envs, h->sm_arrivals, h->state_dev,
                                                                                          actions, obs, ag, reward, success);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_env_get_state(bmi_env* h, float* st, bmi_stream_t stream) {
  BMI_REQUIRE(h && st, "bmi_env_get_state: null pointer");
  const int n = h->n_envs * BMI_ENV_STATE_DIM;
  copy_state_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(st, h->state_dev, n);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_env_set_state(bmi_env* h, const float* st, bmi_stream_t stream) {
  BMI_REQUIRE(h && st, "bmi_env_set_state: null pointer");
  const int n = h->n_envs * BMI_ENV_STATE_DIM;
  copy_state_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(h->state_dev, st, n);
  BMI_LAUNCHED();
  return BMI_OK;
}

extern "C" int bmi_actor_transpose(const float* actor_params, int32_t obs_dim, int32_t goal_dim, int32_t act_dim,
                                   int32_t hidden, float* actor_t, bmi_stream_t stream) {
  BMI_REQUIRE(actor_params && actor_t, "bmi_actor_transpose: null pointer");
  BMI_REQUIRE(hidden == HID && obs_dim == BMI_OBS_DIM && goal_dim == BMI_GOAL_DIM && act_dim == BMI_ACT_DIM,
              "bmi_actor_transpose: the fused rollout is compiled for 27+3 -> 256 -> 256 -> 256 -> 4");
  cudaStream_t st = as_stream(stream);
  const int ins[4] = {obs_dim + goal_dim, hidden, hidden, hidden}, outs[4] = {hidden, hidden, hidden, act_dim};
  size_t off = 0;
  for (int l = 0; l < 4; ++l) {
    const int n = ins[l] * outs[l];
    transpose_kernel<<<(n + 255) / 256, 256, 0, st>>>(actor_params + off, actor_t + off, outs[l], ins[l]);
    BMI_LAUNCHED();
    off += n;
    BMI_CUDA_CHECK(cudaMemcpyAsync(actor_t + off, actor_params + off, outs[l] * sizeof(float), cudaMemcpyDeviceToDevice, st));
    off += outs[l];
  }
  return BMI_OK;
}

extern "C" int bmi_env_rollout(bmi_env* h, const bmi_rollout_args* a, bmi_stream_t stream) {
  BMI_REQUIRE(h && a, "bmi_env_rollout: null pointer");
  BMI_REQUIRE(a->T > 0 && a->actor_t && a->o_mean && a->o_std && a->g_mean && a->g_std, "bmi_env_rollout: missing policy inputs");
  BMI_REQUIRE(!a->explore || a->counter, "bmi_env_rollout: exploration needs a Philox counter");
  RolloutArgs ra;
  ra.T = a->T; ra.explore = a->explore; ra.actor_t = a->actor_t;
  ra.o_mean = a->o_mean; ra.o_std = a->o_std; ra.g_mean = a->g_mean; ra.g_std = a->g_std;
  ra.clip_range = a->clip_range; ra.action_max = a->action_max; ra.noise_eps = a->noise_eps;
  ra.random_eps = a->random_eps; ra.late_clip = a->late_clip; ra.seed = a->seed; ra.counter = (const unsigned long long*)a->counter;
  ra.ep_obs = ra.ep_ag = ra.ep_g = ra.ep_act = nullptr;
  if (a->episodes) {
    const bmi_episodes* e = a->episodes;
    BMI_REQUIRE(e->dtype == BMI_F32 && e->T == a->T && e->n_episodes == h->n_envs && e->obs_dim == BMI_OBS_DIM &&
                    e->goal_dim == BMI_GOAL_DIM && e->act_dim == BMI_ACT_DIM,
                "bmi_env_rollout: episodes must be float32 [n_envs][T(+1)][27|3|3|4]");
    ra.ep_obs = (float*)e->obs; ra.ep_ag = (float*)e->ag; ra.ep_g = (float*)e->g; ra.ep_act = (float*)e->actions;
  }
  ra.init = a->init; ra.obs = a->obs; ra.ag = a->ag; ra.g = a->g; ra.success = a->success;
  cudaStream_t st = as_stream(stream);
  rollout_kernel<<<(h->n_envs + ENVW - 1) / ENVW, 32 * WARPS, sizeof(BlockSmem), st>>>(h->model_dev, h->ep, h->n_envs, h->sm_arrivals, h->state_dev, ra);
  BMI_LAUNCHED();
  if (a->explore) {
    advance_counter_kernel3<<<1, 1, 0, st>>>(a->counter, (uint64_t)a->T * (uint64_t)h->n_envs);
    BMI_LAUNCHED();
  }
  return BMI_OK;
}

#ifdef BMI_PROF
extern "C" int bmi_debug_prof(unsigned long long* host_out, int n_words, int reset) {
  if (host_out) BMI_CUDA_CHECK(cudaMemcpyFromSymbol(host_out, bmi::g_prof, (size_t)n_words * 8));
  if (reset) {
    void* p = nullptr;
    BMI_CUDA_CHECK(cudaGetSymbolAddress(&p, bmi::g_prof));
    BMI_CUDA_CHECK(cudaMemset(p, 0, sizeof(unsigned long long) * 8192 * 8));
  }
  return BMI_OK;
}
#endif
