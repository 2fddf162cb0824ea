// Vectorised bmirobot environment: one WARP per env instance, 28 envs per thread block, one block per SM
// (4096 envs = 147 blocks = one resident wave on a B200).
//
// Reference behaviour restated (paths relative to the reference tree; the arithmetic itself
// lives in PyBullet, so the algorithm follows oracle/bmi_physics_oracle.c):
//   bmirobot_env/bmirobot_env_push_F.py:92-108   step: clip, action[3]=0, IK + motors, 20 sub-steps
//   bmirobot_env/bmirobot_env_push_F.py:110-165  reset (block / goal placement ranges)
//   bmirobot_env/bmirobot_env_push_F.py:169-237  27-float observation
//   bmirobot_env/bmirobot.py:129-191             applyAction / sent_hand_moving
//   bmirobot_env/bmirobot_inverse_kinematics.py:28-33  position-only DLS IK of link 11
//   bmirobot_env/bmirobot_env_pickandplace_v2.py:92-95,116-131  pick task deltas
//
// Per sub-step: forward kinematics -> mass matrix by composite rigid bodies + bias by one
// Newton-Euler pass (joint_space_dynamics) -> register-resident Cholesky -> M^-1 -> unconstrained
// velocities -> contact generation (arm self-collision from the baked two-joint pair tables, bmirobot.py:58 flags=9; block /
// table / arm contacts: lane = vertex, bounding-sphere broadphase) -> constraint rows
// (lane = row) and their coupling table -> projected Gauss-Seidel in constraint space (lane = joint /
// block velocity component / contact; one shuffle per row update, substep_solve) -> semi-implicit
// Euler.  fp32 throughout, no tensor cores.
// The WHOLE model blob (solver constants, joint tree, collision polytopes: 5.5 KB) is staged into
// shared memory by one TMA bulk copy per block: the env working sets fill the SM's shared memory, so
// there is no L1 left and anything read from global memory would be an L2 round trip.
#include <vector>

#include "common.cuh"
#include "../../include/bmi_model.h"

namespace bmi {

constexpr int NL = 9;          // links / joints of the right arm
constexpr int NU = 15;         // generalized velocities: 9 joints + block linear 3 + angular 3
constexpr int EE = 8;          // right_hand2
constexpr int MAXC = 9;        // contacts per sub-step (one solver lane each)
constexpr int MAXA = 6;        // ... of which at most 6 involve an arm link
constexpr int MAXR = 3 * MAXC; // contact rows: normal + two friction directions per contact
// The device copy of the model keeps the link records at a stride of 33 floats (the file format's 32 would put the same
// field of all nine links into ONE shared-memory bank: every "lane = link" read was a 9-way conflict); bmi_env_create
// re-strides the blob and shifts the shape / pool offsets by LINK_SHIFT.
constexpr int LINK_STRIDE_DEV = BMI_LINK_STRIDE + 1;
constexpr int LINK_REGION_DEV = ((BMI_MAX_LINKS * LINK_STRIDE_DEV + 3) / 4) * 4;          // 300 floats (keeps 16-byte alignment)
constexpr int LINK_SHIFT = LINK_REGION_DEV - BMI_MAX_LINKS * BMI_LINK_STRIDE;             // 12
constexpr int STAGED = BMI_MODEL_HDR + LINK_REGION_DEV;  // floats staged by TMA (reset kernel: joint tree only)
constexpr int STAGED_FULL = 1408;  // env kernels stage the WHOLE blob (joint tree + collision polytopes); capacity in floats
constexpr int HID = 256;         // hidden width of the actor (models.py:15-17)
constexpr unsigned FULL = 0xffffffffu;
#ifndef BMI_BLOCKS_PER_SM
#define BMI_BLOCKS_PER_SM 1
#endif
#ifndef BMI_CONTACT_UNROLL
#define BMI_CONTACT_UNROLL 1   // solver contact loops: 1 = rolled (the unrolled body thrashes the 6 KB L0 I-cache: measured 25 % slower)
#endif
constexpr int kContactUnroll = BMI_CONTACT_UNROLL;
#ifndef BMI_SOLVE_SINGLE_VARIANT
#define BMI_SOLVE_SINGLE_VARIANT 0
#endif
// Block island (block-table rows while no arm link touches the block), see substep_solve (b), (c):
#ifndef BMI_BLK_WARMSTART
#define BMI_BLK_WARMSTART 0   // 1: start its rows from the previous sub-step's impulses.  OFF: four corner contacts are statically
                              // indeterminate, PGS then converges to a different point than Bullet's cold start and a sliding /
                              // spinning block decays differently (measured: 0.59 rad/s after one env-step at 5.8 rad/s)
#endif
#ifndef BMI_BLK_FREEZE
#define BMI_BLK_FREEZE 1      // stop sweeping its rows once converged (0: sweep them to the end like Bullet: 25 % slower)
#endif
#ifndef BMI_BLK_FREEZE_REST
#define BMI_BLK_FREEZE_REST 1e-2f   // "converged" = squared row-velocity change below this fraction of Bullet's threshold ...
#endif
#ifndef BMI_BLK_FREEZE_MOVING
#define BMI_BLK_FREEZE_MOVING 1e-4f // ... 100x tighter while the block moves (> 1 mm/s): there the truncation error accumulates
#endif
#ifndef BMI_MOTOR_UNROLL
#define BMI_MOTOR_UNROLL 3
#endif
constexpr int kMotorUnroll = BMI_MOTOR_UNROLL;
#ifndef BMI_ENVS_PER_BLOCK
#define BMI_ENVS_PER_BLOCK 28
#endif
constexpr int BLOCKS_PER_SM = BMI_BLOCKS_PER_SM;  // x ENVW envs: 28 envs per SM -> 4096 envs resident in one wave

// topology of the right arm: chain 0..6, two fingers on link 6 (checked against the blob)
__host__ __device__ constexpr int parent_of(int i) { return i == 0 ? -1 : (i <= 6 ? i - 1 : 6); }
__host__ __device__ constexpr bool is_ancestor_or_self(int a, int l) {
  return a == l || (a <= 6 && l >= a);  // every chain link j<=6 is an ancestor of all l>=j
}

// Env instances per thread block: ENVW warps, one env each; the envs only share the staged model copy.
constexpr int ENVW = BMI_ENVS_PER_BLOCK;
constexpr int WARPS = ENVW;
static_assert(ENVW >= 1 && ENVW <= 32, "one warp per env, at most 1024 threads per block");

// Solver lane map (solve_substep): lanes 0..8 own the joints, 9..14 the block's six velocity components,
// LANE_CT + c owns contact c (its normal and two friction rows).
constexpr int LANE_BLK = 9, LANE_CT = 16;
static_assert(LANE_CT + MAXC <= 32 && MAXR <= 32, "one lane per contact, one lane per contact row in the set-up");
constexpr int MS = 11;            // row stride of M^-1: lanes i = 0..8 reading entry (i, j) hit banks 11 i + j = {0,11,22,1,12,23,2,13,24} + j,
                                  // which leaves the runs 3..10 and 14..21 free for the contact lanes / the zero row (see Smem)
constexpr int SCOL_BLK = 9;       // coupling-table columns: [0, 9) joints, [9, 15) block velocity, [15, 15 + MAXR) contact rows
constexpr int SCOL_CT = 15;
constexpr int SS = SCOL_CT + MAXR + ((SCOL_CT + MAXR) % 2 == 0 ? 1 : 0);  // odd row stride
constexpr int MAXSC = 3 * MAXA;   // arm-Jacobian scratch rows

struct __align__(16) Smem {      // per-env (per-warp) working set
  const float* model;             // block-shared header params + link records (the TMA destination)
  int nc, na;                     // contacts, contacts on arm links
  float R[NL][9], p[NL][3], z[NL][3], c[NL][3];
  float q[NL], qd[NL], qt[NL], bias[NL], acc[NL];
  float u[16];
  float bp[3], bq[4], bv[3], bw[3], goal[3];
  float Rb[9], Ibinv[9];
  // contacts
  float cx[MAXC][3], cn[MAXC][3], cdist[MAXC], cmu[MAXC];
  int cinfo[MAXC];                // packed: link of body 2 (+1), block flag, arm-Jacobian slot (+1), link of body 1 of a self-contact (+2)
  int pad_c[2];
  // Coupling table of the constraint solver.  Row x (one per contact row, x = 3 c + k) holds what a unit impulse on
  // that row does to every solver variable: [0, 9) the joint velocities (M^-1 Ja^T), [9, 15) the block velocity,
  // [15 + y] the constraint-space velocity of contact row y (the Delassus entry J_y M^-1 J_x^T).
  // Minv, S and zero9 are laid out back to back on purpose: in one solver load the joint lanes read Minv (banks
  // 11 i + j), contact lane c reads S row 3 c (bank 3 + c + j relative to Minv, since 99 = 3 and 3 * 43 = 1 mod 32) and
  // the block / idle lanes read zero9 (bank 16 + j): no two of them share a bank for up to 8 contacts.
  float Minv[NL * MS];            // M^-1 (symmetric), entry (i, j) at [i * MS + j]
  float S[MAXR * SS];
  float zpad[4];
  float zero9[12];                // coefficients of the lanes a joint row does not touch
  union {
    struct {
      union {
        float Rl[NL][9];            // local joint rotations (fk only)
        struct {                    // joint_space_dynamics scratch
          union {
            struct { float cm[NL], cc[NL][3], cI[NL][6]; };           // subtree mass / COM / inertia (mass matrix)
            struct { float w[NL][3], al[NL][3], a[NL][3], vo[NL][3]; };  // link velocities / accelerations (bias)
          };
          float Nk[NL][3], Fk[NL][3];
        } x;
      };
      union {
        struct { float L[NL * NL], Iw[NL][6]; };          // mass matrix / Cholesky factor, world link inertias
        struct { float A[NL * NL], b[NL]; } ik;           // IK scratch (the IK runs before the sub-steps)
      };
    } dyn;                                                // live from fk to the end of contact generation
    struct { float4 sc[MAXR]; float Ja[MAXSC][NL]; } rows;  // row scalars (invd rhs diag mu) + arm Jacobians: set-up only
    struct { float x[32], h[HID]; } pol;                  // policy activations (fused rollout; between env steps)
  };
  float obs[BMI_OBS_DIM + BMI_GOAL_DIM];   // last observation + achieved goal (fused rollout)
  float qik[NL];
  int prof_e;
  // warm start of the block-table island (substep_solve): impulses of the previous sub-step's block-table contacts and
  // the block vertices they belong to (one byte each, 0xff = none); invalidated at the start of every env-step
  float blk_lam[12];
  unsigned blk_ids;
};

// one baked self-collision pair table (include/bmi_model.h SC_*), passed in kernel-parameter space
struct SCPair {
  int la, lb, ja, jb, na, nb;
  unsigned off;
  float a0, b0, inv_h, mu;
};
struct EnvParams {
  int task;
  float bh[3], bmass, binertia[3], bmu;
  int sc_np;                       // self-collision pair tables in use (0: none loaded)
  unsigned long long* drops;       // contacts dropped by the MAXC / MAXA lane budget (statistics)
};
// The pair-table descriptors live in constant memory (one table set per process: the robot model is a process-wide
// asset); kernel-parameter space would be copied to the local stack as soon as EnvParams is passed by reference.
__constant__ SCPair c_scp[BMI_SC_MAX_PAIRS];
__constant__ const float* c_sc_data;

// packed contact record: body 2 = arm link `link` or -1 (static world); hasb: body 1 is the block; arm: slot of the
// contact's arm-Jacobian scratch rows or -1; linkA: body 1 of an arm self-contact (-1 = the base link), -2 otherwise
__device__ __forceinline__ int c_pack(int link, int hasb, int arm, int linkA) {
  return (link + 1) | (hasb << 4) | ((arm + 1) << 5) | ((linkA + 2) << 9);
}
__device__ __forceinline__ int c_link(int i) { return (i & 15) - 1; }
__device__ __forceinline__ int c_hasb(int i) { return (i >> 4) & 1; }
__device__ __forceinline__ int c_arm(int i) { return ((i >> 5) & 15) - 1; }
__device__ __forceinline__ int c_linkA(int i) { return ((i >> 9) & 15) - 2; }
__device__ __forceinline__ int c_vid(int i) { return (i >> 13) & 7; }   // block vertex of a block-table contact

__device__ __forceinline__ float P(const Smem& s, int i) { return s.model[i]; }
__device__ __forceinline__ const float* LK(const Smem& s, int i) { return s.model + BMI_MODEL_HDR + i * LINK_STRIDE_DEV; }

__device__ __forceinline__ float& MINV(Smem& s, int i, int j) { return s.Minv[i * MS + j]; }


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

#ifdef BMI_PROF
// Debug build only (tools/prof_rollout_phases.py): per-env cycle counters of the fused rollout.
__device__ unsigned long long g_prof[8192 * 8];
#define PROF_T0() long long prof_t0 = clock64()
#define PROF_ADD(e, k) do { const long long t1_ = clock64(); if (lane == 0 && (e) < 8192) g_prof[(e) * 8 + (k)] += (unsigned long long)(t1_ - prof_t0); prof_t0 = t1_; } while (0)
#define PROF_CNT(e, k, v) do { if (lane == 0 && (e) < 8192) g_prof[(e) * 8 + (k)] += (unsigned long long)(v); } while (0)
#define PROF_RESET() prof_t0 = clock64()
#else
#define PROF_T0()
#define PROF_ADD(e, k)
#define PROF_CNT(e, k, v)
#define PROF_RESET()
#endif

// ---- TMA staging of the joint tree ---------------------------------------------------------
__device__ __forceinline__ void stage_model(float* model_s, unsigned long long* mbar_s,
                                            const float* __restrict__ model_g, int tid, unsigned bytes) {
  const int lane = tid;  // thread 0 of the block issues the copy; every thread of every warp waits on the barrier
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(mbar_s);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(model_s);
  // bytes: a multiple of 16 (TMA bulk copies move 16-byte units)
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
        s.dyn.Rl[lane][3 * r + cc] = Jr[3 * r] * Rq[cc] + Jr[3 * r + 1] * Rq[3 + cc] + Jr[3 * r + 2] * Rq[6 + cc];
  }
  __syncwarp();
#pragma unroll 1
  for (int i = 0; i < NL; ++i) {
    const int pa = parent_of(i);
    if (lane < 9) {
      const int r = lane / 3, cc = lane % 3;
      float v;
      if (pa < 0) v = s.dyn.Rl[i][lane];
      else v = s.R[pa][3 * r] * s.dyn.Rl[i][cc] + s.R[pa][3 * r + 1] * s.dyn.Rl[i][3 + cc] + s.R[pa][3 * r + 2] * s.dyn.Rl[i][6 + cc];
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

// ---- joint-space dynamics: mass matrix by composite rigid bodies, bias by one recursive Newton-Euler pass ---------
// Same quantities as the oracle's rnea() calls (oracle/bmi_physics_oracle.c: the mass matrix is ten RNEA sweeps there);
// here the work is spread over the warp instead of repeated per lane:
//   M_jk = z_j . [ Ic_k z_k + (cc_k - p_j) x mc_k (z_k x (cc_k - p_k)) ]   for j on the path base -> k
// (mc, cc, Ic: mass, centre of mass and inertia about it of the subtree hanging off joint k), and
//   bias_j = z_j . sum_{l in subtree(j)} [ N_l + (c_l - p_j) x F_l ]
// with the per-link wrenches F_l, N_l (velocity, gravity and Bullet link-damping terms) evaluated one link per lane
// after a single serial pass for the link velocities / accelerations.
__device__ __forceinline__ void sym_mat_vec(float* o, const float* I6, const float* v) {  // I6: xx xy xz yy yz zz
  const float x = I6[0] * v[0] + I6[1] * v[1] + I6[2] * v[2], y = I6[1] * v[0] + I6[3] * v[1] + I6[4] * v[2],
              z = I6[2] * v[0] + I6[4] * v[1] + I6[5] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}

__device__ __noinline__ void joint_space_dynamics(Smem& s, float gz, float kl, float ka, int lane) {
  auto& X = s.dyn.x;
  // (1) world inertia of every link about its own COM, R diag(I) R^T: lane = (link, component)
  for (int t = lane; t < NL * 6; t += 32) {
    const int i = t / 6, ci = t - 6 * i;
    const int a = ci < 3 ? 0 : (ci < 5 ? 1 : 2), b = ci < 3 ? ci : (ci < 5 ? ci - 2 : 2);
    const float* R = s.R[i];
    const float* I = LK(s, i) + ML_INERTIA;
    s.dyn.Iw[i][ci] = R[3 * a] * I[0] * R[3 * b] + R[3 * a + 1] * I[1] * R[3 * b + 1] + R[3 * a + 2] * I[2] * R[3 * b + 2];
  }
  // (2) mass and centre of mass of every subtree: lane = joint
  if (lane < NL) {
    float m = 0.f, h[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int l = lane; l < NL; ++l) {
      if (!is_ancestor_or_self(lane, l)) continue;
      const float ml = LK(s, l)[ML_MASS];
      m += ml;
      h[0] = fmaf(ml, s.c[l][0], h[0]); h[1] = fmaf(ml, s.c[l][1], h[1]); h[2] = fmaf(ml, s.c[l][2], h[2]);
    }
    const float im = 1.f / m;
    X.cm[lane] = m;
    X.cc[lane][0] = h[0] * im; X.cc[lane][1] = h[1] * im; X.cc[lane][2] = h[2] * im;
  }
  __syncwarp();
  // (3) inertia of every subtree about its centre of mass (parallel-axis sums, all terms positive): lane = (joint, comp)
  for (int t = lane; t < NL * 6; t += 32) {
    const int j = t / 6, ci = t - 6 * j;
    const int a = ci < 3 ? 0 : (ci < 5 ? 1 : 2), b = ci < 3 ? ci : (ci < 5 ? ci - 2 : 2);
    const float c0 = X.cc[j][0], c1 = X.cc[j][1], c2 = X.cc[j][2];
    float acc = 0.f;
#pragma unroll 1
    for (int l = j; l < NL; ++l) {
      if (!is_ancestor_or_self(j, l)) continue;
      const float d0 = s.c[l][0] - c0, d1 = s.c[l][1] - c1, d2 = s.c[l][2] - c2;
      const float da = a == 0 ? d0 : (a == 1 ? d1 : d2), db = b == 0 ? d0 : (b == 1 ? d1 : d2);
      const float dd = a == b ? d0 * d0 + d1 * d1 + d2 * d2 : 0.f;
      acc += s.dyn.Iw[l][ci] + LK(s, l)[ML_MASS] * (dd - da * db);
    }
    X.cI[j][ci] = acc;
  }
  __syncwarp();
  // (4) wrench of subtree k under a unit acceleration of joint k: lane = k
  if (lane < NL) {
    const float zk[3] = {s.z[lane][0], s.z[lane][1], s.z[lane][2]};
    const float r[3] = {X.cc[lane][0] - s.p[lane][0], X.cc[lane][1] - s.p[lane][1], X.cc[lane][2] - s.p[lane][2]};
    float N[3], ak[3];
    sym_mat_vec(N, X.cI[lane], zk);
    cross3(ak, zk, r);
    const float m = X.cm[lane];
    X.Nk[lane][0] = N[0]; X.Nk[lane][1] = N[1]; X.Nk[lane][2] = N[2];
    X.Fk[lane][0] = m * ak[0]; X.Fk[lane][1] = m * ak[1]; X.Fk[lane][2] = m * ak[2];
  }
  __syncwarp();
  // (5) lower triangle of M: lane = pair (k, j <= k)
  for (int t = lane; t < NL * (NL + 1) / 2; t += 32) {
    const int k = (t >= 1) + (t >= 3) + (t >= 6) + (t >= 10) + (t >= 15) + (t >= 21) + (t >= 28) + (t >= 36);
    const int j = t - k * (k + 1) / 2;
    float v = 0.f;
    if (is_ancestor_or_self(j, k)) {
      const float r[3] = {X.cc[k][0] - s.p[j][0], X.cc[k][1] - s.p[j][1], X.cc[k][2] - s.p[j][2]};
      float m[3];
      cross3(m, r, X.Fk[k]);
      v = s.z[j][0] * (X.Nk[k][0] + m[0]) + s.z[j][1] * (X.Nk[k][1] + m[1]) + s.z[j][2] * (X.Nk[k][2] + m[2]);
    }
    s.dyn.L[k * NL + j] = v;
  }
  __syncwarp();
  // (6) link velocities and accelerations (zero joint accelerations): one serial pass down the tree
  if (lane == 0) {
    float w[3] = {0.f, 0.f, 0.f}, al[3] = {0.f, 0.f, 0.f}, a[3] = {0.f, 0.f, -gz}, vo[3] = {0.f, 0.f, 0.f};
    float w6[3], al6[3], a6[3], vo6[3];
#pragma unroll 1
    for (int i = 0; i < NL; ++i) {
      const int pa = parent_of(i);
      if (i == 7 || i == 8) {  // fingers hang off link 6
#pragma unroll
        for (int k = 0; k < 3; ++k) { w[k] = w6[k]; al[k] = al6[k]; a[k] = a6[k]; vo[k] = vo6[k]; }
      }
      if (pa >= 0) {
        const float r[3] = {s.p[i][0] - s.p[pa][0], s.p[i][1] - s.p[pa][1], s.p[i][2] - s.p[pa][2]};
        float t[3], t2[3];
        cross3(t, w, r);
        vo[0] += t[0]; vo[1] += t[1]; vo[2] += t[2];
        cross3(t2, w, t);
        cross3(t, al, r);
        a[0] += t[0] + t2[0]; a[1] += t[1] + t2[1]; a[2] += t[2] + t2[2];
      }
      const float qdi = s.qd[i];
      const float zi[3] = {s.z[i][0], s.z[i][1], s.z[i][2]};
      {
        float t[3];
        cross3(t, w, zi);
        al[0] = fmaf(qdi, t[0], al[0]); al[1] = fmaf(qdi, t[1], al[1]); al[2] = fmaf(qdi, t[2], al[2]);
        w[0] = fmaf(qdi, zi[0], w[0]); w[1] = fmaf(qdi, zi[1], w[1]); w[2] = fmaf(qdi, zi[2], w[2]);
      }
      if (i == 6) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { w6[k] = w[k]; al6[k] = al[k]; a6[k] = a[k]; vo6[k] = vo[k]; }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) { X.w[i][k] = w[k]; X.al[i][k] = al[k]; X.a[i][k] = a[k]; X.vo[i][k] = vo[k]; }
    }
  }
  __syncwarp();
  // (7) wrench of every link, moment taken about the world origin: lane = link
  if (lane < NL) {
    const int i = lane;
    const float w[3] = {X.w[i][0], X.w[i][1], X.w[i][2]}, al[3] = {X.al[i][0], X.al[i][1], X.al[i][2]};
    const float ci[3] = {s.c[i][0], s.c[i][1], s.c[i][2]};
    const float rc[3] = {ci[0] - s.p[i][0], ci[1] - s.p[i][1], ci[2] - s.p[i][2]};
    const float mass = LK(s, i)[ML_MASS];
    float t[3], t2[3], t3[3];
    cross3(t, w, rc);    // velocity of the COM relative to the link origin
    cross3(t2, w, t);
    cross3(t3, al, rc);
    const float ac[3] = {X.a[i][0] + t3[0] + t2[0], X.a[i][1] + t3[1] + t2[1], X.a[i][2] + t3[2] + t2[2]};
    const float vc[3] = {X.vo[i][0] + t[0], X.vo[i][1] + t[1], X.vo[i][2] + t[2]};
    const float kd = mass * (kl + kl * sqrtf(dot3(vc, vc)));
    const float F[3] = {mass * ac[0] + kd * vc[0], mass * ac[1] + kd * vc[1], mass * ac[2] + kd * vc[2]};
    float Iw_w[3], N[3];
    sym_mat_vec(Iw_w, s.dyn.Iw[i], w);
    sym_mat_vec(N, s.dyn.Iw[i], al);
    cross3(t, w, Iw_w);
    const float kk = ka + ka * sqrtf(dot3(w, w));
    cross3(t2, ci, F);
    X.Fk[i][0] = F[0]; X.Fk[i][1] = F[1]; X.Fk[i][2] = F[2];
    X.Nk[i][0] = N[0] + t[0] + kk * Iw_w[0] + t2[0];
    X.Nk[i][1] = N[1] + t[1] + kk * Iw_w[1] + t2[1];
    X.Nk[i][2] = N[2] + t[2] + kk * Iw_w[2] + t2[2];
  }
  __syncwarp();
  // (8) bias torque of joint j: z_j . (sum of subtree moments about the origin - p_j x sum of subtree forces)
  if (lane < NL) {
    float FS[3] = {0.f, 0.f, 0.f}, NS[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
    for (int l = lane; l < NL; ++l) {
      if (!is_ancestor_or_self(lane, l)) continue;
      FS[0] += X.Fk[l][0]; FS[1] += X.Fk[l][1]; FS[2] += X.Fk[l][2];
      NS[0] += X.Nk[l][0]; NS[1] += X.Nk[l][1]; NS[2] += X.Nk[l][2];
    }
    float t[3];
    cross3(t, s.p[lane], FS);
    s.bias[lane] = s.z[lane][0] * (NS[0] - t[0]) + s.z[lane][1] * (NS[1] - t[1]) + s.z[lane][2] * (NS[2] - t[2]);
  }
  __syncwarp();
}

// Cholesky of the 9x9 SPD matrix in Lm (lower triangle valid; factor written back in place, reciprocal diagonal).
// Lane i < 9 keeps row i in registers; column j is finished with one shuffle for the pivot and one per trailing row
// (45 shuffles in all, no shared-memory round trips between columns).
__device__ __noinline__ void chol9(float* Lm, int lane) {
  const int row = lane < NL ? lane : NL - 1;
  float a[NL];
#pragma unroll
  for (int k = 0; k < NL; ++k) a[k] = lane < NL ? Lm[row * NL + k] : 0.f;   // entries k > row are never used; idle lanes must not
                                                                            // read what lane 8 writes back below (racecheck)
#pragma unroll
  for (int j = 0; j < NL; ++j) {
    const float pj = __shfl_sync(FULL, a[j], j);
    const float d = rsqrtf(fmaxf(pj, 1e-20f));   // the diagonal stores 1 / L_jj: the solves multiply instead of divide
    const float Lij = a[j] * d;
    a[j] = lane == j ? d : Lij;
#pragma unroll
    for (int k = j + 1; k < NL; ++k) {
      const float Lkj = __shfl_sync(FULL, Lij, k);
      a[k] = fmaf(-Lij, Lkj, a[k]);               // meaningful for rows i >= k
    }
  }
  if (lane < NL) {
#pragma unroll
    for (int k = 0; k < NL; ++k) Lm[lane * NL + k] = a[k];   // the strict upper triangle is never read
  }
  __syncwarp();
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
    // Damped least squares step  dq = J^T (J J^T + damp I)^-1 e : by the push-through identity this IS the oracle's
    // (J^T J + damp I)^-1 J^T e, with a 3x3 system instead of a 9x9 one (the form BussIK's CalcDeltaThetasDLS uses).
    float U[6] = {Jc[0] * Jc[0], Jc[0] * Jc[1], Jc[0] * Jc[2], Jc[1] * Jc[1], Jc[1] * Jc[2], Jc[2] * Jc[2]};
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {   // lanes 0..15 end up with identical sums (columns live in lanes 0..8)
#pragma unroll
      for (int k = 0; k < 6; ++k) U[k] += __shfl_xor_sync(FULL, U[k], o);
    }
    // LDL^T of [[a b c] [b d f] [c f g]] + damp I, then two triangular solves
    const float a_ = U[0] + damp, d0i = 1.f / a_;
    const float l10 = U[1] * d0i, l20 = U[2] * d0i;
    const float d1 = U[3] + damp - l10 * U[1], d1i = 1.f / d1;
    const float t21 = U[4] - l20 * U[1];
    const float l21 = t21 * d1i;
    const float d2 = U[5] + damp - l20 * U[2] - l21 * t21, d2i = 1.f / d2;
    const float y0 = e[0], y1 = e[1] - l10 * y0, y2 = e[2] - l20 * y0 - l21 * y1;
    const float x2 = y2 * d2i, x1 = y1 * d1i - l21 * x2, x0 = y0 * d0i - l10 * x1 - l20 * x2;
    const float dq = Jc[0] * x0 + Jc[1] * x1 + Jc[2] * x2;   // zero for lanes that own no column
    const float mx = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(lane < NL ? fabsf(dq) : 0.f)));
    const float sc = mx > maxang ? maxang / mx : 1.f;
    if (lane < NL) s.qik[lane] += sc * dq;
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

// linkA_of: per-lane body 1 of a self-contact (>= -1), or -2 for contacts against the table / the block
__device__ __noinline__ void push_contacts(Smem& s, const EnvParams& ep, unsigned mask, int lane, int link, int hasb, const float* x,
                                              const float* n, float dist, float mu, int linkA_of = -2) {
  if (mask == 0) return;
  const int base = s.nc, abase = s.na;
  // capacity: MAXC contacts in total, MAXA of them on arm links; later candidates (lane order) are dropped
  int allowed = MAXC - base;
  const bool arm = link >= 0 || linkA_of > -2;   // uniform per call (self-contacts come in their own call)
  if (arm) allowed = min(allowed, MAXA - abase);
  const int rank = __popc(mask & ((1u << lane) - 1));
  const int n_add = min(__popc(mask), max(allowed, 0));
  if (((mask >> lane) & 1u) && rank < n_add) {
    const int slot = base + rank;
    s.cx[slot][0] = x[0]; s.cx[slot][1] = x[1]; s.cx[slot][2] = x[2];
    s.cn[slot][0] = n[0]; s.cn[slot][1] = n[1]; s.cn[slot][2] = n[2];
    s.cdist[slot] = dist; s.cmu[slot] = mu;
    s.cinfo[slot] = c_pack(link, hasb, arm ? abase + rank : -1, linkA_of) | ((hasb && link < 0) ? (lane << 13) : 0);   // block vertex id
  }
  if (lane == 0 && n_add < __popc(mask) && ep.drops) atomicAdd(ep.drops, (unsigned long long)(__popc(mask) - n_add));
  __syncwarp();
  if (lane == 0) {
    s.nc = base + n_add;
    if (arm) s.na = abase + n_add;
  }
  __syncwarp();
}

// ---- self-collision of the arm: baked pair tables (tools/bake_selfcol.py + tools/geom; oracle: find_self_contacts / sc_lookup) -
// Every link pair that can touch is two joints apart, so the whole narrow phase (signed distance of the two hulls,
// normal, witness point) is a function of two joint angles, sampled off line with the oracle's GJK + EPA on a 5 mrad
// grid.  Here: lane = (pair p = lane / 8, node c = (lane / 2) % 4 of the surrounding cell, half = lane % 2 of the
// node's 32-byte record), so ONE coalesced 16-byte load per lane fetches all four cells of all four pairs; bilinear when
// the four nodes describe the same hull features (normals within 1.8 degrees, witness points within 2 mm), else the
// nearest node (feature switches are kinks of the distance function; the wrist rests on one).
__device__ __noinline__ void self_contacts(Smem& s, const EnvParams& ep, int lane) {
  const int p = lane >> 3, c = (lane >> 1) & 3, half = lane & 1;
  const bool live = p < ep.sc_np;
  const SCPair& d = c_scp[live ? p : 0];
  float fa = (s.q[d.ja] - d.a0) * d.inv_h, fb = (s.q[d.jb] - d.b0) * d.inv_h;
  fa = fminf(fmaxf(fa, 0.f), (float)(d.na - 1));
  fb = fminf(fmaxf(fb, 0.f), (float)(d.nb - 1));
  const int i = min((int)fa, d.na - 2), j = min((int)fb, d.nb - 2);
  const float wa = fa - (float)i, wb = fb - (float)j;
  const int near = (wa >= 0.5f ? 2 : 0) + (wb >= 0.5f ? 1 : 0);
  float4 v = make_float4(1e3f, 0.f, 0.f, 0.f);
  if (live) v = __ldg(reinterpret_cast<const float4*>(c_sc_data + d.off) + 2 * ((size_t)(i + (c >> 1)) * d.nb + (j + (c & 1))) + half);
  const int src = (lane & ~7) | (near << 1) | half;
  const float4 vn = make_float4(__shfl_sync(FULL, v.x, src), __shfl_sync(FULL, v.y, src), __shfl_sync(FULL, v.z, src), __shfl_sync(FULL, v.w, src));
  bool ok;
  if (half == 0) ok = v.x < 100.f && v.y * vn.y + v.z * vn.z + v.w * vn.w >= 0.9995f;
  else { const float dx = v.x - vn.x, dy = v.y - vn.y, dz = v.z - vn.z; ok = dx * dx + dy * dy + dz * dz <= 4e-6f; }
  const bool same = ((__ballot_sync(FULL, ok) >> (lane & ~7)) & 0xffu) == 0xffu;
  const float w = ((c & 2) ? wa : 1.f - wa) * ((c & 1) ? wb : 1.f - wb);
  float4 a = make_float4(w * v.x, w * v.y, w * v.z, w * v.w);
#pragma unroll
  for (int o = 2; o <= 4; o <<= 1) {
    a.x += __shfl_xor_sync(FULL, a.x, o); a.y += __shfl_xor_sync(FULL, a.y, o);
    a.z += __shfl_xor_sync(FULL, a.z, o); a.w += __shfl_xor_sync(FULL, a.w, o);
  }
  if (!same) a = vn;
  // the leader lane of each pair (lane % 8 == 0: half 0) collects the witness point from its neighbour (half 1)
  const float xl0 = __shfl_down_sync(FULL, a.x, 1), xl1 = __shfl_down_sync(FULL, a.y, 1), xl2 = __shfl_down_sync(FULL, a.z, 1);
  const float far = __shfl_sync(FULL, vn.x, (lane & ~7) | (near << 1));   // nearest node's distance (sentinel test)
  float dist = 0.f, nw[3] = {0.f, 0.f, 0.f}, xw[3] = {0.f, 0.f, 0.f};
  bool valid = false;
  if (live && (lane & 7) == 0 && far < 100.f) {
    float nl[3] = {a.y, a.z, a.w};
    if (same) { const float r = rsqrtf(dot3(nl, nl)); nl[0] *= r; nl[1] *= r; nl[2] *= r; }
    const float xl[3] = {xl0, xl1, xl2};
    dist = a.x - 2.f * P(s, MP_HULL_MARGIN);
    if (!(dist > P(s, MP_SELF_NEAR))) {
      valid = true;
      if (d.la >= 0) {
        mat_vec(nw, s.R[d.la], nl);
        mat_vec(xw, s.R[d.la], xl);
        xw[0] += s.p[d.la][0]; xw[1] += s.p[d.la][1]; xw[2] += s.p[d.la][2];
      } else {  // right_link1 is rigid with the base: identity rotation at the base position
        nw[0] = nl[0]; nw[1] = nl[1]; nw[2] = nl[2];
        xw[0] = xl[0] + P(s, MP_BASE_PX); xw[1] = xl[1] + P(s, MP_BASE_PY); xw[2] = xl[2] + P(s, MP_BASE_PZ);
      }
    }
  }
  const unsigned m = __ballot_sync(FULL, valid);
  push_contacts(s, ep, m, lane, d.lb, 0, xw, nw, dist, d.mu, live ? d.la : -2);
}

__device__ __noinline__ void find_contacts(Smem& s, const EnvParams& ep, const float* __restrict__ model_g,
                                           float block_margin, bool with_self, int lane) {
  if (lane == 0) { s.nc = 0; s.na = 0; }
  __syncwarp();
  // the arm's self-contacts come first (Bullet keeps its manifolds in creation order: the robot is loaded before the
  // table and the block), which also keeps them inside the lane budget
  if (with_self && ep.sc_np > 0 && P(s, MP_SELF_COLLISION) > 0.5f) self_contacts(s, ep, lane);
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
  }
  __syncwarp();
  const float up[3] = {0.f, 0.f, 1.f};
  {  // block vertices vs table plane, up to 4 deepest
    const float d = bvx[2] - tz;
    unsigned m = select_deepest(d, lane < 8, P(s, MP_TABLE_MARGIN), 4, lane);
    push_contacts(s, ep, m, lane, -1, 1, bvx, up, d, ep.bmu * P(s, MP_MU_TABLE));
  }
  const int ns = (int)P(s, MP_N_SHAPES);
  const float* shapes = s.model + (int)P(s, MP_SHAPES_OFF);   // staged in shared memory with the joint tree
  const float* pool = s.model + (int)P(s, MP_POOL_OFF);
  const float brad = sqrtf(ep.bh[0] * ep.bh[0] + ep.bh[1] * ep.bh[1] + ep.bh[2] * ep.bh[2]);
  for (int si = 0; si < ns; ++si) {
    const float* sh = shapes + si * BMI_SHAPE_STRIDE;
    const int l = (int)(*(sh + MS_LINK)), nv = (int)(*(sh + MS_NVERTS)), np = (int)(*(sh + MS_NPLANES));
    const float* verts = pool + (int)(*(sh + MS_VERT_OFF));
    const float* planes = pool + (int)(*(sh + MS_PLANE_OFF));
    const float smu = (*(sh + MS_MU));
    // broadphase on the link's bounding sphere: against the table plane and against the block's bounding sphere
    // (exact: the sphere contains every hull vertex, 1e-4 m of slack covers the fp32 rounding of the transform)
    float lc[3] = {(*(sh + MS_SPHERE_C)), (*(sh + MS_SPHERE_C + 1)), (*(sh + MS_SPHERE_C + 2))}, cw[3];
    mat_vec(cw, s.R[l], lc);
    cw[0] += s.p[l][0]; cw[1] += s.p[l][1]; cw[2] += s.p[l][2];
    const float srad = (*(sh + MS_SPHERE_R));
    const bool near_table = cw[2] - srad <= tz + P(s, MP_CONTACT_MARGIN) + 1e-4f;   // uniform
    float dd[3] = {s.bp[0] - cw[0], s.bp[1] - cw[1], s.bp[2] - cw[2]};
    const bool near_block = !(sqrtf(dot3(dd, dd)) > srad + brad + block_margin);       // uniform
    if (!near_table && !near_block) continue;
    float wv[3] = {0, 0, 0};
    if (lane < nv) {
      float lv[3] = {(*(verts + 3 * lane)), (*(verts + 3 * lane + 1)), (*(verts + 3 * lane + 2))};
      mat_vec(wv, s.R[l], lv);
      wv[0] += s.p[l][0]; wv[1] += s.p[l][1]; wv[2] += s.p[l][2];
    }
    if (near_table) {  // hull vertices vs table plane, up to 2 deepest
      const float d = wv[2] - tz;
      unsigned m = select_deepest(d, lane < nv, P(s, MP_CONTACT_MARGIN), 2, lane);
      push_contacts(s, ep, m, lane, l, 0, wv, up, d, smu * P(s, MP_MU_TABLE));
    }
    if (!near_block) continue;
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
        const float4 pl = reinterpret_cast<const float4*>(planes)[pi];
        const float sd = pl.x * xl[0] + pl.y * xl[1] + pl.z * xl[2] + pl.w;
        if (sd > best) { best = sd; bpi = pi; }
      }
      const float4 pl = reinterpret_cast<const float4*>(planes)[bpi];
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
    push_contacts(s, ep, m, lane, l, 1, x, nrm, d, ep.bmu * smu);
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
// substep_dynamics: M, bias, M^-1, predicted velocities, contacts.  substep_solve: constraint rows, coupling table,
// projected Gauss-Seidel, integration.  All on the env's own warp.
__device__ __noinline__ void substep_dynamics(Smem& s, const EnvParams& ep, const float* __restrict__ model_g, int lane) {
  const float dt = P(s, MP_DT), gz = P(s, MP_GRAVITY), kl = P(s, MP_LIN_DAMP), ka = P(s, MP_ANG_DAMP);
  fk(s, s.q, lane);
  PROF_T0();
  joint_space_dynamics(s, gz, kl, ka, lane);
  PROF_ADD(s.prof_e, 2);
  chol9(s.dyn.L, lane);
  // M^-1 columns (lanes 0..8) and unconstrained acceleration (lane 9)
  if (lane <= NL) {
    float b[NL], x[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) b[i] = lane < NL ? (i == lane ? 1.f : 0.f) : (-LK(s, i)[ML_DAMPING] * s.qd[i] - s.bias[i]);
    chol9_solve(s.dyn.L, b, x);
    if (lane < NL) {
#pragma unroll
      for (int i = 0; i < NL; ++i) MINV(s, i, lane) = x[i];
    } else {
#pragma unroll
      for (int i = 0; i < NL; ++i) s.acc[i] = x[i];
    }
  }
  __syncwarp();
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
  PROF_ADD(s.prof_e, 3);
  find_contacts(s, ep, model_g, P(s, MP_BLOCK_MARGIN), true, lane);
  PROF_ADD(s.prof_e, 4);
  if (lane < 9) {  // world-frame inverse inertia of the block: R diag(1/I) R^T
    const int r = lane / 3, cc = lane % 3;
    s.Ibinv[lane] = s.Rb[3 * r] * s.Rb[3 * cc] / ep.binertia[0] + s.Rb[3 * r + 1] * s.Rb[3 * cc + 1] / ep.binertia[1] +
                    s.Rb[3 * r + 2] * s.Rb[3 * cc + 2] / ep.binertia[2];
  }
  __syncwarp();
}

// Shared-memory loads the compiler must not hoist out of the iteration loop (the coupling coefficients are loop
// invariant; hoisting them into registers spills to LOCAL memory).
__device__ __forceinline__ float lds_f(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// MUFU.RSQ without the denormal-range rescaling of rsqrtf (the argument is a squared impulse norm well inside the
// normal range whenever the result is used)
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// ---- constraint rows + projected Gauss-Seidel + integration ------------------------------------------------------
// Same rows, row order, clamps and residual exit as pgs_solve in oracle/bmi_physics_oracle.c (Bullet's order: motors,
// violated joint limits, contact normals, friction cones), but iterated in CONSTRAINT space so that the sequential
// chain per row is   local clamp -> one warp shuffle -> one FMA per lane:
//   * lane j < 9 owns joint j: its variable is the joint's velocity delta dv_j; its rows are the motor of joint j and
//     (if violated) the joint's limit;
//   * lanes 9..14 own the block's six velocity deltas (no rows of their own);
//   * lane LANE_CT + c owns contact c: its variables are v_k = J_k . dv for its normal and two friction rows.
// A row update computes its impulse change d from the owner's registers only, broadcasts d with one shuffle, and every
// lane adds (coupling coefficient) x d to its variables.  The coefficients are M^-1 (joint -> joint), M^-1 Ja^T
// (joint <-> contact row), the block's response (contact row -> block velocity) and the Delassus entries J_y M^-1 J_x^T
// (contact row -> contact row), precomputed once per sub-step into s.Minv / s.S.  In exact arithmetic the iterates are
// those of the velocity-space solver.
__device__ __noinline__ void substep_solve(Smem& s, const EnvParams& ep, int lane) {
  PROF_T0();
  const float dt = P(s, MP_DT);
  const int nc = (int)__reduce_max_sync(FULL, (unsigned)s.nc), nrows = 3 * nc;   // REDUX: provably warp-uniform
  const bool has_arm = __ballot_sync(FULL, s.na > 0) != 0u;   // warp-uniform by construction (ballot)
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;   // solver variables of this lane
  float lam0 = 0.f, lam1 = 0.f, lam2 = 0.f, dl0 = 0.f, dl1 = 0.f, dl2 = 0.f;
  float rhs0 = 0.f, rhs1 = 0.f, rhs2 = 0.f, inv0 = 0.f, inv1 = 0.f, inv2 = 0.f, dg0 = 0.f, dg1 = 0.f, dg2 = 0.f;
  float mu = 0.f, lsgn = 1.f;
  const float max_imp = P(s, MP_MOTOR_FORCE) * dt, lim_hi = P(s, MP_JOINT_LIMIT_IMPULSE);
  bool viol = false;
  // ---- joint lanes: motor row (always) and limit row (when violated); at most one side can be violated ------------
  if (lane < NL) {
    const float w = MINV(s, lane, lane);
    const float target = P(s, MP_MOTOR_KP) * (s.qt[lane] - s.q[lane]) / dt + (1.f - P(s, MP_MOTOR_KD)) * s.qd[lane];
    inv0 = 1.f / w; dg0 = w;
    rhs0 = (target - s.u[lane]) / w;
    const float pen_lo = s.q[lane] - LK(s, lane)[ML_LO], pen_hi = LK(s, lane)[ML_HI] - s.q[lane];
    const bool vlo = !(pen_lo > 0.f), vhi = !(pen_hi > 0.f);
    viol = vlo || vhi;
    if (viol) {
      const float pen = vlo ? pen_lo : pen_hi;
      lsgn = vlo ? 1.f : -1.f;
      rhs1 = (-pen * P(s, MP_ERP_JOINT) / dt - lsgn * s.u[lane]) / w;
      inv1 = lsgn / w;   // d = rhs1 - (sgn dv_j) / w
      dg1 = w;
    }
  }
  unsigned limit_mask = __ballot_sync(FULL, viol);
  // ---- contact rows: lane = row x = 3 c + k (k = 0 normal, 1 / 2 the friction directions) --------------------------
  float Jb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int my_as = -1;  // arm-Jacobian scratch row of this lane's contact row, or -1
  if (lane < nrows) {
    const int ci = lane / 3, kind = lane - 3 * ci;
    float n[3] = {s.cn[ci][0], s.cn[ci][1], s.cn[ci][2]}, dir[3], t1[3], t2[3];
    plane_space(n, t1, t2);
#pragma unroll
    for (int a = 0; a < 3; ++a) dir[a] = kind == 0 ? n[a] : (kind == 1 ? t1[a] : t2[a]);
    const float x[3] = {s.cx[ci][0], s.cx[ci][1], s.cx[ci][2]};
    const int info = s.cinfo[ci];
    const int link = c_link(info), hasb = c_hasb(info), linkA = c_linkA(info);
    float* Srow = s.S + lane * SS;
    float diag = 0.f, rel = 0.f;
    if (link >= 0) {  // arm part, compact runtime loops through this row's scratch slot
      const float sgn = hasb ? -1.f : 1.f;
      my_as = c_arm(info) * 3 + kind;
      float* Jr = s.rows.Ja[my_as];
      float diag_sum = 0.f;
      // Self-contact (linkA > -2): body 1 = linkA at the witness x, body 2 = link at x2 = x - n * (core distance).
      // Two passes over the same scratch row: first J_A + J_B (only its quadratic form is needed), then the row's real
      // Jacobian J_A - J_B.  Bullet's diagonal for two links of ONE multibody is J_A M^-1 J_A^T + J_B M^-1 J_B^T (no
      // cross term, btMultiBodyConstraintSolver::setupMultiBodyContactConstraint) = the mean of the two quadratic forms.
      const bool self = linkA > -2;
      const float dcore = s.cdist[ci] + 2.f * P(s, MP_HULL_MARGIN);
      const float x2[3] = {x[0] - n[0] * dcore, x[1] - n[1] * dcore, x[2] - n[2] * dcore};
#pragma unroll 1
      for (int pass = self ? 0 : 1; pass < 2; ++pass) {
        for (int j = 0; j < NL; ++j) Jr[j] = 0.f;
        const float sb = self ? (pass == 0 ? 1.f : -1.f) : sgn;   // sign of body 2's part
#pragma unroll 1
        for (int j = link; j >= 0; j = parent_of(j)) {  // joints on the path base -> link
          const float* xx = self ? x2 : x;
          float r[3] = {xx[0] - s.p[j][0], xx[1] - s.p[j][1], xx[2] - s.p[j][2]}, cr[3];
          cross3(cr, s.z[j], r);
          Jr[j] = sb * dot3(dir, cr);
        }
        if (self) {
#pragma unroll 1
          for (int j = linkA; j >= 0; j = parent_of(j)) {
            float r[3] = {x[0] - s.p[j][0], x[1] - s.p[j][1], x[2] - s.p[j][2]}, cr[3];
            cross3(cr, s.z[j], r);
            Jr[j] += dot3(dir, cr);
          }
        }
        diag = 0.f; rel = 0.f;
#pragma unroll 1
        for (int i = 0; i < NL; ++i) {
          float acc = 0.f;
          for (int j = 0; j < NL; ++j) acc += MINV(s, i, j) * Jr[j];
          Srow[i] = acc;
          diag += Jr[i] * acc;
          rel += Jr[i] * s.u[i];
        }
        if (pass == 0) diag_sum = diag;
      }
      if (self && P(s, MP_SELF_SPLIT_DIAG) > 0.5f) diag = 0.5f * (diag + diag_sum);
    } else {
#pragma unroll
      for (int i = 0; i < NL; ++i) Srow[i] = 0.f;
    }
    float Wb[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (hasb) {
      float r[3] = {x[0] - s.bp[0], x[1] - s.bp[1], x[2] - s.bp[2]}, t[3];
      cross3(t, r, dir);
#pragma unroll
      for (int a = 0; a < 3; ++a) { Jb[a] = dir[a]; Jb[3 + a] = t[a]; Wb[a] = dir[a] / ep.bmass; }
      mat_vec(Wb + 3, s.Ibinv, Jb + 3);
#pragma unroll
      for (int a = 0; a < 6; ++a) { diag += Jb[a] * Wb[a]; rel += Jb[a] * s.u[9 + a]; }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) Srow[SCOL_BLK + a] = Wb[a];
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
    s.rows.sc[lane] = make_float4(invd, rhs, diag, s.cmu[ci]);
  }
  if (lane < 12) s.zero9[lane] = 0.f;
  __syncwarp();
  // ---- Delassus entries: lane y fills column SCOL_CT + y of every row x:  J_y . (M^-1 J_x^T) ------------------------
  {
    const unsigned arm_rows = __ballot_sync(FULL, my_as >= 0);   // bit x: row x has an arm part
    float Ja[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) Ja[j] = my_as >= 0 ? s.rows.Ja[my_as][j] : 0.f;
    const bool mine_arm = my_as >= 0;
#pragma unroll 1
    for (int x = 0; x < nrows; ++x) {
      const float* W = s.S + x * SS;
      float a = Jb[0] * W[SCOL_BLK] + Jb[1] * W[SCOL_BLK + 1] + Jb[2] * W[SCOL_BLK + 2] + Jb[3] * W[SCOL_BLK + 3] +
                Jb[4] * W[SCOL_BLK + 4] + Jb[5] * W[SCOL_BLK + 5];
      if ((arm_rows >> x) & 1u) {   // uniform
        float b = 0.f;
#pragma unroll
        for (int j = 0; j < NL; ++j) b = fmaf(Ja[j], W[j], b);
        a += mine_arm ? b : 0.f;
      }
      if (lane < nrows) s.S[x * SS + SCOL_CT + lane] = a;
    }
  }
  __syncwarp();
  // ---- owner lanes pick up their rows' scalars; per-lane coefficient addresses --------------------------------------
  // joint events (source joint j):   coefficient k of this lane at ja + 4 j + k * SS * 4
  // contact events (source row x):   coefficient k of this lane at ca + x * SS * 4 + 4 k
  // Only contact lanes own three variables; for the others k = 1, 2 read in-bounds don't-care words into v1 / v2,
  // which they never use (that keeps ONE address register per event type and immediate offsets for k).
  const unsigned s_base = (unsigned)__cvta_generic_to_shared(s.S);
  unsigned ja = (unsigned)__cvta_generic_to_shared(s.zero9), ca = s_base;   // default: the zero row / column 0 (idle lanes)
  const int myc = lane - LANE_CT;
  if (lane < NL) {
    ja = (unsigned)__cvta_generic_to_shared(s.Minv) + (unsigned)lane * (MS * 4u);
    ca = s_base + (unsigned)lane * 4u;
  } else if (lane < LANE_BLK + 6) {
    ca = s_base + (unsigned)lane * 4u;   // columns 9..14
  } else if (myc >= 0 && myc < nc) {
    const float4 a = s.rows.sc[3 * myc], b = s.rows.sc[3 * myc + 1], c = s.rows.sc[3 * myc + 2];
    inv0 = a.x; rhs0 = a.y; dg0 = a.z; mu = a.w;
    inv1 = b.x; rhs1 = b.y; dg1 = b.z;
    inv2 = c.x; rhs2 = c.y; dg2 = c.z;
    ja = s_base + (unsigned)(3 * myc) * (SS * 4u);
    ca = s_base + (unsigned)(SCOL_CT + 3 * myc) * 4u;
  }
  int max_it = (int)P(s, MP_SOLVER_ITERS);
  const float thresh = P(s, MP_RESIDUAL_THRESH);
  const bool alternate = __reduce_max_sync(FULL, P(s, MP_SWEEP_ALTERNATE) > 0.5f ? 1u : 0u) != 0u;
  // ---- iteration schedule ------------------------------------------------------------------------------------------
  // (a) Compression (oracle: substep(), MP_PGS_COMPRESS).  The arm's self-contact rows carry Bullet's same-multibody
  //     diagonal, 1e3 .. 3e3 times the true one: over the 150 iterations their impulses grow as an almost linear ramp
  //     (the loop never converges: that softness IS the reference's wrist) and every other row tracks the ramp with one
  //     iteration of lag.  n_c iterations with K-fold steps on those rows followed by `tail` plain ones land on the same
  //     point: K n_c + tail = MP_SOLVER_ITERS equivalent iterations.
  // (b) Block island: while no arm link touches the block, the block-table rows form a separate island that converges
  //     in 20-50 iterations; once its residual is below the threshold its rows are skipped (Bullet keeps sweeping them
  //     with impulse changes below 3e-4 x the row diagonal).  The threshold is 100x below Bullet's for a block at rest
  //     and 10 000x below while the block moves (there a truncated solve biases the friction every sub-step).
  unsigned self_mask = 0u, blk_mask = 0u, armblk_mask = 0u;   // per contact slot
  {
    const int info = lane < nc ? s.cinfo[lane] : 0;
    self_mask = __ballot_sync(FULL, lane < nc && c_linkA(info) > -2);
    blk_mask = __ballot_sync(FULL, lane < nc && c_hasb(info) && c_link(info) < 0);
    armblk_mask = __ballot_sync(FULL, lane < nc && c_hasb(info) && c_link(info) >= 0);
  }
  const float Kc = P(s, MP_PGS_COMPRESS);
  int n_c = 0;
  if (Kc > 1.f && self_mask) {
    n_c = (int)((float)(max_it - (int)P(s, MP_PGS_TAIL)) / Kc);
    max_it = n_c + (max_it - (int)((float)n_c * Kc));
  }
  max_it = (int)__reduce_max_sync(FULL, (unsigned)max_it);   // provably warp-uniform loop bounds
  n_c = (int)__reduce_max_sync(FULL, (unsigned)n_c);
  const bool own_ct = myc >= 0 && myc < nc;
  const float kself = (own_ct && ((self_mask >> myc) & 1u)) ? Kc : 1.f;
  const bool blk_lane = own_ct && ((blk_mask >> myc) & 1u);
  const bool blk_island = armblk_mask == 0u && blk_mask != 0u;
  unsigned skip_mask = 0u;      // contact slots whose rows are no longer swept
  // freeze threshold of the island: tangential + angular speed of the block before the solve (warp-uniform)
  const float blk_speed2 = s.u[9] * s.u[9] + s.u[10] * s.u[10] +
                           9e-4f * (s.u[12] * s.u[12] + s.u[13] * s.u[13] + s.u[14] * s.u[14]);   // 3 cm lever arm
  const float blk_freeze = (blk_speed2 > 1e-6f ? BMI_BLK_FREEZE_MOVING : BMI_BLK_FREEZE_REST) * thresh;
  // (c) Warm start of the block island (compile-time option, OFF): starting its rows from the previous sub-step's impulses
  //     ends the island after 1-3 sweeps instead of ~45, but four corner contacts are statically indeterminate and PGS then
  //     converges to a different point than from Bullet's cold start (the multibody solver does not warm-start): a sliding
  //     or spinning block decayed measurably differently from the oracle.  Cold start costs ~1 % of the rollout.
  if (blk_island && BMI_BLK_WARMSTART) {
    const unsigned ids = s.blk_ids;
    if (blk_lane) {
      const int vid = c_vid(s.cinfo[myc]);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if ((int)((ids >> (8 * k)) & 0xffu) == vid) { lam0 = s.blk_lam[3 * k]; lam1 = s.blk_lam[3 * k + 1]; lam2 = s.blk_lam[3 * k + 2]; }
    }
    if (__ballot_sync(FULL, blk_lane && lam0 != 0.f)) {
      for (unsigned m = blk_mask; m; m &= m - 1u) {
        const int c = __ffs(m) - 1;
        const float l0 = __shfl_sync(FULL, lam0, LANE_CT + c), l1 = __shfl_sync(FULL, lam1, LANE_CT + c), l2 = __shfl_sync(FULL, lam2, LANE_CT + c);
        const unsigned a = ca + (unsigned)c * (3u * SS * 4u);
        v0 = fmaf(lds_f(a), l0, v0); v1 = fmaf(lds_f(a + 4u), l0, v1); v2 = fmaf(lds_f(a + 8u), l0, v2);
        v0 = fmaf(lds_f(a + SS * 4u), l1, v0); v1 = fmaf(lds_f(a + SS * 4u + 4u), l1, v1); v2 = fmaf(lds_f(a + SS * 4u + 8u), l1, v2);
        v0 = fmaf(lds_f(a + 2u * SS * 4u), l2, v0); v1 = fmaf(lds_f(a + 2u * SS * 4u + 4u), l2, v1); v2 = fmaf(lds_f(a + 2u * SS * 4u + 8u), l2, v2);
      }
    }
  }
  int it = 0;
  PROF_ADD(s.prof_e, 5);
  // one joint-source update: every lane adds its coefficient x d (contact lanes: all three rows, only when some
  // contact touches the arm — otherwise their coefficients are zero)
#define BMI_JOINT_EVENT(ARM, joff, dj)                                                    \
  do {                                                                                    \
    const float c0_ = lds_f(ja + (joff));                                                 \
    if (ARM) {                                                                            \
      const float c1_ = lds_f(ja + (joff) + SS * 4u), c2_ = lds_f(ja + (joff) + 2u * SS * 4u); \
      v1 = fmaf(c1_, (dj), v1); v2 = fmaf(c2_, (dj), v2);                                 \
    }                                                                                     \
    v0 = fmaf(c0_, (dj), v0);                                                             \
  } while (0)
  // motors (J = e_j, bounds +-max_imp), then violated joint limits (J = +-e_j, bounds [0, hi]).  Bullet sweeps these
  // non-contact rows BACKWARDS on even iterations (btMultiBodyConstraintSolver::solveSingleIteration: index =
  // iteration & 1 ? j : size - 1 - j): limits from the highest joint down, then motors 8 .. 0.
#define BMI_LIMIT_ROWS(ARM, FWD, TRACK)                                                        \
  do {                                                                                    \
    for (unsigned m = limit_mask; m;) {                                                   \
      const int j = (FWD) ? __ffs(m) - 1 : 31 - __clz(m);                                 \
      m &= ~(1u << j);                                                                    \
      const float cand = fminf(fmaxf(lam1 + fmaf(-v0, inv1, rhs1), 0.f), lim_hi);         \
      const float d = cand - lam1;                                                        \
      const float dj = __shfl_sync(FULL, d * lsgn, j);                                    \
      if (lane == j) { lam1 = cand; if (TRACK) dl1 = d; }                                 \
      BMI_JOINT_EVENT(ARM, 4u * (unsigned)j, dj);                                         \
    }                                                                                     \
  } while (0)
#define BMI_JOINT_ROWS(ARM, TRACK)                                                             \
  do {                                                                                    \
    const bool fwd_ = !alternate || (it & 1);                                             \
    if (!fwd_ && limit_mask) BMI_LIMIT_ROWS(ARM, false, TRACK);                               \
    int j = fwd_ ? 0 : NL - 1;                                                            \
    const int jstep_ = fwd_ ? 1 : -1;                                                     \
    /* clamp(lam + x, -m, m) - lam == clamp(x, -m - lam, m - lam): a joint lane's motor impulse only changes at its own */ \
    /* row, so the two bounds are per-sweep constants and a row is FFMA + 2 FMNMX instead of FFMA + 2 FADD + 2 FMNMX */   \
    const float mlo_ = -max_imp - lam0, mhi_ = max_imp - lam0;                            \
    _Pragma("unroll (kMotorUnroll)")                                                      \
    for (int jj = 0; jj < NL; ++jj, j += jstep_) {                                        \
      const float d = fminf(fmaxf(fmaf(-v0, inv0, rhs0), mlo_), mhi_);                    \
      const float dj = __shfl_sync(FULL, d, j);                                           \
      if (lane == j) { lam0 += d; if (TRACK) dl0 = d; }                                   \
      BMI_JOINT_EVENT(ARM, 4u * (unsigned)j, dj);                                         \
    }                                                                                     \
    if (fwd_ && limit_mask) BMI_LIMIT_ROWS(ARM, true, TRACK);                                 \
  } while (0)
  unsigned active = nc >= 32 ? 0xffffffffu : ((1u << nc) - 1u);   // contact slots still swept (bit c)
  float kf = n_c > 0 ? kself : 1.f;
  // The loop (macro: three instances so that the common case carries no residual bookkeeping at all -- TRACKJ: joint rows
  // record their last impulse change (models without under-relaxed rows: Bullet's residual exit); TRACKC: contact rows do
  // (needed until the block island is frozen)):
  //   joint rows (motors, violated limits; swept backwards on even iterations)
  //   contact normals, then friction cones, of the slots still swept (rolled loops over a bit mask: the body must stay
  //     inside the L0 instruction cache)
  //   residual: with under-relaxed rows in the system the loop never meets Bullet's threshold (the oracle runs all
  //     iterations too), so the global residual is only evaluated for models without them (TRACKJ); the block island
  //     is frozen 100x below Bullet's threshold (3e-5 m/s of row velocity change).
#define BMI_PGS_LOOP(TRACKJ, TRACKC) \
_Pragma("unroll 1") \
  while (true) { \
    if (it == n_c) kf = 1.f; \
    BMI_JOINT_ROWS(true, TRACKJ); \
_Pragma("unroll 1") \
    for (unsigned m = active; m; m &= m - 1u) { \
      const int c = __ffs(m) - 1; \
      const unsigned a = ca + (unsigned)c * (3u * SS * 4u); \
      const float c0 = lds_f(a), c1 = lds_f(a + 4u), c2 = lds_f(a + 8u); \
      const float d = fmaxf(kf * fmaf(-v0, inv0, rhs0), -lam0);   /* max(lam + k x, 0) - lam */ \
      const float dc = __shfl_sync(FULL, d, LANE_CT + c); \
      if (myc == c) { lam0 += d; if (TRACKC) dl0 = d; } \
      v0 = fmaf(c0, dc, v0); v1 = fmaf(c1, dc, v1); v2 = fmaf(c2, dc, v2); \
    } \
_Pragma("unroll 1") \
    for (unsigned m = active; m; m &= m - 1u) { \
      const int c = __ffs(m) - 1; \
      const unsigned a = ca + (unsigned)c * (3u * SS * 4u) + SS * 4u; \
      const float p0 = lds_f(a), p1 = lds_f(a + 4u), p2 = lds_f(a + 8u); \
      const float q0 = lds_f(a + SS * 4u), q1 = lds_f(a + SS * 4u + 4u), q2 = lds_f(a + SS * 4u + 8u); \
      const float lim = mu * lam0; \
      float sa = fmaf(kf, fmaf(-v1, inv1, rhs1), lam1), sb = fmaf(kf, fmaf(-v2, inv2, rhs2), lam2); \
      const float n2 = sa * sa + sb * sb; \
      const float sc = n2 > lim * lim ? lim * rsqrt_fast(n2) : 1.f; \
      sa *= sc; sb *= sc; \
      const float da = sa - lam1, db = sb - lam2; \
      const float dac = __shfl_sync(FULL, da, LANE_CT + c), dbc = __shfl_sync(FULL, db, LANE_CT + c); \
      if (myc == c) { lam1 = sa; lam2 = sb; if (TRACKC) { dl1 = da; dl2 = db; } } \
      v0 = fmaf(p0, dac, v0); v1 = fmaf(p1, dac, v1); v2 = fmaf(p2, dac, v2); \
      v0 = fmaf(q0, dbc, v0); v1 = fmaf(q1, dbc, v1); v2 = fmaf(q2, dbc, v2); \
    } \
    ++it; \
    if (it >= max_it) break; \
    if (!(TRACKC)) continue; \
    const bool want_blk = blk_island && skip_mask == 0u; \
    if (!(TRACKJ) && !want_blk) break; \
    if (self_mask == 0u || want_blk) { \
      const float r0 = dl0 * dg0, r1 = dl1 * dg1, r2 = dl2 * dg2; \
      const float rl = fmaxf(r0 * r0, fmaxf(r1 * r1, r2 * r2)); \
      if (self_mask == 0u) { \
        const float resid = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(rl))); \
        if (resid <= thresh) break; \
      } \
      if (want_blk) { \
        const float rb = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(blk_lane ? rl : 0.f))); \
        if (BMI_BLK_FREEZE && rb <= blk_freeze) { skip_mask = blk_mask; active &= ~blk_mask; } \
      } \
    } \
  }
  if (self_mask != 0u) {
    if (blk_island) { BMI_PGS_LOOP(false, true); }      // until the block island is frozen (2 sweeps on average) ...
    if (it < max_it) { BMI_PGS_LOOP(false, false); }     // ... then without any residual bookkeeping
  } else {
    BMI_PGS_LOOP(true, true);
  }
#undef BMI_PGS_LOOP
#undef BMI_JOINT_ROWS
#undef BMI_LIMIT_ROWS
#undef BMI_JOINT_EVENT
  PROF_CNT(s.prof_e, 7, it);
  __syncwarp();   // the warm-start reads above are ordered before the stores below (shuffles are not memory barriers)
  if (BMI_BLK_WARMSTART) {  // impulses of the block island for the next sub-step's warm start
    const int slot = __popc(blk_mask & ((1u << (myc & 31)) - 1u));   // rank of this lane's contact among the block-table contacts
    if (blk_island && blk_lane && slot < 4) { s.blk_lam[3 * slot] = lam0; s.blk_lam[3 * slot + 1] = lam1; s.blk_lam[3 * slot + 2] = lam2; }
    unsigned idb = blk_island && blk_lane && slot < 4 ? ((unsigned)c_vid(s.cinfo[myc]) << (8 * slot)) | ~(0xffu << (8 * slot)) : 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) idb &= __shfl_xor_sync(FULL, idb, o);
    if (lane == 0) s.blk_ids = idb;
  }
  // ---- integrate --------------------------------------------------------------------------------------------------
  const float unew = lane < NU ? s.u[lane] + v0 : 0.f;
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
  PROF_ADD(s.prof_e, 6);
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
  float model_s[STAGED_FULL];
  unsigned long long mbar;
  int pad_[2];
  Smem sw[ENVW];
};
static_assert(sizeof(Smem) % 16 == 0, "Smem slots must be 16-byte aligned");
static_assert((offsetof(Smem, S) - offsetof(Smem, Minv)) / 4 % 32 == 3 && (offsetof(Smem, zero9) - offsetof(Smem, Minv)) / 4 % 32 == 16,
              "bank layout of the solver tables (see Smem)");
static_assert(offsetof(BlockSmem, sw) % 16 == 0, "Smem slots must be 16-byte aligned");
static_assert((sizeof(BlockSmem) + 1024) * BLOCKS_PER_SM <= 228 * 1024, "BLOCKS_PER_SM blocks (+1 KB reserved each) must fit the SM's 228 KB");
template <int N> struct PrintSize;
#ifdef BMI_PRINT_SIZES
PrintSize<sizeof(Smem)> print_smem_size;
#endif
extern __shared__ __align__(16) unsigned char bmi_dyn_smem[];

// Block prologue: stage the model (one TMA bulk copy, shared by the block's envs), point every env slot at it.
__device__ __forceinline__ void block_begin(BlockSmem& bs, const float* __restrict__ model_g, unsigned model_bytes) {
  if (threadIdx.x < ENVW) bs.sw[threadIdx.x].model = bs.model_s;
  stage_model(bs.model_s, &bs.mbar, model_g, threadIdx.x, model_bytes);  // mbarrier-init fence + __syncthreads inside
}

// clip, (pick: auto-grip), IK, motor set-points  (bmirobot_env_push_F.py:92-101) — one warp per env
__device__ __noinline__ void env_step_begin(Smem& s, const EnvParams& ep, const float* __restrict__ model_g,
                                            const float* a_in, int lane) {
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = fminf(fmaxf(a_in[i], -0.5f), 0.5f);
  if (ep.task == BMI_TASK_PUSH) a[3] = 0.f;  // bmirobot_env_push_F.py:94
  if (lane == 0) s.blk_ids = 0xffffffffu;    // every env-step starts the block island cold (step-wise == fused rollout)
  fk(s, s.q, lane);
  if (ep.task == BMI_TASK_PICK) {  // auto-grip (bmirobot_env_pickandplace_v2.py:94-95)
    find_contacts(s, ep, model_g, 1e-4f, false, lane);
    bool touch = false;
    for (int c = 0; c < s.nc; ++c) touch |= (c_hasb(s.cinfo[c]) && c_link(s.cinfo[c]) >= 0 && s.cdist[c] < 1e-4f);
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

// One env step of this warp's env: n_substeps x [dynamics + contacts | rows + solve + integrate].
__device__ __forceinline__ void env_step_warp(Smem& s, const EnvParams& ep, const float* __restrict__ model_g,
                                              const float* a_in, int lane, int e) {
  PROF_T0();
#ifdef BMI_PROF
  if (lane == 0) s.prof_e = e;
  __syncwarp();
#endif
  env_step_begin(s, ep, model_g, a_in, lane);
  PROF_ADD(e, 1);
  const int nsub = (int)P(s, MP_N_SUBSTEPS);
  for (int i = 0; i < nsub; ++i) {
    substep_dynamics(s, ep, model_g, lane);
    substep_solve(s, ep, lane);
    PROF_RESET();
  }
}

// ---- kernels ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WARPS, BLOCKS_PER_SM)
env_step_kernel(const float* __restrict__ model_g, unsigned model_bytes, EnvParams ep, int n_envs,
                float* __restrict__ state, const float* __restrict__ actions, float* __restrict__ obs,
                float* __restrict__ ag, float* __restrict__ reward, float* __restrict__ success) {
  BlockSmem& bs = *reinterpret_cast<BlockSmem*>(bmi_dyn_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  block_begin(bs, model_g, model_bytes);
  const int e = blockIdx.x * ENVW + warp;
  if (e >= n_envs) return;  // whole warp
  Smem& s = bs.sw[warp];
  load_state(s, state + (size_t)e * BMI_ENV_STATE_DIM, lane);
  float a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = actions[e * 4 + i];
  env_step_warp(s, ep, model_g, a, lane, e);
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
  __syncwarp();   // out may alias x
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

// actor parameters inside the transposed flat buffer (bmi_actor_transpose)
struct PolicyW {
  const float *Wt1, *b1, *Wt2, *b2, *Wt3, *b3, *Wt4, *b4;
};
__device__ __forceinline__ PolicyW policy_weights(const RolloutArgs& ra) {
  constexpr int Dx = BMI_OBS_DIM + BMI_GOAL_DIM, Da = BMI_ACT_DIM;
  PolicyW w;
  w.Wt1 = ra.actor_t;        w.b1 = w.Wt1 + Dx * HID;
  w.Wt2 = w.b1 + HID;        w.b2 = w.Wt2 + HID * HID;
  w.Wt3 = w.b2 + HID;        w.b3 = w.Wt3 + HID * HID;
  w.Wt4 = w.b3 + HID;        w.b4 = w.Wt4 + HID * Da;
  return w;
}

// One step of the rollout loop of ddpg_agent.learn() (ddpg_agent.py:111-141) for env e at time t: record obs / ag / g,
// policy, exploration noise, record the action, env step, new observation into s.obs.
__device__ __forceinline__ void rollout_step(Smem& s, const EnvParams& ep, const float* __restrict__ model_g,
                                             const RolloutArgs& ra, const PolicyW& pw, unsigned long long ctr0, int n_envs,
                                             int e, int t, int lane) {
  constexpr int Do = BMI_OBS_DIM, Dg = BMI_GOAL_DIM, Da = BMI_ACT_DIM, Dx = Do + Dg;
  const float *Wt1 = pw.Wt1, *b1 = pw.b1, *Wt2 = pw.Wt2, *b2 = pw.b2, *Wt3 = pw.Wt3, *b3 = pw.b3, *Wt4 = pw.Wt4, *b4 = pw.b4;
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
  policy_layer(Wt1, b1, s.pol.x, Dx, s.pol.h, lane);
  policy_layer(Wt2, b2, s.pol.h, HID, s.pol.h, lane);   // in place: every lane has read all inputs before any stores
  policy_layer(Wt3, b3, s.pol.h, HID, s.pol.h, lane);
  float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    const int k = lane * 8 + kk;
    const float hk = s.pol.h[k];
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
  // ---- env step -----------------------------------------------------------------------------------------------------
  PROF_ADD(e, 0);
  env_step_warp(s, ep, model_g, a, lane, e);
  observe(s, lane, s.obs, s.obs + Do);
}

__global__ void __launch_bounds__(32 * WARPS, BLOCKS_PER_SM)
rollout_kernel(const float* __restrict__ model_g, unsigned model_bytes, EnvParams ep, int n_envs,
               float* __restrict__ state, RolloutArgs ra) {
  BlockSmem& bs = *reinterpret_cast<BlockSmem*>(bmi_dyn_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  block_begin(bs, model_g, model_bytes);
  const int e = blockIdx.x * ENVW + warp;
  if (e >= n_envs) return;  // whole warp
  Smem& s = bs.sw[warp];
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
  constexpr int Do = BMI_OBS_DIM, Dg = BMI_GOAL_DIM;
  const PolicyW pw = policy_weights(ra);
  const unsigned long long ctr0 = ra.explore ? *ra.counter : 0ull;
  for (int t = 0; t < ra.T; ++t) rollout_step(s, ep, model_g, ra, pw, ctr0, n_envs, e, t, lane);
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
  stage_model(model_s, &mbar_s, model_g, threadIdx.x, STAGED * sizeof(float));
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
  unsigned model_bytes = 0;     // staged size: the blob rounded up to 16 bytes
  bool self_collision = false;  // MP_SELF_COLLISION of the blob
  float* sc_dev = nullptr;          // self-collision pair tables (bmi_env_set_selfcol)
  unsigned long long* drops_dev = nullptr;
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
  BMI_REQUIRE(n + LINK_SHIFT <= STAGED_FULL, "bmi_env_create: model blob (%lld floats) exceeds the shared-memory staging area (%d)",
              (long long)n, STAGED_FULL - LINK_SHIFT);
  BMI_REQUIRE((int)b[MP_SHAPES_OFF] >= BMI_MODEL_HDR + BMI_MAX_LINKS * BMI_LINK_STRIDE, "bmi_env_create: shapes overlap the link records");
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
  h->ep.sc_np = 0; h->ep.drops = nullptr;
  const float lx = 2 * h->ep.bh[0], ly = 2 * h->ep.bh[1], lz = 2 * h->ep.bh[2], mm = h->ep.bmass / 12.f;
  h->ep.binertia[0] = mm * (ly * ly + lz * lz);
  h->ep.binertia[1] = mm * (lx * lx + lz * lz);
  h->ep.binertia[2] = mm * (lx * lx + ly * ly);
  // device copy with the link records re-strided (LINK_STRIDE_DEV) and everything behind them shifted by LINK_SHIFT
  const int64_t nd = n + LINK_SHIFT;
  std::vector<float> dev((size_t)((nd + 3) / 4) * 4, 0.f);
  for (int i = 0; i < BMI_MODEL_HDR; ++i) dev[i] = b[i];
  for (int i = 0; i < NL; ++i)
    for (int k = 0; k < BMI_LINK_STRIDE; ++k) dev[BMI_MODEL_HDR + i * LINK_STRIDE_DEV + k] = b[BMI_MODEL_HDR + i * BMI_LINK_STRIDE + k];
  const int tail0 = BMI_MODEL_HDR + BMI_MAX_LINKS * BMI_LINK_STRIDE;
  for (int64_t i = tail0; i < n; ++i) dev[i + LINK_SHIFT] = b[i];
  dev[MP_SHAPES_OFF] += (float)LINK_SHIFT;
  dev[MP_POOL_OFF] += (float)LINK_SHIFT;
  dev[MP_TOTAL] = (float)nd;
  h->model_floats = nd;
  h->model_bytes = (unsigned)(dev.size() * sizeof(float));
  if (cudaMalloc(&h->model_dev, h->model_bytes) != cudaSuccess ||
      cudaMalloc(&h->state_dev, (size_t)n_envs * BMI_ENV_STATE_DIM * sizeof(float)) != cudaSuccess) {
    set_error("bmi_env_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    bmi_env_destroy(h);
    return BMI_ERR_CUDA;
  }
  BMI_CUDA_CHECK(cudaMemcpy(h->model_dev, dev.data(), h->model_bytes, cudaMemcpyHostToDevice));
  BMI_CUDA_CHECK(cudaMemset(h->state_dev, 0, (size_t)n_envs * BMI_ENV_STATE_DIM * sizeof(float)));
  if (cudaMalloc(&h->drops_dev, sizeof(unsigned long long)) == cudaSuccess) {
    BMI_CUDA_CHECK(cudaMemset(h->drops_dev, 0, sizeof(unsigned long long)));
    h->ep.drops = h->drops_dev;
  }
  h->self_collision = b[MP_SELF_COLLISION] > 0.5f;
  *out = h;
  return BMI_OK;
}

extern "C" int bmi_env_set_selfcol(bmi_env* h, const void* table, int64_t bytes) {
  BMI_REQUIRE(h && table, "bmi_env_set_selfcol: null pointer");
  BMI_REQUIRE(bytes >= (int64_t)((SC_HDR + SC_DESC) * sizeof(float)) && bytes % 16 == 0, "bmi_env_set_selfcol: bad table size");
  const float* t = (const float*)table;
  const int np = (int)t[SC_NPAIRS];
  BMI_REQUIRE(t[SC_MAGIC] == BMI_SC_MAGIC && (int64_t)t[SC_TOTAL] * 4 == bytes && np >= 1 && np <= BMI_SC_MAX_PAIRS,
              "bmi_env_set_selfcol: table magic / size / pair count mismatch");
  SCPair scp[BMI_SC_MAX_PAIRS] = {};
  for (int p = 0; p < np; ++p) {
    const float* d = t + SC_HDR + SC_DESC * p;
    SCPair& o = scp[p];
    o.la = (int)d[SC_LA]; o.lb = (int)d[SC_LB]; o.ja = (int)d[SC_JA]; o.jb = (int)d[SC_JB];
    o.na = (int)d[SC_NA]; o.nb = (int)d[SC_NB]; o.off = (unsigned)d[SC_OFF];
    o.a0 = d[SC_A0]; o.b0 = d[SC_B0]; o.inv_h = 1.f / d[SC_H]; o.mu = d[SC_MU];
    BMI_REQUIRE(o.la >= -1 && o.la < NL && o.lb >= 0 && o.lb < NL && o.ja >= 0 && o.ja < NL && o.jb >= 0 && o.jb < NL &&
                    o.na >= 2 && o.nb >= 2 && o.off % 4 == 0 && (int64_t)o.off + 8ll * o.na * o.nb <= bytes / 4,
                "bmi_env_set_selfcol: bad descriptor of pair %d", p);
  }
  if (h->sc_dev) cudaFree(h->sc_dev);
  h->sc_dev = nullptr;
  BMI_CUDA_CHECK(cudaMalloc(&h->sc_dev, (size_t)bytes));
  BMI_CUDA_CHECK(cudaMemcpy(h->sc_dev, table, (size_t)bytes, cudaMemcpyHostToDevice));
  const float* data = h->sc_dev;
  BMI_CUDA_CHECK(cudaMemcpyToSymbol(c_scp, scp, sizeof(scp)));
  BMI_CUDA_CHECK(cudaMemcpyToSymbol(c_sc_data, &data, sizeof(data)));
  h->ep.sc_np = np;
  return BMI_OK;
}

extern "C" int bmi_env_contact_drops(bmi_env* h, uint64_t* host_out, int32_t reset) {
  BMI_REQUIRE(h && h->drops_dev, "bmi_env_contact_drops: no counter");
  if (host_out) BMI_CUDA_CHECK(cudaMemcpy(host_out, h->drops_dev, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  if (reset) BMI_CUDA_CHECK(cudaMemset(h->drops_dev, 0, sizeof(uint64_t)));
  return BMI_OK;
}

// a model with self-collision switched on must not be stepped without its pair tables (no silent degradation)
static int require_tables(const bmi_env* h, const char* who) {
  BMI_REQUIRE(!h->self_collision || h->ep.sc_np > 0, "%s: the model enables self-collision but no pair tables are loaded (bmi_env_set_selfcol)", who);
  return BMI_OK;
}

extern "C" int bmi_env_destroy(bmi_env* h) {
  if (!h) return BMI_OK;
  if (h->model_dev) cudaFree(h->model_dev);
  if (h->state_dev) cudaFree(h->state_dev);
  if (h->sc_dev) cudaFree(h->sc_dev);
  if (h->drops_dev) cudaFree(h->drops_dev);
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
  if (int rc = require_tables(h, "bmi_env_step")) return rc;
  env_step_kernel<<<(h->n_envs + ENVW - 1) / ENVW, 32 * WARPS, sizeof(BlockSmem), as_stream(stream)>>>(h->model_dev, h->model_bytes, h->ep, h->n_envs, h->state_dev,
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

static int fill_rollout_args(bmi_env* h, const bmi_rollout_args* a, RolloutArgs& ra, const char* who) {
  BMI_REQUIRE(h && a, "%s: null pointer", who);
  BMI_REQUIRE(a->T > 0 && a->actor_t && a->o_mean && a->o_std && a->g_mean && a->g_std, "%s: missing policy inputs", who);
  BMI_REQUIRE(!a->explore || a->counter, "%s: exploration needs a Philox counter", who);
  ra.T = a->T; ra.explore = a->explore; ra.actor_t = a->actor_t;
  ra.o_mean = a->o_mean; ra.o_std = a->o_std; ra.g_mean = a->g_mean; ra.g_std = a->g_std;
  ra.clip_range = a->clip_range; ra.action_max = a->action_max; ra.noise_eps = a->noise_eps;
  ra.random_eps = a->random_eps; ra.late_clip = a->late_clip; ra.seed = a->seed; ra.counter = (const unsigned long long*)a->counter;
  ra.ep_obs = ra.ep_ag = ra.ep_g = ra.ep_act = nullptr;
  if (a->episodes) {
    const bmi_episodes* e = a->episodes;
    BMI_REQUIRE(e->dtype == BMI_F32 && e->T == a->T && e->n_episodes == h->n_envs && e->obs_dim == BMI_OBS_DIM &&
                    e->goal_dim == BMI_GOAL_DIM && e->act_dim == BMI_ACT_DIM,
                "%s: episodes must be float32 [n_envs][T(+1)][27|3|3|4]", who);
    ra.ep_obs = (float*)e->obs; ra.ep_ag = (float*)e->ag; ra.ep_g = (float*)e->g; ra.ep_act = (float*)e->actions;
  }
  ra.init = a->init; ra.obs = a->obs; ra.ag = a->ag; ra.g = a->g; ra.success = a->success;
  return BMI_OK;
}

extern "C" int bmi_env_rollout(bmi_env* h, const bmi_rollout_args* a, bmi_stream_t stream) {
  RolloutArgs ra;
  const int rc = fill_rollout_args(h, a, ra, "bmi_env_rollout");
  if (rc != BMI_OK) return rc;
  if (int rc2 = require_tables(h, "bmi_env_rollout")) return rc2;
  cudaStream_t st = as_stream(stream);
  rollout_kernel<<<(h->n_envs + ENVW - 1) / ENVW, 32 * WARPS, sizeof(BlockSmem), st>>>(h->model_dev, h->model_bytes, h->ep, h->n_envs, h->state_dev, ra);
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
