/*
 * bmi_physics_oracle.c — CPU (double precision) restatement of the bmirobot environment step.
 * TEST INFRASTRUCTURE ONLY: linked by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; the product never loads it.
 *
 * PARITY UNPINNED against PyBullet itself: the reference delegates the arithmetic of this path
 * to PyBullet 3.1.7 (README.md:10; Bullet btMultiBodyDynamicsWorld + BussIK), which is neither
 * in the reference tree nor installable here.  This file restates the published algorithm
 * (Featherstone forward dynamics, velocity-level MLCP solved by projected Gauss-Seidel with
 * Bullet's row set-up and row order, GJK / EPA self-collision of the arm's convex hulls with
 * Bullet's same-multibody row diagonal, DLS inverse kinematics) and is pinned on what the
 * reference's own recorded trajectories (bmirobot_1000_*_demo.npz) determine: reset pose, block
 * drop / settle transient (contact ERP 0.08, slop 1e-5, matched to 1e-6), sliding friction, the
 * 10-step arm trajectory of the fresh-process episode 0 (EE within 0.9 mm after step 1, 14.3 mm
 * after step 10; wrist hold angle and elbow stall of the permanent self-contacts)
 * and the open-loop replay of whole recorded push / pick episodes (tests/test_oracle_physics.py,
 * profiles/r02_reference_replay.md).
 *
 * Reference call sites restated (paths relative to the reference tree):
 *   bmirobot_env/bmirobot_env_push_F.py:92-108   step: clip, action[3]=0, IK+motors, 20 sub-steps
 *   bmirobot_env/bmirobot_env_push_F.py:110-165  reset
 *   bmirobot_env/bmirobot_env_push_F.py:169-237  27-float observation
 *   bmirobot_env/bmirobot.py:129-191             applyAction / sent_hand_moving (motor set-points)
 *   bmirobot_env/bmirobot_inverse_kinematics.py:28-33  position-only IK of link 11
 *   bmirobot_env/bmirobot_env_pickandplace_v2.py:92-95 auto-grip rule (pick task)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/bmi_model.h"
#include "../tools/geom/convex_epa.h"

#define NL 9
#define NU 15 /* generalized velocities: 9 joints + block linear 3 + block angular 3 */
#define MAX_ROWS 192
#define MAX_CONTACTS 48      /* storage only: the oracle keeps every contact it finds */
#define MAX_ARM_CONTACTS 48
#define MAX_SHAPE_V 32
#define MAX_SHAPE_P 64
#define MAX_HULLS 12         /* full-resolution convex hulls of the collision meshes (self-collision, bmo_set_hulls) */

typedef struct {
  int nl, ns;
  int parent[NL], shape_of[NL];
  double jpos[NL][3], jrot[NL][9], axis[NL][3], lo[NL], hi[NL], damp[NL], mass[NL], com[NL][3], inertia[NL][3], mu[NL];
  int s_link[BMI_MAX_SHAPES], s_nv[BMI_MAX_SHAPES], s_np[BMI_MAX_SHAPES];
  double s_v[BMI_MAX_SHAPES][MAX_SHAPE_V][3], s_p[BMI_MAX_SHAPES][MAX_SHAPE_P][4], s_c[BMI_MAX_SHAPES][3], s_r[BMI_MAX_SHAPES],
      s_mu[BMI_MAX_SHAPES];
  double P[BMI_MODEL_HDR];
  /* full convex hulls (every hull vertex of the collision mesh) used by the arm's self-collision */
  int nh, h_link[MAX_HULLS], h_nv[MAX_HULLS];
  double* h_v[MAX_HULLS];
  double h_c[MAX_HULLS][3], h_r[MAX_HULLS], h_mu[MAX_HULLS];
  /* baked pair tables of the kernel (include/bmi_model.h SC_*): used instead of GJK / EPA when MP_SELF_TABLE is set, so
   * that the kernel-vs-oracle tests compare like with like; the table-vs-EPA deviation is measured on the CPU */
  const float* sc_table; int64_t sc_n;
} Model;

typedef struct {
  double q[NL], qd[NL], qt[NL];
  double bp[3], bq[4], bv[3], bw[3];
  double goal[3];
  /* warm-start cache: contacts of the previous sub-step (slot id, normal + two friction impulses) */
  int wn, wid[MAX_CONTACTS];
  double wl[MAX_CONTACTS][3];
  double wm[NL]; /* motor impulses of the previous sub-step */
} State;

typedef struct {
  Model m;
  int task;
  double bh[3], bmass, binertia[3], bmu; /* block half extents / mass / diagonal inertia / friction */
  int cap_c, cap_a;                      /* optional contact caps (kernel-vs-oracle tests), 0 = unbounded */
  int kernel_schedule;                   /* 1: honour MP_PGS_COMPRESS / MP_PGS_TAIL (the kernel's iteration schedule) */
  /* statistics of the last step (for tests / tuning) */
  int last_rows, last_iters, last_contacts;
  double resid_hist[256]; int worst_row[256];
} Env;

/* ------------------------------------------------------------------ small vector helpers */
static void v3set(double* a, double x, double y, double z) { a[0] = x; a[1] = y; a[2] = z; }
static void v3cpy(double* a, const double* b) { a[0] = b[0]; a[1] = b[1]; a[2] = b[2]; }
static void v3add(double* o, const double* a, const double* b) { o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2]; }
static void v3sub(double* o, const double* a, const double* b) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v3cross(double* o, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static void v3axpy(double* o, double s, const double* a) { o[0] += s * a[0]; o[1] += s * a[1]; o[2] += s * a[2]; }
static double v3norm(const double* a) { return sqrt(v3dot(a, a)); }
static void m3mul(double* o, const double* a, const double* b) { /* o = a b, row-major */
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(o, t, sizeof(t));
}
static void m3vec(double* o, const double* a, const double* v) {
  double x = a[0] * v[0] + a[1] * v[1] + a[2] * v[2], y = a[3] * v[0] + a[4] * v[1] + a[5] * v[2],
         z = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void m3tvec(double* o, const double* a, const double* v) { /* o = a^T v */
  double x = a[0] * v[0] + a[3] * v[1] + a[6] * v[2], y = a[1] * v[0] + a[4] * v[1] + a[7] * v[2],
         z = a[2] * v[0] + a[5] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
/* rotation about a unit axis (Rodrigues) */
static void axis_angle(double* R, const double* u, double th) {
  double c = cos(th), s = sin(th), C = 1 - c;
  R[0] = c + u[0] * u[0] * C;        R[1] = u[0] * u[1] * C - u[2] * s; R[2] = u[0] * u[2] * C + u[1] * s;
  R[3] = u[1] * u[0] * C + u[2] * s; R[4] = c + u[1] * u[1] * C;        R[5] = u[1] * u[2] * C - u[0] * s;
  R[6] = u[2] * u[0] * C - u[1] * s; R[7] = u[2] * u[1] * C + u[0] * s; R[8] = c + u[2] * u[2] * C;
}
static void quat_to_mat(double* R, const double* q) { /* q = x,y,z,w */
  double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}
/* roll, pitch, yaw of a rotation matrix (PyBullet getEulerFromQuaternion convention) */
static void mat_to_euler(double* e, const double* R) {
  double sarg = -R[6];
  if (sarg <= -0.99999) { e[0] = 0; e[1] = -0.5 * M_PI; e[2] = atan2(-R[1], -R[2]); }   /* gimbal lock */
  else if (sarg >= 0.99999) { e[0] = 0; e[1] = 0.5 * M_PI; e[2] = atan2(-R[1], R[2]); }
  else { e[0] = atan2(R[7], R[8]); e[1] = asin(sarg); e[2] = atan2(R[3], R[0]); }
}

/* ------------------------------------------------------------------ model loading */
static int load_model(Model* m, const float* b, int64_t n) {
  if (n < BMI_MODEL_HDR || b[MP_MAGIC] != BMI_MODEL_MAGIC) return -1;
  if ((int64_t)b[MP_TOTAL] != n) return -2;
  memset(m, 0, sizeof(*m));
  for (int i = 0; i < BMI_MODEL_HDR; ++i) m->P[i] = b[i];
  m->nl = (int)b[MP_N_LINKS];
  m->ns = (int)b[MP_N_SHAPES];
  if (m->nl != NL || m->ns > BMI_MAX_SHAPES) return -3;
  const float* lk = b + (int)b[MP_LINKS_OFF];
  for (int i = 0; i < NL; ++i, lk += BMI_LINK_STRIDE) {
    m->parent[i] = (int)lk[ML_PARENT];
    for (int k = 0; k < 3; ++k) { m->jpos[i][k] = lk[ML_JPOS + k]; m->axis[i][k] = lk[ML_AXIS + k]; m->com[i][k] = lk[ML_COM + k]; m->inertia[i][k] = lk[ML_INERTIA + k]; }
    for (int k = 0; k < 9; ++k) m->jrot[i][k] = lk[ML_JROT + k];
    m->lo[i] = lk[ML_LO]; m->hi[i] = lk[ML_HI]; m->damp[i] = lk[ML_DAMPING]; m->mass[i] = lk[ML_MASS];
    m->shape_of[i] = (int)lk[ML_SHAPE]; m->mu[i] = lk[ML_MU];
  }
  const float* sh = b + (int)b[MP_SHAPES_OFF];
  const float* pool = b + (int)b[MP_POOL_OFF];
  for (int s = 0; s < m->ns; ++s, sh += BMI_SHAPE_STRIDE) {
    m->s_link[s] = (int)sh[MS_LINK]; m->s_nv[s] = (int)sh[MS_NVERTS]; m->s_np[s] = (int)sh[MS_NPLANES];
    if (m->s_nv[s] > MAX_SHAPE_V || m->s_np[s] > MAX_SHAPE_P) return -4;
    const float* v = pool + (int)sh[MS_VERT_OFF];
    const float* p = pool + (int)sh[MS_PLANE_OFF];
    for (int k = 0; k < m->s_nv[s]; ++k) v3set(m->s_v[s][k], v[3 * k], v[3 * k + 1], v[3 * k + 2]);
    for (int k = 0; k < m->s_np[s]; ++k) for (int c = 0; c < 4; ++c) m->s_p[s][k][c] = p[4 * k + c];
    v3set(m->s_c[s], sh[MS_SPHERE_C], sh[MS_SPHERE_C + 1], sh[MS_SPHERE_C + 2]);
    m->s_r[s] = sh[MS_SPHERE_R]; m->s_mu[s] = sh[MS_MU];
  }
  return 0;
}

/* ------------------------------------------------------------------ kinematics */
typedef struct {
  double R[NL][9], p[NL][3], z[NL][3], c[NL][3]; /* link rotation, origin, joint axis (world), COM */
} Kin;

static void fk(const Model* m, const double* q, Kin* k) {
  for (int i = 0; i < NL; ++i) {
    double Rq[9], Rl[9];
    axis_angle(Rq, m->axis[i], q[i]);
    m3mul(Rl, m->jrot[i], Rq); /* parent-from-child */
    int pa = m->parent[i];
    if (pa < 0) {
      memcpy(k->R[i], Rl, sizeof(Rl));
      v3set(k->p[i], m->P[MP_BASE_PX] + m->jpos[i][0], m->P[MP_BASE_PY] + m->jpos[i][1], m->P[MP_BASE_PZ] + m->jpos[i][2]);
    } else {
      m3mul(k->R[i], k->R[pa], Rl);
      double t[3];
      m3vec(t, k->R[pa], m->jpos[i]);
      v3add(k->p[i], k->p[pa], t);
    }
    m3vec(k->z[i], k->R[i], m->axis[i]);
    double t[3];
    m3vec(t, k->R[i], m->com[i]);
    v3add(k->c[i], k->p[i], t);
  }
}

/* translational Jacobian (3 x 9, row-major) of world point x rigidly attached to link l */
static void point_jacobian(const Model* m, const Kin* k, int l, const double* x, double* J) {
  memset(J, 0, sizeof(double) * 27);
  for (int j = l; j >= 0; j = m->parent[j]) {
    double r[3], c[3];
    v3sub(r, x, k->p[j]);
    v3cross(c, k->z[j], r);
    J[j] = c[0]; J[9 + j] = c[1]; J[18 + j] = c[2];
  }
}

/* recursive Newton-Euler in world coordinates: tau = M(q) qdd + bias(q, qd) with gravity gz and
 * Bullet's per-link velocity damping (force m v (k + k|v|), torque I w (k + k|w|) at the COM). */
static void rnea(const Model* m, const Kin* k, const double* qd, const double* qdd, double gz, double kl, double ka,
                 double* tau) {
  double w[NL][3], al[NL][3], a[NL][3], vo[NL][3]; /* ang vel, ang acc, origin acc, origin vel */
  double f[NL][3], n[NL][3];
  for (int i = 0; i < NL; ++i) {
    int pa = m->parent[i];
    double wp[3] = {0, 0, 0}, alp[3] = {0, 0, 0}, ap[3] = {0, 0, -gz}, vp[3] = {0, 0, 0}, r[3] = {0, 0, 0};
    if (pa >= 0) {
      v3cpy(wp, w[pa]); v3cpy(alp, al[pa]); v3cpy(ap, a[pa]); v3cpy(vp, vo[pa]);
      v3sub(r, k->p[i], k->p[pa]);
    }
    double t[3], t2[3];
    /* origin velocity / acceleration of link i (its origin is rigidly attached to the parent) */
    v3cross(t, wp, r); v3add(vo[i], vp, t);
    v3cross(t, alp, r); v3add(a[i], ap, t);
    v3cross(t, wp, r); v3cross(t2, wp, t); v3add(a[i], a[i], t2);
    /* angular */
    v3cpy(w[i], wp); v3axpy(w[i], qd[i], k->z[i]);
    v3cpy(al[i], alp); v3axpy(al[i], qdd[i], k->z[i]);
    v3cross(t, wp, k->z[i]); v3axpy(al[i], qd[i], t);
    /* COM acceleration, force and moment */
    double rc[3], ac[3], vc[3];
    v3sub(rc, k->c[i], k->p[i]);
    v3cross(t, al[i], rc); v3add(ac, a[i], t);
    v3cross(t, w[i], rc); v3cross(t2, w[i], t); v3add(ac, ac, t2);
    v3cross(t, w[i], rc); v3add(vc, vo[i], t);
    double F[3], N[3], Iw[3], wl[3], all_[3], tl[3];
    for (int c = 0; c < 3; ++c) F[c] = m->mass[i] * ac[c];
    double vn = v3norm(vc);
    v3axpy(F, m->mass[i] * (kl + kl * vn), vc);
    /* I (world) x = R diag(I) R^T x */
    m3tvec(wl, k->R[i], w[i]); m3tvec(all_, k->R[i], al[i]);
    for (int c = 0; c < 3; ++c) tl[c] = m->inertia[i][c] * wl[c];
    m3vec(Iw, k->R[i], tl);
    for (int c = 0; c < 3; ++c) tl[c] = m->inertia[i][c] * all_[c];
    m3vec(N, k->R[i], tl);
    v3cross(t, w[i], Iw); v3add(N, N, t);
    double wn = v3norm(w[i]);
    v3axpy(N, (ka + ka * wn), Iw);
    v3cpy(f[i], F);
    v3cross(t, rc, F); v3add(n[i], N, t); /* moment about the link origin */
  }
  for (int i = NL - 1; i >= 0; --i) {
    tau[i] = v3dot(k->z[i], n[i]);
    int pa = m->parent[i];
    if (pa >= 0) {
      double r[3], t[3];
      v3sub(r, k->p[i], k->p[pa]);
      v3add(f[pa], f[pa], f[i]);
      v3cross(t, r, f[i]);
      v3add(n[pa], n[pa], n[i]);
      v3add(n[pa], n[pa], t);
    }
  }
}

/* Cholesky of an n x n SPD matrix (row-major, ld NL), in place (lower); returns 0 on success */
static int chol9(double* A) {
  for (int j = 0; j < NL; ++j) {
    double d = A[j * NL + j];
    for (int k = 0; k < j; ++k) d -= A[j * NL + k] * A[j * NL + k];
    if (d <= 0) return -1;
    d = sqrt(d);
    A[j * NL + j] = d;
    for (int i = j + 1; i < NL; ++i) {
      double s = A[i * NL + j];
      for (int k = 0; k < j; ++k) s -= A[i * NL + k] * A[j * NL + k];
      A[i * NL + j] = s / d;
    }
  }
  return 0;
}
static void chol9_solve(const double* Lm, const double* b, double* x) {
  double y[NL];
  for (int i = 0; i < NL; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= Lm[i * NL + k] * y[k];
    y[i] = s / Lm[i * NL + i];
  }
  for (int i = NL - 1; i >= 0; --i) {
    double s = y[i];
    for (int k = i + 1; k < NL; ++k) s -= Lm[k * NL + i] * x[k];
    x[i] = s / Lm[i * NL + i];
  }
}

/* ------------------------------------------------------------------ inverse kinematics
 * PyBullet calculateInverseKinematics without null-space arrays = BussIK damped least squares
 * (IK2_VEL_DLS): up to ik_iters steps of dq = (J^T J + d I)^-1 J^T e from the current q, each
 * step limited to ik_max_angle, stopping when |e| < ik_thresh.  Position only. */
static void solve_sym9(double* A, double* b) { /* Gaussian elimination with partial pivoting, in place */
  for (int c = 0; c < NL; ++c) {
    int p = c;
    for (int r = c + 1; r < NL; ++r) if (fabs(A[r * NL + c]) > fabs(A[p * NL + c])) p = r;
    if (p != c) {
      for (int k = 0; k < NL; ++k) { double t = A[c * NL + k]; A[c * NL + k] = A[p * NL + k]; A[p * NL + k] = t; }
      double t = b[c]; b[c] = b[p]; b[p] = t;
    }
    double d = A[c * NL + c];
    for (int r = c + 1; r < NL; ++r) {
      double f = A[r * NL + c] / d;
      for (int k = c; k < NL; ++k) A[r * NL + k] -= f * A[c * NL + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = NL - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < NL; ++k) s -= A[r * NL + k] * b[k];
    b[r] = s / A[r * NL + r];
  }
}

static void ee_point(const Model* m, const Kin* k, double* x) {
  int ee = (int)m->P[MP_EE_LINK];
  v3cpy(x, k->p[ee]); /* link-frame origin, getLinkState(...)[4] (bmirobot.py:144-145) */
}

static void solve_ik(const Model* m, const double* q0, const double* target, double* qout) {
  double q[NL];
  memcpy(q, q0, sizeof(q));
  int iters = (int)m->P[MP_IK_ITERS];
  int ee = (int)m->P[MP_EE_LINK];
  for (int it = 0; it < iters; ++it) {
    Kin k;
    fk(m, q, &k);
    double x[3], e[3], J[27];
    ee_point(m, &k, x);
    v3sub(e, target, x);
    if (v3norm(e) <= m->P[MP_IK_THRESH]) break;
    point_jacobian(m, &k, ee, x, J);
    double A[NL * NL], b[NL];
    for (int i = 0; i < NL; ++i) {
      for (int j = 0; j < NL; ++j) A[i * NL + j] = J[i] * J[j] + J[9 + i] * J[9 + j] + J[18 + i] * J[18 + j];
      A[i * NL + i] += m->P[MP_IK_DAMPING];
      b[i] = J[i] * e[0] + J[9 + i] * e[1] + J[18 + i] * e[2];
    }
    solve_sym9(A, b);
    double mx = 0;
    for (int i = 0; i < NL; ++i) if (fabs(b[i]) > mx) mx = fabs(b[i]);
    double sc = mx > m->P[MP_IK_MAX_ANGLE] ? m->P[MP_IK_MAX_ANGLE] / mx : 1.0;
    for (int i = 0; i < NL; ++i) q[i] += sc * b[i];
  }
  memcpy(qout, q, sizeof(q));
}

/* ------------------------------------------------------------------ constraint rows */
typedef struct {
  double J[NU], W[NU]; /* Jacobian row and M^-1 J^T */
  double inv_diag, rhs, lo, hi, lambda;
  int friction_of;     /* index of the normal row for friction rows, else -1 */
  double mu;
  int under;           /* 1: row uses Bullet's same-multibody diagonal (strongly under-relaxed) */
} Row;

typedef struct {
  int link;     /* arm link index of body 2, or -1 (static world) */
  int has_block;/* 1: body 1 is the block */
  int link1;    /* self-collision: arm link index of body 1 (>= 0; -1 = the static base link), unused otherwise */
  int self;     /* 1: contact between two links of the arm */
  double x[3], n[3], dist, mu;  /* x: contact point (on body 1 for self-collision), n: normal from body 2 towards body 1 */
  double x2[3]; /* self-collision: witness point on body 2 */
  int id;       /* candidate slot: block vertex 0..7 | 8 + 32 shape + hull vertex | 136 + 32 shape + candidate | 1000 + 16 hullA + hullB */
} Contact;

static void block_vertices(const Env* e, const State* s, const double* Rb, double v[8][3]) {
  for (int k = 0; k < 8; ++k) {
    double l[3] = {(k & 1 ? 1 : -1) * e->bh[0], (k & 2 ? 1 : -1) * e->bh[1], (k & 4 ? 1 : -1) * e->bh[2]}, t[3];
    m3vec(t, Rb, l);
    v3add(v[k], s->bp, t);
  }
}

/* keep the `cap` candidates with the smallest distance (ties: lowest index), in index order */
static int select_deepest(const double* dist, int n, double margin, int cap, int* out) {
  int picked[64] = {0}, cnt = 0;
  for (int r = 0; r < cap; ++r) {
    int best = -1;
    for (int i = 0; i < n; ++i)
      if (!picked[i] && dist[i] < margin && (best < 0 || dist[i] < dist[best])) best = i;
    if (best < 0) break;
    picked[best] = 1;
    ++cnt;
  }
  int o = 0;
  for (int i = 0; i < n; ++i) if (picked[i]) out[o++] = i;
  return cnt;
}

static int find_contacts(const Env* e, const State* s, const Kin* k, Contact* C, int nc0) {
  const Model* m = &e->m;
  /* cap_c / cap_a: optional lane budget of the CUDA kernel (bmo_set_caps; default = storage bound = keep everything) */
  const int MAXC_ = e->cap_c > 0 ? e->cap_c : MAX_CONTACTS, MAXA_ = e->cap_a > 0 ? e->cap_a : MAX_ARM_CONTACTS;
  int nc = nc0, na = nc0;   /* the contacts already in the list are the arm's self-contacts */
  double Rb[9], bv[8][3];
  quat_to_mat(Rb, s->bq);
  block_vertices(e, s, Rb, bv);
  const double tz = m->P[MP_TABLE_Z];
  /* block vertices vs table plane: up to 4 deepest */
  {
    double d[8];
    int sel[8];
    for (int i = 0; i < 8; ++i) d[i] = bv[i][2] - tz;
    int n = select_deepest(d, 8, m->P[MP_TABLE_MARGIN], 4, sel);
    for (int i = 0; i < n && nc < MAXC_; ++i) {
      Contact* c = &C[nc++];
      c->link = -1; c->link1 = -1; c->self = 0; c->has_block = 1; c->id = sel[i]; v3cpy(c->x, bv[sel[i]]); v3set(c->n, 0, 0, 1);
      c->dist = d[sel[i]]; c->mu = e->bmu * m->P[MP_MU_TABLE];
    }
  }
  double brad = v3norm(e->bh);
  if (m->P[MP_FULL_HULLS] > 0.5 && m->nh > 0) {
    /* faithful geometry: every arm link's full convex hull against the table top and against the block (box hull);
     * narrow phase = closest points / penetration depth (GJK + EPA), one point per pair and sub-step, Bullet's margins
     * (hull MP_HULL_MARGIN; the block is a btBoxShape whose margin is inside its extents) */
    const double margin = m->P[MP_HULL_MARGIN];
    double bl[24];
    for (int i = 0; i < 8; ++i) { bl[3 * i] = (i & 1 ? 1 : -1) * e->bh[0]; bl[3 * i + 1] = (i & 2 ? 1 : -1) * e->bh[1]; bl[3 * i + 2] = (i & 4 ? 1 : -1) * e->bh[2]; }
    Cvx B; B.nv = 8; B.v = bl; memcpy(B.R, Rb, sizeof(Rb)); v3cpy(B.p, s->bp);
    for (int h = 0; h < m->nh; ++h) {
      int l = m->h_link[h];
      if (l < 0) continue;
      Cvx A; A.nv = m->h_nv[h]; A.v = m->h_v[h]; memcpy(A.R, k->R[l], sizeof(A.R)); v3cpy(A.p, k->p[l]);
      /* table: the two deepest hull vertices below the contact margin */
      {
        double cw[3], t[3];
        m3vec(t, A.R, m->h_c[h]); v3add(cw, A.p, t);
        if (cw[2] - m->h_r[h] - margin < tz + m->P[MP_CONTACT_MARGIN]) {
          int b0 = -1, b1 = -1; double d0 = 1e30, d1 = 1e30;
          for (int i = 0; i < A.nv; ++i) {
            double z = A.R[6] * A.v[3 * i] + A.R[7] * A.v[3 * i + 1] + A.R[8] * A.v[3 * i + 2] + A.p[2] - margin - tz;
            if (z < d0) { d1 = d0; b1 = b0; d0 = z; b0 = i; } else if (z < d1) { d1 = z; b1 = i; }
          }
          int sel[2] = {b0 < b1 ? b0 : b1, b0 < b1 ? b1 : b0}; double dd[2] = {b0 < b1 ? d0 : d1, b0 < b1 ? d1 : d0};
          for (int i = 0; i < 2; ++i) if (sel[i] >= 0 && dd[i] < m->P[MP_CONTACT_MARGIN] && nc < MAX_CONTACTS) {
            Contact* c = &C[nc++];
            double w[3]; m3vec(w, A.R, A.v + 3 * sel[i]); v3add(c->x, A.p, w); c->x[2] -= margin;
            c->link = l; c->link1 = -1; c->self = 0; c->has_block = 0; c->id = 8 + 1024 * (h + 1) + sel[i]; v3set(c->n, 0, 0, 1);
            c->dist = dd[i]; c->mu = m->h_mu[h] * m->P[MP_MU_TABLE];
          }
        }
      }
      /* block */
      double ca[3], t[3], d[3];
      m3vec(t, A.R, m->h_c[h]); v3add(ca, A.p, t); v3sub(d, ca, s->bp);
      if (v3norm(d) > m->h_r[h] + brad + margin + m->P[MP_SELF_NEAR]) continue;
      double n[3], pa[3], pb[3];
      double depth = epa_penetration(&B, &A, n, pa, pb);   /* body 1 = block (A of the EPA), body 2 = link */
      if (depth < 0) {
        double gap = gjk_distance(&B, &A, pa, pb);
        if (gap < 0 || gap > margin + m->P[MP_SELF_NEAR]) continue;
        for (int r = 0; r < 3; ++r) n[r] = -(pa[r] - pb[r]) / gap;
        depth = -gap;
      }
      if (nc >= MAX_CONTACTS) break;
      Contact* c = &C[nc++];
      c->link = l; c->link1 = -1; c->self = 0; c->has_block = 1; c->id = 500 + h;
      for (int r = 0; r < 3; ++r) { c->n[r] = -n[r]; c->x[r] = pb[r]; }   /* Bullet takes the point on body B; both sides use it here */
      c->dist = -(depth + margin);
      c->mu = e->bmu * m->h_mu[h];
    }
    return nc;
  }
  for (int sidx = 0; sidx < m->ns; ++sidx) {
    int l = m->s_link[sidx], nv = m->s_nv[sidx], np = m->s_np[sidx];
    double wv[MAX_SHAPE_V][3];
    for (int i = 0; i < nv; ++i) { double t[3]; m3vec(t, k->R[l], m->s_v[sidx][i]); v3add(wv[i], k->p[l], t); }
    /* hull vertices vs table plane: up to 2 deepest */
    {
      double d[MAX_SHAPE_V];
      int sel[MAX_SHAPE_V];
      for (int i = 0; i < nv; ++i) d[i] = wv[i][2] - tz;
      int n = select_deepest(d, nv, m->P[MP_CONTACT_MARGIN], 2, sel);
      for (int i = 0; i < n && nc < MAXC_ && na < MAXA_; ++i) {
        Contact* c = &C[nc++];
        ++na;
        c->link = l; c->link1 = -1; c->self = 0; c->has_block = 0; c->id = 8 + 32 * sidx + sel[i]; v3cpy(c->x, wv[sel[i]]); v3set(c->n, 0, 0, 1);
        c->dist = d[sel[i]]; c->mu = m->s_mu[sidx] * m->P[MP_MU_TABLE];
      }
    }
    /* block vs hull: broadphase on bounding spheres */
    double cw[3], t[3], dd[3];
    m3vec(t, k->R[l], m->s_c[sidx]); v3add(cw, k->p[l], t);
    v3sub(dd, s->bp, cw);
    if (v3norm(dd) > m->s_r[sidx] + brad + m->P[MP_BLOCK_MARGIN]) continue;
    double d[8 + MAX_SHAPE_V], nrm[8 + MAX_SHAPE_V][3];
    /* (a) block vertices against the hull planes: signed distance = max over planes */
    for (int i = 0; i < 8; ++i) {
      double xl[3], r[3];
      v3sub(r, bv[i], k->p[l]); m3tvec(xl, k->R[l], r);
      double best = -1e30; int bp = 0;
      for (int p = 0; p < np; ++p) {
        double sd = m->s_p[sidx][p][0] * xl[0] + m->s_p[sidx][p][1] * xl[1] + m->s_p[sidx][p][2] * xl[2] + m->s_p[sidx][p][3];
        if (sd > best) { best = sd; bp = p; }
      }
      d[i] = best;
      m3vec(nrm[i], k->R[l], m->s_p[sidx][bp]); /* outward hull normal: from link towards block */
    }
    /* (b) hull vertices against the block box: signed distance = max over the 6 faces */
    for (int i = 0; i < nv; ++i) {
      double xb[3], r[3];
      v3sub(r, wv[i], s->bp); m3tvec(xb, Rb, r);
      double best = -1e30; int ba = 0; double sg = 1;
      for (int a = 0; a < 3; ++a) {
        double sd = fabs(xb[a]) - e->bh[a];
        if (sd > best) { best = sd; ba = a; sg = xb[a] >= 0 ? 1 : -1; }
      }
      d[8 + i] = best;
      double ln[3] = {0, 0, 0}, wn[3];
      ln[ba] = sg;
      m3vec(wn, Rb, ln); /* outward block normal: from block towards link; flip -> from link to block */
      v3set(nrm[8 + i], -wn[0], -wn[1], -wn[2]);
    }
    int sel[8 + MAX_SHAPE_V];
    int n = select_deepest(d, 8 + nv, m->P[MP_BLOCK_MARGIN], 3, sel);
    for (int i = 0; i < n && nc < MAXC_ && na < MAXA_; ++i) {
      int ci = sel[i];
      Contact* c = &C[nc++];
      ++na;
      c->link = l; c->link1 = -1; c->self = 0; c->has_block = 1; c->id = 136 + 32 * sidx + ci;
      v3cpy(c->x, ci < 8 ? bv[ci] : wv[ci - 8]);
      v3cpy(c->n, nrm[ci]); /* normal from the link (body 2) towards the block (body 1) */
      c->dist = d[ci]; c->mu = e->bmu * m->s_mu[sidx];
    }
  }
  return nc;
}

/* Self-collision of the arm (bmirobot.py:58 flags=9 -> URDF_USE_SELF_COLLISION, Bullet's default of that flag: every
 * pair of links collides except a link with its DIRECT parent; links rigidly attached to the fixed base are static and
 * do not collide with each other).  Narrow phase = penetration depth of the two convex hulls (GJK + EPA as in Bullet's
 * btGjkEpaPenetrationDepthSolver, convex_epa.h), one point per pair and sub-step; the hulls carry Bullet's collision
 * margin (MP_HULL_MARGIN per shape).  Only penetrating pairs produce a row. */
static int hull_parent_link(const Model* m, int hl) { return hl < 0 ? -2 : m->parent[hl]; }
static void hull_world(const Model* m, const Kin* k, int h, Cvx* c) {
  int l = m->h_link[h];
  c->nv = m->h_nv[h]; c->v = m->h_v[h];
  if (l >= 0) { memcpy(c->R, k->R[l], sizeof(c->R)); v3cpy(c->p, k->p[l]); }
  else { /* right_link1: rigidly attached to the fixed base; link 0's joint origin is expressed in its frame */
    double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    memcpy(c->R, I, sizeof(I));
    v3set(c->p, m->P[MP_BASE_PX], m->P[MP_BASE_PY], m->P[MP_BASE_PZ]);
  }
}
/* signed core distance of hulls a, b (gap > 0, or -penetration depth), unit normal n from b towards a and the witness
 * points; returns 0 when the cores are farther apart than `far` */
static int hull_pair(const Model* m, const Kin* k, int a, int b, double far, double* dist, double* n, double* pa, double* pb) {
  Cvx A, B;
  hull_world(m, k, a, &A); hull_world(m, k, b, &B);
  double ca[3], cb[3], t[3], d[3];
  m3vec(t, A.R, m->h_c[a]); v3add(ca, A.p, t);
  m3vec(t, B.R, m->h_c[b]); v3add(cb, B.p, t);
  v3sub(d, ca, cb);
  if (v3norm(d) > m->h_r[a] + m->h_r[b] + far) return 0;
  double ne[3];
  double depth = epa_penetration(&A, &B, ne, pa, pb);
  if (depth >= 0) { for (int r = 0; r < 3; ++r) n[r] = -ne[r]; *dist = -depth; return 1; }
  double gap = gjk_distance(&A, &B, pa, pb);
  if (gap < 0 || gap > far) return 0;
  for (int r = 0; r < 3; ++r) n[r] = (pa[r] - pb[r]) / gap;
  *dist = gap;
  return 1;
}
/* table look-up shared (as an algorithm) with csrc/physics.cu self_contacts(): bilinear over the four surrounding nodes
 * when they describe the same hull features (normals within 1.8 degrees, witness points within 2 mm), else the nearest
 * node; out = dist, n(3), xa(3) in the frame of link A.  Returns 0 when there is no contact information. */
static int sc_lookup(const float* T, const float* d, double qa, double qb, double* out7) {
  const double h = d[SC_H];
  const int na = (int)d[SC_NA], nb = (int)d[SC_NB];
  double fa = (qa - d[SC_A0]) / h, fb = (qb - d[SC_B0]) / h;
  fa = fmin(fmax(fa, 0.0), na - 1.0);
  fb = fmin(fmax(fb, 0.0), nb - 1.0);
  int i = (int)floor(fa), j = (int)floor(fb);
  if (i > na - 2) i = na - 2;
  if (j > nb - 2) j = nb - 2;
  const double wa = fa - i, wb = fb - j;
  const float* node[4]; double w[4] = {(1 - wa) * (1 - wb), (1 - wa) * wb, wa * (1 - wb), wa * wb};
  const float* base = T + (int64_t)d[SC_OFF];
  node[0] = base + 8 * ((int64_t)i * nb + j); node[1] = node[0] + 8; node[2] = node[0] + 8 * (int64_t)nb; node[3] = node[2] + 8;
  int near = (wa >= 0.5 ? 2 : 0) + (wb >= 0.5 ? 1 : 0);
  if (node[near][0] > 100.f) return 0;
  int same = 1;
  for (int c = 0; c < 4; ++c) {
    if (node[c][0] > 100.f) { same = 0; break; }
    double dn = node[c][1] * node[near][1] + node[c][2] * node[near][2] + node[c][3] * node[near][3];
    double dx = 0; for (int r = 0; r < 3; ++r) { double t = node[c][4 + r] - node[near][4 + r]; dx += t * t; }
    if (dn < 0.9995 || dx > 4e-6) { same = 0; break; }
  }
  if (!same) { for (int r = 0; r < 7; ++r) out7[r] = node[near][r]; return 1; }
  for (int r = 0; r < 7; ++r) out7[r] = w[0] * node[0][r] + w[1] * node[1][r] + w[2] * node[2][r] + w[3] * node[3][r];
  double nn = sqrt(out7[1] * out7[1] + out7[2] * out7[2] + out7[3] * out7[3]);
  for (int r = 1; r < 4; ++r) out7[r] /= nn;
  return 1;
}
static int find_self_contacts(const Env* e, const State* s, const Kin* k, Contact* C, int nc) {
  const Model* m = &e->m;
  const double margin = m->P[MP_HULL_MARGIN], near = m->P[MP_SELF_NEAR];
  if (m->P[MP_SELF_TABLE] > 0.5 && m->sc_table) { /* the kernel's path: the baked 2-joint pair tables */
    const float* T = m->sc_table;
    int np = (int)T[SC_NPAIRS];
    for (int p = 0; p < np; ++p) {
      const float* d = T + SC_HDR + SC_DESC * p;
      double o[7];
      if (!sc_lookup(T, d, s->q[(int)d[SC_JA]], s->q[(int)d[SC_JB]], o)) continue;
      double dist = o[0] - 2 * margin;
      if (dist > near) continue;
      if (nc >= (e->cap_a > 0 ? e->cap_a : MAX_CONTACTS)) break;
      Contact* c = &C[nc++];
      int la = (int)d[SC_LA], lb = (int)d[SC_LB];
      c->link1 = la; c->link = lb; c->self = 1; c->has_block = 0; c->id = 1000 + p;
      double Ra[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pa[3] = {m->P[MP_BASE_PX], m->P[MP_BASE_PY], m->P[MP_BASE_PZ]};
      if (la >= 0) { memcpy(Ra, k->R[la], sizeof(Ra)); v3cpy(pa, k->p[la]); }
      double t[3];
      m3vec(c->n, Ra, o + 1);
      m3vec(t, Ra, o + 4); v3add(c->x, pa, t);
      for (int r = 0; r < 3; ++r) c->x2[r] = c->x[r] - c->n[r] * o[0];
      c->dist = dist; c->mu = d[SC_MU];
    }
    return nc;
  }
  for (int a = 0; a < m->nh; ++a)
    for (int b = a + 1; b < m->nh; ++b) {
      int la = m->h_link[a], lb = m->h_link[b];
      if (hull_parent_link(m, lb) == la || hull_parent_link(m, la) == lb) continue; /* parent-child pairs are filtered */
      if (la < 0 && lb < 0) continue;
      double dist, n[3], pa[3], pb[3];
      /* Bullet reports closest points while the gap is below the two margins plus the manifold's contact breaking
       * threshold; a separated row only limits the approach velocity (rhs -= dist / dt) */
      if (!hull_pair(m, k, a, b, 2 * margin + near, &dist, n, pa, pb)) continue;
      if (nc >= MAX_CONTACTS) break;
      Contact* c = &C[nc++];
      c->link1 = la; c->link = lb; c->self = 1; c->has_block = 0; c->id = 1000 + 16 * a + b;
      for (int r = 0; r < 3; ++r) { c->n[r] = n[r]; c->x[r] = pa[r]; c->x2[r] = pb[r]; }
      c->dist = dist - 2 * margin;
      c->mu = m->h_mu[a] * m->h_mu[b];
      if (c->mu > 10.0) c->mu = 10.0; /* Bullet calculateCombinedFriction clamps the product to 10 */
    }
  return nc;
}

/* two unit tangents orthogonal to n (Bullet btPlaneSpace1) */
static void plane_space(const double* n, double* p, double* q) {
  if (fabs(n[2]) > 0.7071067811865475244) {
    double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
    v3set(p, 0, -n[2] * k, n[1] * k);
    v3set(q, a * k, -n[0] * p[2], n[0] * p[1]);
  } else {
    double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
    v3set(p, -n[1] * k, n[0] * k, 0);
    v3set(q, -n[2] * p[1], n[2] * p[0], a * k);
  }
}

/* fill J for direction dir at contact c (body 1 = block and/or body 2 = link; world static) */
static void contact_jacobian(const Env* e, const State* s, const Kin* k, const Contact* c, const double* dir, double* J) {
  memset(J, 0, sizeof(double) * NU);
  if (c->self) { /* self-collision: body 1 = link1 at x, body 2 = link at x2 */
    double Jp[27];
    if (c->link1 >= 0) {
      point_jacobian(&e->m, k, c->link1, c->x, Jp);
      for (int j = 0; j < NL; ++j) J[j] += dir[0] * Jp[j] + dir[1] * Jp[9 + j] + dir[2] * Jp[18 + j];
    }
    if (c->link >= 0) {
      point_jacobian(&e->m, k, c->link, c->x2, Jp);
      for (int j = 0; j < NL; ++j) J[j] -= dir[0] * Jp[j] + dir[1] * Jp[9 + j] + dir[2] * Jp[18 + j];
    }
    return;
  }
  if (c->has_block) {
    double r[3], t[3];
    v3sub(r, c->x, s->bp);
    v3cross(t, r, dir);
    for (int a = 0; a < 3; ++a) { J[9 + a] = dir[a]; J[12 + a] = t[a]; }
  }
  if (c->link >= 0) {
    double Jp[27];
    point_jacobian(&e->m, k, c->link, c->x, Jp);
    double sgn = c->has_block ? -1.0 : 1.0; /* link is body 2 when the block is present */
    for (int j = 0; j < NL; ++j) J[j] = sgn * (dir[0] * Jp[j] + dir[1] * Jp[9 + j] + dir[2] * Jp[18 + j]);
  }
}

static void apply_minv(const Env* e, const double* Lm, const double* Ib_inv_world, const double* J, double* W) {
  chol9_solve(Lm, J, W);
  for (int a = 0; a < 3; ++a) W[9 + a] = J[9 + a] / e->bmass;
  m3vec(W + 12, Ib_inv_world, J + 12);
}

static double dotn(const double* a, const double* b, int n);
/* Bullet's diagonal for a row between two links of the SAME multibody (btMultiBodyConstraintSolver::
 * setupMultiBodyContactConstraint): jacDiagABInv = 1 / (J_A M^-1 J_A^T + J_B M^-1 J_B^T) -- the two sides are treated as
 * if they were separate bodies, the cross term -2 J_A M^-1 J_B^T is not included.  For neighbouring links of one chain
 * this is far larger than the true diagonal, i.e. the row is strongly under-relaxed and does not converge within the
 * iteration budget: the arm's permanent self-contacts are soft because of it. */
static double self_contact_split_diag(const Env* e, const Kin* k, const double* Lm, const Contact* c, const double* dir) {
  double d = 0;
  for (int side = 0; side < 2; ++side) {
    int l = side == 0 ? c->link1 : c->link;
    if (l < 0) continue;
    double Jp[27], J[NL], W[NL];
    point_jacobian(&e->m, k, l, side == 0 ? c->x : c->x2, Jp);
    for (int j = 0; j < NL; ++j) J[j] = dir[0] * Jp[j] + dir[1] * Jp[9 + j] + dir[2] * Jp[18 + j];
    chol9_solve(Lm, J, W);
    d += dotn(J, W, NL);
  }
  return d;
}
static double dotn(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; ++i) s += a[i] * b[i]; return s; }

/* one simulation sub-step (Bullet btMultiBodyDynamicsWorld::stepSimulation restated) */
static void substep(Env* e, State* s) {
  const Model* m = &e->m;
  const double dt = m->P[MP_DT], gz = m->P[MP_GRAVITY], kl = m->P[MP_LIN_DAMP], ka = m->P[MP_ANG_DAMP];
  Kin k;
  fk(m, s->q, &k);
  /* mass matrix column by column and bias */
  double M[NL * NL], bias[NL], zero[NL] = {0}, ej[NL];
  rnea(m, &k, s->qd, zero, gz, kl, ka, bias);
  for (int j = 0; j < NL; ++j) {
    memset(ej, 0, sizeof(ej));
    ej[j] = 1;
    double col[NL];
    rnea(m, &k, zero, ej, 0.0, 0.0, 0.0, col);
    for (int i = 0; i < NL; ++i) M[i * NL + j] = col[i];
  }
  for (int i = 0; i < NL; ++i) for (int j = 0; j < i; ++j) M[i * NL + j] = M[j * NL + i] = 0.5 * (M[i * NL + j] + M[j * NL + i]);
  double Lm[NL * NL];
  memcpy(Lm, M, sizeof(M));
  chol9(Lm);
  /* unconstrained velocities */
  double u[NU], tau[NL], acc[NL];
  for (int i = 0; i < NL; ++i) tau[i] = -m->damp[i] * s->qd[i] - bias[i];
  chol9_solve(Lm, tau, acc);
  for (int i = 0; i < NL; ++i) u[i] = s->qd[i] + dt * acc[i];
  double Rb[9];
  quat_to_mat(Rb, s->bq);
  {
    double vn = v3norm(s->bv), wn = v3norm(s->bw);
    for (int a = 0; a < 3; ++a) {
      u[9 + a] = s->bv[a] + dt * (-(kl + kl * vn) * s->bv[a]);
      u[12 + a] = s->bw[a] + dt * (-(ka + ka * wn) * s->bw[a]);
    }
    u[11] += dt * gz;
  }
  double Ibinv[9];
  { /* R diag(1/I) R^T */
    double t[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[3 * r + c] = Rb[3 * r + c] / e->binertia[c];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
      Ibinv[3 * r + c] = t[3 * r] * Rb[3 * c] + t[3 * r + 1] * Rb[3 * c + 1] + t[3 * r + 2] * Rb[3 * c + 2];
  }
  /* ---- rows ---- */
  static Row rows[MAX_ROWS];
  int nr = 0;
  const double max_imp = m->P[MP_MOTOR_FORCE] * dt;
  for (int j = 0; j < NL; ++j) { /* position motors: drive velocity to kp (q* - q)/dt + (1-kd) qd */
    Row* r = &rows[nr++];
    memset(r, 0, sizeof(*r));
    r->J[j] = 1;
    apply_minv(e, Lm, Ibinv, r->J, r->W);
    r->inv_diag = 1.0 / r->W[j];
    double target = m->P[MP_MOTOR_KP] * (s->qt[j] - s->q[j]) / dt + (1.0 - m->P[MP_MOTOR_KD]) * s->qd[j];
    r->rhs = (target - u[j]) * r->inv_diag;
    r->lo = -max_imp; r->hi = max_imp; r->friction_of = -1;
  }
  for (int j = 0; j < NL; ++j) { /* joint limits (only when violated) */
    for (int side = 0; side < 2; ++side) {
      double pen = side == 0 ? s->q[j] - m->lo[j] : m->hi[j] - s->q[j];
      if (pen > 0) continue;
      Row* r = &rows[nr++];
      memset(r, 0, sizeof(*r));
      r->J[j] = side == 0 ? 1 : -1;
      apply_minv(e, Lm, Ibinv, r->J, r->W);
      r->inv_diag = 1.0 / fabs(r->W[j]);
      double rel = r->J[j] * u[j];
      r->rhs = (-pen * m->P[MP_ERP_JOINT] / dt - rel) * r->inv_diag;
      r->lo = 0; r->hi = m->P[MP_JOINT_LIMIT_IMPULSE]; r->friction_of = -1;
    }
  }
  const int n_noncontact = nr;
  Contact C[MAX_CONTACTS];
  /* the arm's self-contacts first: Bullet keeps its manifolds in creation order and the robot is loaded before the table
   * and the block (bmirobot.py:53-77, bmirobot_env_push_F.py:145-159) */
  int nc = 0;
  if (m->P[MP_SELF_COLLISION] > 0.5) nc = find_self_contacts(e, s, &k, C, nc);
  nc = find_contacts(e, s, &k, C, nc);
  int normal_row[MAX_CONTACTS];
  for (int ci = 0; ci < nc; ++ci) { /* normal rows */
    Row* r = &rows[nr];
    memset(r, 0, sizeof(*r));
    contact_jacobian(e, s, &k, &C[ci], C[ci].n, r->J);
    apply_minv(e, Lm, Ibinv, r->J, r->W);
    r->inv_diag = 1.0 / dotn(r->J, r->W, NU);
    if (C[ci].self && m->P[MP_SELF_SPLIT_DIAG] > 0.5) { r->inv_diag = 1.0 / self_contact_split_diag(e, &k, Lm, &C[ci], C[ci].n); r->under = 1; }
    double rel = dotn(r->J, u, NU);
    double pen = C[ci].dist + m->P[MP_LINEAR_SLOP];
    double pos_err = 0, vel_err = -rel;
    if (pen > 0) vel_err -= pen / dt; else pos_err = -pen * m->P[MP_ERP_CONTACT] / dt;
    r->rhs = (pos_err + vel_err) * r->inv_diag;
    r->lo = 0; r->hi = 1e10; r->friction_of = -1;
    normal_row[ci] = nr++;
  }
  const int n_normal_end = nr;
  for (int ci = 0; ci < nc; ++ci) { /* two friction rows per contact, coupled by a cone */
    double t1[3], t2[3];
    plane_space(C[ci].n, t1, t2);
    for (int d = 0; d < 2; ++d) {
      Row* r = &rows[nr++];
      memset(r, 0, sizeof(*r));
      contact_jacobian(e, s, &k, &C[ci], d == 0 ? t1 : t2, r->J);
      apply_minv(e, Lm, Ibinv, r->J, r->W);
      r->inv_diag = 1.0 / dotn(r->J, r->W, NU);
      if (C[ci].self && m->P[MP_SELF_SPLIT_DIAG] > 0.5) { r->inv_diag = 1.0 / self_contact_split_diag(e, &k, Lm, &C[ci], d == 0 ? t1 : t2); r->under = 1; }
      r->rhs = -dotn(r->J, u, NU) * r->inv_diag;
      r->friction_of = normal_row[ci]; r->mu = C[ci].mu;
    }
  }
  /* ---- projected Gauss-Seidel (Bullet order: non-contact, normals, friction cones) ---- */
  double dv[NU] = {0};
  { /* warm start (Bullet SOLVER_USE_WARMSTARTING): contacts that persist from the previous sub-step start from
       warmstart_factor x their last impulses; motors and limits start from zero */
    const double wf = m->P[MP_WARMSTART];
    for (int j = 0; j < NL && wf > 0; ++j) { /* motor rows are rows 0..8 */
      Row* r = &rows[j];
      r->lambda = fmin(fmax(wf * s->wm[j], r->lo), r->hi);
      for (int a = 0; a < NU; ++a) dv[a] += r->W[a] * r->lambda;
    }
    for (int ci = 0; ci < nc && wf > 0; ++ci)
      for (int w = 0; w < s->wn; ++w)
        if (s->wid[w] == C[ci].id) {
          Row* rn = &rows[normal_row[ci]];
          Row* ra = &rows[n_normal_end + 2 * ci];
          Row* rb = ra + 1;
          rn->lambda = wf * s->wl[w][0]; ra->lambda = wf * s->wl[w][1]; rb->lambda = wf * s->wl[w][2];
          for (int a = 0; a < NU; ++a) dv[a] += rn->W[a] * rn->lambda + ra->W[a] * ra->lambda + rb->W[a] * rb->lambda;
          break;
        }
  }
  int max_it = (int)m->P[MP_SOLVER_ITERS];
  /* Iteration compression (the CUDA kernel's schedule, OFF in the oracle's default = Bullet's plain loop).  The
   * same-multibody rows are under-relaxed by 3e-4 .. 1e-3, so over the 150 iterations their impulses grow as a nearly
   * linear ramp that the other rows track with one iteration of lag.  A ramp of K-fold steps reaches the same point in
   * 1/K of the iterations; the last MP_PGS_TAIL iterations run with the true step so that the tracking lag at the end is
   * Bullet's.  Equivalent iteration count = K n_c + tail = MP_SOLVER_ITERS.  Error: O((K - 1) / N) of the ramp's
   * second-order term, measured in tests/test_oracle_physics.py and tests/test_gpu_physics.py. */
  double Kc = e->kernel_schedule ? m->P[MP_PGS_COMPRESS] : 0.0;
  int n_c = 0, has_under = 0;
  for (int ri = 0; ri < nr; ++ri) has_under |= rows[ri].under;
  if (Kc > 1.0 && has_under) {
    int tail = (int)m->P[MP_PGS_TAIL];
    n_c = (int)((max_it - tail) / Kc);
    max_it = n_c + (max_it - (int)(n_c * Kc));
  }
  int it;
  for (it = 0; it < max_it; ++it) {
    double resid = 0;
    const double kf_it = it < n_c ? Kc : 1.0;
    for (int rj = 0; rj < n_normal_end; ++rj) {
      /* Bullet sweeps the non-contact rows backwards on even iterations (btMultiBodyConstraintSolver::solveSingleIteration) */
      int ri = rj;
      if (rj < n_noncontact && m->P[MP_SWEEP_ALTERNATE] > 0.5 && (it & 1) == 0) ri = n_noncontact - 1 - rj;
      Row* r = &rows[ri];
      double d = r->rhs - dotn(r->J, dv, NU) * r->inv_diag;
      if (r->under) d *= kf_it;
      double sum = r->lambda + d;
      if (sum < r->lo) { d = r->lo - r->lambda; sum = r->lo; }
      else if (sum > r->hi) { d = r->hi - r->lambda; sum = r->hi; }
      r->lambda = sum;
      for (int a = 0; a < NU; ++a) dv[a] += r->W[a] * d;
      double res = d / r->inv_diag;
      if (res * res > resid) resid = res * res;
    }
    for (int ri = n_normal_end; ri < nr; ri += 2) {
      Row *ra = &rows[ri], *rb = &rows[ri + 1];
      double lim = ra->mu * rows[ra->friction_of].lambda;
      double da = ra->rhs - dotn(ra->J, dv, NU) * ra->inv_diag;
      double db = rb->rhs - dotn(rb->J, dv, NU) * rb->inv_diag;
      if (ra->under) { da *= kf_it; db *= kf_it; }
      double sa = ra->lambda + da, sb = rb->lambda + db;
      double nrm = sqrt(sa * sa + sb * sb);
      if (nrm > lim) { double sc = nrm > 0 ? lim / nrm : 0; sa *= sc; sb *= sc; }
      da = sa - ra->lambda; db = sb - rb->lambda;
      ra->lambda = sa; rb->lambda = sb;
      for (int a = 0; a < NU; ++a) dv[a] += ra->W[a] * da + rb->W[a] * db;
      double r1 = da / ra->inv_diag, r2 = db / rb->inv_diag;
      if (r1 * r1 > resid) resid = r1 * r1;
      if (r2 * r2 > resid) resid = r2 * r2;
    }
    if (it < 256) e->resid_hist[it] = resid;
    if (resid <= m->P[MP_RESIDUAL_THRESH]) { ++it; break; }
  }
  e->last_rows = nr; e->last_iters = it; e->last_contacts = nc;
  s->wn = nc;
  for (int j = 0; j < NL; ++j) s->wm[j] = rows[j].lambda;
  for (int ci = 0; ci < nc; ++ci) {
    s->wid[ci] = C[ci].id;
    s->wl[ci][0] = rows[normal_row[ci]].lambda;
    s->wl[ci][1] = rows[n_normal_end + 2 * ci].lambda;
    s->wl[ci][2] = rows[n_normal_end + 2 * ci + 1].lambda;
  }
  /* ---- integrate ---- */
  for (int i = 0; i < NL; ++i) { s->qd[i] = u[i] + dv[i]; s->q[i] += dt * s->qd[i]; }
  for (int a = 0; a < 3; ++a) { s->bv[a] = u[9 + a] + dv[9 + a]; s->bw[a] = u[12 + a] + dv[12 + a]; s->bp[a] += dt * s->bv[a]; }
  { /* quaternion exponential map (btTransformUtil::integrateTransform) */
    double wn = v3norm(s->bw), ax[3], th = wn * dt;
    if (wn < 1e-12) { ax[0] = s->bw[0] * 0.5 * dt; ax[1] = s->bw[1] * 0.5 * dt; ax[2] = s->bw[2] * 0.5 * dt; }
    else { double sc = sin(0.5 * th) / wn; ax[0] = s->bw[0] * sc; ax[1] = s->bw[1] * sc; ax[2] = s->bw[2] * sc; }
    double dq[4] = {ax[0], ax[1], ax[2], cos(0.5 * th)}, q0[4] = {s->bq[0], s->bq[1], s->bq[2], s->bq[3]};
    double r[4];
    r[3] = dq[3] * q0[3] - dq[0] * q0[0] - dq[1] * q0[1] - dq[2] * q0[2];
    r[0] = dq[3] * q0[0] + dq[0] * q0[3] + dq[1] * q0[2] - dq[2] * q0[1];
    r[1] = dq[3] * q0[1] - dq[0] * q0[2] + dq[1] * q0[3] + dq[2] * q0[0];
    r[2] = dq[3] * q0[2] + dq[0] * q0[1] - dq[1] * q0[0] + dq[2] * q0[3];
    double nn = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
    for (int a = 0; a < 4; ++a) s->bq[a] = r[a] / nn;
  }
}

/* ------------------------------------------------------------------ observation (27 floats) */
static void observe(const Env* e, const State* s, double* obs, double* ag) {
  const Model* m = &e->m;
  Kin k;
  fk(m, s->q, &k);
  int ee = (int)m->P[MP_EE_LINK];
  double w[3] = {0, 0, 0}, vo[3] = {0, 0, 0};
  /* velocity of the EE link: accumulate along the chain */
  int chain[NL], n = 0;
  for (int i = ee; i >= 0; i = m->parent[i]) chain[n++] = i;
  double pprev[3] = {0, 0, 0};
  int first = 1;
  for (int c = n - 1; c >= 0; --c) {
    int i = chain[c];
    if (!first) { double r[3], t[3]; v3sub(r, k.p[i], pprev); v3cross(t, w, r); v3add(vo, vo, t); }
    v3axpy(w, s->qd[i], k.z[i]);
    v3cpy(pprev, k.p[i]);
    first = 0;
  }
  double rc[3], t[3], vcom[3], eul[3];
  v3sub(rc, k.c[ee], k.p[ee]);
  v3cross(t, w, rc);
  v3add(vcom, vo, t); /* PyBullet reports the link velocity at the COM */
  mat_to_euler(eul, k.R[ee]);
  for (int a = 0; a < 3; ++a) {
    obs[a] = k.p[ee][a];
    obs[3 + a] = eul[a];
    obs[6 + a] = vcom[a];
    obs[9 + a] = w[a];
    obs[12 + a] = s->bp[a];
    obs[15 + a] = eul[a]; /* the reference's blockOrn slot repeats the gripper euler (push_F.py:188) */
    obs[18 + a] = s->bp[a] - k.p[ee][a];
    obs[21 + a] = s->bv[a];
    obs[24 + a] = s->bw[a];
    ag[a] = s->bp[a];
  }
}

/* ------------------------------------------------------------------ public API (ctypes) */
typedef struct { Env env; State st; } Handle;

void* bmo_create(const float* blob, int64_t n, int task) {
  Handle* h = (Handle*)calloc(1, sizeof(Handle));
  if (load_model(&h->env.m, blob, n) != 0) { free(h); return NULL; }
  h->env.task = task;
  const double* P = h->env.m.P;
  int o = task == 0 ? MP_PUSH_HX : MP_PICK_HX;
  for (int a = 0; a < 3; ++a) h->env.bh[a] = P[o + a];
  h->env.bmass = P[o + 3];
  h->env.bmu = P[o + 4];
  double lx = 2 * h->env.bh[0], ly = 2 * h->env.bh[1], lz = 2 * h->env.bh[2], mm = h->env.bmass / 12.0;
  h->env.binertia[0] = mm * (ly * ly + lz * lz);
  h->env.binertia[1] = mm * (lx * lx + lz * lz);
  h->env.binertia[2] = mm * (lx * lx + ly * ly);
  h->st.bq[3] = 1;
  return h;
}
void bmo_destroy(void* hp) {
  Handle* h = (Handle*)hp;
  for (int i = 0; i < h->env.m.nh; ++i) free(h->env.m.h_v[i]);
  free(hp);
}
/* full-resolution hulls: data = [n_hulls, then per hull: link (-1 = right_link1, rigid with the base), friction, n_verts,
 * 3 n_verts coordinates in the link frame] */
int bmo_set_hulls(void* hp, const float* data, int64_t n) {
  Model* m = &((Handle*)hp)->env.m;
  int64_t o = 0;
  if (n < 1) return -1;
  int nh = (int)data[o++];
  if (nh > MAX_HULLS) return -2;
  for (int i = 0; i < nh; ++i) {
    if (o + 3 > n) return -3;
    m->h_link[i] = (int)data[o++]; m->h_mu[i] = data[o++]; m->h_nv[i] = (int)data[o++];
    int nv = m->h_nv[i];
    if (o + 3 * (int64_t)nv > n) return -3;
    m->h_v[i] = (double*)malloc(sizeof(double) * 3 * nv);
    double lo[3] = {1e30, 1e30, 1e30}, hi[3] = {-1e30, -1e30, -1e30};
    for (int k = 0; k < 3 * nv; ++k) {
      double v = data[o++];
      m->h_v[i][k] = v;
      if (v < lo[k % 3]) lo[k % 3] = v;
      if (v > hi[k % 3]) hi[k % 3] = v;
    }
    double r = 0;
    for (int c = 0; c < 3; ++c) m->h_c[i][c] = 0.5 * (lo[c] + hi[c]);
    for (int k = 0; k < nv; ++k) { double d[3]; v3sub(d, m->h_v[i] + 3 * k, m->h_c[i]); if (v3norm(d) > r) r = v3norm(d); }
    m->h_r[i] = r;
  }
  m->nh = nh;
  return 0;
}
void bmo_set_param(void* hp, int idx, double v) { ((Handle*)hp)->env.m.P[idx] = v; }
double bmo_get_param(void* hp, int idx) { return ((Handle*)hp)->env.m.P[idx]; }

/* init8 = block x,y,z,yaw, goal x,y,z, unused   (bmirobot_env_push_F.py:113-160) */
void bmo_reset(void* hp, const double* init8, double* obs, double* ag, double* g) {
  Handle* h = (Handle*)hp;
  State* s = &h->st;
  memset(s, 0, sizeof(*s));
  v3set(s->bp, init8[0], init8[1], init8[2]);
  s->bq[2] = sin(0.5 * init8[3]); s->bq[3] = cos(0.5 * init8[3]);
  v3set(s->goal, init8[4], init8[5], init8[6]);
  observe(&h->env, s, obs, ag);
  v3cpy(g, s->goal);
}

void bmo_step(void* hp, const double* action, double* obs, double* ag, double* reward, double* success) {
  Handle* h = (Handle*)hp;
  Env* e = &h->env;
  State* s = &h->st;
  const Model* m = &e->m;
  double a[4];
  for (int i = 0; i < 4; ++i) a[i] = fmin(fmax(action[i], -0.5), 0.5);
  if (e->task == 0) a[3] = 0; /* push: bmirobot_env_push_F.py:94 */
  Kin k;
  fk(m, s->q, &k);
  int ee = (int)m->P[MP_EE_LINK];
  if (e->task == 1) { /* pick: auto-grip when any arm shape is within 1e-4 of the block (pickandplace_v2.py:94-95) */
    Contact C[MAX_CONTACTS];
    double save = e->m.P[MP_BLOCK_MARGIN];
    e->m.P[MP_BLOCK_MARGIN] = 1e-4;
    int nc = find_contacts(e, s, &k, C, 0);
    e->m.P[MP_BLOCK_MARGIN] = save;
    for (int i = 0; i < nc; ++i) if (C[i].has_block && C[i].link >= 0 && C[i].dist < 1e-4) { a[3] = -1; break; }
  }
  /* applyAction (bmirobot.py:129-162) */
  double target[3] = {fmin(fmax(k.p[ee][0] + a[0], -1.0), 1.0), fmin(fmax(k.p[ee][1] + a[1], -1.0), 1.0),
                      fmin(fmax(k.p[ee][2] + a[2], 0.0), 1.0)};
  double qik[NL];
  solve_ik(m, s->q, target, qik);
  for (int j = 0; j < 7; ++j) s->qt[j] = qik[j];
  s->qt[7] = s->q[7] + a[3]; /* sent_hand_moving (bmirobot.py:163-191) */
  s->qt[8] = s->q[8] - a[3];
  int nsub = (int)m->P[MP_N_SUBSTEPS];
  for (int i = 0; i < nsub; ++i) substep(e, s);
  observe(e, s, obs, ag);
  double d[3];
  v3sub(d, ag, s->goal);
  double dist = sqrt((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
  *success = dist < m->P[MP_DIST_THRESHOLD] ? 1.0 : 0.0;
  *reward = dist > m->P[MP_DIST_THRESHOLD] ? -1.0 : -0.0;
}

void bmo_get_state(void* hp, double* st48) {
  State* s = &((Handle*)hp)->st;
  memset(st48, 0, sizeof(double) * 48);
  memcpy(st48 + ST_Q, s->q, sizeof(s->q)); memcpy(st48 + ST_QD, s->qd, sizeof(s->qd)); memcpy(st48 + ST_QT, s->qt, sizeof(s->qt));
  memcpy(st48 + ST_BPOS, s->bp, 24); memcpy(st48 + ST_BQUAT, s->bq, 32); memcpy(st48 + ST_BVEL, s->bv, 24);
  memcpy(st48 + ST_BANG, s->bw, 24); memcpy(st48 + ST_GOAL, s->goal, 24);
}
void bmo_set_state(void* hp, const double* st48) {
  State* s = &((Handle*)hp)->st;
  memcpy(s->q, st48 + ST_Q, sizeof(s->q)); memcpy(s->qd, st48 + ST_QD, sizeof(s->qd)); memcpy(s->qt, st48 + ST_QT, sizeof(s->qt));
  memcpy(s->bp, st48 + ST_BPOS, 24); memcpy(s->bq, st48 + ST_BQUAT, 32); memcpy(s->bv, st48 + ST_BVEL, 24);
  memcpy(s->bw, st48 + ST_BANG, 24); memcpy(s->goal, st48 + ST_GOAL, 24);
}
void bmo_stats(void* hp, int* out3) {
  Env* e = &((Handle*)hp)->env;
  out3[0] = e->last_rows; out3[1] = e->last_iters; out3[2] = e->last_contacts;
}
/* expose pieces for unit tests */
void bmo_ik(void* hp, const double* q0, const double* target, double* qout) { solve_ik(&((Handle*)hp)->env.m, q0, target, qout); }
void bmo_fk_ee(void* hp, const double* q, double* pos3, double* euler3) {
  Kin k;
  fk(&((Handle*)hp)->env.m, q, &k);
  int ee = (int)((Handle*)hp)->env.m.P[MP_EE_LINK];
  v3cpy(pos3, k.p[ee]);
  mat_to_euler(euler3, k.R[ee]);
}
void bmo_mass_matrix(void* hp, const double* q, double* M81) {
  const Model* m = &((Handle*)hp)->env.m;
  Kin k;
  fk(m, q, &k);
  double zero[NL] = {0}, ej[NL], col[NL];
  for (int j = 0; j < NL; ++j) {
    memset(ej, 0, sizeof(ej)); ej[j] = 1;
    rnea(m, &k, zero, ej, 0, 0, 0, col);
    for (int i = 0; i < NL; ++i) M81[i * NL + j] = col[i];
  }
}

void bmo_resid_hist(void* hp, double* out256) { memcpy(out256, ((Handle*)hp)->env.resid_hist, sizeof(double) * 256); }

/* debug / tests: contacts of the current state: out[i*12..] = link1, link, has_block, dist, n(3), x(3), mu, id */
int bmo_contacts(void* hp, double* out, int cap) {
  Handle* h = (Handle*)hp;
  Kin k;
  fk(&h->env.m, h->st.q, &k);
  Contact C[MAX_CONTACTS];
  int nc = 0;
  if (h->env.m.P[MP_SELF_COLLISION] > 0.5) nc = find_self_contacts(&h->env, &h->st, &k, C, nc);
  nc = find_contacts(&h->env, &h->st, &k, C, nc);
  for (int i = 0; i < nc && i < cap; ++i) {
    double* o = out + 12 * i;
    o[0] = C[i].link1; o[1] = C[i].link; o[2] = C[i].has_block; o[3] = C[i].dist;
    for (int a = 0; a < 3; ++a) { o[4 + a] = C[i].n[a]; o[7 + a] = C[i].x[a]; }
    o[10] = C[i].mu; o[11] = C[i].id;
  }
  return nc;
}

/* the kernel's baked pair tables (format: include/bmi_model.h SC_*); the pointer must stay valid */
void bmo_set_selfcol_table(void* hp, const float* data, int64_t n) {
  Model* m = &((Handle*)hp)->env.m;
  m->sc_table = data; m->sc_n = n;
}
/* table baker: exact pair evaluation at joint angles q9 -> out8 = core distance, n(3) and witness xa(3) in the frame of
 * hull a's link, pad; distance 1e3 when the cores are farther apart than `far` */
void bmo_pair_query(void* hp, int a, int b, const double* q9, double far, float* out8) {
  Model* m = &((Handle*)hp)->env.m;
  Kin k;
  fk(m, q9, &k);
  double dist, n[3], pa[3], pb[3];
  for (int r = 0; r < 8; ++r) out8[r] = 0.f;
  out8[0] = 1e3f;
  if (!hull_pair(m, &k, a, b, far, &dist, n, pa, pb)) return;
  Cvx A; hull_world(m, &k, a, &A);
  double t[3], nl[3], xl[3];
  v3sub(t, pa, A.p); m3tvec(xl, A.R, t); m3tvec(nl, A.R, n);
  out8[0] = (float)dist;
  for (int r = 0; r < 3; ++r) { out8[1 + r] = (float)nl[r]; out8[4 + r] = (float)xl[r]; }
}

/* kernel-vs-oracle tests: apply the CUDA kernel's lane budget (contacts in total / on arm links), 0 = keep everything */
void bmo_set_caps(void* hp, int cap_contacts, int cap_arm) {
  Env* e = &((Handle*)hp)->env;
  e->cap_c = cap_contacts; e->cap_a = cap_arm;
}
/* kernel-vs-oracle tests: run the solver with the kernel's compressed iteration schedule (model params MP_PGS_COMPRESS /
 * MP_PGS_TAIL) instead of Bullet's plain MP_SOLVER_ITERS loop */
void bmo_set_kernel_schedule(void* hp, int on) { ((Handle*)hp)->env.kernel_schedule = on; }
