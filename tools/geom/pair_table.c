/* Build-time sampler of the arm's self-collision pair tables (tools/bake_selfcol.py -> assets/bmirobot_selfcol.bin).
 *
 * Product tooling, independent of oracle/: forward kinematics of the baked joint tree (include/bmi_model.h) for a grid of
 * two joint angles, then the exact narrow phase of two convex hulls (GJK distance / EPA penetration, convex_epa.h -- the
 * counterpart of Bullet's btGjkPairDetector + btGjkEpaPenetrationDepthSolver behind p.loadURDF(..., flags=9),
 * bmirobot.py:58).  Per node: core distance (gap > 0 or -penetration depth, 1e3 = far), unit normal from hull B towards
 * hull A and the witness point on A, both in the frame of link A.
 *
 *   gcc -O2 -shared -fPIC -o libpairtable.so pair_table.c -lm
 */
#include <math.h>
#include <string.h>

#include "../../include/bmi_model.h"
#include "convex_epa.h"

#define PT_NL BMI_MAX_LINKS

static void pt_mul(double* o, const double* a, const double* b) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  memcpy(o, t, sizeof(t));
}
static void pt_vec(double* o, const double* a, const double* v) {
  const double x = a[0] * v[0] + a[1] * v[1] + a[2] * v[2], y = a[3] * v[0] + a[4] * v[1] + a[5] * v[2],
               z = a[6] * v[0] + a[7] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void pt_tvec(double* o, const double* a, const double* v) {
  const double x = a[0] * v[0] + a[3] * v[1] + a[6] * v[2], y = a[1] * v[0] + a[4] * v[1] + a[7] * v[2],
               z = a[2] * v[0] + a[5] * v[1] + a[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
static void pt_rodrigues(double* R, const double* u, double th) {
  const double c = cos(th), s = sin(th), C = 1 - c;
  R[0] = c + u[0] * u[0] * C;        R[1] = u[0] * u[1] * C - u[2] * s; R[2] = u[0] * u[2] * C + u[1] * s;
  R[3] = u[1] * u[0] * C + u[2] * s; R[4] = c + u[1] * u[1] * C;        R[5] = u[1] * u[2] * C - u[0] * s;
  R[6] = u[2] * u[0] * C - u[1] * s; R[7] = u[2] * u[1] * C + u[0] * s; R[8] = c + u[2] * u[2] * C;
}

/* world pose (R, p) of every link of the blob's joint tree at joint angles q */
static void pt_fk(const float* blob, const double* q, double (*R)[9], double (*p)[3]) {
  const int off = (int)blob[MP_LINKS_OFF], nl = (int)blob[MP_N_LINKS];
  for (int i = 0; i < nl && i < PT_NL; ++i) {
    const float* L = blob + off + BMI_LINK_STRIDE * i;
    double axis[3], jrot[9], jpos[3], Rq[9], Rl[9];
    for (int r = 0; r < 3; ++r) { axis[r] = L[ML_AXIS + r]; jpos[r] = L[ML_JPOS + r]; }
    for (int r = 0; r < 9; ++r) jrot[r] = L[ML_JROT + r];
    pt_rodrigues(Rq, axis, q[i]);
    pt_mul(Rl, jrot, Rq);
    const int pa = (int)L[ML_PARENT];
    if (pa < 0) {
      memcpy(R[i], Rl, sizeof(Rl));
      p[i][0] = blob[MP_BASE_PX] + jpos[0]; p[i][1] = blob[MP_BASE_PY] + jpos[1]; p[i][2] = blob[MP_BASE_PZ] + jpos[2];
    } else {
      double t[3];
      pt_mul(R[i], R[pa], Rl);
      pt_vec(t, R[pa], jpos);
      for (int r = 0; r < 3; ++r) p[i][r] = p[pa][r] + t[r];
    }
  }
}

static void pt_pose(const float* blob, double (*R)[9], double (*p)[3], int link, Cvx* c) {
  if (link >= 0) { memcpy(c->R, R[link], sizeof(c->R)); memcpy(c->p, p[link], sizeof(c->p)); return; }
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   /* right_link1: rigid with the fixed base */
  memcpy(c->R, I, sizeof(I));
  c->p[0] = blob[MP_BASE_PX]; c->p[1] = blob[MP_BASE_PY]; c->p[2] = blob[MP_BASE_PZ];
}

static void pt_sphere(const double* v, int nv, double* c, double* r) {
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int k = 0; k < nv; ++k)
    for (int a = 0; a < 3; ++a) { if (v[3 * k + a] < lo[a]) lo[a] = v[3 * k + a]; if (v[3 * k + a] > hi[a]) hi[a] = v[3 * k + a]; }
  for (int a = 0; a < 3; ++a) c[a] = 0.5 * (lo[a] + hi[a]);
  *r = 0;
  for (int k = 0; k < nv; ++k) {
    double d[3]; s3(d, v + 3 * k, c);
    const double l = sqrt(d3(d, d));
    if (l > *r) *r = l;
  }
}

/* out[(i * nqb + j) * 8 ..]: node (qa_vals[i], qb_vals[j]) of the pair (hull A on link_a, hull B on link_b; -1 = base link);
 * every other joint at zero */
void pt_bake_rows(const float* blob, const double* va, int na, int link_a, const double* vb, int nb, int link_b, int ja,
                  int jb, const double* qa_vals, int nqa, const double* qb_vals, int nqb, double far, float* out) {
  double ca[3], cb[3], ra, rb;
  pt_sphere(va, na, ca, &ra);
  pt_sphere(vb, nb, cb, &rb);
  for (int i = 0; i < nqa; ++i)
    for (int j = 0; j < nqb; ++j) {
      float* o = out + ((long)i * nqb + j) * 8;
      for (int r = 0; r < 8; ++r) o[r] = 0.f;
      o[0] = 1e3f;
      double q[PT_NL] = {0}, R[PT_NL][9], p[PT_NL][3];
      q[ja] = qa_vals[i]; q[jb] = qb_vals[j];
      pt_fk(blob, q, R, p);
      Cvx A, B;
      A.nv = na; A.v = va; B.nv = nb; B.v = vb;
      pt_pose(blob, R, p, link_a, &A);
      pt_pose(blob, R, p, link_b, &B);
      double wa[3], wb[3], t[3], d[3];
      pt_vec(t, A.R, ca); for (int r = 0; r < 3; ++r) wa[r] = A.p[r] + t[r];
      pt_vec(t, B.R, cb); for (int r = 0; r < 3; ++r) wb[r] = B.p[r] + t[r];
      s3(d, wa, wb);
      if (sqrt(d3(d, d)) > ra + rb + far) continue;
      double n[3], ne[3], pa[3], pb[3], dist;
      const double depth = epa_penetration(&A, &B, ne, pa, pb);
      if (depth >= 0) {
        for (int r = 0; r < 3; ++r) n[r] = -ne[r];
        dist = -depth;
      } else {
        const double gap = gjk_distance(&A, &B, pa, pb);
        if (gap < 0 || gap > far) continue;
        for (int r = 0; r < 3; ++r) n[r] = (pa[r] - pb[r]) / gap;
        dist = gap;
      }
      double xl[3], nl[3];
      s3(t, pa, A.p); pt_tvec(xl, A.R, t); pt_tvec(nl, A.R, n);
      o[0] = (float)dist;
      for (int r = 0; r < 3; ++r) { o[1 + r] = (float)nl[r]; o[4 + r] = (float)xl[r]; }
    }
}

/* Narrow phase of two posed hulls (world = R v + p), for tests: returns 1 and dist / n (from B towards A) / witness points,
 * or 0 when the hulls are farther apart than `far`. */
int pt_pair_world(const double* va, int na, const double* Ra, const double* pa3, const double* vb, int nb, const double* Rb,
                  const double* pb3, double far, double* dist, double* n, double* wa, double* wb) {
  Cvx A, B;
  A.nv = na; A.v = va; memcpy(A.R, Ra, sizeof(A.R)); memcpy(A.p, pa3, sizeof(A.p));
  B.nv = nb; B.v = vb; memcpy(B.R, Rb, sizeof(B.R)); memcpy(B.p, pb3, sizeof(B.p));
  double ne[3];
  const double depth = epa_penetration(&A, &B, ne, wa, wb);
  if (depth >= 0) {
    for (int r = 0; r < 3; ++r) n[r] = -ne[r];
    *dist = -depth;
    return 1;
  }
  const double gap = gjk_distance(&A, &B, wa, wb);
  if (gap < 0 || gap > far) return 0;
  for (int r = 0; r < 3; ++r) n[r] = (wa[r] - wb[r]) / gap;
  *dist = gap;
  return 1;
}
