/* GJK intersection test + EPA penetration depth for two convex vertex sets (double precision).
 * Standard algorithms (Gilbert-Johnson-Keerthi; expanding polytope, van den Bergen 2001). */
#include <math.h>
#include <string.h>
typedef struct { int nv; const double* v; double R[9], p[3]; } Cvx;   /* world = R v + p */
typedef struct { double w[3], a[3], b[3]; } SV;                     /* support vertex of A - B and its witnesses */

static inline double d3(const double* a, const double* b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
static inline void c3(double* o, const double* a, const double* b) {
  double x=a[1]*b[2]-a[2]*b[1], y=a[2]*b[0]-a[0]*b[2], z=a[0]*b[1]-a[1]*b[0]; o[0]=x;o[1]=y;o[2]=z; }
static inline void s3(double* o, const double* a, const double* b) { o[0]=a[0]-b[0];o[1]=a[1]-b[1];o[2]=a[2]-b[2]; }

static void cvx_support(const Cvx* c, const double* d, double* out) {
  double dl[3] = { c->R[0]*d[0]+c->R[3]*d[1]+c->R[6]*d[2], c->R[1]*d[0]+c->R[4]*d[1]+c->R[7]*d[2], c->R[2]*d[0]+c->R[5]*d[1]+c->R[8]*d[2] };
  int best = 0; double bd = -1e300;
  for (int i = 0; i < c->nv; ++i) { double t = d3(c->v + 3*i, dl); if (t > bd) { bd = t; best = i; } }
  const double* v = c->v + 3*best;
  for (int r = 0; r < 3; ++r) out[r] = c->R[3*r]*v[0] + c->R[3*r+1]*v[1] + c->R[3*r+2]*v[2] + c->p[r];
}
static void mk_support(const Cvx* A, const Cvx* B, const double* d, SV* s) {
  double nd[3] = {-d[0], -d[1], -d[2]};
  cvx_support(A, d, s->a); cvx_support(B, nd, s->b); s3(s->w, s->a, s->b);
}

/* GJK: returns 1 and a tetrahedron enclosing the origin when A and B intersect, else 0 */
static int gjk_intersect(const Cvx* A, const Cvx* B, SV* simp /*4*/) {
  double d[3]; s3(d, B->p, A->p);
  if (d3(d,d) < 1e-20) { d[0]=1; d[1]=0; d[2]=0; }
  int n = 0;
  mk_support(A, B, d, &simp[0]); n = 1;
  d[0] = -simp[0].w[0]; d[1] = -simp[0].w[1]; d[2] = -simp[0].w[2];
  for (int it = 0; it < 128; ++it) {
    if (d3(d,d) < 1e-30) { /* origin on the simplex: treat as touching -> expand with an arbitrary direction */ d[0]=1e-3; d[1]=2e-3; d[2]=3e-3; }
    SV nw; mk_support(A, B, d, &nw);
    if (d3(nw.w, d) < 0) return 0;
    /* push front */
    for (int k = n; k > 0; --k) simp[k] = simp[k-1];
    simp[0] = nw; ++n;
    double ao[3] = {-simp[0].w[0], -simp[0].w[1], -simp[0].w[2]};
    if (n == 2) {
      double ab[3]; s3(ab, simp[1].w, simp[0].w);
      if (d3(ab, ao) > 0) { double t[3]; c3(t, ab, ao); c3(d, t, ab); }
      else { n = 1; memcpy(d, ao, sizeof(ao)); }
    } else if (n == 3) {
      double ab[3], ac[3], abc[3], t[3];
      s3(ab, simp[1].w, simp[0].w); s3(ac, simp[2].w, simp[0].w); c3(abc, ab, ac);
      c3(t, abc, ac);
      if (d3(t, ao) > 0) {
        if (d3(ac, ao) > 0) { simp[1] = simp[2]; n = 2; double u[3]; c3(u, ac, ao); c3(d, u, ac); }
        else { if (d3(ab, ao) > 0) { n = 2; double u[3]; c3(u, ab, ao); c3(d, u, ab); } else { n = 1; memcpy(d, ao, sizeof(ao)); } }
      } else {
        c3(t, ab, abc);
        if (d3(t, ao) > 0) { if (d3(ab, ao) > 0) { n = 2; double u[3]; c3(u, ab, ao); c3(d, u, ab); } else { n = 1; memcpy(d, ao, sizeof(ao)); } }
        else {
          if (d3(abc, ao) > 0) memcpy(d, abc, sizeof(abc));
          else { SV tmp = simp[1]; simp[1] = simp[2]; simp[2] = tmp; d[0]=-abc[0]; d[1]=-abc[1]; d[2]=-abc[2]; }
        }
      }
    } else { /* tetrahedron a=0,b=1,c=2,d=3 */
      double ab[3], ac[3], ad[3], abc[3], acd[3], adb[3];
      s3(ab, simp[1].w, simp[0].w); s3(ac, simp[2].w, simp[0].w); s3(ad, simp[3].w, simp[0].w);
      c3(abc, ab, ac); c3(acd, ac, ad); c3(adb, ad, ab);
      /* orient the face normals outward (away from the opposite vertex) */
      if (d3(abc, ad) > 0) { abc[0]=-abc[0]; abc[1]=-abc[1]; abc[2]=-abc[2]; }
      if (d3(acd, ab) > 0) { acd[0]=-acd[0]; acd[1]=-acd[1]; acd[2]=-acd[2]; }
      if (d3(adb, ac) > 0) { adb[0]=-adb[0]; adb[1]=-adb[1]; adb[2]=-adb[2]; }
      if (d3(abc, ao) > 0) { n = 3; memcpy(d, abc, sizeof(abc)); }                              /* keep a,b,c */
      else if (d3(acd, ao) > 0) { simp[1] = simp[2]; simp[2] = simp[3]; n = 3; memcpy(d, acd, sizeof(acd)); }
      else if (d3(adb, ao) > 0) { simp[2] = simp[1]; simp[1] = simp[3]; n = 3; memcpy(d, adb, sizeof(adb)); }
      else return 1;
      /* re-run the triangle case logic next iteration through the general support step: d already points to the origin side */
    }
  }
  return 0;
}

#define EPA_MAXV 512
#define EPA_MAXF 1024
typedef struct { int v[3]; double n[3], d; int alive; } EFace;

static int epa_make_face(const SV* V, EFace* f, int a, int b, int c, const double* inside) {
  double ab[3], ac[3]; s3(ab, V[b].w, V[a].w); s3(ac, V[c].w, V[a].w);
  c3(f->n, ab, ac);
  double l = sqrt(d3(f->n, f->n));
  if (l < 1e-30) return -1;
  f->n[0] /= l; f->n[1] /= l; f->n[2] /= l;
  f->v[0] = a; f->v[1] = b; f->v[2] = c;
  f->d = d3(f->n, V[a].w);
  double t[3]; s3(t, inside, V[a].w);
  if (d3(f->n, t) > 0) { /* normal points to the interior: flip */
    f->n[0]=-f->n[0]; f->n[1]=-f->n[1]; f->n[2]=-f->n[2]; f->d = -f->d; int s = f->v[1]; f->v[1] = f->v[2]; f->v[2] = s;
  }
  f->alive = 1;
  return 0;
}

/* returns depth >= 0 (A must move by -n*depth, i.e. B by +n*depth... see below) or -1 when separated.
 * n: unit vector such that translating A by -depth*n separates the bodies (n points from B's side towards A? no:
 * w = a - b; the face normal n of the A-B polytope closest to the origin: moving A by -n*depth puts the origin on the boundary).
 * pa, pb: witness points on A and B in world coordinates (pb = pa - n*depth). */
static double epa_penetration(const Cvx* A, const Cvx* B, double* n, double* pa, double* pb) {
  static __thread SV V[EPA_MAXV]; static __thread EFace F[EPA_MAXF];
  SV simp[5];
  if (!gjk_intersect(A, B, simp)) return -1.0;
  int nv = 4, nf = 0;
  for (int i = 0; i < 4; ++i) V[i] = simp[i];
  double inside[3] = {0,0,0};
  for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) inside[k] += 0.25 * V[i].w[k];
  static const int T[4][3] = {{0,1,2},{0,2,3},{0,3,1},{1,3,2}};
  for (int i = 0; i < 4; ++i) if (epa_make_face(V, &F[nf], T[i][0], T[i][1], T[i][2], inside) == 0) ++nf;
  if (nf < 4) return -1.0; /* degenerate tetrahedron */
  int best = -1;
  for (int it = 0; it < 400; ++it) {
    best = -1;
    for (int i = 0; i < nf; ++i) if (F[i].alive && (best < 0 || F[i].d < F[best].d)) best = i;
    SV nw; mk_support(A, B, F[best].n, &nw);
    double sd = d3(nw.w, F[best].n);
    if (sd - F[best].d < 1e-10 || nv >= EPA_MAXV || nf >= EPA_MAXF - 64) break;
    /* remove faces visible from nw, collect the horizon */
    int E[256][2], ne = 0;
    V[nv] = nw;
    for (int i = 0; i < nf; ++i) if (F[i].alive) {
      double t[3]; s3(t, nw.w, V[F[i].v[0]].w);
      if (d3(F[i].n, t) > 1e-14) {
        F[i].alive = 0;
        for (int e = 0; e < 3; ++e) {
          int a = F[i].v[e], b = F[i].v[(e+1)%3], found = -1;
          for (int k = 0; k < ne; ++k) if (E[k][0] == b && E[k][1] == a) { found = k; break; }
          if (found >= 0) { E[found][0] = E[ne-1][0]; E[found][1] = E[ne-1][1]; --ne; }
          else if (ne < 256) { E[ne][0] = a; E[ne][1] = b; ++ne; }
        }
      }
    }
    if (ne == 0) break;
    for (int k = 0; k < ne; ++k) {
      /* reuse dead slots? keep it simple: append */
      if (epa_make_face(V, &F[nf], E[k][0], E[k][1], nv, inside) == 0) ++nf;
    }
    ++nv;
  }
  const EFace* f = &F[best];
  /* projection of the origin on the face, barycentric coordinates */
  const double *p0 = V[f->v[0]].w, *p1 = V[f->v[1]].w, *p2 = V[f->v[2]].w;
  double P[3] = {f->n[0]*f->d, f->n[1]*f->d, f->n[2]*f->d};
  double v0[3], v1[3], v2[3]; s3(v0, p1, p0); s3(v1, p2, p0); s3(v2, P, p0);
  double d00 = d3(v0,v0), d01 = d3(v0,v1), d11 = d3(v1,v1), d20 = d3(v2,v0), d21 = d3(v2,v1);
  double den = d00*d11 - d01*d01, l1 = (d11*d20 - d01*d21)/den, l2 = (d00*d21 - d01*d20)/den, l0 = 1 - l1 - l2;
  for (int k = 0; k < 3; ++k) {
    pa[k] = l0*V[f->v[0]].a[k] + l1*V[f->v[1]].a[k] + l2*V[f->v[2]].a[k];
    pb[k] = l0*V[f->v[0]].b[k] + l1*V[f->v[1]].b[k] + l2*V[f->v[2]].b[k];
    n[k] = f->n[k];
  }
  return f->d;
}

/* ---- GJK distance between two SEPARATED convex vertex sets (closest points), Ericson-style sub-simplex search ---- */
static void gjk_closest_on_segment(const double* a, const double* b, double* l) {
  double ab[3]; s3(ab, b, a);
  double t = -d3(a, ab) / d3(ab, ab);
  if (t <= 0) { l[0] = 1; l[1] = 0; } else if (t >= 1) { l[0] = 0; l[1] = 1; } else { l[0] = 1 - t; l[1] = t; }
}
static void gjk_closest_on_triangle(const double* a, const double* b, const double* c, double* l) {
  double ab[3], ac[3], ap[3] = {-a[0], -a[1], -a[2]};
  s3(ab, b, a); s3(ac, c, a);
  double d1 = d3(ab, ap), d2 = d3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { l[0] = 1; l[1] = 0; l[2] = 0; return; }
  double bp[3] = {-b[0], -b[1], -b[2]};
  double d3_ = d3(ab, bp), d4 = d3(ac, bp);
  if (d3_ >= 0 && d4 <= d3_) { l[0] = 0; l[1] = 1; l[2] = 0; return; }
  double vc = d1 * d4 - d3_ * d2;
  if (vc <= 0 && d1 >= 0 && d3_ <= 0) { double v = d1 / (d1 - d3_); l[0] = 1 - v; l[1] = v; l[2] = 0; return; }
  double cp[3] = {-c[0], -c[1], -c[2]};
  double d5 = d3(ab, cp), d6 = d3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { l[0] = 0; l[1] = 0; l[2] = 1; return; }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); l[0] = 1 - w; l[1] = 0; l[2] = w; return; }
  double va = d3_ * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3_) >= 0 && (d5 - d6) >= 0) { double w = (d4 - d3_) / ((d4 - d3_) + (d5 - d6)); l[0] = 0; l[1] = 1 - w; l[2] = w; return; }
  double den = 1.0 / (va + vb + vc);
  l[1] = vb * den; l[2] = vc * den; l[0] = 1 - l[1] - l[2];
}
/* returns the distance (> 0) and the closest points, or -1 when the sets touch / intersect */
static double gjk_distance(const Cvx* A, const Cvx* B, double* pa, double* pb) {
  SV S[4]; double l[4] = {1, 0, 0, 0}; int n = 1;
  double d[3]; s3(d, A->p, B->p);
  if (d3(d, d) < 1e-20) { d[0] = 1; d[1] = 0; d[2] = 0; }
  double nd[3] = {-d[0], -d[1], -d[2]};
  mk_support(A, B, nd, &S[0]);
  double v[3] = {S[0].w[0], S[0].w[1], S[0].w[2]};
  for (int it = 0; it < 200; ++it) {
    double vv = d3(v, v);
    if (vv < 1e-24) return -1.0;
    SV w; double nv_[3] = {-v[0], -v[1], -v[2]};
    mk_support(A, B, nv_, &w);
    if (vv - d3(v, w.w) <= 1e-14 * vv + 1e-18) break;        /* no progress possible: v is the closest point */
    int dup = 0;
    for (int i = 0; i < n; ++i) { double t[3]; s3(t, w.w, S[i].w); if (d3(t, t) < 1e-24) dup = 1; }
    if (dup) break;
    S[n++] = w;
    /* closest point of the simplex to the origin; reduce to the supporting sub-simplex */
    if (n == 2) { gjk_closest_on_segment(S[0].w, S[1].w, l); }
    else if (n == 3) { gjk_closest_on_triangle(S[0].w, S[1].w, S[2].w, l); }
    else {
      /* tetrahedron: origin inside -> intersecting; else the closest of the four faces */
      static const int Fc[4][3] = {{0,1,2},{0,1,3},{0,2,3},{1,2,3}};
      double best = 1e300, bl[4] = {0,0,0,0}; int inside = 1;
      for (int f = 0; f < 4; ++f) {
        const double *a = S[Fc[f][0]].w, *b = S[Fc[f][1]].w, *c = S[Fc[f][2]].w, *o = S[6 - Fc[f][0] - Fc[f][1] - Fc[f][2]].w;
        double ab[3], ac[3], nn[3], ao[3]; s3(ab, b, a); s3(ac, c, a); c3(nn, ab, ac); s3(ao, o, a);
        double so = d3(nn, ao), sp = -d3(nn, a);
        if (so * sp < 0 || fabs(so) < 1e-30) { /* origin on the outer side of this face */
          inside = 0;
          double tl[3]; gjk_closest_on_triangle(a, b, c, tl);
          double pt[3]; for (int k = 0; k < 3; ++k) pt[k] = tl[0]*a[k] + tl[1]*b[k] + tl[2]*c[k];
          double dd = d3(pt, pt);
          if (dd < best) { best = dd; bl[0]=bl[1]=bl[2]=bl[3]=0; bl[Fc[f][0]] = tl[0]; bl[Fc[f][1]] = tl[1]; bl[Fc[f][2]] = tl[2]; }
        }
      }
      if (inside) return -1.0;
      memcpy(l, bl, sizeof(bl));
    }
    int m = 0;
    for (int i = 0; i < n; ++i) if (l[i] > 0) { S[m] = S[i]; l[m] = l[i]; ++m; }
    n = m;
    v[0] = v[1] = v[2] = 0;
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) v[k] += l[i] * S[i].w[k];
  }
  for (int k = 0; k < 3; ++k) { pa[k] = 0; pb[k] = 0; for (int i = 0; i < n; ++i) { pa[k] += l[i] * S[i].a[k]; pb[k] += l[i] * S[i].b[k]; } }
  return sqrt(d3(v, v));
}
