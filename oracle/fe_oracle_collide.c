/* fe_oracle_collide.c -- CPU ORACLE collision stage (test infrastructure, NOT product code).
 *
 * Restates mj_collision for the geom types the furniture scenes use:
 *   filtering : compile-time pair list (same body / welded / parent-child / static-static / <exclude>) plus the
 *               run-time (contype1 & conaffinity2) || (contype2 & conaffinity1) test on the per-env masks that
 *               the reference rewrites (furniture.py:869-878, :1441-1461) and a bounding-sphere test;
 *   narrow    : plane-sphere, plane-cylinder, plane-box, sphere-sphere, sphere-box (analytic, as MuJoCo),
 *               box-box (SAT + face clipping / edge-edge; OWN algorithm -- MuJoCo's mjc_BoxBox is not restated),
 *               sphere-cylinder, cylinder-cylinder, cylinder-box through Minkowski Portal Refinement (the algorithm
 *               MuJoCo 2.0 delegates to libccd's ccdMPRPenetration: tolerance 1e-6, 50 iterations), one contact;
 *   parameters: condim = max, friction = max, solref / solimp mixed 50/50, margin = max (MuJoCo mj_contactParam).
 * Contact convention: frame[0:3] = normal pointing from geom1 to geom2, dist < 0 = penetration, pos = mid-surface.
 */
#include <math.h>
#include <string.h>

#include "fe_oracle.h"

#define MINVAL 1e-15
enum { G_PLANE = 0, G_SPHERE = 2, G_CAPSULE = 3, G_CYLINDER = 5, G_BOX = 6, G_MESH = 7 };

static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void sub3(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void addscl3(double* r, const double* a, const double* b, double s) { r[0] = a[0] + s * b[0]; r[1] = a[1] + s * b[1]; r[2] = a[2] + s * b[2]; }
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static inline void col3(double* r, const double* R, int k) { r[0] = R[k]; r[1] = R[3 + k]; r[2] = R[6 + k]; }

/* mju_makeFrame: complete the normal to a right-handed orthonormal frame */
static void make_frame(double* frame) {
  double* x = frame;
  double* y = frame + 3;
  double* z = frame + 6;
  normalize3(x);
  if (fabs(x[1]) < 0.5) { y[0] = 0; y[1] = 1; y[2] = 0; } else { y[0] = 0; y[1] = 0; y[2] = 1; }
  double dt = dot3(x, y);
  addscl3(y, y, x, -dt);
  normalize3(y);
  cross3(z, x, y);
}
static void set_contact(om_contact* c, double dist, const double* pos, const double* normal) {
  c->dist = dist;
  memcpy(c->pos, pos, 3 * sizeof(double));
  memcpy(c->frame, normal, 3 * sizeof(double));
  memset(c->frame + 3, 0, 6 * sizeof(double));
  make_frame(c->frame);
}

/* ---------------------------------------------------------------- analytic pairs */
static int plane_sphere(const double* pp, const double* pR, const double* c, double r, double margin, om_contact* out) {
  double n[3], t[3], pos[3];
  col3(n, pR, 2);
  sub3(t, c, pp);
  double dist = dot3(t, n) - r;
  if (dist >= margin) return 0;
  addscl3(pos, c, n, -(r + 0.5 * dist));
  set_contact(out, dist, pos, n);
  return 1;
}
static int plane_box(const double* pp, const double* pR, const double* c, const double* R, const double* s, double margin, om_contact* out) {
  double n[3];
  col3(n, pR, 2);
  int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; i++) {
    double loc[3] = {(i & 1 ? s[0] : -s[0]), (i & 2 ? s[1] : -s[1]), (i & 4 ? s[2] : -s[2])}, corner[3], t[3], pos[3];
    for (int k = 0; k < 3; k++) corner[k] = c[k] + R[3 * k] * loc[0] + R[3 * k + 1] * loc[1] + R[3 * k + 2] * loc[2];
    sub3(t, corner, pp);
    double dist = dot3(t, n);
    if (dist >= margin) continue;
    addscl3(pos, corner, n, -0.5 * dist);
    set_contact(out + cnt, dist, pos, n);
    cnt++;
  }
  return cnt;
}
static int plane_cylinder(const double* pp, const double* pR, const double* c, const double* R, double r, double h, double margin, om_contact* out) {
  double n[3], axis[3], vec[3], t[3], pos[3], p[3];
  col3(n, pR, 2);
  col3(axis, R, 2);
  double prj = dot3(n, axis);
  if (prj > 0) { axis[0] = -axis[0]; axis[1] = -axis[1]; axis[2] = -axis[2]; prj = -prj; } /* axis points toward the plane */
  for (int k = 0; k < 3; k++) vec[k] = -n[k] + axis[k] * prj; /* -n projected on the disc plane */
  double len = norm3(vec);
  if (len < 1e-12) { col3(vec, R, 0); for (int k = 0; k < 3; k++) vec[k] *= r; } /* disc parallel to plane */
  else for (int k = 0; k < 3; k++) vec[k] *= r / len;
  int cnt = 0;
  /* 1: deepest rim point of the near disc */
  for (int k = 0; k < 3; k++) p[k] = c[k] + axis[k] * h + vec[k];
  sub3(t, p, pp);
  double dist = dot3(t, n);
  if (dist >= margin) return 0;
  addscl3(pos, p, n, -0.5 * dist);
  set_contact(out + cnt++, dist, pos, n);
  /* 2: same side of the far disc (cylinder lying on its side) */
  for (int k = 0; k < 3; k++) p[k] = c[k] - axis[k] * h + vec[k];
  sub3(t, p, pp);
  dist = dot3(t, n);
  if (dist < margin) { addscl3(pos, p, n, -0.5 * dist); set_contact(out + cnt++, dist, pos, n); }
  /* 3,4: triangle points on the near disc (cylinder standing on its cap) */
  double w[3];
  cross3(w, vec, axis);
  for (int sgn = -1; sgn <= 1; sgn += 2) {
    for (int k = 0; k < 3; k++) p[k] = c[k] + axis[k] * h - 0.5 * vec[k] + sgn * 0.8660254037844386 * w[k];
    sub3(t, p, pp);
    dist = dot3(t, n);
    if (dist < margin) { addscl3(pos, p, n, -0.5 * dist); set_contact(out + cnt++, dist, pos, n); }
  }
  return cnt;
}
/* plane - mesh: hull vertices below the margin, the four deepest (ties: lowest index first) */
static int plane_mesh(const double* pp, const double* pR, const double* c, const double* R, const double* verts, int nvert, double margin, om_contact* out) {
  double n[3];
  col3(n, pR, 2);
  int idx[4], cnt = 0;
  double dep[4];
  for (int i = 0; i < nvert; i++) {
    double w[3], t[3];
    for (int k = 0; k < 3; k++) w[k] = c[k] + R[3 * k] * verts[3 * i] + R[3 * k + 1] * verts[3 * i + 1] + R[3 * k + 2] * verts[3 * i + 2];
    sub3(t, w, pp);
    double dist = dot3(t, n);
    if (dist >= margin) continue;
    int pos = cnt < 4 ? cnt : 4;
    while (pos > 0 && dist < dep[pos - 1]) --pos; /* insertion point in the ascending list of depths (strict: ties keep index order) */
    if (pos >= 4) continue;
    for (int j = (cnt < 4 ? cnt : 3); j > pos; --j) { dep[j] = dep[j - 1]; idx[j] = idx[j - 1]; }
    dep[pos] = dist; idx[pos] = i;
    if (cnt < 4) cnt++;
  }
  for (int q = 0; q < cnt; q++) {
    double w[3], pos[3];
    const double* v = verts + 3 * idx[q];
    for (int k = 0; k < 3; k++) w[k] = c[k] + R[3 * k] * v[0] + R[3 * k + 1] * v[1] + R[3 * k + 2] * v[2];
    addscl3(pos, w, n, -0.5 * dep[q]);
    set_contact(out + q, dep[q], pos, n);
  }
  return cnt;
}
/* plane - capsule: the two end spheres of the segment (mjc_PlaneCapsule) */
static int plane_capsule(const double* pp, const double* pR, const double* c, const double* R, double r, double h, double margin, om_contact* out) {
  double n[3], axis[3];
  col3(n, pR, 2);
  col3(axis, R, 2);
  int cnt = 0;
  for (int sg = 1; sg >= -1; sg -= 2) {
    double e[3], t[3], pos[3];
    addscl3(e, c, axis, sg * h);
    sub3(t, e, pp);
    double dist = dot3(t, n) - r;
    if (dist >= margin) continue;
    addscl3(pos, e, n, -(r + 0.5 * dist));
    set_contact(out + cnt++, dist, pos, n);
  }
  return cnt;
}
static int sphere_sphere(const double* c1, double r1, const double* c2, double r2, double margin, om_contact* out) {
  double n[3], pos[3];
  sub3(n, c2, c1);
  double d = norm3(n);
  double dist = d - r1 - r2;
  if (dist >= margin) return 0;
  if (d < MINVAL) { n[0] = 1; n[1] = n[2] = 0; } else { n[0] /= d; n[1] /= d; n[2] /= d; }
  addscl3(pos, c1, n, r1 + 0.5 * dist);
  set_contact(out, dist, pos, n);
  return 1;
}
static int sphere_box(const double* c, double r, const double* bc, const double* R, const double* s, double margin, om_contact* out) {
  double t[3], loc[3], cl[3], n[3], pos[3];
  sub3(t, c, bc);
  for (int k = 0; k < 3; k++) loc[k] = R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]; /* R^T t */
  int inside = 1;
  for (int k = 0; k < 3; k++) {
    cl[k] = loc[k] < -s[k] ? -s[k] : (loc[k] > s[k] ? s[k] : loc[k]);
    if (cl[k] != loc[k]) inside = 0;
  }
  double dist;
  if (!inside) {
    double dl[3] = {cl[0] - loc[0], cl[1] - loc[1], cl[2] - loc[2]}; /* sphere centre -> closest point (box frame) */
    double d = norm3(dl);
    dist = d - r;
    if (dist >= margin) return 0;
    for (int k = 0; k < 3; k++) n[k] = (R[3 * k] * dl[0] + R[3 * k + 1] * dl[1] + R[3 * k + 2] * dl[2]) / d;
  } else { /* centre inside the box: exit through the nearest face */
    int best = 0;
    double bd = 1e300;
    for (int k = 0; k < 3; k++) { double dd = s[k] - fabs(loc[k]); if (dd < bd) { bd = dd; best = k; } }
    double sg = loc[best] >= 0 ? 1.0 : -1.0;
    for (int k = 0; k < 3; k++) n[k] = -sg * R[3 * k + best]; /* from sphere toward the box interior */
    dist = -bd - r;
  }
  addscl3(pos, c, n, r + 0.5 * dist);
  set_contact(out, dist, pos, n);
  return 1;
}

/* ---------------------------------------------------------------- box-box: SAT + clipping (own algorithm) */
static int clip_poly(double (*poly)[3], int n, const double* cR, const double* ax, double lim, double sgn) {
  /* keep the half space sgn * ((p - cR) . ax) <= lim */
  double outp[16][3];
  int m = 0;
  for (int i = 0; i < n; i++) {
    const double* P = poly[i];
    const double* Q = poly[(i + 1) % n];
    double t[3];
    sub3(t, P, cR);
    double dp = sgn * dot3(t, ax) - lim;
    sub3(t, Q, cR);
    double dq = sgn * dot3(t, ax) - lim;
    if (dp <= 0) { memcpy(outp[m++], P, 3 * sizeof(double)); }
    if ((dp <= 0) != (dq <= 0)) {
      double u = dp / (dp - dq);
      for (int k = 0; k < 3; k++) outp[m][k] = P[k] + u * (Q[k] - P[k]);
      m++;
    }
  }
  for (int i = 0; i < m; i++) memcpy(poly[i], outp[i], 3 * sizeof(double));
  return m;
}
static int box_box(const double* cA, const double* RA, const double* a, const double* cB, const double* RB, const double* b, double margin, om_contact* out) {
  double A[3][3], B[3][3], d[3], C[3][3], AC[3][3], dA[3], dB[3];
  for (int k = 0; k < 3; k++) { col3(A[k], RA, k); col3(B[k], RB, k); }
  sub3(d, cB, cA);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { C[i][j] = dot3(A[i], B[j]); AC[i][j] = fabs(C[i][j]); }
  for (int i = 0; i < 3; i++) { dA[i] = dot3(d, A[i]); dB[i] = dot3(d, B[i]); }
  double best_face = -1e300; int face = -1;
  for (int i = 0; i < 3; i++) {
    double sep = fabs(dA[i]) - (a[i] + b[0] * AC[i][0] + b[1] * AC[i][1] + b[2] * AC[i][2]);
    if (sep > margin) return 0;
    if (sep > best_face) { best_face = sep; face = i; }
  }
  for (int j = 0; j < 3; j++) {
    double sep = fabs(dB[j]) - (b[j] + a[0] * AC[0][j] + a[1] * AC[1][j] + a[2] * AC[2][j]);
    if (sep > margin) return 0;
    if (sep > best_face) { best_face = sep; face = 3 + j; }
  }
  double best_edge = -1e300; int ei = -1, ej = -1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double l2 = 1 - C[i][j] * C[i][j];
      if (l2 < 1e-6) continue;
      double l = sqrt(l2);
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      double dl = dA[i2] * C[i1][j] - dA[i1] * C[i2][j];
      double ra = a[i1] * AC[i2][j] + a[i2] * AC[i1][j];
      double rb = b[j1] * AC[i][j2] + b[j2] * AC[i][j1];
      double sep = (fabs(dl) - ra - rb) / l;
      if (sep > margin) return 0;
      if (sep > best_edge) { best_edge = sep; ei = i; ej = j; }
    }
  if (ei >= 0 && -best_edge < 0.95 * (-best_face) - 1e-5) {
    /* edge-edge: closest points between the supporting edges */
    double L[3], pa[3], pb[3], n[3];
    cross3(L, A[ei], B[ej]);
    normalize3(L);
    if (dot3(L, d) < 0) { L[0] = -L[0]; L[1] = -L[1]; L[2] = -L[2]; }
    memcpy(pa, cA, sizeof pa); memcpy(pb, cB, sizeof pb);
    for (int k = 0; k < 3; k++) {
      if (k != ei) { double sg = dot3(L, A[k]) > 0 ? 1 : -1; addscl3(pa, pa, A[k], sg * a[k]); }
      if (k != ej) { double sg = dot3(L, B[k]) > 0 ? -1 : 1; addscl3(pb, pb, B[k], sg * b[k]); }
    }
    /* lines pa + s uA, pb + t uB */
    double w[3];
    sub3(w, pa, pb);
    double uu = 1, vv = 1, uv = C[ei][ej], uw = dot3(A[ei], w), vw = dot3(B[ej], w);
    double den = uu * vv - uv * uv;
    double s = (uv * vw - vv * uw) / den, t = (uu * vw - uv * uw) / den;
    double qa[3], qb[3], pos[3];
    addscl3(qa, pa, A[ei], s);
    addscl3(qb, pb, B[ej], t);
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (qa[k] + qb[k]);
    memcpy(n, L, sizeof n);
    set_contact(out, best_edge, pos, n);
    return 1;
  }
  /* face contact: clip the incident face against the side planes of the reference face */
  const double *cR, *cI, *hR, *hI;
  double (*Rax)[3], (*Iax)[3];
  double nref[3];
  int ri, refIsA = face < 3;
  if (refIsA) { ri = face; cR = cA; cI = cB; hR = a; hI = b; Rax = A; Iax = B; double sg = dA[ri] >= 0 ? 1 : -1; for (int k = 0; k < 3; k++) nref[k] = sg * A[ri][k]; }
  else { ri = face - 3; cR = cB; cI = cA; hR = b; hI = a; Rax = B; Iax = A; double sg = dB[ri] >= 0 ? -1 : 1; for (int k = 0; k < 3; k++) nref[k] = sg * B[ri][k]; }
  int ik = 0; double bestd = -1;
  for (int k = 0; k < 3; k++) { double v = fabs(dot3(Iax[k], nref)); if (v > bestd) { bestd = v; ik = k; } }
  double sgI = dot3(Iax[ik], nref) > 0 ? -1.0 : 1.0; /* incident face: outward normal most anti-parallel to nref */
  int iu = (ik + 1) % 3, iv = (ik + 2) % 3;
  double fc[3], poly[16][3];
  addscl3(fc, cI, Iax[ik], sgI * hI[ik]);
  const double su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
  for (int q = 0; q < 4; q++) for (int k = 0; k < 3; k++) poly[q][k] = fc[k] + su[q] * hI[iu] * Iax[iu][k] + sv[q] * hI[iv] * Iax[iv][k];
  int np = 4, ru = (ri + 1) % 3, rv = (ri + 2) % 3;
  np = clip_poly(poly, np, cR, Rax[ru], hR[ru], 1);
  if (np) np = clip_poly(poly, np, cR, Rax[ru], hR[ru], -1);
  if (np) np = clip_poly(poly, np, cR, Rax[rv], hR[rv], 1);
  if (np) np = clip_poly(poly, np, cR, Rax[rv], hR[rv], -1);
  int cnt = 0;
  double nrm[3];
  for (int k = 0; k < 3; k++) nrm[k] = refIsA ? nref[k] : -nref[k];
  for (int q = 0; q < np && cnt < 8; q++) {
    double t[3], pos[3];
    sub3(t, poly[q], cR);
    double depth = hR[ri] - dot3(t, nref);
    if (depth <= -margin) continue;
    addscl3(pos, poly[q], nref, 0.5 * depth);
    set_contact(out + cnt, -depth, pos, nrm);
    cnt++;
  }
  return cnt;
}

/* ---------------------------------------------------------------- MPR for convex pairs (sphere/cylinder/box) */
typedef struct { int type; const double *pos, *mat, *size; double inflate; /* half the contact margin, added along the query direction (mjccd_support) */
                 const double* verts; int nvert; /* convex-hull vertices of a mesh geom (geom frame) */ } cvx;
static void support(const cvx* g, const double* dir, double* out) {
  double l[3] = {g->mat[0] * dir[0] + g->mat[3] * dir[1] + g->mat[6] * dir[2], g->mat[1] * dir[0] + g->mat[4] * dir[1] + g->mat[7] * dir[2],
                 g->mat[2] * dir[0] + g->mat[5] * dir[1] + g->mat[8] * dir[2]};
  double p[3];
  if (g->type == G_SPHERE) {
    double n = norm3(l);
    for (int k = 0; k < 3; k++) p[k] = n > MINVAL ? l[k] / n * g->size[0] : 0;
  } else if (g->type == G_BOX) {
    for (int k = 0; k < 3; k++) p[k] = l[k] >= 0 ? g->size[k] : -g->size[k];
  } else if (g->type == G_MESH) { /* hull vertex furthest along the direction (first one on ties) */
    int best = 0;
    double bd = -1e300;
    for (int i = 0; i < g->nvert; i++) { double dd = dot3(l, g->verts + 3 * i); if (dd > bd) { bd = dd; best = i; } }
    for (int k = 0; k < 3; k++) p[k] = g->verts[3 * best + k];
  } else if (g->type == G_CAPSULE) { /* sphere swept along the local z segment */
    double n = norm3(l);
    for (int k = 0; k < 3; k++) p[k] = n > MINVAL ? l[k] / n * g->size[0] : 0;
    p[2] += l[2] >= 0 ? g->size[1] : -g->size[1];
  } else { /* cylinder */
    double n = sqrt(l[0] * l[0] + l[1] * l[1]);
    p[0] = n > MINVAL ? l[0] / n * g->size[0] : 0;
    p[1] = n > MINVAL ? l[1] / n * g->size[0] : 0;
    p[2] = l[2] >= 0 ? g->size[1] : -g->size[1];
  }
  for (int k = 0; k < 3; k++) out[k] = g->pos[k] + g->mat[3 * k] * p[0] + g->mat[3 * k + 1] * p[1] + g->mat[3 * k + 2] * p[2] + g->inflate * dir[k];
}
typedef struct { double v[3], v1[3], v2[3]; } sup;
static void mink(const cvx* g1, const cvx* g2, const double* dir, sup* s) { /* point of (g1 - g2) furthest along dir */
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  support(g1, dir, s->v1);
  support(g2, nd, s->v2);
  sub3(s->v, s->v1, s->v2);
}
/* closest point on triangle abc to the origin (Ericson); returns squared distance, writes the point */
static double origin_tri_dist2(const double* a, const double* b, const double* c, double* w) {
  double ab[3], ac[3], ap[3], bp[3], cp[3];
  sub3(ab, b, a); sub3(ac, c, a);
  for (int k = 0; k < 3; k++) { ap[k] = -a[k]; bp[k] = -b[k]; cp[k] = -c[k]; }
  double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { memcpy(w, a, 24); return dot3(w, w); }
  double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { memcpy(w, b, 24); return dot3(w, w); }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); addscl3(w, a, ab, v); return dot3(w, w); }
  double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  if (d6 >= 0 && d5 <= d6) { memcpy(w, c, 24); return dot3(w, w); }
  double vb = d5 * d2 - d1 * d6;
  if (vb <= 0 && d2 >= 0 && d6 <= 0) { double v = d2 / (d2 - d6); addscl3(w, a, ac, v); return dot3(w, w); }
  double va = d3 * d6 - d5 * d4;
  if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    double v = (d4 - d3) / ((d4 - d3) + (d5 - d6)), bc[3];
    sub3(bc, c, b); addscl3(w, b, bc, v); return dot3(w, w);
  }
  double den = 1 / (va + vb + vc), v = vb * den, u = vc * den;
  for (int k = 0; k < 3; k++) w[k] = a[k] + ab[k] * v + ac[k] * u;
  return dot3(w, w);
}
static void portal_dir(const sup* p, double* dir) {
  double e1[3], e2[3];
  sub3(e1, p[2].v, p[1].v); sub3(e2, p[3].v, p[1].v);
  cross3(dir, e1, e2);
  normalize3(dir);
}
static void expand_portal(sup* p, const sup* v4) {
  double v4v0[3];
  cross3(v4v0, v4->v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) { if (dot3(p[2].v, v4v0) > 0) p[1] = *v4; else p[3] = *v4; }
  else { if (dot3(p[3].v, v4v0) > 0) p[2] = *v4; else p[1] = *v4; }
}
static void find_pos(const sup* p, double* pos) {
  double b[4], t[3], sum;
  cross3(t, p[1].v, p[2].v); b[0] = dot3(t, p[3].v);
  cross3(t, p[3].v, p[2].v); b[1] = dot3(t, p[0].v);
  cross3(t, p[0].v, p[1].v); b[2] = dot3(t, p[3].v);
  cross3(t, p[2].v, p[1].v); b[3] = dot3(t, p[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 0) {
    double dir[3];
    b[0] = 0;
    portal_dir(p, dir);
    cross3(t, p[2].v, p[3].v); b[1] = dot3(t, dir);
    cross3(t, p[3].v, p[1].v); b[2] = dot3(t, dir);
    cross3(t, p[1].v, p[2].v); b[3] = dot3(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  double inv = 1 / sum, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  for (int i = 0; i < 4; i++) for (int k = 0; k < 3; k++) { p1[k] += b[i] * p[i].v1[k]; p2[k] += b[i] * p[i].v2[k]; }
  for (int k = 0; k < 3; k++) pos[k] = 0.5 * inv * (p1[k] + p2[k]);
}
/* MPR on the Minkowski difference g1 - g2 seen from v0 = c1 - c2 (libccd's ccdMPRPenetration convention: the
   returned direction points from g1 to g2). Returns 1 and fills depth/dir/pos when the shapes penetrate. */
static int mpr_penetration(const cvx* g1, const cvx* g2, double* depth, double* dir_out, double* pos) {
  const double tol = 1e-6, eps = 1e-10;
  const int max_iter = 50;
  sup p[4], v4;
  double dir[3], va[3], vb[3];
  sub3(p[0].v, g1->pos, g2->pos);
  memcpy(p[0].v1, g1->pos, 24); memcpy(p[0].v2, g2->pos, 24);
  if (norm3(p[0].v) < eps) p[0].v[0] = 1e-5;
  for (int k = 0; k < 3; k++) dir[k] = -p[0].v[k];
  normalize3(dir);
  mink(g1, g2, dir, &p[1]);
  if (dot3(p[1].v, dir) <= 0) return 0;
  cross3(dir, p[0].v, p[1].v);
  if (norm3(dir) < eps) { /* origin on the segment v0-v1 */
    memcpy(dir_out, p[1].v, 24);
    *depth = normalize3(dir_out);
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (p[1].v1[k] + p[1].v2[k]);
    return *depth > 0;
  }
  normalize3(dir);
  mink(g1, g2, dir, &p[2]);
  if (dot3(p[2].v, dir) <= 0) return 0;
  sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
  cross3(dir, va, vb);
  normalize3(dir);
  if (dot3(dir, p[0].v) > 0) { sup t = p[1]; p[1] = p[2]; p[2] = t; for (int k = 0; k < 3; k++) dir[k] = -dir[k]; }
  for (int it = 0;; it++) { /* discover a portal the origin ray passes through */
    if (it > 100) return 0;
    mink(g1, g2, dir, &p[3]);
    if (dot3(p[3].v, dir) <= 0) return 0;
    int cont = 0;
    cross3(va, p[1].v, p[3].v);
    if (dot3(va, p[0].v) < -eps) { p[2] = p[3]; cont = 1; }
    if (!cont) { cross3(va, p[3].v, p[2].v); if (dot3(va, p[0].v) < -eps) { p[1] = p[3]; cont = 1; } }
    if (!cont) break;
    sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
    cross3(dir, va, vb);
    normalize3(dir);
  }
  for (int it = 0;; it++) { /* refine until the portal passes the origin */
    portal_dir(p, dir);
    if (dot3(p[1].v, dir) >= 0) break;
    mink(g1, g2, dir, &v4);
    double dv4 = dot3(v4.v, dir);
    double m1 = dv4 - dot3(p[1].v, dir), m2 = dv4 - dot3(p[2].v, dir), m3 = dv4 - dot3(p[3].v, dir);
    double mn = m1 < m2 ? m1 : m2; mn = mn < m3 ? mn : m3;
    if (dv4 < 0 || mn <= tol || it > max_iter) return 0;
    expand_portal(p, &v4);
  }
  for (int it = 0;; it++) { /* push the portal out to the surface of the difference */
    portal_dir(p, dir);
    mink(g1, g2, dir, &v4);
    double dv4 = dot3(v4.v, dir);
    double m1 = dv4 - dot3(p[1].v, dir), m2 = dv4 - dot3(p[2].v, dir), m3 = dv4 - dot3(p[3].v, dir);
    double mn = m1 < m2 ? m1 : m2; mn = mn < m3 ? mn : m3;
    if (mn <= tol || it > max_iter) {
      double w[3];
      double d2 = origin_tri_dist2(p[1].v, p[2].v, p[3].v, w);
      *depth = sqrt(d2);
      if (*depth < eps) memcpy(dir_out, dir, 24);
      else for (int k = 0; k < 3; k++) dir_out[k] = w[k] / *depth;
      find_pos(p, pos);
      return 1;
    }
    expand_portal(p, &v4);
  }
}
static int convex_pair(const cvx* g1, const cvx* g2, om_contact* out) {
  double depth, dir[3], pos[3];
  if (!mpr_penetration(g1, g2, &depth, dir, pos)) return 0;
  if (depth <= 0) return 0;
  set_contact(out, -depth + g1->inflate + g2->inflate, pos, dir);
  return 1;
}

/* ---------------------------------------------------------------- dispatch */
int om_collide_pair(const om_model* m, const om_data* d, int g1, int g2, om_contact* out) {
  if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  const double *p1 = d->geom_xpos + 3 * g1, *R1 = d->geom_xmat + 9 * g1, *s1 = m->geom_size + 3 * g1;
  const double *p2 = d->geom_xpos + 3 * g2, *R2 = d->geom_xmat + 9 * g2, *s2 = m->geom_size + 3 * g2;
  int n = 0;
  const double margin = m->geom_margin[g1] > m->geom_margin[g2] ? m->geom_margin[g1] : m->geom_margin[g2]; /* mj_contactParam: max */
  if (t1 == G_PLANE) {
    double t[3], nrm[3];
    col3(nrm, R1, 2);
    sub3(t, p2, p1);
    if (dot3(t, nrm) > m->geom_rbound[g2] + margin) return 0;
    if (t2 == G_SPHERE) n = plane_sphere(p1, R1, p2, s2[0], margin, out);
    else if (t2 == G_CAPSULE) n = plane_capsule(p1, R1, p2, R2, s2[0], s2[1], margin, out);
    else if (t2 == G_CYLINDER) n = plane_cylinder(p1, R1, p2, R2, s2[0], s2[1], margin, out);
    else if (t2 == G_BOX) n = plane_box(p1, R1, p2, R2, s2, margin, out);
    else if (t2 == G_MESH) n = plane_mesh(p1, R1, p2, R2, m->mesh_vert + 3 * m->geom_meshadr[g2], m->geom_meshnum[g2], margin, out);
    else return 0;
  } else {
    double t[3];
    sub3(t, p2, p1);
    double bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
    if (dot3(t, t) > bound * bound) return 0;
    if (t1 == G_SPHERE && t2 == G_SPHERE) n = sphere_sphere(p1, s1[0], p2, s2[0], margin, out);
    else if (t1 == G_SPHERE && t2 == G_BOX) n = sphere_box(p1, s1[0], p2, R2, s2, margin, out);
    else if (t1 == G_BOX && t2 == G_BOX) n = box_box(p1, R1, s1, p2, R2, s2, margin, out);
    else { /* a cylinder or a capsule on one side: MPR on shapes inflated by half the margin each */
      cvx a = {t1, p1, R1, s1, 0.5 * margin, t1 == G_MESH ? m->mesh_vert + 3 * m->geom_meshadr[g1] : NULL, t1 == G_MESH ? m->geom_meshnum[g1] : 0};
      cvx b = {t2, p2, R2, s2, 0.5 * margin, t2 == G_MESH ? m->mesh_vert + 3 * m->geom_meshadr[g2] : NULL, t2 == G_MESH ? m->geom_meshnum[g2] : 0};
      n = convex_pair(&a, &b, out);
    }
  }
  /* mj_contactParam */
  for (int i = 0; i < n; i++) {
    om_contact* c = out + i;
    c->geom1 = g1; c->geom2 = g2;
    c->dim = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
    for (int k = 0; k < 3; k++) {
      double f1 = m->geom_friction[3 * g1 + k], f2 = m->geom_friction[3 * g2 + k];
      double f = f1 > f2 ? f1 : f2;
      if (k == 0) { c->friction[0] = c->friction[1] = f; } else if (k == 1) c->friction[2] = f; else { c->friction[3] = c->friction[4] = f; }
    }
    const double *sr1 = m->geom_solref + 2 * g1, *sr2 = m->geom_solref + 2 * g2;
    if (sr1[0] > 0 && sr2[0] > 0) { c->solref[0] = 0.5 * (sr1[0] + sr2[0]); c->solref[1] = 0.5 * (sr1[1] + sr2[1]); }
    else { c->solref[0] = sr1[0] < sr2[0] ? sr1[0] : sr2[0]; c->solref[1] = sr1[1] < sr2[1] ? sr1[1] : sr2[1]; }
    for (int k = 0; k < 5; k++) c->solimp[k] = 0.5 * (m->geom_solimp[5 * g1 + k] + m->geom_solimp[5 * g2 + k]);
    c->mu = 0; c->efc_address = -1; c->margin = margin;
    c->gap = m->geom_gap[g1] > m->geom_gap[g2] ? m->geom_gap[g1] : m->geom_gap[g2];
    for (int k = 0; k < 5; k++) if (c->friction[k] < 1e-5) c->friction[k] = 1e-5; /* mjMINMU */
    if (c->dim != 1) c->dim = 3; /* condim 1 and 3 only in these scenes */
  }
  return n;
}

void om_collision(const om_model* m, om_data* d) {
  d->ncon = 0;
  for (int k = 0; k < m->npair; k++) {
    int g1 = m->collision_pairs[2 * k], g2 = m->collision_pairs[2 * k + 1];
    int ct1 = d->geom_contype[g1], ca1 = d->geom_conaffinity[g1], ct2 = d->geom_contype[g2], ca2 = d->geom_conaffinity[g2];
    if (!((ct1 & ca2) || (ct2 & ca1))) continue;
    om_contact tmp[8];
    int n = om_collide_pair(m, d, g1, g2, tmp);
    for (int i = 0; i < n; i++) {
      if (d->ncon >= OM_MAXCON) { d->warning |= 1; return; }
      d->contact[d->ncon++] = tmp[i];
    }
  }
}
