// fe_collide.h -- narrow-phase primitives (fp32, one lane per candidate pair).
//
// Geometry definitions follow MuJoCo's mj_collision for the pair types the furniture scenes contain:
// analytic plane-{sphere,cylinder,box}, sphere-sphere, sphere-box; Minkowski-portal-refinement (the libccd algorithm
// MuJoCo 2.0 calls for cylinder pairs) for sphere/cylinder/box vs cylinder; box-box by SAT + reference-face clipping.
// Contact convention: normal points from geom1 to geom2, dist < 0 is penetration, pos lies mid-way between surfaces.
#pragma once
#include "fe_model.h"
#include "fe_warp.h"

struct FeCon {
  float dist, pos[3], n[3];
};

FE_HD void fe_col(float* r, const float* R, int k) { r[0] = R[k]; r[1] = R[3 + k]; r[2] = R[6 + k]; }

FE_HD int fe_plane_sphere(const float* pp, const float* pR, const float* c, float r, float margin, FeCon* out) {
  float n[3], t[3];
  fe_col(n, pR, 2);
  v3sub(t, c, pp);
  float dist = v3dot(t, n) - r;
  if (dist >= margin) return 0;
  out->dist = dist;
  v3madd(out->pos, c, n, -(r + 0.5f * dist));
  v3cpy(out->n, n);
  return 1;
}

FE_HD int fe_plane_box(const float* pp, const float* pR, const float* c, const float* R, const float* s, float margin, FeCon* out) {
  float n[3];
  fe_col(n, pR, 2);
  int cnt = 0;
  for (int i = 0; i < 8 && cnt < 4; ++i) {
    float lx = (i & 1) ? s[0] : -s[0], ly = (i & 2) ? s[1] : -s[1], lz = (i & 4) ? s[2] : -s[2];
    float corner[3] = {c[0] + R[0] * lx + R[1] * ly + R[2] * lz, c[1] + R[3] * lx + R[4] * ly + R[5] * lz, c[2] + R[6] * lx + R[7] * ly + R[8] * lz};
    float t[3];
    v3sub(t, corner, pp);
    float dist = v3dot(t, n);
    if (dist >= margin) continue;
    out[cnt].dist = dist;
    v3madd(out[cnt].pos, corner, n, -0.5f * dist);
    v3cpy(out[cnt].n, n);
    ++cnt;
  }
  return cnt;
}

FE_HD int fe_plane_cylinder(const float* pp, const float* pR, const float* c, const float* R, float r, float h, float margin, FeCon* out) {
  float n[3], ax[3], vec[3], p[3], t[3];
  fe_col(n, pR, 2);
  fe_col(ax, R, 2);
  float prj = v3dot(n, ax);
  if (prj > 0.f) { ax[0] = -ax[0]; ax[1] = -ax[1]; ax[2] = -ax[2]; prj = -prj; }
  for (int k = 0; k < 3; ++k) vec[k] = -n[k] + ax[k] * prj;
  float len = v3norm(vec);
  if (len < 1e-6f) { fe_col(vec, R, 0); for (int k = 0; k < 3; ++k) vec[k] *= r; }
  else { float s = r / len; for (int k = 0; k < 3; ++k) vec[k] *= s; }
  for (int k = 0; k < 3; ++k) p[k] = c[k] + ax[k] * h + vec[k];
  v3sub(t, p, pp);
  float dist = v3dot(t, n);
  if (dist >= margin) return 0;
  int cnt = 0;
  out[cnt].dist = dist; v3madd(out[cnt].pos, p, n, -0.5f * dist); v3cpy(out[cnt].n, n); ++cnt;
  for (int k = 0; k < 3; ++k) p[k] = c[k] - ax[k] * h + vec[k];
  v3sub(t, p, pp);
  dist = v3dot(t, n);
  if (dist < margin) { out[cnt].dist = dist; v3madd(out[cnt].pos, p, n, -0.5f * dist); v3cpy(out[cnt].n, n); ++cnt; }
  float w[3];
  v3cross(w, vec, ax);
  for (int sg = -1; sg <= 1; sg += 2) {
    for (int k = 0; k < 3; ++k) p[k] = c[k] + ax[k] * h - 0.5f * vec[k] + (float)sg * 0.8660254037844386f * w[k];
    v3sub(t, p, pp);
    dist = v3dot(t, n);
    if (dist < margin) { out[cnt].dist = dist; v3madd(out[cnt].pos, p, n, -0.5f * dist); v3cpy(out[cnt].n, n); ++cnt; }
  }
  return cnt;
}

// plane - mesh: hull vertices below the margin, the four deepest (ties: lowest index first)
FE_HD int fe_plane_mesh(const float* pp, const float* pR, const float* c, const float* R, const float* verts, int nvert, float margin, FeCon* out) {
  float n[3];
  fe_col(n, pR, 2);
  int i0 = -1, i1 = -1, i2 = -1, i3 = -1, cnt = 0; // kept in scalars: a sorted insertion without indexed local arrays
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
  for (int i = 0; i < nvert; ++i) {
    float w[3], t[3];
    m3mulv(w, R, verts + 3 * i);
    v3add(w, w, c);
    v3sub(t, w, pp);
    const float dist = v3dot(t, n);
    if (dist >= margin) continue;
    if (cnt < 1 || dist < d0) { d3 = d2; i3 = i2; d2 = d1; i2 = i1; d1 = d0; i1 = i0; d0 = dist; i0 = i; }
    else if (cnt < 2 || dist < d1) { d3 = d2; i3 = i2; d2 = d1; i2 = i1; d1 = dist; i1 = i; }
    else if (cnt < 3 || dist < d2) { d3 = d2; i3 = i2; d2 = dist; i2 = i; }
    else if (cnt < 4 || dist < d3) { d3 = dist; i3 = i; }
    if (cnt < 4) ++cnt;
  }
  for (int q = 0; q < cnt; ++q) {
    const int iv = q == 0 ? i0 : (q == 1 ? i1 : (q == 2 ? i2 : i3));
    const float dq = q == 0 ? d0 : (q == 1 ? d1 : (q == 2 ? d2 : d3));
    float w[3];
    m3mulv(w, R, verts + 3 * iv);
    v3add(w, w, c);
    out[q].dist = dq; v3madd(out[q].pos, w, n, -0.5f * dq); v3cpy(out[q].n, n);
  }
  return cnt;
}

// plane - capsule: the two end spheres of the segment (mjc_PlaneCapsule)
FE_HD int fe_plane_capsule(const float* pp, const float* pR, const float* c, const float* R, float r, float h, float margin, FeCon* out) {
  float n[3], ax[3];
  fe_col(n, pR, 2);
  fe_col(ax, R, 2);
  int cnt = 0;
  for (int sg = 1; sg >= -1; sg -= 2) {
    float e[3], t[3];
    v3madd(e, c, ax, (float)sg * h);
    v3sub(t, e, pp);
    const float dist = v3dot(t, n) - r;
    if (dist >= margin) continue;
    out[cnt].dist = dist; v3madd(out[cnt].pos, e, n, -(r + 0.5f * dist)); v3cpy(out[cnt].n, n); ++cnt;
  }
  return cnt;
}

FE_HD int fe_sphere_sphere(const float* c1, float r1, const float* c2, float r2, float margin, FeCon* out) {
  float n[3];
  v3sub(n, c2, c1);
  float d = v3norm(n), dist = d - r1 - r2;
  if (dist >= margin) return 0;
  if (d < 1e-20f) { n[0] = 1.f; n[1] = n[2] = 0.f; } else { float s = 1.f / d; n[0] *= s; n[1] *= s; n[2] *= s; }
  out->dist = dist;
  v3madd(out->pos, c1, n, r1 + 0.5f * dist);
  v3cpy(out->n, n);
  return 1;
}

FE_HD int fe_sphere_box(const float* c, float r, const float* bc, const float* R, const float* s, float margin, FeCon* out) {
  float t[3], loc[3], cl[3], n[3];
  v3sub(t, c, bc);
  m3tmulv(loc, R, t);
  bool inside = true;
  for (int k = 0; k < 3; ++k) {
    cl[k] = fminf(fmaxf(loc[k], -s[k]), s[k]);
    if (cl[k] != loc[k]) inside = false;
  }
  float dist;
  if (!inside) {
    float dl[3] = {cl[0] - loc[0], cl[1] - loc[1], cl[2] - loc[2]};
    float d = v3norm(dl);
    dist = d - r;
    if (dist >= margin) return 0;
    m3mulv(n, R, dl);
    float inv = 1.f / d;
    n[0] *= inv; n[1] *= inv; n[2] *= inv;
  } else {
    int best = 0;
    float bd = 1e30f;
    for (int k = 0; k < 3; ++k) { float dd = s[k] - fabsf(loc[k]); if (dd < bd) { bd = dd; best = k; } }
    float sg = loc[best] >= 0.f ? 1.f : -1.f;
    for (int k = 0; k < 3; ++k) n[k] = -sg * R[3 * k + best];
    dist = -bd - r;
  }
  out->dist = dist;
  v3madd(out->pos, c, n, r + 0.5f * dist);
  v3cpy(out->n, n);
  return 1;
}

// ---- box-box
FE_HD int fe_clip(float (*poly)[3], int n, const float* cR, const float* ax, float lim, float sgn) {
  float o[12][3];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const float* P = poly[i];
    const float* Q = poly[(i + 1 == n) ? 0 : i + 1];
    float t[3];
    v3sub(t, P, cR);
    float dp = sgn * v3dot(t, ax) - lim;
    v3sub(t, Q, cR);
    float dq = sgn * v3dot(t, ax) - lim;
    if (dp <= 0.f && m < 12) { v3cpy(o[m], P); ++m; }
    if ((dp <= 0.f) != (dq <= 0.f) && m < 12) {
      float u = dp / (dp - dq);
      for (int k = 0; k < 3; ++k) o[m][k] = P[k] + u * (Q[k] - P[k]);
      ++m;
    }
  }
  for (int i = 0; i < m; ++i) v3cpy(poly[i], o[i]);
  return m;
}

FE_HD int fe_box_box(const float* cA, const float* RA, const float* a, const float* cB, const float* RB, const float* b, float margin, FeCon* out) {
  float A[3][3], B[3][3], d[3], Cm[3][3], AC[3][3], dA[3], dB[3];
  for (int k = 0; k < 3; ++k) { fe_col(A[k], RA, k); fe_col(B[k], RB, k); }
  v3sub(d, cB, cA);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { Cm[i][j] = v3dot(A[i], B[j]); AC[i][j] = fabsf(Cm[i][j]); }
  for (int i = 0; i < 3; ++i) { dA[i] = v3dot(d, A[i]); dB[i] = v3dot(d, B[i]); }
  float best_face = -1e30f;
  int face = -1;
  for (int i = 0; i < 3; ++i) {
    float sep = fabsf(dA[i]) - (a[i] + b[0] * AC[i][0] + b[1] * AC[i][1] + b[2] * AC[i][2]);
    if (sep > margin) return 0;
    if (sep > best_face) { best_face = sep; face = i; }
  }
  for (int j = 0; j < 3; ++j) {
    float sep = fabsf(dB[j]) - (b[j] + a[0] * AC[0][j] + a[1] * AC[1][j] + a[2] * AC[2][j]);
    if (sep > margin) return 0;
    if (sep > best_face) { best_face = sep; face = 3 + j; }
  }
  float best_edge = -1e30f;
  int ei = -1, ej = -1;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float l2 = 1.f - Cm[i][j] * Cm[i][j];
      if (l2 < 1e-6f) continue;
      float l = sqrtf(l2);
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      float dl = dA[i2] * Cm[i1][j] - dA[i1] * Cm[i2][j];
      float ra = a[i1] * AC[i2][j] + a[i2] * AC[i1][j];
      float rb = b[j1] * AC[i][j2] + b[j2] * AC[i][j1];
      float sep = (fabsf(dl) - ra - rb) / l;
      if (sep > margin) return 0;
      if (sep > best_edge) { best_edge = sep; ei = i; ej = j; }
    }
  if (ei >= 0 && -best_edge < 0.95f * (-best_face) - 1e-5f) {
    float L[3], pa[3], pb[3];
    v3cross(L, A[ei], B[ej]);
    v3normalize(L);
    if (v3dot(L, d) < 0.f) { L[0] = -L[0]; L[1] = -L[1]; L[2] = -L[2]; }
    v3cpy(pa, cA); v3cpy(pb, cB);
    for (int k = 0; k < 3; ++k) {
      if (k != ei) { float sg = v3dot(L, A[k]) > 0.f ? 1.f : -1.f; v3madd(pa, pa, A[k], sg * a[k]); }
      if (k != ej) { float sg = v3dot(L, B[k]) > 0.f ? -1.f : 1.f; v3madd(pb, pb, B[k], sg * b[k]); }
    }
    float w[3];
    v3sub(w, pa, pb);
    float uv = Cm[ei][ej], uw = v3dot(A[ei], w), vw = v3dot(B[ej], w);
    float den = 1.f - uv * uv;
    float s = (uv * vw - uw) / den, t = (vw - uv * uw) / den;
    float qa[3], qb[3];
    v3madd(qa, pa, A[ei], s);
    v3madd(qb, pb, B[ej], t);
    out->dist = best_edge;
    for (int k = 0; k < 3; ++k) out->pos[k] = 0.5f * (qa[k] + qb[k]);
    v3cpy(out->n, L);
    return 1;
  }
  const float *cR, *cI, *hR, *hI;
  float(*Rax)[3];
  float(*Iax)[3];
  float nref[3];
  int ri;
  bool refIsA = face < 3;
  if (refIsA) { ri = face; cR = cA; cI = cB; hR = a; hI = b; Rax = A; Iax = B; float sg = dA[ri] >= 0.f ? 1.f : -1.f; for (int k = 0; k < 3; ++k) nref[k] = sg * A[ri][k]; }
  else { ri = face - 3; cR = cB; cI = cA; hR = b; hI = a; Rax = B; Iax = A; float sg = dB[ri] >= 0.f ? -1.f : 1.f; for (int k = 0; k < 3; ++k) nref[k] = sg * B[ri][k]; }
  int ik = 0;
  float bestd = -1.f;
  for (int k = 0; k < 3; ++k) { float v = fabsf(v3dot(Iax[k], nref)); if (v > bestd) { bestd = v; ik = k; } }
  float sgI = v3dot(Iax[ik], nref) > 0.f ? -1.f : 1.f;
  int iu = (ik + 1) % 3, iv = (ik + 2) % 3;
  float fc[3], poly[12][3];
  v3madd(fc, cI, Iax[ik], sgI * hI[ik]);
  const float su[4] = {1.f, -1.f, -1.f, 1.f}, sv[4] = {1.f, 1.f, -1.f, -1.f};
  for (int q = 0; q < 4; ++q)
    for (int k = 0; k < 3; ++k) poly[q][k] = fc[k] + su[q] * hI[iu] * Iax[iu][k] + sv[q] * hI[iv] * Iax[iv][k];
  int np = 4, ru = (ri + 1) % 3, rv = (ri + 2) % 3;
  np = fe_clip(poly, np, cR, Rax[ru], hR[ru], 1.f);
  if (np) np = fe_clip(poly, np, cR, Rax[ru], hR[ru], -1.f);
  if (np) np = fe_clip(poly, np, cR, Rax[rv], hR[rv], 1.f);
  if (np) np = fe_clip(poly, np, cR, Rax[rv], hR[rv], -1.f);
  int cnt = 0;
  for (int q = 0; q < np && cnt < 8; ++q) {
    float t[3];
    v3sub(t, poly[q], cR);
    float depth = hR[ri] - v3dot(t, nref);
    if (depth <= -margin) continue;
    out[cnt].dist = -depth;
    v3madd(out[cnt].pos, poly[q], nref, 0.5f * depth);
    for (int k = 0; k < 3; ++k) out[cnt].n[k] = refIsA ? nref[k] : -nref[k];
    ++cnt;
  }
  return cnt;
}

// ---- Minkowski portal refinement on (g1 - g2), v0 = c1 - c2; direction returned points from g1 to g2
// `size` of a mesh collider (type 7 | nvert << 8) points to its convex-hull vertices (geom frame) instead of the size vector.
// `infl` = half the contact margin: the support function pushes the surface out by it (mjccd_support), fe_mpr takes it back.
struct FeCvx {
  int type;
  const float *pos, *mat, *size;
};
// FULL = false: sphere / cylinder / box only (the common scenes); FULL = true adds the mesh-hull and capsule supports
template <bool FULL>
FE_HD void fe_support(const FeCvx& g, const float* dir, float infl, float* out) {
  float l[3], p[3];
  m3tmulv(l, g.mat, dir);
  const int type = g.type & 255;
  if (type == 2) {
    float n = v3norm(l);
    float s = n > 1e-20f ? g.size[0] / n : 0.f;
    p[0] = l[0] * s; p[1] = l[1] * s; p[2] = l[2] * s;
  } else if (type == 6) {
    p[0] = l[0] >= 0.f ? g.size[0] : -g.size[0];
    p[1] = l[1] >= 0.f ? g.size[1] : -g.size[1];
    p[2] = l[2] >= 0.f ? g.size[2] : -g.size[2];
  } else if (FULL && type == 7) { // mesh: hull vertex furthest along the direction (first one on ties)
    const int nvert = g.type >> 8;
    int best = 0;
    float bd = -1e30f;
    for (int i = 0; i < nvert; ++i) { const float dd = l[0] * g.size[3 * i] + l[1] * g.size[3 * i + 1] + l[2] * g.size[3 * i + 2]; if (dd > bd) { bd = dd; best = i; } }
    p[0] = g.size[3 * best]; p[1] = g.size[3 * best + 1]; p[2] = g.size[3 * best + 2];
  } else if (FULL && type == 3) { // capsule: a sphere swept along the local z segment
    float n = v3norm(l);
    float s = n > 1e-20f ? g.size[0] / n : 0.f;
    p[0] = l[0] * s; p[1] = l[1] * s; p[2] = l[2] * s + (l[2] >= 0.f ? g.size[1] : -g.size[1]);
  } else {
    float n = sqrtf(l[0] * l[0] + l[1] * l[1]);
    float s = n > 1e-20f ? g.size[0] / n : 0.f;
    p[0] = l[0] * s; p[1] = l[1] * s;
    p[2] = l[2] >= 0.f ? g.size[1] : -g.size[1];
  }
  m3mulv(out, g.mat, p);
  v3add(out, out, g.pos);
  if (infl != 0.f) v3madd(out, out, dir, infl);
}
struct FeSup {
  float v[3], v1[3], v2[3];
};
template <bool FULL>
FE_HD void fe_mink(const FeCvx& g1, const FeCvx& g2, const float* dir, float infl, FeSup* s) {
  float nd[3] = {-dir[0], -dir[1], -dir[2]};
  fe_support<FULL>(g1, dir, infl, s->v1);
  fe_support<FULL>(g2, nd, infl, s->v2);
  v3sub(s->v, s->v1, s->v2);
}
// Closest point of triangle (a, b, c) to the origin (Ericson's region tests); returns its squared distance.  Evaluated in
// float64: the final MPR portal is a tiny triangle (1e-4) some 1e-3 away from the origin, and the region tests are products of
// differences of nearly equal numbers -- in fp32 they misclassify the region and the witness lands off the triangle (seen: a
// normal 40 degrees off for a 0.5 mm cylinder-box penetration).  One call per MPR contact: the cost is nil.
FE_HD float fe_origin_tri(const float* af, const float* bf, const float* cf, float* w) {
  const double a[3] = {af[0], af[1], af[2]}, b[3] = {bf[0], bf[1], bf[2]}, c[3] = {cf[0], cf[1], cf[2]};
  const double ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ac[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
#define FE_D3(x, y) ((x)[0] * (y)[0] + (x)[1] * (y)[1] + (x)[2] * (y)[2])
  double r[3];
  const double d1 = -FE_D3(ab, a), d2 = -FE_D3(ac, a);
  const double d3 = -FE_D3(ab, b), d4 = -FE_D3(ac, b);
  const double d5 = -FE_D3(ab, c), d6 = -FE_D3(ac, c);
  const double vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
  if (d1 <= 0.0 && d2 <= 0.0) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
  else if (d3 >= 0.0 && d4 <= d3) { r[0] = b[0]; r[1] = b[1]; r[2] = b[2]; }
  else if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) { const double v = d1 / (d1 - d3); for (int k = 0; k < 3; ++k) r[k] = a[k] + ab[k] * v; }
  else if (d6 >= 0.0 && d5 <= d6) { r[0] = c[0]; r[1] = c[1]; r[2] = c[2]; }
  else if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) { const double v = d2 / (d2 - d6); for (int k = 0; k < 3; ++k) r[k] = a[k] + ac[k] * v; }
  else if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) { const double v = (d4 - d3) / ((d4 - d3) + (d5 - d6)); for (int k = 0; k < 3; ++k) r[k] = b[k] + (c[k] - b[k]) * v; }
  else { const double den = 1.0 / (va + vb + vc), v = vb * den, u = vc * den; for (int k = 0; k < 3; ++k) r[k] = a[k] + ab[k] * v + ac[k] * u; }
  const double dd = FE_D3(r, r);
#undef FE_D3
  w[0] = (float)r[0]; w[1] = (float)r[1]; w[2] = (float)r[2];
  return (float)dd;
}
FE_HD void fe_portal_dir(const FeSup* p, float* dir) {
  float e1[3], e2[3];
  v3sub(e1, p[2].v, p[1].v); v3sub(e2, p[3].v, p[1].v);
  v3cross(dir, e1, e2);
  v3normalize(dir);
}
FE_HD void fe_expand_portal(FeSup* p, const FeSup& v4) {
  float x[3];
  v3cross(x, v4.v, p[0].v);
  if (v3dot(p[1].v, x) > 0.f) { if (v3dot(p[2].v, x) > 0.f) p[1] = v4; else p[3] = v4; }
  else { if (v3dot(p[3].v, x) > 0.f) p[2] = v4; else p[1] = v4; }
}
FE_HD void fe_find_pos(const FeSup* p, float* pos) {
  float b[4], t[3];
  v3cross(t, p[1].v, p[2].v); b[0] = v3dot(t, p[3].v);
  v3cross(t, p[3].v, p[2].v); b[1] = v3dot(t, p[0].v);
  v3cross(t, p[0].v, p[1].v); b[2] = v3dot(t, p[3].v);
  v3cross(t, p[2].v, p[1].v); b[3] = v3dot(t, p[0].v);
  float sum = b[0] + b[1] + b[2] + b[3];
  if (sum <= 0.f) {
    float dir[3];
    b[0] = 0.f;
    fe_portal_dir(p, dir);
    v3cross(t, p[2].v, p[3].v); b[1] = v3dot(t, dir);
    v3cross(t, p[3].v, p[1].v); b[2] = v3dot(t, dir);
    v3cross(t, p[1].v, p[2].v); b[3] = v3dot(t, dir);
    sum = b[1] + b[2] + b[3];
  }
  float inv = 0.5f / sum;
  for (int k = 0; k < 3; ++k) {
    float s = 0.f;
    for (int i = 0; i < 4; ++i) s += b[i] * (p[i].v1[k] + p[i].v2[k]);
    pos[k] = s * inv;
  }
}
template <bool FULL>
FE_HD int fe_mpr(const FeCvx& g1, const FeCvx& g2, float infl, FeCon* out) {
  const float tol = 1e-6f, eps = 1e-9f;
  FeSup p[4], v4;
  float dir[3], va[3], vb[3];
  v3sub(p[0].v, g1.pos, g2.pos);
  v3cpy(p[0].v1, g1.pos); v3cpy(p[0].v2, g2.pos);
  if (v3norm(p[0].v) < 1e-7f) p[0].v[0] = 1e-5f;
  dir[0] = -p[0].v[0]; dir[1] = -p[0].v[1]; dir[2] = -p[0].v[2];
  v3normalize(dir);
  fe_mink<FULL>(g1, g2, dir, infl, &p[1]);
  if (v3dot(p[1].v, dir) <= 0.f) return 0;
  v3cross(dir, p[0].v, p[1].v);
  if (v3dot(dir, dir) < eps * eps) {
    v3cpy(out->n, p[1].v);
    float depth = v3normalize(out->n);
    if (!(depth > 0.f)) return 0;
    out->dist = -depth + 2.f * infl;
    for (int k = 0; k < 3; ++k) out->pos[k] = 0.5f * (p[1].v1[k] + p[1].v2[k]);
    return 1;
  }
  v3normalize(dir);
  fe_mink<FULL>(g1, g2, dir, infl, &p[2]);
  if (v3dot(p[2].v, dir) <= 0.f) return 0;
  v3sub(va, p[1].v, p[0].v); v3sub(vb, p[2].v, p[0].v);
  v3cross(dir, va, vb);
  v3normalize(dir);
  if (v3dot(dir, p[0].v) > 0.f) { FeSup t = p[1]; p[1] = p[2]; p[2] = t; dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2]; }
  for (int it = 0;; ++it) {
    if (it > 100) return 0;
    fe_mink<FULL>(g1, g2, dir, infl, &p[3]);
    if (v3dot(p[3].v, dir) <= 0.f) return 0;
    bool cont = false;
    v3cross(va, p[1].v, p[3].v);
    if (v3dot(va, p[0].v) < -eps) { p[2] = p[3]; cont = true; }
    if (!cont) { v3cross(va, p[3].v, p[2].v); if (v3dot(va, p[0].v) < -eps) { p[1] = p[3]; cont = true; } }
    if (!cont) break;
    v3sub(va, p[1].v, p[0].v); v3sub(vb, p[2].v, p[0].v);
    v3cross(dir, va, vb);
    v3normalize(dir);
  }
  for (int it = 0;; ++it) {
    fe_portal_dir(p, dir);
    if (v3dot(p[1].v, dir) >= 0.f) break;
    fe_mink<FULL>(g1, g2, dir, infl, &v4);
    float dv4 = v3dot(v4.v, dir);
    float mn = fminf(fminf(dv4 - v3dot(p[1].v, dir), dv4 - v3dot(p[2].v, dir)), dv4 - v3dot(p[3].v, dir));
    if (dv4 < 0.f || mn <= tol || it > 50) return 0;
    fe_expand_portal(p, v4);
  }
  for (int it = 0;; ++it) {
    fe_portal_dir(p, dir);
    fe_mink<FULL>(g1, g2, dir, infl, &v4);
    float dv4 = v3dot(v4.v, dir);
    float mn = fminf(fminf(dv4 - v3dot(p[1].v, dir), dv4 - v3dot(p[2].v, dir)), dv4 - v3dot(p[3].v, dir));
    if (mn <= tol || it > 50) {
      float w[3];
      float depth = sqrtf(fe_origin_tri(p[1].v, p[2].v, p[3].v, w));
      if (depth < 1e-9f) v3cpy(out->n, dir);
      else { float s = 1.f / depth; out->n[0] = w[0] * s; out->n[1] = w[1] * s; out->n[2] = w[2] * s; }
      if (!(depth > 0.f)) return 0;
      out->dist = -depth + 2.f * infl;
      fe_find_pos(p, out->pos);
      return 1;
    }
    fe_expand_portal(p, v4);
  }
}

// the rarely used colliders (mesh hulls, capsules), kept out of the common dispatch so that its register budget stays small
FE_HDN int fe_narrow_rare(const fe_model* m, int g1, int g2, const float* p1, const float* R1, const float* p2, const float* R2, float margin, FeCon* out) {
  const int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  const float *s1 = m->geom_size[g1], *s2 = m->geom_size[g2];
  if (t1 == 0) {
    if (t2 == 3) return fe_plane_capsule(p1, R1, p2, R2, s2[0], s2[1], margin, out);
    return fe_plane_mesh(p1, R1, p2, R2, &m->mesh_vert[m->geom_meshadr[g2]][0], m->geom_meshnum[g2], margin, out);
  }
  // MuJoCo has analytic routines for the capsule pairs (mjc_CapsuleBox ...), which can return two points where MPR returns the deepest one
  FeCvx a = {t1, p1, R1, s1}, b = {t2, p2, R2, s2};
  if (t1 == 7) { a.type = 7 | (m->geom_meshnum[g1] << 8); a.size = &m->mesh_vert[m->geom_meshadr[g1]][0]; }
  if (t2 == 7) { b.type = 7 | (m->geom_meshnum[g2] << 8); b.size = &m->mesh_vert[m->geom_meshadr[g2]][0]; }
  return fe_mpr<true>(a, b, 0.5f * margin, out);
}

// dispatch on the (ordered) type pair; geometry in world frame; contacts closer than `margin` are reported (mj_collision:
// dist < margin, margin = max of the two geoms').  Returns contact count (<= 8).
FE_HDN int fe_narrowphase(const fe_model* m, int g1, int g2, const float* p1, const float* R1, const float* p2, const float* R2, float margin, FeCon* out) {
  const int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  if (t1 == 3 || t2 == 3 || t2 == 7) return fe_narrow_rare(m, g1, g2, p1, R1, p2, R2, margin, out); // t1 <= t2: a mesh is always on side 2
  const float *s1 = m->geom_size[g1], *s2 = m->geom_size[g2];
  if (t1 == 0) {
    if (t2 == 2) return fe_plane_sphere(p1, R1, p2, s2[0], margin, out);
    if (t2 == 5) return fe_plane_cylinder(p1, R1, p2, R2, s2[0], s2[1], margin, out);
    if (t2 == 6) return fe_plane_box(p1, R1, p2, R2, s2, margin, out);
    return 0;
  }
  if (t1 == 2 && t2 == 2) return fe_sphere_sphere(p1, s1[0], p2, s2[0], margin, out);
  if (t1 == 2 && t2 == 6) return fe_sphere_box(p1, s1[0], p2, R2, s2, margin, out);
  if (t1 == 6 && t2 == 6) return fe_box_box(p1, R1, s1, p2, R2, s2, margin, out);
  FeCvx a = {t1, p1, R1, s1}, b = {t2, p2, R2, s2}; // a cylinder on one side: MPR
  return fe_mpr<false>(a, b, 0.5f * margin, out);
}
