/* fe_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See fe_oracle.h for scope and pinning status.
 *
 * Restates, in plain serial double-precision C, the mj_step pipeline the reference drives through
 * MjSim.step() (furniture/env/furniture.py:2878-2879):
 *   mj_fwdPosition  : kinematics -> inertia (CRBA) -> factor -> collision -> constraint rows
 *   mj_fwdVelocity  : body velocities, passive damping, RNE bias force
 *   mj_fwdActuation : motor / position / velocity actuators, ctrl + force clamps
 *   mj_fwdAcceleration : qacc_smooth = M^-1 (passive - bias + applied + actuator)
 *   mj_fwdConstraint   : primal Newton solver with exact line search, elliptic friction cones
 *   mj_Euler           : semi-implicit Euler with implicit joint damping, quaternion integration
 * Spatial quantities are expressed in the world frame about the world origin (motion = [omega; v_O],
 * force = [torque_O; f]); this differs from MuJoCo's subtree-CoM frame but M, qfrc_bias and J are frame independent.
 */
#include "fe_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999

/* ------------------------------------------------------------------ small math */
static inline void v3_set(double* r, double a, double b, double c) { r[0] = a; r[1] = b; r[2] = c; }
static inline void v3_copy(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void v3_add(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void v3_sub(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v3_addscl(double* r, const double* a, const double* b, double s) { r[0] = a[0] + s * b[0]; r[1] = a[1] + s * b[1]; r[2] = a[2] + s * b[2]; }
static inline double v3_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v3_cross(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double v3_norm(const double* a) { return sqrt(v3_dot(a, a)); }
/* r = R (3x3 row-major) * a */
static inline void m3_mulv(double* r, const double* R, const double* a) {
  double x = R[0] * a[0] + R[1] * a[1] + R[2] * a[2], y = R[3] * a[0] + R[4] * a[1] + R[5] * a[2], z = R[6] * a[0] + R[7] * a[1] + R[8] * a[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void m3_tmulv(double* r, const double* R, const double* a) {
  double x = R[0] * a[0] + R[3] * a[1] + R[6] * a[2], y = R[1] * a[0] + R[4] * a[1] + R[7] * a[2], z = R[2] * a[0] + R[5] * a[1] + R[8] * a[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void q_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void q_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void q_to_mat(double* R, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static void q_axis_angle(double* q, const double* axis, double angle) {
  double s = sin(0.5 * angle);
  q[0] = cos(0.5 * angle); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* spatial cross products, vectors are [w(3); v(3)] */
static void sp_crossm(double* r, const double* V, const double* S) { /* V x_m S */
  double a[3], b[3], c[3];
  v3_cross(a, V, S);
  v3_cross(b, V, S + 3);
  v3_cross(c, V + 3, S);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void sp_crossf(double* r, const double* V, const double* F) { /* V x* F, F = [n; f] */
  double a[3], b[3], c[3];
  v3_cross(a, V, F);
  v3_cross(b, V + 3, F + 3);
  v3_cross(c, V, F + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
static void m6_mulv(double* r, const double* I, const double* v) {
  double t[6];
  for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += I[6 * i + j] * v[j]; t[i] = s; }
  memcpy(r, t, sizeof t);
}

/* ------------------------------------------------------------------ model / data plumbing */
om_model* om_model_new(void) { return (om_model*)calloc(1, sizeof(om_model)); }
void om_model_free(om_model* m) {
  if (!m) return;
#define X(n) free(m->n);
  OM_INT_ARRAYS(X)
  OM_DBL_ARRAYS(X)
#undef X
  free(m);
}
int om_model_set_int(om_model* m, const char* name, const int* v, int n) {
#define X(f) if (!strcmp(name, #f)) { m->f = v[0]; return 0; }
  OM_INT_SCALARS(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { free(m->f); m->f = (int*)malloc(sizeof(int) * (n > 0 ? n : 1)); memcpy(m->f, v, sizeof(int) * n); return 0; }
  OM_INT_ARRAYS(X)
#undef X
  return -1;
}
int om_model_set_dbl(om_model* m, const char* name, const double* v, int n) {
#define X(f) if (!strcmp(name, #f)) { m->f = v[0]; return 0; }
  OM_DBL_SCALARS(X)
#undef X
#define X(f) if (!strcmp(name, #f)) { free(m->f); m->f = (double*)malloc(sizeof(double) * (n > 0 ? n : 1)); memcpy(m->f, v, sizeof(double) * n); return 0; }
  OM_DBL_ARRAYS(X)
#undef X
  return -1;
}

om_data* om_data_new(const om_model* m) {
  om_data* d = (om_data*)calloc(1, sizeof(om_data));
#define X(n, sz) d->n = (double*)calloc((size_t)((sz) > 0 ? (sz) : 1), sizeof(double));
  OM_DATA_DBL_ARRAYS(X)
#undef X
#define X(n, sz) d->n = (int*)calloc((size_t)((sz) > 0 ? (sz) : 1), sizeof(int));
  OM_DATA_INT_ARRAYS(X)
#undef X
  d->nefc_cap = 6 * m->neq + m->njnt + 3 * OM_MAXCON;
  int nv = m->nv > 0 ? m->nv : 1;
  d->efc_J = (double*)calloc((size_t)d->nefc_cap * nv, sizeof(double));
#define A(f, k) d->f = (double*)calloc((size_t)d->nefc_cap * (k), sizeof(double));
  A(efc_pos, 1) A(efc_margin, 1) A(efc_diagApprox, 1) A(efc_R, 1) A(efc_D, 1) A(efc_aref, 1) A(efc_vel, 1) A(efc_force, 1) A(efc_KBIP, 4)
#undef A
  d->efc_type = (int*)calloc((size_t)d->nefc_cap, sizeof(int));
  d->efc_id = (int*)calloc((size_t)d->nefc_cap, sizeof(int));
  om_reset_data(m, d);
  return d;
}
void om_data_free(om_data* d) {
  if (!d) return;
#define X(n, sz) free(d->n);
  OM_DATA_DBL_ARRAYS(X)
  OM_DATA_INT_ARRAYS(X)
#undef X
  free(d->efc_J); free(d->efc_pos); free(d->efc_margin); free(d->efc_diagApprox); free(d->efc_R); free(d->efc_D);
  free(d->efc_aref); free(d->efc_vel); free(d->efc_force); free(d->efc_KBIP); free(d->efc_type); free(d->efc_id);
  free(d);
}
double* om_data_dbl(om_data* d, const om_model* m, const char* name, int* n) {
#define X(f, sz) if (!strcmp(name, #f)) { *n = (sz); return d->f; }
  OM_DATA_DBL_ARRAYS(X)
#undef X
#define E(f, k) if (!strcmp(name, #f)) { *n = d->nefc * (k); return d->f; }
  E(efc_pos, 1) E(efc_margin, 1) E(efc_diagApprox, 1) E(efc_R, 1) E(efc_D, 1) E(efc_aref, 1) E(efc_vel, 1) E(efc_force, 1) E(efc_KBIP, 4)
#undef E
  if (!strcmp(name, "efc_J")) { *n = d->nefc * m->nv; return d->efc_J; }
  *n = 0;
  return NULL;
}
int* om_data_int(om_data* d, const om_model* m, const char* name, int* n) {
#define X(f, sz) if (!strcmp(name, #f)) { *n = (sz); return d->f; }
  OM_DATA_INT_ARRAYS(X)
#undef X
  if (!strcmp(name, "efc_type")) { *n = d->nefc; return d->efc_type; }
  if (!strcmp(name, "efc_id")) { *n = d->nefc; return d->efc_id; }
  *n = 0;
  return NULL;
}
int om_data_scalar(const om_data* d, const char* name) {
  if (!strcmp(name, "ncon")) return d->ncon;
  if (!strcmp(name, "nefc")) return d->nefc;
  if (!strcmp(name, "ne")) return d->ne;
  if (!strcmp(name, "nl")) return d->nl;
  if (!strcmp(name, "nc")) return d->nc;
  if (!strcmp(name, "solver_niter")) return d->solver_niter;
  if (!strcmp(name, "warning")) return d->warning;
  return -1;
}
om_contact* om_data_contacts(om_data* d) { return d->contact; }
void om_clear_warning(om_data* d) { d->warning = 0; }
void om_reset_data(const om_model* m, om_data* d) { /* MjSim.reset(): qpos=qpos0, everything else zero */
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nq);
  memset(d->qvel, 0, sizeof(double) * m->nv);
  memset(d->ctrl, 0, sizeof(double) * m->nu);
  memset(d->qfrc_applied, 0, sizeof(double) * m->nv);
  memset(d->xfrc_applied, 0, sizeof(double) * 6 * m->nbody);
  memset(d->qacc_warmstart, 0, sizeof(double) * m->nv);
  memset(d->qacc, 0, sizeof(double) * m->nv);
  memcpy(d->eq_data, m->eq_data, sizeof(double) * 7 * m->neq);
  memcpy(d->geom_contype, m->geom_contype, sizeof(int) * m->ngeom);
  memcpy(d->geom_conaffinity, m->geom_conaffinity, sizeof(int) * m->ngeom);
  memcpy(d->eq_active, m->eq_active, sizeof(int) * m->neq);
  d->time = 0; d->ncon = 0; d->nefc = 0; d->warning = 0;
}

/* ------------------------------------------------------------------ mj_kinematics (+ joint motion subspaces) */
void om_kinematics(const om_model* m, om_data* d) {
  v3_set(d->xpos, 0, 0, 0);
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  q_to_mat(d->xmat, d->xquat);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b], ja = m->body_jntadr[b], jn = m->body_jntnum[b];
    double pos[3], quat[4], R[9];
    if (jn == 1 && m->jnt_type[ja] == 0) { /* free joint: pose straight from qpos (quaternion normalised) */
      int qa = m->jnt_qposadr[ja];
      v3_copy(pos, d->qpos + qa);
      memcpy(quat, d->qpos + qa + 3, sizeof quat);
      q_normalize(quat);
      v3_copy(d->xanchor + 3 * ja, pos);
      q_to_mat(R, quat);
      v3_set(d->xaxis + 3 * ja, R[2], R[5], R[8]);
    } else {
      m3_mulv(pos, d->xmat + 9 * p, m->body_pos + 3 * b);
      v3_add(pos, pos, d->xpos + 3 * p);
      q_mul(quat, d->xquat + 4 * p, m->body_quat + 4 * b);
      for (int j = ja; j < ja + jn; j++) {
        double anchor[3], axis[3], t[3];
        q_to_mat(R, quat);
        m3_mulv(anchor, R, m->jnt_pos + 3 * j); v3_add(anchor, anchor, pos);
        m3_mulv(axis, R, m->jnt_axis + 3 * j);
        v3_copy(d->xanchor + 3 * j, anchor);
        v3_copy(d->xaxis + 3 * j, axis);
        double q = d->qpos[m->jnt_qposadr[j]] - m->qpos0[m->jnt_qposadr[j]];
        if (m->jnt_type[j] == 3) { /* hinge: rotate about the joint axis, keep the anchor fixed */
          double ql[4], qn[4];
          q_axis_angle(ql, m->jnt_axis + 3 * j, q);
          q_mul(qn, quat, ql);
          memcpy(quat, qn, sizeof quat);
          q_to_mat(R, quat);
          m3_mulv(t, R, m->jnt_pos + 3 * j);
          v3_sub(pos, anchor, t);
        } else { /* slide */
          v3_addscl(pos, pos, axis, q);
        }
      }
      q_normalize(quat);
    }
    v3_copy(d->xpos + 3 * b, pos);
    memcpy(d->xquat + 4 * b, quat, sizeof quat);
    q_to_mat(d->xmat + 9 * b, quat);
  }
  for (int b = 0; b < m->nbody; b++) {
    double t[3], q[4];
    m3_mulv(t, d->xmat + 9 * b, m->body_ipos + 3 * b);
    v3_add(d->xipos + 3 * b, t, d->xpos + 3 * b);
    q_mul(q, d->xquat + 4 * b, m->body_iquat + 4 * b);
    q_to_mat(d->ximat + 9 * b, q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double t[3], q[4];
    m3_mulv(t, d->xmat + 9 * b, m->geom_pos + 3 * g);
    v3_add(d->geom_xpos + 3 * g, t, d->xpos + 3 * b);
    q_mul(q, d->xquat + 4 * b, m->geom_quat + 4 * g);
    q_to_mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    double t[3], q[4];
    m3_mulv(t, d->xmat + 9 * b, m->site_pos + 3 * s);
    v3_add(d->site_xpos + 3 * s, t, d->xpos + 3 * b);
    q_mul(q, d->xquat + 4 * b, m->site_quat + 4 * s);
    q_to_mat(d->site_xmat + 9 * s, q);
  }
  /* motion subspace of every dof: S = [axis; anchor x axis] (hinge), [0; axis] (slide / free translation) */
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j], b = m->jnt_bodyid[j];
    if (m->jnt_type[j] == 0) {
      for (int k = 0; k < 3; k++) {
        double* S = d->dofS + 6 * (da + k);
        memset(S, 0, 6 * sizeof(double));
        S[3 + k] = 1;
        double* Sr = d->dofS + 6 * (da + 3 + k);
        const double* R = d->xmat + 9 * b;
        double ax[3] = {R[k], R[3 + k], R[6 + k]};
        v3_copy(Sr, ax);
        v3_cross(Sr + 3, d->xpos + 3 * b, ax);
      }
    } else if (m->jnt_type[j] == 3) {
      double* S = d->dofS + 6 * da;
      v3_copy(S, d->xaxis + 3 * j);
      v3_cross(S + 3, d->xanchor + 3 * j, d->xaxis + 3 * j);
    } else {
      double* S = d->dofS + 6 * da;
      v3_set(S, 0, 0, 0);
      v3_copy(S + 3, d->xaxis + 3 * j);
    }
  }
}

/* dense Cholesky A = L L^T in place (lower); returns 0 ok, -1 not positive definite */
static int chol_factor(double* A, int n) {
  for (int k = 0; k < n; k++) {
    double s = A[k * n + k];
    for (int j = 0; j < k; j++) s -= A[k * n + j] * A[k * n + j];
    if (!(s > MINVAL)) return -1;
    double l = sqrt(s);
    A[k * n + k] = l;
    for (int i = k + 1; i < n; i++) {
      double t = A[i * n + k];
      for (int j = 0; j < k; j++) t -= A[i * n + j] * A[k * n + j];
      A[i * n + k] = t / l;
    }
  }
  return 0;
}
static void chol_solve(const double* L, int n, double* x) {
  for (int i = 0; i < n; i++) { double s = x[i]; for (int j = 0; j < i; j++) s -= L[i * n + j] * x[j]; x[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int j = i + 1; j < n; j++) s -= L[j * n + i] * x[j]; x[i] = s / L[i * n + i]; }
}

/* spatial inertia of body b about the world origin, 6x6 row-major acting on [w; v] -> [n; f] */
static void body_inertia6(const om_model* m, const om_data* d, int b, double* I) {
  double mass = m->body_mass[b];
  const double* R = d->ximat + 9 * b;
  const double* c = d->xipos + 3 * b;
  const double* di = m->body_inertia + 3 * b;
  double Ic[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Ic[3 * i + j] = R[3 * i] * di[0] * R[3 * j] + R[3 * i + 1] * di[1] * R[3 * j + 1] + R[3 * i + 2] * di[2] * R[3 * j + 2];
  double cc = v3_dot(c, c);
  memset(I, 0, 36 * sizeof(double));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) I[6 * i + j] = Ic[3 * i + j] + mass * ((i == j ? cc : 0.0) - c[i] * c[j]);
  double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { I[6 * i + 3 + j] = mass * cx[3 * i + j]; I[6 * (3 + i) + j] = -mass * cx[3 * i + j]; }
  for (int i = 0; i < 3; i++) I[6 * (3 + i) + 3 + i] = mass;
}

void om_smooth(const om_model* m, om_data* d) {
  int nv = m->nv, nb = m->nbody;
  double* I6 = (double*)malloc(sizeof(double) * 36 * nb);
  double* Ic = (double*)malloc(sizeof(double) * 36 * nb);
  double* bacc = (double*)calloc(6 * nb, sizeof(double));
  double* bfrc = (double*)calloc(6 * nb, sizeof(double));
  for (int b = 0; b < nb; b++) body_inertia6(m, d, b, I6 + 36 * b);
  memcpy(Ic, I6, sizeof(double) * 36 * nb);
  for (int b = nb - 1; b > 0; b--) { /* composite rigid body inertia */
    int p = m->body_parentid[b];
    for (int k = 0; k < 36; k++) Ic[36 * p + k] += Ic[36 * b + k];
  }
  memset(d->qM, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    double f[6];
    m6_mulv(f, Ic + 36 * m->dof_bodyid[i], d->dofS + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += d->dofS[6 * j + k] * f[k];
      d->qM[i * nv + j] = d->qM[j * nv + i] = s;
    }
    d->qM[i * nv + i] += m->dof_armature[i];
  }
  memcpy(d->qLD, d->qM, sizeof(double) * nv * nv);
  if (chol_factor(d->qLD, nv)) d->warning |= 2;

  /* velocities + bias accelerations (forward pass of RNE with qacc = 0, gravity as base acceleration) */
  memset(d->bvel, 0, 6 * sizeof(double));
  bacc[3] = -m->opt_gravity[0]; bacc[4] = -m->opt_gravity[1]; bacc[5] = -m->opt_gravity[2];
  for (int b = 1; b < nb; b++) {
    int p = m->body_parentid[b], ja = m->body_jntadr[b], jn = m->body_jntnum[b];
    double V[6], A[6];
    memcpy(V, d->bvel + 6 * p, sizeof V);
    memcpy(A, bacc + 6 * p, sizeof A);
    if (jn == 1 && m->jnt_type[ja] == 0) {
      int da = m->jnt_dofadr[ja];
      double Vp[6];
      memcpy(Vp, V, sizeof V);
      for (int k = 0; k < 6; k++)
        for (int c = 0; c < 6; c++) V[c] += d->dofS[6 * (da + k) + c] * d->qvel[da + k];
      for (int k = 0; k < 3; k++) sp_crossm(d->dofSdot + 6 * (da + k), Vp, d->dofS + 6 * (da + k));
      for (int k = 3; k < 6; k++) sp_crossm(d->dofSdot + 6 * (da + k), V, d->dofS + 6 * (da + k)); /* axes rotate with the body */
      for (int k = 0; k < 6; k++)
        for (int c = 0; c < 6; c++) A[c] += d->dofSdot[6 * (da + k) + c] * d->qvel[da + k];
    } else {
      for (int j = ja; j < ja + jn; j++) {
        int da = m->jnt_dofadr[j];
        sp_crossm(d->dofSdot + 6 * da, V, d->dofS + 6 * da);
        for (int c = 0; c < 6; c++) { A[c] += d->dofSdot[6 * da + c] * d->qvel[da]; }
        for (int c = 0; c < 6; c++) V[c] += d->dofS[6 * da + c] * d->qvel[da];
      }
    }
    memcpy(d->bvel + 6 * b, V, sizeof V);
    memcpy(bacc + 6 * b, A, sizeof A);
    double IA[6], IV[6], VxIV[6];
    m6_mulv(IA, I6 + 36 * b, A);
    m6_mulv(IV, I6 + 36 * b, V);
    sp_crossf(VxIV, V, IV);
    for (int c = 0; c < 6; c++) bfrc[6 * b + c] = IA[c] + VxIV[c];
  }
  for (int b = nb - 1; b > 0; b--) {
    int da = m->body_dofadr[b];
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      double s = 0;
      for (int c = 0; c < 6; c++) s += d->dofS[6 * (da + k) + c] * bfrc[6 * b + c];
      d->qfrc_bias[da + k] = s;
    }
    int p = m->body_parentid[b];
    for (int c = 0; c < 6; c++) bfrc[6 * p + c] += bfrc[6 * b + c];
  }
  /* passive: joint damping only (no springs in these scenes) */
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
  /* actuation */
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int u = 0; u < m->nu; u++) {
    int j = m->actuator_jntid[u];
    double c = d->ctrl[u];
    if (m->actuator_ctrllimited[u]) { double lo = m->actuator_ctrlrange[2 * u], hi = m->actuator_ctrlrange[2 * u + 1]; c = c < lo ? lo : (c > hi ? hi : c); }
    double gear = m->actuator_gear[u];
    double len = gear * d->qpos[m->jnt_qposadr[j]], vel = gear * d->qvel[m->jnt_dofadr[j]];
    double f = m->actuator_gainprm[u] * c + m->actuator_biasprm[3 * u] + m->actuator_biasprm[3 * u + 1] * len + m->actuator_biasprm[3 * u + 2] * vel;
    if (m->actuator_forcelimited[u]) { double lo = m->actuator_forcerange[2 * u], hi = m->actuator_forcerange[2 * u + 1]; f = f < lo ? lo : (f > hi ? hi : f); }
    d->actuator_force[u] = f;
    d->qfrc_actuator[m->jnt_dofadr[j]] += gear * f;
  }
  /* Cartesian forces applied at body CoMs: wrench about the origin, projected on supporting dofs */
  memset(bfrc, 0, sizeof(double) * 6 * nb);
  for (int b = 1; b < nb; b++) {
    const double* x = d->xfrc_applied + 6 * b;
    double t[3];
    v3_cross(t, d->xipos + 3 * b, x);
    bfrc[6 * b + 0] = x[3] + t[0]; bfrc[6 * b + 1] = x[4] + t[1]; bfrc[6 * b + 2] = x[5] + t[2];
    bfrc[6 * b + 3] = x[0]; bfrc[6 * b + 4] = x[1]; bfrc[6 * b + 5] = x[2];
  }
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
  for (int b = nb - 1; b > 0; b--) {
    int da = m->body_dofadr[b];
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      double s = 0;
      for (int c = 0; c < 6; c++) s += d->dofS[6 * (da + k) + c] * bfrc[6 * b + c];
      d->qfrc_smooth[da + k] += s;
    }
    int p = m->body_parentid[b];
    for (int c = 0; c < 6; c++) bfrc[6 * p + c] += bfrc[6 * b + c];
  }
  memcpy(d->qacc_smooth, d->qfrc_smooth, sizeof(double) * nv);
  chol_solve(d->qLD, nv, d->qacc_smooth);
  free(I6); free(Ic); free(bacc); free(bfrc);
}

void om_site_velocity(const om_model* m, const om_data* d, int site, double* out6) {
  const double* V = d->bvel + 6 * m->site_bodyid[site];
  double t[3];
  v3_cross(t, V, d->site_xpos + 3 * site);
  out6[0] = V[3] + t[0]; out6[1] = V[4] + t[1]; out6[2] = V[5] + t[2];
  out6[3] = V[0]; out6[4] = V[1]; out6[5] = V[2];
}

/* ------------------------------------------------------------------ constraint rows */
/* translational (rows 0-2) and rotational (rows 3-5) Jacobian of a point fixed to `body`; each row has nv entries */
static void jac_point(const om_model* m, const om_data* d, int body, const double* p, double* jp, double* jr) {
  int nv = m->nv;
  if (jp) memset(jp, 0, sizeof(double) * 3 * nv);
  if (jr) memset(jr, 0, sizeof(double) * 3 * nv);
  int b = body;
  while (b > 0 && m->body_dofnum[b] == 0) b = m->body_parentid[b];
  if (b == 0) return;
  int i = m->body_dofadr[b] + m->body_dofnum[b] - 1;
  for (; i >= 0; i = m->dof_parentid[i]) {
    const double* S = d->dofS + 6 * i;
    double t[3];
    v3_cross(t, S, p);
    if (jp) for (int k = 0; k < 3; k++) jp[k * nv + i] = S[3 + k] + t[k];
    if (jr) for (int k = 0; k < 3; k++) jr[k * nv + i] = S[k];
  }
}

static void get_solparam(const double* solref, const double* solimp_in, double timestep, double* sr, double* si) {
  sr[0] = solref[0]; sr[1] = solref[1];
  if (sr[0] > 0 && sr[0] < 2 * timestep) sr[0] = 2 * timestep; /* refsafe */
  memcpy(si, solimp_in, 5 * sizeof(double));
  for (int k = 0; k < 2; k++) si[k] = si[k] < MINIMP ? MINIMP : (si[k] > MAXIMP ? MAXIMP : si[k]);
  if (si[2] < MINVAL) si[2] = MINVAL;
  si[3] = si[3] < MINIMP ? MINIMP : (si[3] > MAXIMP ? MAXIMP : si[3]);
  if (si[4] < 1) si[4] = 1;
}
static int add_row(om_data* d, int nv, int type, int id, const double* Jrow, double pos, double margin, double diagApprox) {
  int i = d->nefc;
  if (i >= d->nefc_cap) { d->warning |= 1; return -1; }
  memcpy(d->efc_J + (size_t)i * nv, Jrow, sizeof(double) * nv);
  d->efc_type[i] = type; d->efc_id[i] = id; d->efc_pos[i] = pos; d->efc_margin[i] = margin;
  d->efc_diagApprox[i] = diagApprox < MINVAL ? MINVAL : diagApprox;
  d->nefc++;
  return i;
}

void om_make_constraint(const om_model* m, om_data* d) {
  int nv = m->nv;
  double* jp1 = (double*)malloc(sizeof(double) * 3 * nv * 4);
  double *jr1 = jp1 + 3 * nv, *jp2 = jr1 + 3 * nv, *jr2 = jp2 + 3 * nv;
  double* row = (double*)malloc(sizeof(double) * nv);
  d->nefc = d->ne = d->nl = d->nc = 0;
  /* ---- weld equalities: 3 position rows (anchor on body1 vs origin of body2), 3 orientation rows */
  for (int e = 0; e < m->neq; e++) {
    if (!d->eq_active[e]) continue;
    int b1 = m->eq_obj1id[e], b2 = m->eq_obj2id[e];
    const double* data = d->eq_data + 7 * e;
    double p1[3], cpos[6];
    m3_mulv(p1, d->xmat + 9 * b1, data);
    v3_add(p1, p1, d->xpos + 3 * b1);
    v3_sub(cpos, p1, d->xpos + 3 * b2);
    jac_point(m, d, b1, p1, jp1, jr1);
    jac_point(m, d, b2, d->xpos + 3 * b2, jp2, jr2);
    double quat[4], q2c[4], qe[4];
    q_mul(quat, d->xquat + 4 * b1, data + 3); /* desired orientation of body2 */
    q2c[0] = d->xquat[4 * b2]; q2c[1] = -d->xquat[4 * b2 + 1]; q2c[2] = -d->xquat[4 * b2 + 2]; q2c[3] = -d->xquat[4 * b2 + 3];
    q_mul(qe, q2c, quat);
    cpos[3] = qe[1]; cpos[4] = qe[2]; cpos[5] = qe[3];
    double wt = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    double wr = m->body_invweight0[2 * b1 + 1] + m->body_invweight0[2 * b2 + 1];
    for (int k = 0; k < 3; k++) {
      for (int i = 0; i < nv; i++) row[i] = jp1[k * nv + i] - jp2[k * nv + i];
      add_row(d, nv, OM_EFC_EQUALITY, e, row, cpos[k], 0, wt);
    }
    /* d/dt vec(conj(q2) q1 qrel) = 0.5 * vec(conj(q2) * (w1 - w2) * q1 qrel) */
    double* R3 = (double*)malloc(sizeof(double) * 3 * nv);
    for (int i = 0; i < nv; i++) {
      double ax[4] = {0, jr1[i] - jr2[i], jr1[nv + i] - jr2[nv + i], jr1[2 * nv + i] - jr2[2 * nv + i]};
      double t1[4], t2[4];
      q_mul(t1, q2c, ax);
      q_mul(t2, t1, quat);
      R3[i] = 0.5 * t2[1]; R3[nv + i] = 0.5 * t2[2]; R3[2 * nv + i] = 0.5 * t2[3];
    }
    for (int k = 0; k < 3; k++) add_row(d, nv, OM_EFC_EQUALITY, e, R3 + k * nv, cpos[3 + k], 0, wr);
    free(R3);
  }
  d->ne = d->nefc;
  /* ---- joint limits */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == 0) continue;
    double q = d->qpos[m->jnt_qposadr[j]];
    int da = m->jnt_dofadr[j];
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q - m->jnt_range[2 * j] : m->jnt_range[2 * j + 1] - q;
      if (dist < 0) { /* margin 0 */
        memset(row, 0, sizeof(double) * nv);
        row[da] = side == 0 ? 1 : -1;
        add_row(d, nv, OM_EFC_LIMIT, j, row, dist, 0, m->dof_invweight0[da]);
      }
    }
  }
  d->nl = d->nefc - d->ne;
  /* ---- contacts (condim 3 elliptic: normal + 2 tangents; condim 1: normal only) */
  for (int c = 0; c < d->ncon; c++) {
    om_contact* con = d->contact + c;
    int b1 = m->geom_bodyid[con->geom1], b2 = m->geom_bodyid[con->geom2];
    jac_point(m, d, b1, con->pos, jp1, NULL);
    jac_point(m, d, b2, con->pos, jp2, NULL);
    double w = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];
    if (con->dist >= con->margin - con->gap) { con->efc_address = -1; continue; } /* reported but inactive (mj_makeConstraint: dist < margin - gap) */
    con->efc_address = d->nefc;
    int dim = con->dim == 1 ? 1 : 3;
    for (int k = 0; k < dim; k++) {
      const double* fr = con->frame + 3 * k;
      for (int i = 0; i < nv; i++)
        row[i] = fr[0] * (jp2[i] - jp1[i]) + fr[1] * (jp2[nv + i] - jp1[nv + i]) + fr[2] * (jp2[2 * nv + i] - jp1[2 * nv + i]);
      add_row(d, nv, dim == 1 ? OM_EFC_CONTACT_FRICTIONLESS : OM_EFC_CONTACT_ELLIPTIC, c, row, k == 0 ? con->dist : 0, k == 0 ? con->margin : 0, w);
    }
  }
  d->nc = d->nefc - d->ne - d->nl;
  /* ---- efc_vel, impedance, R, aref */
  for (int i = 0; i < d->nefc; i++) {
    double s = 0;
    for (int k = 0; k < nv; k++) s += d->efc_J[(size_t)i * nv + k] * d->qvel[k];
    d->efc_vel[i] = s;
  }
  for (int i = 0; i < d->nefc; i++) {
    const double *solref, *solimp;
    int id = d->efc_id[i];
    switch (d->efc_type[i]) {
      case OM_EFC_EQUALITY: solref = m->eq_solref + 2 * id; solimp = m->eq_solimp + 5 * id; break;
      case OM_EFC_LIMIT: solref = m->jnt_solref + 2 * id; solimp = m->jnt_solimp + 5 * id; break;
      default: solref = d->contact[id].solref; solimp = d->contact[id].solimp; break;
    }
    double sr[2], si[5];
    get_solparam(solref, solimp, m->opt_timestep, sr, si);
    double dist = fabs(d->efc_pos[i] - d->efc_margin[i]);
    double x = dist / si[2], imp;
    if (si[0] == si[1]) imp = si[0];
    else if (x >= 1) imp = si[1];
    else {
      double y, mid = si[3], p = si[4];
      if (x <= mid) y = pow(x, p) / pow(mid, p - 1);
      else y = 1 - pow(1 - x, p) / pow(1 - mid, p - 1);
      imp = si[0] + y * (si[1] - si[0]);
    }
    double dmax = si[1];
    double k = 1 / (dmax * dmax * sr[0] * sr[0] * sr[1] * sr[1]);
    double b = 2 / (dmax * sr[0]);
    d->efc_KBIP[4 * i] = k; d->efc_KBIP[4 * i + 1] = b; d->efc_KBIP[4 * i + 2] = imp; d->efc_KBIP[4 * i + 3] = 0;
    double R = (1 - imp) / imp * d->efc_diagApprox[i];
    d->efc_R[i] = R < MINVAL ? MINVAL : R;
    d->efc_aref[i] = -b * d->efc_vel[i] - k * imp * (d->efc_pos[i] - d->efc_margin[i]);
  }
  /* elliptic cones: friction rows get R_normal / impratio; regularised mu */
  for (int c = 0; c < d->ncon; c++) {
    om_contact* con = d->contact + c;
    int i = con->efc_address;
    if (i < 0 || con->dim == 1) { con->mu = 0; continue; }
    d->efc_R[i + 1] = d->efc_R[i] / m->opt_impratio;
    if (d->efc_R[i + 1] < MINVAL) d->efc_R[i + 1] = MINVAL;
    con->mu = con->friction[0] * sqrt(d->efc_R[i + 1] / d->efc_R[i]);
    d->efc_R[i + 2] = d->efc_R[i + 1] * con->friction[0] * con->friction[0] / (con->friction[1] * con->friction[1]);
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
  free(jp1); free(row);
}

/* ------------------------------------------------------------------ primal Newton solver */
typedef struct {
  const om_model* m; om_data* d; int nv, nefc;
  double *Ma, *jar, *force, *grad, *Mgrad, *search, *Mv, *jv, *H;
  int* state; /* 0 satisfied, 1 quadratic, 2 cone */
  double cost, gauss;
} nctx;

/* constraint cost + forces + states for the current jar; if Hc != NULL store cone Hessians (9 per contact) */
static double update_constraint(nctx* c, double* Hc) {
  om_data* d = c->d;
  double cost = 0;
  for (int i = 0; i < c->nefc; i++) {
    int t = d->efc_type[i];
    double D = d->efc_D[i], ja = c->jar[i];
    if (t == OM_EFC_EQUALITY) { c->force[i] = -D * ja; c->state[i] = 1; cost += 0.5 * D * ja * ja; }
    else if (t == OM_EFC_LIMIT || t == OM_EFC_CONTACT_FRICTIONLESS) {
      if (ja < 0) { c->force[i] = -D * ja; c->state[i] = 1; cost += 0.5 * D * ja * ja; } else { c->force[i] = 0; c->state[i] = 0; }
    } else { /* first row of an elliptic contact */
      const om_contact* con = d->contact + d->efc_id[i];
      double mu = con->mu, f1 = con->friction[0], f2 = con->friction[1];
      double U0 = c->jar[i] * mu, U1 = c->jar[i + 1] * f1, U2 = c->jar[i + 2] * f2;
      double N = U0, T = sqrt(U1 * U1 + U2 * U2);
      if (N >= mu * T || (T <= 0 && N >= 0)) { for (int k = 0; k < 3; k++) { c->force[i + k] = 0; c->state[i + k] = 0; } }
      else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        for (int k = 0; k < 3; k++) { c->force[i + k] = -d->efc_D[i + k] * c->jar[i + k]; c->state[i + k] = 1; cost += 0.5 * d->efc_D[i + k] * c->jar[i + k] * c->jar[i + k]; }
      } else {
        double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu));
        double NmT = N - mu * T;
        cost += 0.5 * Dm * NmT * NmT;
        c->force[i] = -Dm * NmT * mu;
        c->force[i + 1] = -c->force[i] / T * U1 * f1;
        c->force[i + 2] = -c->force[i] / T * U2 * f2;
        for (int k = 0; k < 3; k++) c->state[i + k] = 2;
        if (Hc) { /* Hessian of the cone cost w.r.t. jar (3x3) */
          double sc[3] = {mu, f1, f2}, U[3] = {U0, U1, U2}, HU[9];
          HU[0] = Dm;
          for (int a = 1; a < 3; a++) HU[a] = HU[3 * a] = -Dm * mu * U[a] / T;
          for (int a = 1; a < 3; a++)
            for (int b = 1; b < 3; b++)
              HU[3 * a + b] = Dm * mu * mu * U[a] * U[b] / (T * T) - Dm * NmT * mu * ((a == b ? 1.0 : 0.0) / T - U[a] * U[b] / (T * T * T));
          double* H = Hc + 9 * d->efc_id[i];
          for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) H[3 * a + b] = sc[a] * HU[3 * a + b] * sc[b];
        }
      }
      i += 2;
    }
  }
  return cost;
}

/* derivatives of the 1-D cost along the search direction at step alpha */
static void line_eval(nctx* c, double alpha, const double* quadGauss, double* cost, double* d1, double* d2) {
  om_data* d = c->d;
  double p0 = quadGauss[0] + alpha * quadGauss[1] + alpha * alpha * quadGauss[2];
  double p1 = quadGauss[1] + 2 * alpha * quadGauss[2];
  double p2 = 2 * quadGauss[2];
  for (int i = 0; i < c->nefc; i++) {
    int t = d->efc_type[i];
    double D = d->efc_D[i];
    if (t != OM_EFC_CONTACT_ELLIPTIC) {
      double x = c->jar[i] + alpha * c->jv[i];
      if (t == OM_EFC_EQUALITY || x < 0) { p0 += 0.5 * D * x * x; p1 += D * x * c->jv[i]; p2 += D * c->jv[i] * c->jv[i]; }
    } else {
      const om_contact* con = d->contact + d->efc_id[i];
      double mu = con->mu, f1 = con->friction[0], f2 = con->friction[1];
      double x0 = c->jar[i] + alpha * c->jv[i], x1 = c->jar[i + 1] + alpha * c->jv[i + 1], x2 = c->jar[i + 2] + alpha * c->jv[i + 2];
      double N = x0 * mu, U1 = x1 * f1, U2 = x2 * f2, T = sqrt(U1 * U1 + U2 * U2);
      double N1 = c->jv[i] * mu, V1 = c->jv[i + 1] * f1, V2 = c->jv[i + 2] * f2;
      if (N >= mu * T || (T <= 0 && N >= 0)) { /* top zone: nothing */
      } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
        double xs[3] = {x0, x1, x2};
        for (int k = 0; k < 3; k++) { double Dk = d->efc_D[i + k], v = c->jv[i + k]; p0 += 0.5 * Dk * xs[k] * xs[k]; p1 += Dk * xs[k] * v; p2 += Dk * v * v; }
      } else {
        double Dm = D / (mu * mu * (1 + mu * mu));
        double NmT = N - mu * T;
        double T1 = (U1 * V1 + U2 * V2) / T;
        double T2 = (V1 * V1 + V2 * V2 - T1 * T1) / T;
        double a = N1 - mu * T1;
        p0 += 0.5 * Dm * NmT * NmT;
        p1 += Dm * NmT * a;
        p2 += Dm * (a * a - NmT * mu * T2);
      }
      i += 2;
    }
  }
  *cost = p0; *d1 = p1; *d2 = p2;
}

static void build_hessian(nctx* c, const double* Hc) {
  om_data* d = c->d;
  int nv = c->nv;
  memcpy(c->H, d->qM, sizeof(double) * nv * nv);
  for (int i = 0; i < c->nefc; i++) {
    if (c->state[i] == 1) {
      const double* J = d->efc_J + (size_t)i * nv;
      double D = d->efc_D[i];
      for (int a = 0; a < nv; a++) if (J[a] != 0) { double s = D * J[a]; for (int b = 0; b <= a; b++) c->H[a * nv + b] += s * J[b]; }
    } else if (c->state[i] == 2) {
      const double* Hl = Hc + 9 * d->efc_id[i];
      for (int r = 0; r < 3; r++)
        for (int s = 0; s < 3; s++) {
          const double *Jr = d->efc_J + (size_t)(i + r) * nv, *Js = d->efc_J + (size_t)(i + s) * nv;
          double h = Hl[3 * r + s];
          if (h == 0) continue;
          for (int a = 0; a < nv; a++) if (Jr[a] != 0) { double t = h * Jr[a]; for (int b = 0; b <= a; b++) c->H[a * nv + b] += t * Js[b]; }
        }
      i += 2;
    }
  }
  for (int a = 0; a < nv; a++) for (int b = a + 1; b < nv; b++) c->H[a * nv + b] = c->H[b * nv + a];
}

void om_solve(const om_model* m, om_data* d) {
  int nv = m->nv, nefc = d->nefc;
  d->solver_niter = 0;
  if (nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv);
    return;
  }
  nctx c;
  c.m = m; c.d = d; c.nv = nv; c.nefc = nefc;
  double* buf = (double*)calloc((size_t)(6 * nv + 3 * nefc + nv * nv + 9 * (d->ncon + 1)), sizeof(double));
  c.Ma = buf; c.grad = c.Ma + nv; c.Mgrad = c.grad + nv; c.search = c.Mgrad + nv; c.Mv = c.search + nv;
  double* qtry = c.Mv + nv;
  c.jar = qtry + nv; c.force = c.jar + nefc; c.jv = c.force + nefc; c.H = c.jv + nefc;
  double* Hc = c.H + nv * nv;
  c.state = (int*)calloc((size_t)nefc, sizeof(int));
  double scale = 1.0 / (m->stat_meaninertia * (nv > 1 ? nv : 1));

  /* warm start: whichever of qacc_warmstart / qacc_smooth has the lower total cost */
  double best = 0;
  for (int pass = 0; pass < 2; pass++) {
    const double* q = pass == 0 ? d->qacc_warmstart : d->qacc_smooth;
    for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->qM[i * nv + j] * q[j]; c.Ma[i] = s; }
    for (int i = 0; i < nefc; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->efc_J[(size_t)i * nv + j] * q[j]; c.jar[i] = s - d->efc_aref[i]; }
    double cost = update_constraint(&c, NULL);
    for (int i = 0; i < nv; i++) cost += 0.5 * (c.Ma[i] - d->qfrc_smooth[i]) * (q[i] - d->qacc_smooth[i]);
    if (pass == 0) { best = cost; memcpy(d->qacc, q, sizeof(double) * nv); }
    else if (cost < best) memcpy(d->qacc, q, sizeof(double) * nv);
  }
  for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->qM[i * nv + j] * d->qacc[j]; c.Ma[i] = s; }
  for (int i = 0; i < nefc; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->efc_J[(size_t)i * nv + j] * d->qacc[j]; c.jar[i] = s - d->efc_aref[i]; }

  int iter = 0, maxiter = m->opt_iterations;
  double cost = 0, oldcost;
  for (;;) {
    double ccost = update_constraint(&c, Hc);
    double gauss = 0;
    for (int i = 0; i < nv; i++) gauss += 0.5 * (c.Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
    oldcost = cost;
    cost = gauss + ccost;
    for (int i = 0; i < nv; i++) {
      double s = c.Ma[i] - d->qfrc_smooth[i];
      for (int r = 0; r < nefc; r++) s -= d->efc_J[(size_t)r * nv + i] * c.force[r];
      c.grad[i] = s;
    }
    double gnorm = 0;
    for (int i = 0; i < nv; i++) gnorm += c.grad[i] * c.grad[i];
    gnorm = sqrt(gnorm);
    if (iter > 0) {
      double improvement = scale * (oldcost - cost), gradient = scale * gnorm;
      if (improvement < m->opt_tolerance || gradient < m->opt_tolerance) break;
    } else if (scale * gnorm < m->opt_tolerance) break;
    if (iter >= maxiter) break;
    build_hessian(&c, Hc);
    if (chol_factor(c.H, nv)) { d->warning |= 2; break; }
    for (int i = 0; i < nv; i++) c.Mgrad[i] = c.grad[i];
    chol_solve(c.H, nv, c.Mgrad);
    for (int i = 0; i < nv; i++) c.search[i] = -c.Mgrad[i];
    for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->qM[i * nv + j] * c.search[j]; c.Mv[i] = s; }
    for (int i = 0; i < nefc; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->efc_J[(size_t)i * nv + j] * c.search[j]; c.jv[i] = s; }
    double quadGauss[3] = {gauss, 0, 0};
    for (int i = 0; i < nv; i++) { quadGauss[1] += c.search[i] * (c.Ma[i] - d->qfrc_smooth[i]); quadGauss[2] += 0.5 * c.search[i] * c.Mv[i]; }
    /* exact line search: safeguarded Newton on p'(alpha) = 0 (p convex, p'(0) < 0) */
    double lo = 0, hi = -1, alpha = 0, p0, p1, p2, plo;
    line_eval(&c, 0, quadGauss, &p0, &p1, &p2);
    plo = p1;
    if (!(p1 < 0) || !(p2 > 0)) break;
    alpha = -p1 / p2;
    double gtol = 1e-14 * (fabs(plo) > 1 ? fabs(plo) : 1) + 1e-12 * fabs(plo);
    /* p' is piecewise smooth (rows enter and leave the cone zones along the ray), so a Newton step taken from one side of
       the root can land next to the other end of the bracket and back again, shrinking it by almost nothing; the step is
       therefore accepted only while it at least halves the previous one (the rtsafe rule), otherwise bisect */
    double dxold = alpha;
    for (int ls = 0; ls < 100; ls++) {
      line_eval(&c, alpha, quadGauss, &p0, &p1, &p2);
      if (fabs(p1) <= gtol) break;
      if (p1 < 0) lo = alpha; else hi = alpha;
      double next = alpha - p1 / p2;
      if (hi > 0 && (!(next > lo && next < hi) || fabs(2 * p1) > fabs(dxold * p2))) next = 0.5 * (lo + hi);
      if (hi < 0 && !(next > lo)) next = 2 * alpha;
      if (fabs(next - alpha) <= 1e-15 * fabs(alpha)) { alpha = next; break; }
      dxold = fabs(next - alpha);
      alpha = next;
    }
    if (!(alpha > 0)) break;
    for (int i = 0; i < nv; i++) { d->qacc[i] += alpha * c.search[i]; c.Ma[i] += alpha * c.Mv[i]; }
    for (int i = 0; i < nefc; i++) c.jar[i] += alpha * c.jv[i];
    iter++;
  }
  d->solver_niter = iter;
  update_constraint(&c, NULL);
  memcpy(d->efc_force, c.force, sizeof(double) * nefc);
  for (int i = 0; i < nv; i++) { double s = 0; for (int r = 0; r < nefc; r++) s += d->efc_J[(size_t)r * nv + i] * c.force[r]; d->qfrc_constraint[i] = s; }
  d->solver_cost[0] = cost;
  for (int i = 0; i < nv; i++) if (!isfinite(d->qacc[i])) d->warning |= 2;
  free(buf); free(c.state);
}

/* ------------------------------------------------------------------ top level */
void om_forward(const om_model* m, om_data* d) {
  om_kinematics(m, d);
  om_smooth(m, d);
  om_collision(m, d);
  om_make_constraint(m, d);
  om_solve(m, d);
}

void om_step(const om_model* m, om_data* d) {
  int nv = m->nv;
  double h = m->opt_timestep;
  om_forward(m, d);
  /* mj_Euler: (M + h B) qacc_int = qfrc_smooth + qfrc_constraint (joint damping integrated implicitly) */
  double* A = (double*)malloc(sizeof(double) * nv * nv);
  double* qa = (double*)malloc(sizeof(double) * nv);
  int damped = 0;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) damped = 1;
  if (damped) {
    memcpy(A, d->qM, sizeof(double) * nv * nv);
    for (int i = 0; i < nv; i++) { A[i * nv + i] += h * m->dof_damping[i]; qa[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    if (chol_factor(A, nv)) d->warning |= 2;
    chol_solve(A, nv, qa);
  } else memcpy(qa, d->qacc, sizeof(double) * nv);
  /* mj_advance */
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qa[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa_ = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == 0) {
      for (int k = 0; k < 3; k++) d->qpos[qa_ + k] += h * d->qvel[da + k];
      double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
      double n = v3_norm(w);
      if (n * h > MINVAL) {
        double ax[3] = {w[0] / n, w[1] / n, w[2] / n}, dq[4], qn[4];
        q_axis_angle(dq, ax, n * h);
        q_mul(qn, d->qpos + qa_ + 3, dq);
        memcpy(d->qpos + qa_ + 3, qn, sizeof qn);
      }
      q_normalize(d->qpos + qa_ + 3);
    } else d->qpos[qa_] += h * d->qvel[da];
  }
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
  d->time += h;
  for (int i = 0; i < m->nq; i++) if (!isfinite(d->qpos[i]) || fabs(d->qpos[i]) > 1e10) d->warning |= 2;
  for (int i = 0; i < nv; i++) if (!isfinite(d->qvel[i]) || fabs(d->qvel[i]) > 1e10) d->warning |= 2;
  free(A); free(qa);
}
