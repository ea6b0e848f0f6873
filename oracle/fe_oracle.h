/* fe_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Double-precision, single-environment restatement of the physics step that the reference reaches through
 * mujoco-py: MjSim.forward()/MjSim.step() (reference call sites furniture/env/furniture.py:2877-2879, :3079-3082,
 * :1437, models/base.py:113-115).  The arithmetic lives in the closed third-party libmujoco200 (MuJoCo 2.0, reached
 * via the unpinned `mujoco-py` dependency: requirements.txt:12, setup.py:25), which is absent from /root/reference,
 * so this file restates MuJoCo's *published* algorithm (MuJoCo "Computation" chapter + XML reference) for exactly the
 * feature subset the composed furniture scenes use.
 *
 * PARITY UNPINNED for the physics: the reference holds no golden vector / known-answer test for any qpos/qvel/contact
 * value and no MuJoCo binary exists in this container, so this oracle is pinned only by its own
 * physical-consistency tests (tests/test_oracle_physics.py) and by rest states MuJoCo itself produced: the poses at which the
 * swivel-chair parts and the two blocks stand untouched in the reference's demo recordings (demos/*.pkl ->
 * tests/golden/demo_facts.json) are equilibria of this oracle to 1e-7 m over 2000 steps -- equilibria of the soft-contact
 * model, not a trajectory.  The assembly logic (_is_aligned/_connect) and the reset placement sampler ARE pinned bit-exactly
 * against the reference's own Python (oracle/assembly_oracle.py, oracle/ref_env.py place(), tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 */
#ifndef FE_ORACLE_H
#define FE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define OM_INT_SCALARS(X) \
  X(nq) X(nv) X(nu) X(nbody) X(njnt) X(ngeom) X(nsite) X(neq) X(npair) X(opt_cone_elliptic) X(opt_iterations)

#define OM_DBL_SCALARS(X) X(opt_timestep) X(opt_impratio) X(opt_tolerance) X(stat_meaninertia)

#define OM_INT_ARRAYS(X)                                                                                            \
  X(body_parentid) X(body_weldid) X(body_rootid) X(body_jntadr) X(body_jntnum) X(body_dofadr) X(body_dofnum)        \
  X(jnt_type) X(jnt_bodyid) X(jnt_qposadr) X(jnt_dofadr) X(jnt_limited) X(dof_bodyid) X(dof_jntid) X(dof_parentid)  \
  X(geom_type) X(geom_bodyid) X(geom_contype) X(geom_conaffinity) X(geom_condim) X(site_bodyid) X(actuator_type)    \
  X(actuator_jntid) X(actuator_ctrllimited) X(actuator_forcelimited) X(eq_obj1id) X(eq_obj2id) X(eq_active)         \
  X(collision_pairs) X(geom_meshadr) X(geom_meshnum)

#define OM_DBL_ARRAYS(X)                                                                                            \
  X(opt_gravity) X(body_pos) X(body_quat) X(body_ipos) X(body_iquat) X(body_mass) X(body_inertia)                   \
  X(body_invweight0) X(jnt_pos) X(jnt_axis) X(jnt_range) X(jnt_solref) X(jnt_solimp) X(dof_damping)                 \
  X(dof_armature) X(dof_invweight0) X(qpos0) X(geom_size) X(geom_pos) X(geom_quat) X(geom_friction) X(geom_solref)  \
  X(geom_solimp) X(geom_margin) X(geom_gap) X(geom_rbound) X(site_pos) X(site_quat) X(actuator_gear)                \
  X(actuator_gainprm) X(actuator_biasprm) X(actuator_ctrlrange) X(actuator_forcerange) X(eq_solref) X(eq_solimp)    \
  X(eq_data) X(mesh_vert)

typedef struct om_model {
#define X(n) int n;
  OM_INT_SCALARS(X)
#undef X
#define X(n) double n;
  OM_DBL_SCALARS(X)
#undef X
#define X(n) int* n;
  OM_INT_ARRAYS(X)
#undef X
#define X(n) double* n;
  OM_DBL_ARRAYS(X)
#undef X
} om_model;

#define OM_MAXCON 256

typedef struct om_contact {
  double dist;
  double pos[3];
  double frame[9]; /* rows: normal (geom1 -> geom2), tangent1, tangent2 */
  double friction[5];
  double solref[2];
  double solimp[5];
  double mu; /* regularised cone mu */
  double margin; /* max of the two geoms' margins; the constraint acts on dist - margin */
  double gap;    /* max of the two geoms' gaps; a contact with dist >= margin - gap is reported but generates no force */
  int dim;
  int geom1, geom2;
  int efc_address;
} om_contact;

/* per-env data: state in, everything else out. Arrays are owned by the struct. */
#define OM_DATA_DBL_ARRAYS(X)                                                                                     \
  X(qpos, m->nq) X(qvel, m->nv) X(ctrl, m->nu) X(qfrc_applied, m->nv) X(xfrc_applied, 6 * m->nbody)               \
  X(qacc_warmstart, m->nv) X(eq_data, 7 * m->neq) X(xpos, 3 * m->nbody) X(xquat, 4 * m->nbody)                    \
  X(xmat, 9 * m->nbody) X(xipos, 3 * m->nbody) X(ximat, 9 * m->nbody) X(geom_xpos, 3 * m->ngeom)                  \
  X(geom_xmat, 9 * m->ngeom) X(site_xpos, 3 * m->nsite) X(site_xmat, 9 * m->nsite) X(xanchor, 3 * m->njnt)        \
  X(xaxis, 3 * m->njnt) X(dofS, 6 * m->nv) X(dofSdot, 6 * m->nv) X(bvel, 6 * m->nbody) X(qM, m->nv * m->nv)       \
  X(qLD, m->nv * m->nv) X(qfrc_bias, m->nv) X(qfrc_passive, m->nv) X(qfrc_actuator, m->nv) X(qfrc_smooth, m->nv)  \
  X(qacc_smooth, m->nv) X(qacc, m->nv) X(qfrc_constraint, m->nv) X(actuator_force, m->nu) X(solver_cost, 4)

#define OM_DATA_INT_ARRAYS(X) X(geom_contype, m->ngeom) X(geom_conaffinity, m->ngeom) X(eq_active, m->neq)

typedef struct om_data {
#define X(n, sz) double* n;
  OM_DATA_DBL_ARRAYS(X)
#undef X
#define X(n, sz) int* n;
  OM_DATA_INT_ARRAYS(X)
#undef X
  double time;
  int ncon;
  om_contact contact[OM_MAXCON];
  int nefc, nefc_cap;
  int ne, nl, nc; /* equality rows, limit rows, contact rows */
  double *efc_J, *efc_pos, *efc_margin, *efc_diagApprox, *efc_R, *efc_D, *efc_aref, *efc_vel, *efc_force, *efc_KBIP;
  int *efc_type, *efc_id;
  int solver_niter;
  int warning; /* bit0: contact buffer full, bit1: solver NaN/indefinite */
} om_data;

enum { OM_EFC_EQUALITY = 0, OM_EFC_LIMIT = 1, OM_EFC_CONTACT_ELLIPTIC = 2, OM_EFC_CONTACT_FRICTIONLESS = 3 };

om_model* om_model_new(void);
void om_model_free(om_model*);
/* copy a named table into the model; returns 0, or -1 if the name is unknown */
int om_model_set_int(om_model*, const char* name, const int* v, int n);
int om_model_set_dbl(om_model*, const char* name, const double* v, int n);

om_data* om_data_new(const om_model*);
void om_data_free(om_data*);
/* pointer to a named data array (double) / (int); n receives its length */
double* om_data_dbl(om_data*, const om_model*, const char* name, int* n);
int* om_data_int(om_data*, const om_model*, const char* name, int* n);
void om_reset_data(const om_model*, om_data*);
int om_data_scalar(const om_data*, const char* name); /* ncon nefc ne nl nc solver_niter warning */
om_contact* om_data_contacts(om_data*);
void om_clear_warning(om_data*);

void om_forward(const om_model*, om_data*); /* mj_forward */
void om_step(const om_model*, om_data*);    /* mj_step (Euler, implicit joint damping) */

/* stages, exposed for stage-wise parity tests */
void om_kinematics(const om_model*, om_data*);
void om_smooth(const om_model*, om_data*);    /* CRBA + factor + RNE bias + passive + actuation + qacc_smooth */
void om_collision(const om_model*, om_data*); /* broad + narrow phase -> d->contact */
void om_make_constraint(const om_model*, om_data*);
void om_solve(const om_model*, om_data*);     /* primal Newton, elliptic cones */
/* site velocity (mj_objectVelocity, world-aligned frame): out[0:3]=linear, out[3:6]=angular */
void om_site_velocity(const om_model*, const om_data*, int site, double* out6);

/* narrow phase entry used by unit tests: returns number of contacts written (<= 8) */
int om_collide_pair(const om_model*, const om_data*, int g1, int g2, om_contact* out);

#ifdef __cplusplus
}
#endif
#endif
