// fe_model.h -- engine-side compiled model: constant tables shared by all envs of one scene.
//
// Built on the host (furniture_b200/engine_model.py) from the MJCF tables: every body without a joint is fused into
// the link it is welded to (inertias combined, geoms/sites re-expressed in the link frame), the static pair filters of
// mj_collision are resolved into a pair list, and only colliding geoms are kept.  Replaces the MjModel the reference
// builds through load_model_from_xml (furniture/env/models/base.py:113-115).
#pragma once
#include <stdint.h>

#define FE_MODEL_MAGIC 0x46453035 /* "FE05" */
#define FE_MAXLINK 32
#define FE_MAXRDOF 20
#define FE_MAXPART 16
#define FE_MAXDOF 116 /* FE_MAXRDOF + 6 * FE_MAXPART */
#define FE_MAXGEOM 96
#define FE_MAXPAIR 2048
#define FE_MAXSITE 256
#define FE_MAXEQ 40
#define FE_MAXU 20
#define FE_MAXMESHVERT 512 /* convex-hull vertices of all mesh colliders of a scene */

enum { FE_JNT_FREE = 0, FE_JNT_SLIDE = 2, FE_JNT_HINGE = 3 };
enum { FE_GEOM_PLANE = 0, FE_GEOM_SPHERE = 2, FE_GEOM_CAPSULE = 3, FE_GEOM_CYLINDER = 5, FE_GEOM_BOX = 6, FE_GEOM_MESH = 7 };
enum { FE_ACT_MOTOR = 0, FE_ACT_POSITION = 1, FE_ACT_VELOCITY = 2 };
/* geom_tag bits */
enum { FE_TAG_FLOOR = 1, FE_TAG_LFINGER = 2, FE_TAG_RFINGER = 4, FE_TAG_ROBOT = 8, FE_TAG_LFINGER2 = 16, FE_TAG_RFINGER2 = 32 /* fingers of a second arm */, FE_TAG_PART_SHIFT = 8 };

typedef struct fe_model {
  int32_t magic, struct_bytes;
  int32_t nq, nv, nu, nlink, nrlink, nr, npart, ngeom, npair, nsite, neq, maxdepth;
  int32_t has_margin, has_gap; /* any geom with margin > 0 / gap > 0: scenes without skip the per-pair lookups */
  int32_t nmov; /* static geoms whose world position is per-env state (the Cursor agent's two cursors, moved through sim.model.body_pos) */
  float timestep, gravity[3], impratio, meaninertia, robot_ref[3];
  /* links: robot links [0, nrlink) in parent-before-child order (one hinge/slide dof each), then parts (free joint) */
  int32_t link_parent[FE_MAXLINK]; /* -1 = world */
  int32_t link_jtype[FE_MAXLINK], link_qadr[FE_MAXLINK], link_dadr[FE_MAXLINK], link_depth[FE_MAXLINK];
  int32_t link_ancmask[FE_MAXLINK]; /* robot links: bit d set if robot dof d moves this link */
  float link_pos[FE_MAXLINK][3], link_quat[FE_MAXLINK][4]; /* frame in the parent link frame at zero joint value */
  float link_jaxis[FE_MAXLINK][3], link_jpos[FE_MAXLINK][3];
  float link_mass[FE_MAXLINK], link_com[FE_MAXLINK][3];
  float link_inertia_c[FE_MAXLINK][6]; /* about the CoM, link frame: xx yy zz xy xz yz */
  float link_inertia_o[FE_MAXLINK][6]; /* about the link origin, link frame */
  float dof_damping[FE_MAXDOF];
  float rdof_armature[FE_MAXRDOF]; /* added to the diagonal of the joint-space inertia (MJCF joint armature) */
  int32_t rdof_limited[FE_MAXRDOF];
  float rdof_range[FE_MAXRDOF][2], rdof_invweight[FE_MAXRDOF], rdof_solref[FE_MAXRDOF][2], rdof_solimp[FE_MAXRDOF][3];
  /* actuators (joint transmission on robot dofs) */
  int32_t act_type[FE_MAXU], act_dof[FE_MAXU], act_qadr[FE_MAXU], act_ctrllimited[FE_MAXU], act_forcelimited[FE_MAXU];
  float act_gear[FE_MAXU], act_gain[FE_MAXU], act_bias[FE_MAXU][3], act_ctrlrange[FE_MAXU][2], act_forcerange[FE_MAXU][2];
  /* colliding geoms */
  int32_t geom_type[FE_MAXGEOM], geom_link[FE_MAXGEOM], geom_contype0[FE_MAXGEOM], geom_conaffinity0[FE_MAXGEOM], geom_tag[FE_MAXGEOM];
  float geom_pos[FE_MAXGEOM][3], geom_mat[FE_MAXGEOM][9], geom_size[FE_MAXGEOM][3], geom_rbound[FE_MAXGEOM];
  float geom_friction[FE_MAXGEOM], geom_solref[FE_MAXGEOM][2], geom_solimp[FE_MAXGEOM][3], geom_invweight[FE_MAXGEOM];
  float geom_margin[FE_MAXGEOM]; /* contacts are generated below max(margin1, margin2); the constraint acts on dist - margin */
  float geom_gap[FE_MAXGEOM];    /* a contact with dist >= margin - gap is reported (touch flags) but generates no force: gap > 0 marks sensor geoms */
  int32_t geom_mov[FE_MAXGEOM];  /* 1 + slot of a movable static geom, 0 otherwise */
  int32_t geom_meshadr[FE_MAXGEOM], geom_meshnum[FE_MAXGEOM]; /* mesh colliders: their convex-hull vertices in mesh_vert (geom frame) */
  float mesh_vert[FE_MAXMESHVERT][3];
  int32_t pair_g1[FE_MAXPAIR], pair_g2[FE_MAXPAIR]; /* type(g1) <= type(g2) */
  /* sites the env layer reads */
  int32_t site_link[FE_MAXSITE];
  float site_pos[FE_MAXSITE][3], site_quat[FE_MAXSITE][4];
  /* weld equalities between parts */
  int32_t eq_link1[FE_MAXEQ], eq_link2[FE_MAXEQ], eq_active0[FE_MAXEQ];
  float eq_solref[FE_MAXEQ][2], eq_solimp[FE_MAXEQ][3], eq_invw_t[FE_MAXEQ], eq_invw_r[FE_MAXEQ], eq_data0[FE_MAXEQ][7];
} fe_model;
