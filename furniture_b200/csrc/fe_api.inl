// fe_api.inl -- host side of the C-ABI (include/furniture_b200.h), shared by the CUDA library (fe_host.cpp) and the
// lane-emulated test build (tests/emu/fe_emu.cpp).  The including file provides the plat_* functions.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/furniture_b200.h"
#include "fe_ik.h"
#include "fe_ctl.h"

struct FeField {
  std::string name;
  void* ptr;
  int dim, elem;
  bool writable;
};

struct fe_handle {
  int N = 0, device = 0;
  fe_config cfg;
  FeOpt opt;
  fe_model hm;
  fe_scene hs;
  fe_model* dm = nullptr;
  fe_scene* ds = nullptr;
  FeState st;
  FeDebug dbg;
  FeEnvState es;
  FeIkArgs ik = {nullptr, nullptr}; // control_type="ik" when ik.c is set (fe_enable_ik)
  int ik_act_dim = 8;               // 9 for "ik_quaternion"
  FeCtlArgs ctl = {nullptr, nullptr}; // one of the NEW_CONTROLLERS when ctl.c is set (fe_enable_controller)
  int ctl_act_dim = 0;
  int slice_words = 0;
  FeLayout lay;       // where each array of an env's slice starts (same for every env of the handle)
  std::vector<FeField> fields;
  std::vector<void*> allocs;
  std::string err;
  // staging for fe_env_step_host
  void *pin_act = nullptr, *pin_out = nullptr, *dev_act = nullptr, *dev_rew = nullptr, *dev_done = nullptr, *dev_info = nullptr;
  void* plat = nullptr;
};

static std::string g_create_err;
// the handle's device is made current for the duration of every entry point and the caller's restored afterwards
struct FeDevScope {
  fe_handle* h;
  int prev;
  explicit FeDevScope(fe_handle* h_) : h(h_), prev(plat_enter(h_)) {}
  ~FeDevScope() { plat_leave(h, prev); }
};

template <typename T>
static T* h_alloc(fe_handle* h, size_t count) {
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  void* p = plat_alloc(bytes);
  if (!p) return nullptr;
  plat_memset0(p, bytes);
  h->allocs.push_back(p);
  return (T*)p;
}
static void add_field(fe_handle* h, const char* name, void* ptr, int dim, int elem, bool writable) { h->fields.push_back({name, ptr, dim, elem, writable}); }
static FeField* find_field(fe_handle* h, const char* name) {
  for (auto& f : h->fields)
    if (f.name == name) return &f;
  return nullptr;
}
static int fail(fe_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

extern "C" {

size_t fe_model_sizeof(void) { return sizeof(fe_model); }
size_t fe_scene_sizeof(void) { return sizeof(fe_scene); }
size_t fe_config_sizeof(void) { return sizeof(fe_config); }
int fe_is_cuda(void) { return PLAT_IS_CUDA; }

const char* fe_last_error(const fe_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }
int fe_num_envs(const fe_handle* h) { return h->N; }
int fe_obs_dim(const fe_handle* h) { return h->hs.obs_dim; }
int fe_action_dim(const fe_handle* h) { return h->ctl.c ? h->ctl_act_dim : (h->ik.c ? h->ik_act_dim : h->hs.act_dim); }
int fe_info_dim(const fe_handle* h) { return FE_INFO_DIM; }
int fe_smem_bytes_per_env(const fe_handle* h) { return (h->slice_words + FE_ENV_EXTRA_WORDS) * 4; }
const float* fe_obs_dev(const fe_handle* h) { return h->es.obs; }

int fe_create(const void* model_blob, size_t model_bytes, const void* scene_blob, size_t scene_bytes, const fe_config* cfg, int n_envs, int device,
              fe_handle** out) {
  if (!model_blob || model_bytes != sizeof(fe_model)) return fail(nullptr, -1, "fe_create: model blob size mismatch (host/lib layout differ)");
  if (!cfg || cfg->struct_bytes != (int32_t)sizeof(fe_config)) return fail(nullptr, -1, "fe_create: fe_config size mismatch");
  if (n_envs <= 0) return fail(nullptr, -1, "fe_create: n_envs must be positive");
  fe_handle* h = new fe_handle();
  memcpy(&h->hm, model_blob, sizeof(fe_model));
  if (h->hm.magic != FE_MODEL_MAGIC || h->hm.struct_bytes != (int32_t)sizeof(fe_model)) { delete h; return fail(nullptr, -1, "fe_create: bad model magic"); }
  memset(&h->hs, 0, sizeof(fe_scene));
  if (scene_blob) {
    if (scene_bytes != sizeof(fe_scene)) { delete h; return fail(nullptr, -1, "fe_create: scene blob size mismatch"); }
    memcpy(&h->hs, scene_blob, sizeof(fe_scene));
    if (h->hs.magic != FE_SCENE_MAGIC) { delete h; return fail(nullptr, -1, "fe_create: bad scene magic"); }
  }
  h->N = n_envs; h->device = device; h->cfg = *cfg;
  h->opt.maxcon = cfg->maxcon > 0 ? cfg->maxcon : 40;
  h->opt.newton_iters = cfg->newton_iters > 0 ? cfg->newton_iters : 30;
  h->opt.ls_iters = cfg->ls_iters > 0 ? cfg->ls_iters : 20;
  h->opt.tolerance = cfg->tolerance > 0 ? cfg->tolerance : 1e-6f;
  h->opt.lockstep = 31;
  if (const char* e = getenv("FE_LOCKSTEP")) h->opt.lockstep = atoi(e);
  if (h->opt.maxcon > 255) { delete h; return fail(nullptr, -1, "fe_create: maxcon must be <= 255"); }
  if (h->hm.ngeom > 255) { delete h; return fail(nullptr, -1, "fe_create: ngeom must be <= 255"); }
  int rc = plat_init(h);
  if (rc) { std::string e = h->err; delete h; return fail(nullptr, rc, e); }
  FeDevScope dev_scope(h);
  const fe_model& m = h->hm;
  const size_t N = (size_t)n_envs;
  h->dm = (fe_model*)plat_alloc(sizeof(fe_model));
  h->ds = (fe_scene*)plat_alloc(sizeof(fe_scene));
  if (!h->dm || !h->ds) { fe_destroy(h); return fail(nullptr, -2, "fe_create: device allocation failed"); }
  h->allocs.push_back(h->dm); h->allocs.push_back(h->ds);
  plat_upload(h->dm, &h->hm, sizeof(fe_model));
  plat_upload(h->ds, &h->hs, sizeof(fe_scene));
  h->slice_words = fe_layout_build(&h->lay, &h->hm, h->opt);
  h->slice_words = (h->slice_words + 31) & ~31;
  FeState& s = h->st;
  s.N = n_envs;
  const int mc = h->opt.maxcon;
#define ALLOC(dst, T, dim, nm, wr) dst = h_alloc<T>(h, N * (size_t)(dim)); if (!dst) { fe_destroy(h); return fail(nullptr, -2, "fe_create: device allocation failed"); } \
  if (nm) add_field(h, nm, dst, (dim), (int)sizeof(T), wr);
  ALLOC(s.qpos, float, m.nq, "qpos", true) ALLOC(s.qvel, float, m.nv, "qvel", true) ALLOC(s.warm, float, m.nv, "qacc_warmstart", true)
  ALLOC(s.ctrl, float, m.nu, "ctrl", true) ALLOC(s.qfrc_applied, float, m.nr, "qfrc_applied", true) ALLOC(s.gravcomp, float, m.npart, "gravcomp", true)
  ALLOC(s.eq_data, float, 7 * m.neq, "eq_data", true) ALLOC(s.mpos, float, 3 * m.nmov, "static_pos", true) ALLOC(s.contype, int, m.ngeom, "geom_contype", true) ALLOC(s.conaff, int, m.ngeom, "geom_conaffinity", true)
  ALLOC(s.eq_active, int, m.neq, "eq_active", true) ALLOC(s.bias, float, m.nr, "qfrc_bias", false)
  ALLOC(s.lpos, float, 3 * m.nlink, "link_xpos", false) ALLOC(s.lquat, float, 4 * m.nlink, "link_xquat", false) ALLOC(s.lvel, float, 6 * m.nlink, "link_vel", false)
  ALLOC(s.touch, int, m.npart, "touch", false) ALLOC(s.flags, int, 1, "flags", true) ALLOC(s.ncon, int, 1, "ncon", false) ALLOC(s.niter, int, 1, "niter", false) ALLOC(s.stats, int, FE_NSTAT, "stats", false) ALLOC(s.order, int, 1, "order", true)
  FeDebug& d = h->dbg;
  ALLOC(d.Mr, float, m.nr * m.nr, "dbg_Mr", false) ALLOC(d.fs, float, m.nv, "dbg_fs", false) ALLOC(d.as, float, m.nv, "dbg_as", false)
  ALLOC(d.linert, float, 10 * m.nlink, "dbg_linert", false) ALLOC(d.x, float, m.nv, "dbg_x", false) ALLOC(d.fc, float, m.nv, "dbg_fc", false)
  ALLOC(d.lmat, float, 9 * m.nlink, "link_xmat", false) ALLOC(d.S, float, 6 * m.nr, "dbg_S", false)
  ALLOC(d.c_dist, float, mc, "con_dist", false) ALLOC(d.c_pos, float, 3 * mc, "con_pos", false) ALLOC(d.c_frame, float, 9 * mc, "con_frame", false)
  ALLOC(d.c_aref, float, 3 * mc, "con_aref", false) ALLOC(d.c_D, float, 2 * mc, "con_D", false) ALLOC(d.c_f, float, 3 * mc, "con_force", false)
  ALLOC(d.c_geom, int, mc, "con_geom", false) ALLOC(d.c_state, int, mc, "con_state", false)
  FeEnvState& e = h->es;
  const int nsite = m.nsite > 0 ? m.nsite : 1;
  ALLOC(e.obs, float, h->hs.obs_dim > 0 ? h->hs.obs_dim : 1, "obs", false) ALLOC(e.group, int, m.npart, "group", true) ALLOC(e.site_connected, int, nsite, "site_connected", true)
  ALLOC(e.num_connected, int, 1, "num_connected", true) ALLOC(e.prev_num_connected, int, 1, "prev_num_connected", true)
  ALLOC(e.touched, int, m.npart, "touched", true) ALLOC(e.picked, int, m.npart, "picked", true) ALLOC(e.episode_len, int, 1, "episode_length", true)
  ALLOC(e.mt, uint32_t, 624, "mt_state", true) ALLOC(e.mt_pos, int, 1, "mt_pos", true) ALLOC(e.done, int, 1, "done", false) ALLOC(e.robot_contype, int, m.ngeom, nullptr, false)
  ALLOC(e.robot_conaff, int, m.ngeom, nullptr, false) ALLOC(e.episode_reward, float, 1, "episode_reward", false)
#undef ALLOC
  // initial per-env model state: MjSim.reset() semantics (qpos = qpos0 of the XML, masks / welds from the model)
  {
    std::vector<int> ct(N * m.ngeom), ca(N * m.ngeom), ea(N * (m.neq > 0 ? m.neq : 1));
    std::vector<float> ed(N * 7 * (m.neq > 0 ? m.neq : 1));
    for (size_t n = 0; n < N; ++n) {
      for (int g = 0; g < m.ngeom; ++g) { ct[n * m.ngeom + g] = m.geom_contype0[g]; ca[n * m.ngeom + g] = m.geom_conaffinity0[g]; }
      for (int q = 0; q < m.neq; ++q) { ea[n * m.neq + q] = m.eq_active0[q]; for (int k = 0; k < 7; ++k) ed[(n * m.neq + q) * 7 + k] = m.eq_data0[q][k]; }
    }
    plat_upload(s.contype, ct.data(), sizeof(int) * N * m.ngeom);
    plat_upload(s.conaff, ca.data(), sizeof(int) * N * m.ngeom);
    if (m.neq) { plat_upload(s.eq_active, ea.data(), sizeof(int) * N * m.neq); plat_upload(s.eq_data, ed.data(), sizeof(float) * N * 7 * m.neq); }
    if (m.nmov > 0) { // movable static geoms start where the model puts them
      std::vector<float> mp(N * 3 * m.nmov, 0.f);
      for (size_t n = 0; n < N; ++n)
        for (int g = 0; g < m.ngeom; ++g)
          if (m.geom_mov[g]) for (int k = 0; k < 3; ++k) mp[(n * m.nmov + (m.geom_mov[g] - 1)) * 3 + k] = m.geom_pos[g][k];
      plat_upload(s.mpos, mp.data(), sizeof(float) * mp.size());
    }
    std::vector<int> ord(N);
    for (size_t n = 0; n < N; ++n) ord[n] = (int)n;
    plat_upload(s.order, ord.data(), sizeof(int) * N);
    // numpy RandomState(seed + env): init_genrand (Knuth), position at the end of the state so that the first draw regenerates it
    std::vector<uint32_t> mt(N * 624);
    std::vector<int> mpos(N, 624);
    for (size_t n = 0; n < N; ++n) {
      uint32_t* s = mt.data() + n * 624;
      s[0] = (uint32_t)((cfg->seed + n) & 0xffffffffu);
      for (uint32_t i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + i;
    }
    if (cfg->furn_size_rand != 0.f) // the reference draws the size factor while loading the model (furniture.py:1989-1991): one double
      for (size_t n = 0; n < N; ++n) { fe_mt_twist(mt.data() + n * 624); mpos[n] = 2; }
    plat_upload(e.mt, mt.data(), sizeof(uint32_t) * N * 624);
    plat_upload(e.mt_pos, mpos.data(), sizeof(int) * N);
  }
  h->dev_act = h_alloc<float>(h, N * (size_t)(h->hs.act_dim > 0 ? h->hs.act_dim : 1));
  h->dev_rew = h_alloc<float>(h, N);
  h->dev_done = h_alloc<uint8_t>(h, N);
  h->dev_info = h_alloc<int32_t>(h, N * FE_INFO_DIM);
  *out = h;
  return 0;
}

int fe_scene_file_write(const char* path, const void* model_blob, size_t model_bytes, const void* scene_blob, size_t scene_bytes) {
  if (!path || !model_blob || model_bytes != sizeof(fe_model) || !scene_blob || scene_bytes != sizeof(fe_scene))
    return fail(nullptr, -1, "fe_scene_file_write: blob size mismatch");
  FILE* f = fopen(path, "wb");
  if (!f) return fail(nullptr, -7, std::string("fe_scene_file_write: cannot open ") + path);
  const uint32_t hdr[4] = {0x31424546u /* "FEB1" */, (uint32_t)sizeof(fe_model), (uint32_t)sizeof(fe_scene), 0u};
  const bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(model_blob, sizeof(fe_model), 1, f) == 1 && fwrite(scene_blob, sizeof(fe_scene), 1, f) == 1;
  fclose(f);
  return ok ? 0 : fail(nullptr, -7, std::string("fe_scene_file_write: short write to ") + path);
}
int fe_create_from_file(const char* path, const fe_config* cfg, int n_envs, int device, fe_handle** out) {
  FILE* f = path ? fopen(path, "rb") : nullptr;
  if (!f) return fail(nullptr, -7, std::string("fe_create_from_file: cannot open ") + (path ? path : "(null)"));
  uint32_t hdr[4] = {0, 0, 0, 0};
  std::vector<unsigned char> mb(sizeof(fe_model)), sb(sizeof(fe_scene));
  const bool ok = fread(hdr, sizeof(hdr), 1, f) == 1 && hdr[0] == 0x31424546u && hdr[1] == sizeof(fe_model) && hdr[2] == sizeof(fe_scene) &&
                  fread(mb.data(), sizeof(fe_model), 1, f) == 1 && fread(sb.data(), sizeof(fe_scene), 1, f) == 1;
  fclose(f);
  if (!ok) return fail(nullptr, -7, std::string("fe_create_from_file: not a scene file of this library version: ") + path);
  return fe_create(mb.data(), mb.size(), sb.data(), sb.size(), cfg, n_envs, device, out);
}

void fe_destroy(fe_handle* h) {
  if (!h) return;
  plat_fini(h);
  {
    FeDevScope dev_scope(h);
    for (void* p : h->allocs) plat_free(p);
  }
  delete h;
}

int fe_field_dim(fe_handle* h, const char* name, int* dim, int* elem_bytes) {
  FeField* f = find_field(h, name);
  if (!f) return fail(h, -3, std::string("unknown field ") + name);
  if (dim) *dim = f->dim;
  if (elem_bytes) *elem_bytes = f->elem;
  return 0;
}
int fe_get_field(fe_handle* h, const char* name, void* dst, size_t bytes) {
  FeField* f = find_field(h, name);
  if (!f) return fail(h, -3, std::string("unknown field ") + name);
  size_t want = (size_t)h->N * f->dim * f->elem;
  if (bytes != want) return fail(h, -4, std::string("size mismatch for field ") + name);
  FeDevScope dev_scope(h);
  plat_sync(h);
  plat_download(dst, f->ptr, bytes);
  return 0;
}
int fe_set_field(fe_handle* h, const char* name, const void* src, size_t bytes) {
  FeField* f = find_field(h, name);
  if (!f) return fail(h, -3, std::string("unknown field ") + name);
  if (!f->writable) return fail(h, -5, std::string("field is read-only: ") + name);
  size_t want = (size_t)h->N * f->dim * f->elem;
  if (bytes != want) return fail(h, -4, std::string("size mismatch for field ") + name);
  FeDevScope dev_scope(h);
  plat_sync(h);
  plat_upload(f->ptr, src, bytes);
  return 0;
}
int fe_get_state(fe_handle* h, float* qpos, float* qvel) {
  int rc = fe_get_field(h, "qpos", qpos, sizeof(float) * h->N * h->hm.nq);
  return rc ? rc : fe_get_field(h, "qvel", qvel, sizeof(float) * h->N * h->hm.nv);
}
int fe_set_state(fe_handle* h, const float* qpos, const float* qvel) {
  int rc = fe_set_field(h, "qpos", qpos, sizeof(float) * h->N * h->hm.nq);
  return rc ? rc : fe_set_field(h, "qvel", qvel, sizeof(float) * h->N * h->hm.nv);
}

int fe_sim_forward(fe_handle* h, void* stream) { return plat_run_sim(h, 1, 1, stream); }
int fe_sim_step(fe_handle* h, int nsub, void* stream) {
  if (nsub <= 0) return fail(h, -1, "fe_sim_step: nsub must be positive");
  return plat_run_sim(h, nsub, 0, stream);
}

int fe_env_reset(fe_handle* h, const uint8_t* env_mask_dev, float* obs_dev, void* stream) {
  if (h->hs.magic != FE_SCENE_MAGIC) return fail(h, -6, "fe_env_reset: handle was created without a scene blob");
  int rc = plat_run_reset(h, env_mask_dev, stream);
  if (rc) return rc;
  if (obs_dev) plat_copy_d2d(h, obs_dev, h->es.obs, sizeof(float) * (size_t)h->N * h->hs.obs_dim, stream);
  return 0;
}
int fe_env_step(fe_handle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, int32_t* info_dev, void* stream) {
  if (h->hs.magic != FE_SCENE_MAGIC) return fail(h, -6, "fe_env_step: handle was created without a scene blob");
  if (!actions_dev) return fail(h, -1, "fe_env_step: actions is NULL");
  int rc = plat_run_step(h, actions_dev, reward_dev ? reward_dev : (float*)h->dev_rew, done_dev ? done_dev : (uint8_t*)h->dev_done,
                         info_dev ? info_dev : (int32_t*)h->dev_info, stream);
  if (rc) return rc;
  if (obs_dev) plat_copy_d2d(h, obs_dev, h->es.obs, sizeof(float) * (size_t)h->N * h->hs.obs_dim, stream);
  return 0;
}
int fe_env_step_packed(fe_handle* h, const float* actions_dev, float* packed_dev, int32_t* info_dev, void* stream) {
  if (h->hs.magic != FE_SCENE_MAGIC) return fail(h, -6, "fe_env_step_packed: handle was created without a scene blob");
  if (!actions_dev || !packed_dev) return fail(h, -1, "fe_env_step_packed: actions / packed is NULL");
  h->es.packed = packed_dev;
  int rc = plat_run_step(h, actions_dev, (float*)h->dev_rew, (uint8_t*)h->dev_done, info_dev ? info_dev : (int32_t*)h->dev_info, stream);
  h->es.packed = nullptr;
  return rc;
}
int fe_env_step_host(fe_handle* h, const float* actions, float* obs, float* reward, uint8_t* done, int32_t* info) {
  if (h->hs.magic != FE_SCENE_MAGIC) return fail(h, -6, "fe_env_step_host: handle was created without a scene blob");
  return plat_step_host(h, actions, obs, reward, done, info);
}

int fe_set_max_episode_steps(fe_handle* h, int max_episode_steps) {
  if (max_episode_steps <= 0) return fail(h, -1, "fe_set_max_episode_steps: must be positive");
  h->cfg.max_episode_steps = max_episode_steps; // fe_config travels by value with every launch
  return 0;
}

int fe_enable_ik(fe_handle* h, const fe_ik_config* ikc) {
  if (!ikc || ikc->struct_bytes != (int32_t)sizeof(fe_ik_config)) return fail(h, -1, "fe_enable_ik: fe_ik_config size mismatch");
  if (h->hs.magic != FE_SCENE_MAGIC) return fail(h, -6, "fe_enable_ik: handle was created without a scene blob");
  if (ikc->narms != h->hs.narms || ikc->narms < 1 || ikc->narms > 2 || h->hs.narm != 7 * ikc->narms || h->hs.hand_link[0] < 0 || (ikc->narms == 2 && h->hs.hand_link[1] < 0))
    return fail(h, -1, "fe_enable_ik: the IK control type is built for arms of 7 joints (one arm: Sawyer, two: Baxter) and the config must name as many arms as the scene has");
  if (h->ctl.c) return fail(h, -1, "fe_enable_ik: the handle already runs a torque controller");
  if (ikc->action_repeat < 1 || ikc->action_repeat > 16 || ikc->max_iters < 1 || ikc->max_iters > 1000) return fail(h, -1, "fe_enable_ik: action_repeat / max_iters out of range");
  for (int a = 0; a < ikc->narms; ++a)
    for (int k = 0; k < 7; ++k) if (ikc->arm[a].arm_qadr[k] < 0 || ikc->arm[a].arm_qadr[k] >= h->hm.nq) return fail(h, -1, "fe_enable_ik: arm_qadr outside qpos");
  FeDevScope dev_scope(h);
  if (!h->ik.c) {
    fe_ik_config* d = (fe_ik_config*)plat_alloc(sizeof(fe_ik_config));
    FeIkState* st = h_alloc<FeIkState>(h, (size_t)h->N);
    if (!d || !st) return fail(h, -2, "fe_enable_ik: device allocation failed");
    h->allocs.push_back(d);
    h->ik.c = d; h->ik.st = st;
    add_field(h, "ik_state", st, (int)sizeof(FeIkState), 1, true);
  }
  plat_upload((void*)h->ik.c, ikc, sizeof(fe_ik_config));
  h->ik_act_dim = ikc->narms * (ikc->quaternion_mode ? 7 : 6) + ikc->narms + 1;
  return 0;
}
int fe_enable_controller(fe_handle* h, const fe_ctl_config* cc) {
  if (!cc || cc->struct_bytes != (int32_t)sizeof(fe_ctl_config)) return fail(h, -1, "fe_enable_controller: fe_ctl_config size mismatch");
  if (cc->mode < 0 || cc->mode > FE_CTL_POS || cc->control_dim < 1 || cc->control_dim > 7 || !(cc->ramp_steps >= 1.0)) return fail(h, -1, "fe_enable_controller: bad controller parameters");
  if (h->hs.magic != FE_SCENE_MAGIC) return fail(h, -6, "fe_enable_controller: handle was created without a scene blob");
  if (h->hs.narms != 1 || h->hs.narm != 7 || h->hs.hand_link[0] < 0) return fail(h, -1, "fe_enable_controller: built for the one-arm 7-joint (Sawyer) env");
  if (h->ik.c) return fail(h, -1, "fe_enable_controller: the handle already runs the IK control type");
  for (int u = 0; u < h->hm.nu; ++u) // ctrl = qfrc_bias + torques only means torques on motor actuators (robot_torque.xml)
    if (h->hs.act_src[u] < h->hs.narm && (h->hm.act_gain[u] != 1.f || h->hm.act_bias[u][1] != 0.f || h->hm.act_bias[u][2] != 0.f))
      return fail(h, -1, "fe_enable_controller: the arm's actuators are not motors (compose the scene with the torque-actuated robot)");
  FeDevScope dev_scope(h);
  if (!h->ctl.c) {
    fe_ctl_config* d = (fe_ctl_config*)plat_alloc(sizeof(fe_ctl_config));
    FeCtlState* st = h_alloc<FeCtlState>(h, (size_t)h->N);
    if (!d || !st) return fail(h, -2, "fe_enable_controller: device allocation failed");
    h->allocs.push_back(d);
    h->ctl.c = d; h->ctl.st = st;
    add_field(h, "ctl_state", st, (int)sizeof(FeCtlState), 1, true);
  }
  plat_upload((void*)h->ctl.c, cc, sizeof(fe_ctl_config));
  h->ctl_act_dim = cc->control_dim + 2;
  return 0;
}
int fe_ctl_eval(fe_handle* h, const fe_ctl_config* cc, int n_episodes, const int32_t* first, const int32_t* count, int n_records, const uint8_t* reset,
                const uint8_t* policy_step, const double* action, const double* readings, double* torques) {
  if (!cc || cc->struct_bytes != (int32_t)sizeof(fe_ctl_config)) return fail(h, -1, "fe_ctl_eval: fe_ctl_config size mismatch");
  if (cc->mode < 0 || cc->mode > FE_CTL_POS || cc->control_dim < 1 || cc->control_dim > 7 || !(cc->ramp_steps >= 1.0)) return fail(h, -1, "fe_ctl_eval: bad controller parameters");
  static_assert(sizeof(FeCtlIn) == 123 * sizeof(double), "FeCtlIn is the 123-double reading record of the C-ABI");
  if (n_episodes <= 0 || n_records <= 0) return 0;
  return plat_ctl_eval(h, cc, n_episodes, first, count, n_records, reset, policy_step, action, (const FeCtlIn*)readings, torques);
}
int fe_dense_info_dim(void) { return FE_DENSE_INFO; }
size_t fe_dense_recipe_sizeof(void) { return sizeof(fe_dense_recipe); }
int fe_enable_dense_reward(fe_handle* h, const fe_dense_config* dc) {
  if (!dc || dc->struct_bytes != (int32_t)sizeof(fe_dense_config)) return fail(h, -1, "fe_enable_dense_reward: fe_dense_config size mismatch");
  if (h->hs.magic != FE_SCENE_MAGIC) return fail(h, -6, "fe_enable_dense_reward: handle was created without a scene blob");
  const fe_dense_recipe& rc = h->hs.dense;
  if (rc.nsub <= 0 || rc.nsub > FE_DENSE_MAXSUB) return fail(h, -1, "fe_enable_dense_reward: the scene carries no assembly recipe");
  if (h->hs.narms != 1) return fail(h, -1, "fe_enable_dense_reward: the dense reward is defined for the one-arm (Sawyer) env");
  if ((dc->phase_ob != 0) != (h->hs.phase_ob != 0)) return fail(h, -1, "fe_enable_dense_reward: phase_ob disagrees with the scene's obs layout");
  for (int s = 0; s < rc.nsub; ++s) {
    const int ids[5] = {rc.leg_site[s], rc.table_site[s], rc.gl_site[s], rc.gr_site[s], rc.griptip_site};
    for (int k = 0; k < 5; ++k) if (ids[k] < 0 || ids[k] >= h->hm.nsite) return fail(h, -1, "fe_enable_dense_reward: recipe names a site the model does not have");
    if (rc.leg_part[s] < 0 || rc.leg_part[s] >= h->hm.npart) return fail(h, -1, "fe_enable_dense_reward: recipe names a part the model does not have");
  }
  if (rc.grip_site < 0 || rc.grip_site >= h->hm.nsite) return fail(h, -1, "fe_enable_dense_reward: grip_site missing");
  FeDevScope dev_scope(h);
  if (!h->es.dense) {
    fe_dense_config* d = (fe_dense_config*)plat_alloc(sizeof(fe_dense_config));
    FeDenseState* st = h_alloc<FeDenseState>(h, (size_t)h->N);
    float* inf = h_alloc<float>(h, (size_t)h->N * FE_DENSE_INFO);
    if (!d || !st || !inf) return fail(h, -2, "fe_enable_dense_reward: device allocation failed");
    h->allocs.push_back(d);
    h->es.dense = d; h->es.dstate = st; h->es.dinfo = inf;
    add_field(h, "dense_info", inf, FE_DENSE_INFO, (int)sizeof(float), false);
    add_field(h, "dense_state", st, (int)sizeof(FeDenseState), 1, true);
  }
  plat_upload((void*)h->es.dense, dc, sizeof(fe_dense_config));
  return 0;
}
int fe_dense_eval(fe_handle* h, const fe_dense_config* dc, const void* recipe_blob, size_t recipe_bytes, const double* thr4, int n_goal, int n_episodes,
                  const int32_t* first, const int32_t* count, int n_records, int nsite, int npart, int act_dim, const double* site_pos, const double* site_mat,
                  const double* part_pos, const uint8_t* touch, const uint8_t* reset, const uint8_t* connected, const double* ac, double* reward, uint8_t* done,
                  double* info) {
  if (!dc || dc->struct_bytes != (int32_t)sizeof(fe_dense_config)) return fail(h, -1, "fe_dense_eval: fe_dense_config size mismatch");
  if (!recipe_blob || recipe_bytes != sizeof(fe_dense_recipe)) return fail(h, -1, "fe_dense_eval: recipe blob size mismatch");
  if (n_episodes <= 0 || n_records <= 0) return 0;
  return plat_dense_eval(h, dc, (const fe_dense_recipe*)recipe_blob, thr4, n_goal, n_episodes, first, count, n_records, nsite, npart, act_dim, site_pos, site_mat,
                         part_pos, touch, reset, connected, ac, reward, done, info);
}

int fe_is_aligned(fe_handle* h, int n, const double* p1, const double* m1, const double* p2, const double* m2, const double* angles, const int32_t* nangles,
                  const double* thr, uint8_t* aligned, double* tq) {
  if (n <= 0) return 0;
  return plat_is_aligned(h, n, p1, m1, p2, m2, angles, nangles, thr, aligned, tq);
}

} // extern "C"
