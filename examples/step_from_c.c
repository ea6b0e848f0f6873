/* step_from_c.c -- a binder with no Python in it: creates a batch of environments from a compiled scene file, resets and steps them
 * through the C-ABI of include/furniture_b200.h with host buffers, prints what came back.
 *
 *   gcc -O2 -Iinclude examples/step_from_c.c -o /tmp/step_from_c -Lfurniture_b200 -lfurniture_b200 -Wl,-rpath,$PWD/furniture_b200
 *   /tmp/step_from_c furniture_b200/compiled/Sawyer_table_lack_0825.feb 64 3
 *
 * Exit code 0: stepped; 2: the library answered with an error (printed), e.g. no CUDA device. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "furniture_b200.h"

int main(int argc, char** argv) {
  const char* scene = argc > 1 ? argv[1] : "furniture_b200/compiled/Sawyer_table_lack_0825.feb";
  const int n = argc > 2 ? atoi(argv[2]) : 16, steps = argc > 3 ? atoi(argv[3]) : 2;
  fe_config cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.struct_bytes = (int32_t)sizeof cfg; /* the defaults of config/furniture.py */
  cfg.newton_iters = 30; cfg.ls_iters = 20; cfg.tolerance = 1e-6f;
  cfg.nsub = 50; cfg.max_episode_steps = 2000;
  cfg.discrete_grip = 1; cfg.rescale_actions = 1; cfg.auto_align = 1;
  cfg.alignment_pos_dist = 0.1; cfg.alignment_rot_dist_up = 0.9; cfg.alignment_rot_dist_forward = 0.9; cfg.alignment_project_dist = 0.3;
  cfg.ctrl_penalty_coef = 1e-3f; cfg.unstable_penalty_coef = 100.f; cfg.success_reward = 100.f; cfg.touch_reward = 10.f; cfg.pick_reward = 100.f;
  cfg.furn_xyz_rand = 0.02f; cfg.furn_rot_rand = 3.f; cfg.agent_xyz_rand = 0.001f;
  cfg.seed = 123;
  fe_handle* h = NULL;
  if (fe_create_from_file(scene, &cfg, n, 0, &h) != 0) { fprintf(stderr, "fe_create_from_file: %s\n", fe_last_error(NULL)); return 2; }
  const int od = fe_obs_dim(h), ad = fe_action_dim(h), id = fe_info_dim(h);
  float* act = (float*)calloc((size_t)n * ad, sizeof(float));
  float* obs = (float*)malloc(sizeof(float) * (size_t)n * od);
  float* rew = (float*)malloc(sizeof(float) * (size_t)n);
  uint8_t* done = (uint8_t*)malloc((size_t)n);
  int32_t* info = (int32_t*)malloc(sizeof(int32_t) * (size_t)n * id);
  if (fe_env_reset(h, NULL, NULL, NULL) != 0) { fprintf(stderr, "fe_env_reset: %s\n", fe_last_error(h)); return 2; }
  for (int i = 0; i < n; ++i) { act[(size_t)i * ad + ad - 2] = -1.f; act[(size_t)i * ad + ad - 1] = -1.f; } /* gripper open, no connect */
  double sum = 0.0;
  int ndone = 0, len = 0;
  for (int s = 0; s < steps; ++s) {
    if (fe_env_step_host(h, act, obs, rew, done, info) != 0) { fprintf(stderr, "fe_env_step_host: %s\n", fe_last_error(h)); return 2; }
    for (int i = 0; i < n; ++i) { sum += rew[i]; ndone += done[i]; }
    len = info[3];
  }
  printf("envs %d obs_dim %d action_dim %d steps %d episode_length %d done %d mean_reward %.6f obs0 %.6f\n", n, od, ad, steps, len, ndone, sum / ((double)n * steps), obs[0]);
  fe_destroy(h);
  free(act); free(obs); free(rew); free(done); free(info);
  return 0;
}
