// rp_task.cuh — task-layer pieces shared by the device kernels and the host debug entry points.
#pragma once
#include <stdint.h>

#include "rp_math.cuh"
#include "rp_model.h"

#define RP_GOAL_DIM_HD 89
// per-environment scratch of the optimal-transport fingering reward: fingertips 30, key positions 3*88, cost
// matrix + solver doubles 10*88 + 10 + 2*88, solver integers (4*88 + 20) / 2
#define RP_OT_SCRATCH_DOUBLES (30 + 3 * RP_NK + 10 * RP_NK + 10 + 2 * RP_NK + (4 * RP_NK + 20 + 1) / 2)

namespace rp {

// Optimal-transport fingering reward of one environment after a step (piano_with_shadow_hands.py:333-369).
// t: the environment's step index after the step (> 0); qpos: its generalized positions after the step (key k
// at index k); site_xpos: its fingertip sites [hand][site][3], right hand first; goal: the song's goal table
// [T][89].  Fingertips are stacked left hand first, keys in ascending order, as the reference does.
RP_HD double ot_fingering_env(const RpDevModel* dm, int t, const double* qpos, const double* site_xpos, const uint8_t* goal, double* wd) {
  const int trow = t - 1 < dm->T - 1 ? t - 1 : dm->T - 1;  // the goal shown with the previous observation
  const uint8_t* gcur = goal + (size_t)trow * RP_GOAL_DIM_HD;
  double* tips = wd;
  double* keyp = wd + 30;
  for (int f = 0; f < 2 * RP_NSITE; f++) {
    const int src = (1 - f / RP_NSITE) * RP_NSITE + f % RP_NSITE;
    for (int c = 0; c < 3; c++) tips[3 * f + c] = site_xpos[3 * src + c];
  }
  int K = 0;
  for (int k = 0; k < RP_NK; k++) {
    if (!gcur[k]) continue;
    // key geom centre in the world: body origin rotated about the hinge by the key's angle
    const double sn = sin(qpos[k]), cs = cos(qpos[k]);
    const double ax[3] = {dm->key_axis[k][0], dm->key_axis[k][1], dm->key_axis[k][2]};
    const double v[3] = {-dm->key_jpos[k][0], -dm->key_jpos[k][1], -dm->key_jpos[k][2]};
    double axv[3];
    cross3(axv, ax, v);
    const double ad = dot3(ax, v) * (1 - cs);
    for (int c = 0; c < 3; c++) keyp[3 * K + c] = dm->key_pos[k][c] + dm->key_jpos[k][c] + (v[c] * cs + axv[c] * sn + ax[c] * ad);
    keyp[3 * K + 2] += 0.5 * dm->key_half[k][2];
    keyp[3 * K + 0] += 0.35 * dm->key_half[k][0];
    K++;
  }
  double* work = keyp + 3 * RP_NK;
  int* wi = reinterpret_cast<int*>(work + 10 * RP_NK + 10 + 2 * RP_NK);
  return ot_fingering_reward(tips, keyp, K, work, wi);
}

}  // namespace rp
