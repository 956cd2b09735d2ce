// rp_task_kernels.cu — task-layer kernels that run after the step kernel.
//
// rp_ot_fingering_kernel: the optimal-transport fingering reward of PianoWithShadowHands
// (robopianist/suite/tasks/piano_with_shadow_hands.py:333-369), which the reference adds in place of the
// fingering reward whenever that one is disabled — explicitly, or because the song carries no fingering
// (:109-111, :130-141).  One thread per environment: fingertip sites (left hand first, as the reference stacks
// them) x the keys of the goal that was shown with the previous observation, Euclidean distances, the linear
// sum assignment scipy solves there (rp::lsa_rect restates its algorithm, tests/test_hd_math.py pins it to
// scipy), mean of tolerance(d, (0, 0.01), margin 0.1).  The term is written to reward_terms[:, 3] and added to
// the reward the step kernel produced without a fingering term.  Kept out of the step kernel on purpose: that
// kernel's code does not change with it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rp_kernels.h"
#include "rp_task.cuh"

namespace {
__global__ void rp_ot_fingering_kernel(const RpDevModel* __restrict__ dm, int nenv, const int32_t* __restrict__ t_idx,
                                       const double* __restrict__ qpos, int nvtot, const double* __restrict__ site_xpos,
                                       const uint8_t* __restrict__ goal, double* reward, double* reward_terms, double* scratch) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= nenv) return;
  const int t = t_idx[env];
  if (t == 0) return;  // FIRST step of an episode: no reward
  const double r = rp::ot_fingering_env(dm, t, qpos + (size_t)env * nvtot, site_xpos + (size_t)env * 2 * RP_NSITE * 3, goal,
                                        scratch + (size_t)env * RP_OT_SCRATCH_DOUBLES);
  if (reward_terms) reward_terms[(size_t)env * 5 + 3] = r;
  if (reward) reward[env] += r;
}
}  // namespace

cudaError_t rp_launch_ot_fingering(const RpDevModel* dm, int nenv, const int32_t* t_idx, const double* qpos, int nvtot,
                                   const double* site_xpos, const uint8_t* goal, double* reward, double* reward_terms,
                                   double* scratch, cudaStream_t stream) {
  rp_ot_fingering_kernel<<<(nenv + 63) / 64, 64, 0, stream>>>(dm, nenv, t_idx, qpos, nvtot, site_xpos, goal, reward, reward_terms, scratch);
  return cudaGetLastError();
}
