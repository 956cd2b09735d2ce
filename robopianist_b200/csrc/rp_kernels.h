// rp_kernels.h — launch interface between the C-ABI host code and the step kernel.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/roboplay.h"
#include "rp_model.h"

#ifndef RP_EPB
#define RP_EPB RP_BLOCK_ENVS  // environments per thread block of the phase-locked launch shape (tuning builds override it)
#endif

enum { RP_MODE_STEP = 0, RP_MODE_RESET = 1, RP_MODE_PHYSICS = 2 };
enum { RP_STEP_FIRST_ = 0, RP_STEP_MID_ = 1, RP_STEP_LAST_ = 2 };
#define RP_GOAL_DIM_ 89
#define RP_METRICS_STRIDE 16
// MIDI event word: (tick << 9) | (class << 7) | key; tick = physics substeps since the episode start
// (event time = tick * timestep); class 0 NoteOn, 1 NoteOff, 2 SustainOn, 3 SustainOff.  Sorting the
// words of one environment ascending gives the reference's message order (midi_module.py:47-98).
#define RP_MIDI_WORD(tick, cls, key) (((uint32_t)(tick) << 9) | ((uint32_t)(cls) << 7) | (uint32_t)(key))

// debug dump layout (doubles per env), filled for substep 0 when RpKernelArgs::dbg != nullptr
#define RP_DBG_QACC_SMOOTH 0
#define RP_DBG_QACC 160
#define RP_DBG_QFRC_C 320
#define RP_DBG_BIAS 480
#define RP_DBG_SCALARS 640
#define RP_DBG_CONTACTS 656   // 24 x 16
#define RP_DBG_M 1040         // 26 x 26 (hand 0)
#define RP_DBG_MINV 1716      // 26 x 26 (hand 0)
#define RP_DBG_PROFILE 2392      // 2 x 24 phase cycle counters (RP_PROFILE builds)
#define RP_DBG_NPROF 24
#define RP_DBG_STRIDE 2448

struct RpKernelArgs {
  const RpDevModel* dm;
  int nenv, mode, nsub, iterations;
  double tolerance;
  // state, env-major: [N][nv], [N][nu]
  double *qpos, *qvel, *qacc_ws, *ctrl;
  int32_t *t_idx, *step_type_state;
  double* sustain;
  // inputs
  const double* action;      // [N][act_dim]      (STEP)
  const double* ctrl_in;     // [N][nu]           (PHYSICS, optional)
  const uint8_t* reset_mask; // [N]               (RESET, optional)
  const uint8_t* goal;       // [T][89]
  const int8_t* fingering;   // [T][88]
  // outputs (optional)
  double *obs, *reward, *discount, *reward_terms;
  uint8_t* step_type;
  int32_t* diag;             // [N][3] ncon, niter, flags
  double* act_force;         // [N][nu]
  double* site_xpos;         // [N][10][3]
  uint8_t* activation;       // [N][88]
  double* metrics;           // [N][RP_METRICS_STRIDE]
  uint32_t* midi_events;     // [N][midi_cap] packed key / sustain events of the running episode (optional)
  int32_t* midi_count;       // [N] events emitted so far this episode (may exceed midi_cap: overflow)
  int midi_cap, hand_offset_fresh;  // hand_offset_fresh: every episode starts from the unshifted pose (recompile_physics=True)
  double* hand_offset;       // [N] y offset of both hands' attachment (randomize_hand_positions), optional
  uint64_t* rng_state;       // [N] per-environment generator state of the offset draws (with hand_offset)
  int record_cap, pad_rc;    // > 0: contact-record capacity below RP_MAXCON (test hook of the deepest-first selection; hand-collision instantiation)
  int32_t* work;             // [N] work estimate of this step (out)
  const int32_t* order;      // [N] block -> environment (optional)
  double* dbg;
};

size_t rp_kernel_smem_bytes();
cudaError_t rp_launch_step(const RpKernelArgs& a, int nhand, cudaStream_t stream, int block_envs, int hand_collisions);
cudaError_t rp_launch_order(const int32_t* work, int32_t* order, int n, cudaStream_t stream);
// optimal-transport fingering reward (rp_task_kernels.cu; per-environment scratch: RP_OT_SCRATCH_DOUBLES of rp_task.cuh)
cudaError_t rp_launch_ot_fingering(const RpDevModel* dm, int nenv, const int32_t* t_idx, const double* qpos, int nvtot,
                                   const double* site_xpos, const uint8_t* goal, double* reward, double* reward_terms,
                                   double* scratch, cudaStream_t stream);
