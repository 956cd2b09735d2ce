/* roboplay.h — C-ABI of the B200-native batched physics rollout for RoboPianist.
 *
 * This library replaces, for the `piano_with_shadow_hands` / `self_actuated_piano` tasks, the
 * operator boundary the reference reaches through third-party code: MuJoCo's C API
 *   mj_step / mj_step1 / mj_step2 / mj_forward / mj_resetData
 * as driven by dm_control's `Physics.step()` inside `composer.Environment.step()`
 * (reference call sites: robopianist/suite/__init__.py:87-93 builds the environment with
 * legacy_step=True; robopianist/suite/tasks/base.py:28-31,68-70 fixes 10 substeps of 5 ms per
 * control step; robopianist/suite/tasks/piano_with_shadow_hands.py:176-220 is the per-step task
 * logic that is fused into the same launch).  The reference has no native boundary of its own.
 *
 * Conventions: plain-old-data only; every array pointer in rp_flat_model / rp_task is HOST memory
 * and is copied by rp_model_create (the caller may free it afterwards).  Pointers passed to
 * rp_step / rp_reset / rp_get_state / rp_set_state with the suffix _dev are DEVICE pointers owned
 * by the caller (e.g. torch allocations).  All floating point is IEEE double ("f64"), matching
 * MuJoCo's mjtNum.  Functions return 0 on success, non-zero on error; rp_last_error() gives the
 * message for the calling thread.  Launches are asynchronous on the stream given at batch
 * creation; no hidden synchronisation except in the *_host entry points.
 */
#ifndef ROBOPLAY_H_
#define ROBOPLAY_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RP_NKEY 88
#define RP_GOAL_DIM 89 /* 88 keys + sustain */

/* Flattened scene: the subset of mjModel this scene exercises (names follow mjModel). */
typedef struct rp_flat_model {
  int32_t nbody, nv, ngeom, nsite, ntendon, nwrap, nu, nkey, nhand;
  int32_t nsubstep;   /* physics steps per control step (10) */
  int32_t iterations; /* PGS iteration cap (mjOption.iterations) */
  int32_t reserved0;
  double timestep, impratio, tolerance, meaninertia;
  double gravity[3];
  /* bodies [nbody] */
  const int32_t *body_parentid, *body_rootid, *body_dofadr, *body_dofnum;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia;
  const double *body_gravcomp, *body_invweight0; /* invweight0: [nbody,2] */
  /* joints == dofs [nv] (hinge / slide only) */
  const int32_t *jnt_type, *jnt_bodyid, *jnt_limited, *dof_parentid;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_stiffness, *jnt_springref;
  const double *jnt_solref, *jnt_solimp, *jnt_margin;
  const double *dof_damping, *dof_armature, *dof_frictionloss, *dof_solref, *dof_solimp;
  const double *dof_invweight0;
  /* colliding geoms [ngeom] */
  const int32_t *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity, *geom_condim;
  const int32_t *geom_priority, *geom_key, *geom_hand, *geom_forearm, *geom_collide;
  const double *geom_pos, *geom_quat, *geom_size, *geom_friction, *geom_solref, *geom_solimp;
  const double *geom_solmix, *geom_margin, *geom_gap;
  /* fingertip sites [nsite] (right thumb..little, left thumb..little) */
  const int32_t *site_bodyid;
  const double *site_pos;
  /* fixed tendons */
  const int32_t *tendon_adr, *tendon_num, *wrap_dofid;
  const double *wrap_coef;
  /* actuators [nu] */
  const int32_t *actuator_trntype, *actuator_trnid, *actuator_ctrllimited, *actuator_forcelimited;
  const double *actuator_gear, *actuator_gainprm, *actuator_biasprm; /* biasprm: [nu,3] */
  const double *actuator_ctrlrange, *actuator_forcerange;
  /* task bindings */
  const int32_t *key_dofid, *key_geomid, *key_bodyid; /* [nkey] */
  const double *key_qmax;                              /* [nkey] */
  const int32_t *hand_obs_dofid;                       /* [nhand, nobsdof] joints_pos order */
  int32_t nobsdof, reserved1;
  /* <contact><exclude> body pairs [nexclude, 2] (only consulted when hand geoms may collide with each other) */
  const int32_t *exclude_bodies;
  int32_t nexclude, reserved2;
} rp_flat_model;

/* Task description: song tables + the PianoWithShadowHands / SelfActuatedPiano kwargs that
 * change per-step behaviour (piano_with_shadow_hands.py:50-66, self_actuated_piano.py:73-110). */
typedef struct rp_task {
  int32_t kind;               /* 0 = piano_with_shadow_hands, 1 = self_actuated_piano */
  int32_t n_steps;            /* T = len(NoteTrajectory) */
  int32_t n_steps_lookahead;  /* L */
  int32_t wrong_press_termination;
  int32_t disable_fingering_reward;
  int32_t disable_forearm_reward;
  int32_t reward_type;        /* self-actuated: 0 = negative L2, 1 = negative xent */
  int32_t ot_fingering_reward; /* with disable_fingering_reward: add the optimal-transport fingering term in its
                               * place, as PianoWithShadowHands._set_rewards does (piano_with_shadow_hands.py:130-141,
                               * 333-369); 0 = no fingering term at all */
  double energy_penalty_coef;
  const uint8_t *goal;        /* [T, 89] */
  const int8_t *fingering;    /* [T, 88] finger id 0..9 per goal key, -1 none */
} rp_task;

typedef struct rp_model rp_model; /* opaque, host side */
typedef struct rp_batch rp_batch; /* opaque: device-resident SoA state of N environments */

/* step_type values (dm_env.StepType) */
enum { RP_STEP_FIRST = 0, RP_STEP_MID = 1, RP_STEP_LAST = 2 };

const char *rp_last_error(void);
int rp_version(void);

rp_model *rp_model_create(const rp_flat_model *m, const rp_task *task);
void rp_model_destroy(rp_model *);
int rp_obs_dim(const rp_model *);    /* 329 for the default task, 178 self-actuated */
int rp_action_dim(const rp_model *); /* 45 / 89 */
int rp_reward_terms(const rp_model *);

/* device < 0 is rejected: there is no CPU fallback in this library. */
rp_batch *rp_batch_create(const rp_model *, int nenv, int device, void *cuda_stream);
int rp_batch_destroy(rp_batch *);
int rp_set_option(rp_batch *, int iterations, double tolerance);

/* reset(): replaces composer.Environment.reset + mj_resetData + mj_forward for the envs whose
 * mask byte is non-zero (mask_dev == NULL: all). Writes the FIRST observation. */
int rp_reset(rp_batch *, const uint8_t *mask_dev, double *obs_dev /*[N,obs_dim]*/);

/* step(): one control step = before_step + nsubstep x mj_step + after_step + observation +
 * reward + termination, for every env, in one launch.  action_dev [N,action_dim];
 * outputs: obs [N,obs_dim], reward [N], discount [N], step_type [N] (uint8),
 * reward_terms [N,5] optional (NULL to skip).  Envs whose previous step_type was LAST are
 * reset instead (dm_env semantics: reward 0, discount 1, step_type FIRST). */
int rp_step(rp_batch *, const double *action_dev, double *obs_dev, double *reward_dev,
            double *discount_dev, uint8_t *step_type_dev, double *reward_terms_dev);

/* Same call with HOST buffers: copies actions host->device and results device->host on the
 * batch stream and synchronises.  This is the call timed as "e2e". */
int rp_step_host(rp_batch *, const double *action, double *obs, double *reward, double *discount,
                 uint8_t *step_type, double *reward_terms);

/* Staged form of the same call: the caller writes actions into the pinned input buffer and reads the
 * results from the pinned output buffer returned by rp_host_buffers (layout, in doubles: obs
 * [N,obs_dim], reward [N], discount [N], reward_terms [N,5], then step_type as N bytes).  One H2D
 * copy, one launch, one D2H copy, one stream synchronise; no pageable copies. */
int rp_step_staged(rp_batch *);
int rp_host_buffers(rp_batch *, double **action_pinned, double **out_pinned);

/* Physics-only advance (mj_step x nsub with ctrl as given, no task logic); used by parity tests.
 * ctrl_dev [N,nu]. */
int rp_physics_step(rp_batch *, const double *ctrl_dev, int nsub);

/* State access (HOST buffers, synchronous). Layout [N,nv] row-major / [N]. Any pointer may be
 * NULL.  This is mjData.{qpos,qvel,qacc_warmstart} + the task's _t_idx; doubles as checkpoint. */
int rp_get_state(rp_batch *, double *qpos, double *qvel, double *qacc_warmstart, int32_t *t_idx);
int rp_set_state(rp_batch *, const double *qpos, const double *qvel, const double *qacc_warmstart,
                 const int32_t *t_idx);
/* Checkpoint / resume (the reference has none: its episodes always start from qpos0).  One host blob
 * with everything a later step depends on: qpos, qvel, qacc_warmstart, the controls and the sustain state of the last
 * action, t_idx and the episode state, the previous key activation, the musical-metric accumulators, the launch order, and
 * — when enabled — the hand offsets with their generator states and the MIDI event stream.  A batch of the same model,
 * size and optional features restored from it continues bit for bit.  Song tables (rp_set_song) are not part of it. */
size_t rp_checkpoint_bytes(rp_batch *);
int rp_checkpoint_save(rp_batch *, void *host_buf, size_t bytes);
int rp_checkpoint_load(rp_batch *, const void *host_buf, size_t bytes);
/* Diagnostics of the last substep (HOST buffers): ncon [N], niter [N] (PGS sweeps used),
 * flags [N]: bit0 contact overflow (more than RP_MAXCON contact records: a contact across the two hands takes two, every other
 * contact one; with hand collisions enabled the deepest records are kept), bit1 non-finite state, bit2 hand tree too deep,
 * bit3 a hand collider the narrowphase does not cover (box / cylinder / mesh; rp_flat_model.geom_collide == 0) came
 * within reach of the piano during the last step: its contact, if any, was not generated. */
int rp_get_diag(rp_batch *, int32_t *ncon, int32_t *niter, int32_t *flags);
/* Extra sensor-like outputs of the last step (HOST): actuator_force [N,nu], fingertip site
 * positions [N,nsite,3], key activation [N,88] uint8. */
int rp_get_aux(rp_batch *, double *actuator_force, double *site_xpos, uint8_t *activation);

/* Musical metrics of robopianist/wrappers/evaluation.py:40-177 (MidiEvaluationWrapper), accumulated on
 * the device per environment (HOST buffer, synchronous).  out [N,16]: [0..2] running sums over the steps
 * of the current episode of the per-step key-press precision / recall / F1 (sklearn
 * precision_recall_fscore_support, average="binary", zero_division=1, activation vs. the goal keys of
 * that step), [3..5] the same for the sustain pedal, [6] steps in the sums, [7] completed episodes,
 * [8..13] the six means of the last completed episode, [14..15] reserved. */
int rp_get_metrics(rp_batch *, double *out);

/* Device-side MIDI event stream: replaces robopianist/models/piano/midi_module.py:47-98
 * (MidiModule.after_substep, called per physics substep from Piano.after_substep, piano.py:154-162).
 * rp_midi_enable allocates `capacity` event words per environment; from then on every rp_step records,
 * per substep, the keys whose activation changed and pedal changes.  Event word (uint32):
 * (tick << 9) | (class << 7) | key, tick = substeps since the episode start (message time = tick *
 * timestep), class 0 NoteOn (velocity 127 as in the reference), 1 NoteOff, 2 SustainOn, 3 SustainOff,
 * key = 0..87 (MIDI note = key + 21).  Sorting one environment's words ascending yields the reference's
 * message order.  The list restarts when the environment starts a new episode.
 * rp_get_midi (HOST buffers, synchronous): events [N,capacity] (unsorted), counts [N] (> capacity means
 * events were dropped). */
int rp_midi_enable(rp_batch *, int capacity);
int rp_get_midi(rp_batch *, uint32_t *events, int32_t *counts);

/* Replaces the song of a batch (goal table uint8 [n_steps, 89], fingering table int8 [n_steps, 88], HOST buffers;
 * synchronous): what PianoWithShadowHands._maybe_change_midi + _reset_trajectory do at an episode start when the task has
 * augmentations (piano_with_shadow_hands.py:151-165).  Lookahead, reward switches and observation width stay those of
 * the model.  Call rp_reset afterwards: running episodes are not re-indexed. */
int rp_set_song(rp_batch *, const uint8_t *goal, const int8_t *fingering, int n_steps);

/* randomize_hand_positions (robopianist/suite/tasks/piano_with_shadow_hands.py:64,491-499; the shift itself is
 * dm_control's Entity.shift_pose on the hands' attachment frames): from the next reset on, every episode start shifts
 * both hands along y by one draw from U(-0.05, 0.05) m per environment; like the reference (model not recompiled
 * between episodes, suite/__init__.py:56) the shifts accumulate.  One splitmix64 stream per environment seeded from
 * (seed, environment index).  enable = 0 switches it off and clears the offsets; enable = 2 draws every episode's shift
 * from the unshifted pose (suite.load(recompile_physics=True)).  rp_get_hand_offsets: current
 * offsets [N] into a host buffer. */
int rp_randomize_hand_positions(rp_batch *, int enable, uint64_t seed);
int rp_get_hand_offsets(rp_batch *, double *offsets_host);

/* Launch shape of the step kernel (results do not depend on it): environments per thread block, 1 or
 * RP_BLOCK_ENVS (blocks whose environments run each substep in step with each other, which shares their
 * instruction fetches); 0 = chosen from the batch size (default). */
#define RP_BLOCK_ENVS 3
int rp_set_block_envs(rp_batch *, int block_envs);

int64_t rp_launch_count(const rp_batch *); /* kernels launched by this batch so far */

#ifdef __cplusplus
}
#endif
#endif /* ROBOPLAY_H_ */
