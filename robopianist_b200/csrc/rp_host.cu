// rp_host.cu — C-ABI (include/roboplay.h) over the step kernel: model packing, device state,
// launches.  No CPU fallback: every entry point that computes needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/roboplay.h"
#include "rp_kernels.h"
#include "rp_task.cuh"
#include "rp_math.cuh"

static thread_local std::string g_err;
static int fail(const std::string& msg) { g_err = msg; return 1; }
#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

struct rp_model {
  RpDevModel dm;
  std::vector<uint8_t> goal;
  std::vector<int8_t> fingering;
  int nv, nu, nhand, obs_dim, act_dim;
  int nl;      // dofs per hand of the flat model (<= RP_NL; the device layout pads every hand to RP_NL dofs)
  int nv_dev;  // row length of the device state arrays: RP_NK + nhand * RP_NL
};

struct rp_batch {
  const rp_model* model;
  int nenv, device;
  cudaStream_t stream;
  RpDevModel* d_dm;
  uint8_t* d_goal;
  int8_t* d_fingering;
  double *qpos, *qvel, *qacc_ws, *ctrl, *sustain, *act_force, *site_xpos, *dbg;
  int32_t *t_idx, *step_type_state, *diag, *work, *order;
  int order_valid;  // order[] holds a permutation (set after the first full step)
  uint8_t* activation;
  double* metrics;  // [N][RP_METRICS_STRIDE] musical metrics accumulators (see rp_get_metrics)
  uint32_t* midi_events;  // [N][midi_cap], allocated by rp_midi_enable
  double* ot_scratch;     // [N][RP_OT_SCRATCH_DOUBLES], only for tasks with the optimal-transport fingering reward
  int32_t* midi_count;
  int midi_cap;
  double* hand_offset;    // [N] accumulated attachment offset of the hands (rp_randomize_hand_positions)
  uint64_t* rng_state;    // [N] generator state of the offset draws
  int hand_offset_fresh;  // offsets do not accumulate over episodes (the reference with recompile_physics=True)
  int record_cap;         // test hook: contact-record capacity below RP_MAXCON (rp_debug_set_record_cap)
  // staging for rp_step_host
  double *h_action, *h_out;  // pinned
  double *d_action, *d_out;
  int iterations;
  double tolerance;
  int block_envs;  // environments per thread block (0: chosen by batch size)
  int64_t launches;
};

extern "C" const char* rp_last_error(void) { return g_err.c_str(); }
extern "C" int rp_version(void) { return 1; }

static void quat_to_mat(const double* q, double* m) { rp::quat2mat(m, q); }

static void copy_gp(RpGeomParams& gp, const rp_flat_model* m, int g) {
  for (int k = 0; k < 2; k++) gp.solref[k] = m->geom_solref[2 * g + k];
  for (int k = 0; k < 5; k++) gp.solimp[k] = m->geom_solimp[5 * g + k];
  for (int k = 0; k < 3; k++) gp.friction[k] = m->geom_friction[3 * g + k];
  gp.solmix = m->geom_solmix[g]; gp.margin = m->geom_margin[g]; gp.gap = m->geom_gap[g];
  gp.priority = m->geom_priority[g]; gp.contype = m->geom_contype[g]; gp.conaffinity = m->geom_conaffinity[g];
  gp.pad = 0;
}

// ---- host-side helpers of the hand packer -------------------------------------------------------------------------
static void quat_mul_h(const double* a, const double* b, double* o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
struct Pose { double p[3], q[4]; };
// pose of body b in the frame of its ancestor body a (a = 0: world); every body strictly between must be jointless
static bool pose_in_ancestor(const rp_flat_model* m, int b, int a, Pose* out) {
  Pose t;
  for (int c = 0; c < 3; c++) t.p[c] = m->body_pos[3 * b + c];
  for (int c = 0; c < 4; c++) t.q[c] = m->body_quat[4 * b + c];
  int p = m->body_parentid[b];
  while (p != a) {
    if (p <= 0 || m->body_dofnum[p] != 0) return false;
    double R[9], r[3], q2[4];
    rp::quat2mat(R, m->body_quat + 4 * p);
    rp::mat_vec(r, R, t.p);
    for (int c = 0; c < 3; c++) t.p[c] = m->body_pos[3 * p + c] + r[c];
    quat_mul_h(m->body_quat + 4 * p, t.q, q2);
    for (int c = 0; c < 4; c++) t.q[c] = q2[c];
    p = m->body_parentid[p];
  }
  *out = t;
  return true;
}
// nearest ancestor-or-self with joints (0: none below the world), and the pose of b in that body's frame
static int jointed_ancestor(const rp_flat_model* m, int b, Pose* rel) {
  int a = b;
  while (a > 0 && m->body_dofnum[a] == 0) a = m->body_parentid[a];
  if (a == b) { rel->p[0] = rel->p[1] = rel->p[2] = 0; rel->q[0] = 1; rel->q[1] = rel->q[2] = rel->q[3] = 0; return a; }
  if (!pose_in_ancestor(m, b, a, rel)) return -1;
  return a;
}
// eigen-decomposition of a symmetric 3x3 matrix (cyclic Jacobi): A = V diag(w) V^T, V a proper rotation
static void sym_eig3(const double* Ain, double* w, double* V) {
  double A[9];
  for (int i = 0; i < 9; i++) { A[i] = Ain[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = fabs(A[1]) + fabs(A[2]) + fabs(A[5]);
    if (off < 1e-300) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (fabs(A[3 * p + q]) < 1e-300) continue;
        double theta = (A[3 * q + q] - A[3 * p + p]) / (2 * A[3 * p + q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        double c = 1 / sqrt(t * t + 1), sn = t * c;
        for (int k = 0; k < 3; k++) {  // A <- A J
          double akp = A[3 * k + p], akq = A[3 * k + q];
          A[3 * k + p] = c * akp - sn * akq; A[3 * k + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {  // A <- J^T A
          double apk = A[3 * p + k], aqk = A[3 * q + k];
          A[3 * p + k] = c * apk - sn * aqk; A[3 * q + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V[3 * k + p], vkq = V[3 * k + q];
          V[3 * k + p] = c * vkp - sn * vkq; V[3 * k + q] = sn * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; i++) w[i] = A[4 * i];
  double det = V[0] * (V[4] * V[8] - V[5] * V[7]) - V[1] * (V[3] * V[8] - V[5] * V[6]) + V[2] * (V[3] * V[7] - V[4] * V[6]);
  if (det < 0) for (int k = 0; k < 3; k++) V[3 * k + 2] = -V[3 * k + 2];
}
static void mat_to_quat(const double* R, double* q) {  // proper rotation, row major
  double tr = R[0] + R[4] + R[8];
  if (tr > 0) { double s = sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int k = 0; k < 4; k++) q[k] /= n;
}
// One link's inertial from several rigidly attached bodies, each given in the link frame (mass, com, frame quaternion,
// principal moments): total mass, centre of mass, principal axes of the summed tensor about it.
struct Inertial { double mass, pos[3], quat[4], diag[3]; };
static Inertial merge_inertials(const std::vector<Inertial>& parts) {
  Inertial o;
  memset(&o, 0, sizeof(o));
  o.quat[0] = 1;
  for (const Inertial& p : parts) { o.mass += p.mass; for (int c = 0; c < 3; c++) o.pos[c] += p.mass * p.pos[c]; }
  if (o.mass <= 0) { for (int c = 0; c < 3; c++) o.pos[c] = 0; return o; }
  for (int c = 0; c < 3; c++) o.pos[c] /= o.mass;
  double I[9] = {0};
  for (const Inertial& p : parts) {
    double R[9];
    rp::quat2mat(R, p.quat);
    double d[3] = {p.pos[0] - o.pos[0], p.pos[1] - o.pos[1], p.pos[2] - o.pos[2]}, d2 = rp::dot3(d, d);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double v = 0;
        for (int k = 0; k < 3; k++) v += R[3 * i + k] * p.diag[k] * R[3 * j + k];
        I[3 * i + j] += v + p.mass * ((i == j ? d2 : 0.0) - d[i] * d[j]);
      }
  }
  double V[9];
  sym_eig3(I, o.diag, V);
  mat_to_quat(V, o.quat);
  return o;
}

extern "C" rp_model* rp_model_create(const rp_flat_model* m, const rp_task* task) {
  if (!m || !task) { fail("null model/task"); return nullptr; }
  if (m->nkey != RP_NK) { fail("model must have 88 keys"); return nullptr; }
  if (m->nhand != 0 && m->nhand != 2) { fail("nhand must be 0 or 2"); return nullptr; }
  const int nl = m->nhand ? m->nobsdof : 0;  // every hand dof is listed once in hand_obs_dofid (joints_pos)
  if (nl < 0 || nl > RP_NL) { fail("a hand may have at most 26 dofs"); return nullptr; }
  if (m->nv != RP_NK + m->nhand * nl) { fail("unexpected nv (need 88 + nhand * dofs per hand)"); return nullptr; }
  if (!task->goal || !task->fingering) { fail("task goal / fingering tables must not be null"); return nullptr; }
  if (m->nexclude < 0 || (m->nexclude > 0 && !m->exclude_bodies)) { fail("exclude_bodies must list nexclude body pairs"); return nullptr; }
  if (task->n_steps < 1) { fail("task n_steps must be >= 1"); return nullptr; }
  if (task->n_steps_lookahead < 0 || task->n_steps_lookahead > 1024) { fail("task n_steps_lookahead out of range [0, 1024]"); return nullptr; }
  if (task->kind != 0 && task->kind != 1) { fail("task kind must be 0 (piano_with_shadow_hands) or 1 (self_actuated_piano)"); return nullptr; }
  rp_model* M = new rp_model();
  RpDevModel& d = M->dm;
  memset(&d, 0, sizeof(d));
  d.nhand = m->nhand; d.nu = m->nu; d.nsub = m->nsubstep; d.iterations = m->iterations;
  d.task_kind = task->kind;
  d.h = m->timestep; d.tol = m->tolerance; d.impratio = m->impratio;
  d.solver_scale = 1.0 / (m->meaninertia * (m->nv > 1 ? m->nv : 1));
  for (int k = 0; k < 3; k++) d.gravity[k] = m->gravity[k];
  auto bad = [&](const char* msg) -> rp_model* { fail(msg); delete M; return nullptr; };
  // ---------------- keys
  for (int k = 0; k < RP_NK; k++) {
    int j = m->key_dofid[k], b = m->key_bodyid[k], g = m->key_geomid[k];
    if (j != k) return bad("key dofs must be dofs 0..87 in key order");
    if (m->body_parentid[b] != 0 || m->jnt_bodyid[j] != b || m->jnt_type[j] != 3) return bad("key body layout");
    const double* bq = m->body_quat + 4 * b;
    if (fabs(bq[0] - 1) > 1e-12) return bad("key bodies must have identity orientation");
    const double* gq = m->geom_quat + 4 * g;
    if (fabs(gq[0] - 1) > 1e-12 || fabs(m->geom_pos[3 * g]) + fabs(m->geom_pos[3 * g + 1]) + fabs(m->geom_pos[3 * g + 2]) > 1e-12)
      return bad("key geoms must sit at the body origin");
    const double* iq = m->body_iquat + 4 * b;
    if (fabs(iq[0] - 1) > 1e-12) return bad("key inertial frame must be axis aligned");
    for (int c = 0; c < 3; c++) {
      d.key_pos[k][c] = m->body_pos[3 * b + c];
      d.key_jpos[k][c] = m->jnt_pos[3 * j + c];
      d.key_axis[k][c] = m->jnt_axis[3 * j + c];
      d.key_ipos[k][c] = m->body_ipos[3 * b + c];
      d.key_half[k][c] = m->geom_size[3 * g + c];
    }
    d.key_mass[k] = m->body_mass[b];
    // inertia about the hinge: axis' I axis + m |r_perp|^2 + armature
    const double *ax = d.key_axis[k], *I = m->body_inertia + 3 * b;
    double r[3] = {d.key_ipos[k][0] - d.key_jpos[k][0], d.key_ipos[k][1] - d.key_jpos[k][1], d.key_ipos[k][2] - d.key_jpos[k][2]};
    double ra = rp::dot3(r, ax), r2 = rp::dot3(r, r) - ra * ra;
    d.key_M[k] = ax[0] * ax[0] * I[0] + ax[1] * ax[1] * I[1] + ax[2] * ax[2] * I[2] + d.key_mass[k] * r2 + m->dof_armature[j];
    d.key_stiffness[k] = m->jnt_stiffness[j]; d.key_springref[k] = m->jnt_springref[j];
    d.key_damping[k] = m->dof_damping[j];
    if (!m->jnt_limited[j]) return bad("key joints must be limited");
    if (m->dof_frictionloss[j] != 0) return bad("key joints must have no frictionloss");
    d.key_range[k][0] = m->jnt_range[2 * j]; d.key_range[k][1] = m->jnt_range[2 * j + 1];
    rp::ref_KB(m->jnt_solref + 2 * j, m->jnt_solimp[5 * j + 1], m->timestep, &d.key_lim_K[k], &d.key_lim_B[k]);
    for (int c = 0; c < 5; c++) d.key_lim_solimp[k][c] = m->jnt_solimp[5 * j + c];
    d.key_dof_invw[k] = m->dof_invweight0[j];
    d.key_body_invw[k] = m->body_invweight0[2 * b];
    copy_gp(d.key_gp[k], m, g);
  }
  for (int g = 0; g < m->ngeom; g++)
    if (m->geom_key[g] == -2) {
      int b = m->geom_bodyid[g];
      if (m->body_dofnum[b] != 0 || m->body_parentid[b] != 0) return bad("piano base must be static");
      for (int c = 0; c < 3; c++) { d.base_pos[c] = m->body_pos[3 * b + c] + m->geom_pos[3 * g + c]; d.base_half[c] = m->geom_size[3 * g + c]; }
      copy_gp(d.base_gp, m, g);
      d.has_base = 1;
    }
  // white-key lattice
  {
    int nw = 0;
    double ys[RP_NK];
    for (int k = 0; k < RP_NK; k++) ys[k] = d.key_pos[k][1];
    for (int i = 0; i < 52; i++) d.black_after[i] = -1;
    for (int k = 0; k < RP_NK; k++) {
      bool white = d.key_half[k][1] > 0.008;
      if (white) { if (nw >= 52) return bad("more than 52 white keys"); d.white_of_slot[nw++] = k; }
    }
    if (nw != 52) return bad("expected 52 white keys");
    d.lattice_y0 = ys[d.white_of_slot[0]];
    d.lattice_pitch = ys[d.white_of_slot[1]] - ys[d.white_of_slot[0]];
    for (int k = 0; k < RP_NK; k++) {
      bool white = d.key_half[k][1] > 0.008;
      if (white) continue;
      int slot = (int)floor((ys[k] - d.lattice_y0) / d.lattice_pitch);
      if (slot < 0 || slot >= 52 || d.black_after[slot] >= 0) return bad("black key lattice");
      d.black_after[slot] = k;
    }
    // Broadphase slack along y.  With every hinge parallel to y a key's y extent does not change with its angle; white
    // boxes inflated by half the pitch tile the y axis, and a black key's extent lies between the centres of its two
    // white neighbours, so "white slots whose inflated extent overlaps the geom + the blacks next to them" is complete.
    // (Any other hinge direction: the conservative 2 cm of the first version.)
    bool ypar = true;
    double maxmargin = d.has_base ? d.base_gp.margin : 0.0;
    for (int k = 0; k < RP_NK; k++) {
      ypar = ypar && fabs(d.key_axis[k][0]) < 1e-12 && fabs(d.key_axis[k][2]) < 1e-12;
      maxmargin = fmax(maxmargin, d.key_gp[k].margin);
    }
    for (int s2 = 0; s2 < 52 && ypar; s2++) {
      int b = d.black_after[s2];
      if (b < 0) continue;
      double lo = ys[d.white_of_slot[s2]], hi = s2 + 1 < 52 ? ys[d.white_of_slot[s2 + 1]] : lo + d.lattice_pitch;
      if (ys[b] - d.key_half[b][1] < lo || ys[b] + d.key_half[b][1] > hi) ypar = false;
    }
    double hand_margin = 0;
    for (int g = 0; g < m->ngeom; g++) if (m->geom_hand[g] >= 0) hand_margin = fmax(hand_margin, m->geom_margin[g]);
    // geometric classes for the broadphase tables
    d.keygeo_ok = 1;
    bool have[2] = {false, false};
    for (int k = 0; k < RP_NK; k++) {
      const int c = d.key_half[k][1] > 0.008 ? 0 : 1;
      d.key_class[k] = (uint8_t)c;
      double v[10] = {d.key_jpos[k][0], d.key_jpos[k][1], d.key_jpos[k][2], d.key_half[k][0], d.key_half[k][1], d.key_half[k][2],
                      d.key_axis[k][0], d.key_axis[k][1], d.key_axis[k][2], d.key_gp[k].margin};
      if (!have[c]) { memcpy(d.keycls[c], v, sizeof(v)); have[c] = true; }
      else if (memcmp(d.keycls[c], v, sizeof(v)) != 0) d.keygeo_ok = 0;
    }
    d.lattice_slack = ypar ? 0.5 * d.lattice_pitch + fmax(maxmargin, hand_margin) + 1e-9 : 0.02 + fmax(maxmargin, hand_margin);
  }
  // ---------------- actuators on keys (self-actuated piano)
  int nkey_act = 0;
  for (int a = 0; a < m->nu; a++) {
    if (m->actuator_trntype[a] == 0 && m->actuator_trnid[a] < RP_NK) {
      int k = m->actuator_trnid[a];
      if (a != k) return bad("key actuators must be in key order");
      d.key_act_gain[k] = m->actuator_gainprm[a] * m->actuator_gear[a];
      d.key_act_range[k][0] = m->actuator_ctrllimited[a] ? m->actuator_ctrlrange[2 * a] : -1e300;
      d.key_act_range[k][1] = m->actuator_ctrllimited[a] ? m->actuator_ctrlrange[2 * a + 1] : 1e300;
      nkey_act++;
    }
  }
  if (nkey_act != 0 && nkey_act != RP_NK) return bad("either all or no keys may be actuated");
  d.nkey_act = nkey_act;
  // ---------------- hands
  // Every hand occupies RP_NL link slots on the device (lane == link == dof).  A hand of the flat model may have fewer
  // dofs (reduced action space, fewer forearm dofs: shadow_hand.py:162-182,272-311): the unused slots become inert
  // one-dof dummies (unit inertia, no forces, never observed).  Bodies without joints are fused into the link of
  // their nearest jointed ancestor — geoms, sites and inertia re-expressed in that body's frame; MuJoCo keeps them as
  // separate bodies, which is the same rigid body up to rounding.
  int nact_hand = m->nhand ? (m->nu - nkey_act) / m->nhand : 0;
  int order = 0;
  d.nobs[0] = d.nobs[1] = nl;
  int geom_flat[RP_NHMAX][RP_NGH];  // flat-model geom of the hand's n-th colliding primitive
  for (int h = 0; h < m->nhand; h++) {
    int base = RP_NK + h * nl;
    int maxdepth = 0;
    std::vector<std::vector<Inertial>> parts(RP_NL);
    std::vector<double> link_gc(RP_NL, -1.0);
    for (int l = 0; l < nl; l++) {
      int j = base + l, b = m->jnt_bodyid[j];
      bool first = m->body_dofadr[b] == j;
      int parent;
      if (first) {
        int pj = m->dof_parentid[j];
        parent = pj < 0 ? -1 : pj - base;
        Pose rel;
        if (!pose_in_ancestor(m, b, pj < 0 ? 0 : m->jnt_bodyid[pj], &rel)) return bad("hand body tree: a jointed body must descend from the body of its parent dof through jointless bodies only");
        for (int c = 0; c < 3; c++) d.link_pos[h][l][c] = rel.p[c];
        for (int c = 0; c < 4; c++) d.link_quat[h][l][c] = rel.q[c];
      } else {
        parent = l - 1;
        d.link_quat[h][l][0] = 1;
      }
      if (parent >= l || parent < -1) return bad("hand dofs must be in depth-first order");
      d.link_parent[h][l] = parent;
      d.link_depth[h][l] = parent < 0 ? 0 : d.link_depth[h][parent] + 1;
      if (d.link_depth[h][l] >= RP_MAXD) return bad("hand chain too deep");
      maxdepth = d.link_depth[h][l] > maxdepth ? d.link_depth[h][l] : maxdepth;
      for (int dd = 0; dd < RP_MAXD; dd++) d.link_anc[h][l][dd] = -1;
      if (parent >= 0) for (int dd = 0; dd < d.link_depth[h][l]; dd++) d.link_anc[h][l][dd] = d.link_anc[h][parent][dd];
      d.link_anc[h][l][d.link_depth[h][l]] = l;
      d.link_jtype[h][l] = m->jnt_type[j];
      d.link_limited[h][l] = m->jnt_limited[j];
      for (int c = 0; c < 3; c++) { d.link_jpos[h][l][c] = m->jnt_pos[3 * j + c]; d.link_jaxis[h][l][c] = m->jnt_axis[3 * j + c]; }
      d.link_iquat[h][l][0] = 1;
      d.link_body_invw[h][l] = m->body_invweight0[2 * b];
      d.link_damping[h][l] = m->dof_damping[j]; d.link_armature[h][l] = m->dof_armature[j];
      d.link_frictionloss[h][l] = m->dof_frictionloss[j];
      d.link_stiffness[h][l] = m->jnt_stiffness[j]; d.link_springref[h][l] = m->jnt_springref[j];
      d.link_range[h][l][0] = m->jnt_range[2 * j]; d.link_range[h][l][1] = m->jnt_range[2 * j + 1];
      rp::ref_KB(m->jnt_solref + 2 * j, m->jnt_solimp[5 * j + 1], m->timestep, &d.link_lim_K[h][l], &d.link_lim_B[h][l]);
      for (int c = 0; c < 5; c++) d.link_lim_solimp[h][l][c] = m->jnt_solimp[5 * j + c];
      d.link_dof_invw[h][l] = m->dof_invweight0[j];
      double imp = rp::impedance(m->dof_solimp + 5 * j, 0.0, 0.0), K;
      d.link_fr_R[h][l] = fmax(RP_MINVAL, (1 - imp) * m->dof_invweight0[j] / imp);
      rp::ref_KB(m->dof_solref + 2 * j, m->dof_solimp[5 * j + 1], m->timestep, &K, &d.link_fr_B[h][l]);
      d.dof_act[h][l] = -1;
    }
    for (int l = nl; l < RP_NL; l++) {  // inert padding links
      d.link_parent[h][l] = -1; d.link_depth[h][l] = 0;
      for (int dd = 0; dd < RP_MAXD; dd++) d.link_anc[h][l][dd] = -1;
      d.link_anc[h][l][0] = l;
      d.link_jtype[h][l] = 3; d.link_jaxis[h][l][2] = 1;
      d.link_quat[h][l][0] = 1; d.link_iquat[h][l][0] = 1;
      d.link_armature[h][l] = 1.0; d.link_dof_invw[h][l] = 1.0; d.link_body_invw[h][l] = 1.0; d.link_fr_R[h][l] = 1.0;
      d.dof_act[h][l] = -1;
    }
    // link of every hand body (-1: rigidly attached to the world) and its pose in that link's body frame
    auto body_link = [&](int b, Pose* rel) -> int {
      int a = jointed_ancestor(m, b, rel);
      if (a <= 0) return -1;
      return m->body_dofadr[a] + m->body_dofnum[a] - 1 - base;
    };
    // inertials: the link's own body plus the jointless bodies hanging off it
    for (int b = 1; b < m->nbody; b++) {
      int jb = b;
      while (jb > 0 && m->body_dofnum[jb] == 0) jb = m->body_parentid[jb];
      if (jb <= 0 || m->body_dofadr[jb] < base || m->body_dofadr[jb] >= base + nl) continue;  // not a (moving) body of this hand
      Pose rel;
      int l = body_link(b, &rel);
      if (l < 0 || l >= nl) return bad("hand body outside its hand's dof range");
      Inertial in;
      in.mass = m->body_mass[b];
      double R[9], r[3];
      rp::quat2mat(R, rel.q);
      rp::mat_vec(r, R, m->body_ipos + 3 * b);
      for (int c = 0; c < 3; c++) { in.pos[c] = rel.p[c] + r[c]; in.diag[c] = m->body_inertia[3 * b + c]; }
      quat_mul_h(rel.q, m->body_iquat + 4 * b, in.quat);
      if (b == jb) {  // the jointed body itself: kept verbatim when nothing is fused into it
        for (int c = 0; c < 3; c++) in.pos[c] = m->body_ipos[3 * b + c];
        for (int c = 0; c < 4; c++) in.quat[c] = m->body_iquat[4 * b + c];
        parts[l].insert(parts[l].begin(), in);
      } else if (in.mass > 0) {
        parts[l].push_back(in);
      }
      if (link_gc[l] >= 0 && link_gc[l] != m->body_gravcomp[b] && in.mass > 0) return bad("bodies fused into one link must share their gravcomp");
      if (in.mass > 0 || b == jb) link_gc[l] = m->body_gravcomp[b];
    }
    for (int l = 0; l < nl; l++) {
      if (parts[l].empty()) continue;  // all but the last dof of a multi-dof body: massless
      Inertial in = parts[l].size() == 1 ? parts[l][0] : merge_inertials(parts[l]);
      for (int c = 0; c < 3; c++) { d.link_ipos[h][l][c] = in.pos[c]; d.link_inertia[h][l][c] = in.diag[c]; }
      for (int c = 0; c < 4; c++) d.link_iquat[h][l][c] = in.quat[c];
      d.link_mass[h][l] = in.mass;
      d.link_gravcomp[h][l] = link_gc[l] < 0 ? 0.0 : link_gc[l];
    }
    d.maxdepth[h] = maxdepth;
    // subtree sizes, sibling ranks
    for (int l = RP_NL - 1; l >= 0; l--) {
      d.link_subtree[h][l] += 1;
      if (d.link_parent[h][l] >= 0) d.link_subtree[h][d.link_parent[h][l]] += d.link_subtree[h][l];
    }
    {  // sweep schedule: pivots are taken leaves first (k = NL-1 .. 0); runs are stored in pivot order k = 0 .. NL-1
      int np = 0;
      for (int k = 0; k < RP_NL; k++) {
        d.sw_off[h][k] = np;
        int S[RP_NL], n = 0;
        for (int dd = 0; dd < d.link_depth[h][k]; dd++) S[n++] = d.link_anc[h][k][dd];
        for (int i = k + 1; i < k + d.link_subtree[h][k]; i++) S[n++] = i;
        for (int a = 0; a < n; a++)
          for (int b2 = a; b2 < n; b2++) {
            if (np >= RP_SW_MAXPAIRS) return bad("sweep schedule too long");
            d.sw_pairs[h][np++] = (uint16_t)(S[a] * 32 + S[b2]);
          }
      }
      d.sw_off[h][RP_NL] = np;
      if (np > RP_SW_NPAIRS_SMEM) return bad("sweep schedule exceeds the kernel's shared-memory copy");
      if (h > 0 && (memcmp(d.sw_off[h], d.sw_off[0], sizeof(d.sw_off[0])) || memcmp(d.sw_pairs[h], d.sw_pairs[0], sizeof(uint16_t) * np)))
        return bad("the hands must have the same kinematic tree (one sweep schedule is shared)");
    }
    {
      int nchild[RP_NL] = {0};
      for (int l = 0; l < RP_NL; l++) {
        int p = d.link_parent[h][l];
        if (p >= 0) {
          d.link_childrank[h][l] = nchild[p]++;
          int dd = d.link_depth[h][l];
          if (nchild[p] > d.link_nchild_slots[h][dd]) d.link_nchild_slots[h][dd] = nchild[p];
        }
      }
    }
    // actuators of this hand
    d.nact[h] = nact_hand;
    if (nact_hand > RP_NAH) return bad("too many actuators per hand");
    for (int ai = 0; ai < nact_hand; ai++) {
      int a = nkey_act + h * nact_hand + ai;
      double gear = m->actuator_gear[a];
      int d0, d1 = -1;
      double c0 = gear, c1 = 0;
      if (m->actuator_trntype[a] == 0) d0 = m->actuator_trnid[a] - base;
      else {
        int t = m->actuator_trnid[a];
        if (m->tendon_num[t] != 2) return bad("only two-joint fixed tendons are supported");
        d0 = m->wrap_dofid[m->tendon_adr[t]] - base; d1 = m->wrap_dofid[m->tendon_adr[t] + 1] - base;
        c0 = gear * m->wrap_coef[m->tendon_adr[t]]; c1 = gear * m->wrap_coef[m->tendon_adr[t] + 1];
        if (d1 < 0 || d1 >= nl) return bad("actuator does not belong to its hand");
      }
      if (d0 < 0 || d0 >= nl) return bad("actuator does not belong to its hand (hand actuators must be contiguous)");
      d.act_dof0[h][ai] = d0; d.act_dof1[h][ai] = d1; d.act_coef0[h][ai] = c0; d.act_coef1[h][ai] = c1;
      d.act_gain[h][ai] = m->actuator_gainprm[a];
      for (int c = 0; c < 3; c++) d.act_bias[h][ai][c] = m->actuator_biasprm[3 * a + c];
      d.act_ctrllimited[h][ai] = m->actuator_ctrllimited[a]; d.act_forcelimited[h][ai] = m->actuator_forcelimited[a];
      for (int c = 0; c < 2; c++) { d.act_ctrlrange[h][ai][c] = m->actuator_ctrlrange[2 * a + c]; d.act_forcerange[h][ai][c] = m->actuator_forcerange[2 * a + c]; }
      if (d.dof_act[h][d0] >= 0 || (d1 >= 0 && d.dof_act[h][d1] >= 0)) return bad("more than one actuator on a dof");
      d.dof_act[h][d0] = ai; d.dof_act_coef[h][d0] = c0;
      if (d1 >= 0) { d.dof_act[h][d1] = ai; d.dof_act_coef[h][d1] = c1; }
    }
    // colliding primitives
    int ng = 0;
    for (int g = 0; g < m->ngeom; g++) {
      if (m->geom_hand[g] != h || !m->geom_collide[g]) continue;
      if (ng >= RP_NGH) return bad("too many colliding hand geoms");
      int b = m->geom_bodyid[g];
      Pose rel;
      int l = body_link(b, &rel);
      if (l < 0 || l >= nl) return bad("a colliding hand geom sits on a body that is welded to the world");
      if (m->geom_type[g] != 2 && m->geom_type[g] != 3) return bad("only sphere/capsule hand colliders are supported");
      d.geom_link[h][ng] = l; d.geom_type[h][ng] = m->geom_type[g]; d.geom_order[h][ng] = order++;
      geom_flat[h][ng] = g;
      double R[9], Rb[9], r[3], ax[3];
      quat_to_mat(m->geom_quat + 4 * g, R);
      rp::quat2mat(Rb, rel.q);
      rp::mat_vec(r, Rb, m->geom_pos + 3 * g);
      double axl[3] = {R[2], R[5], R[8]};
      rp::mat_vec(ax, Rb, axl);
      for (int c = 0; c < 3; c++) { d.geom_pos[h][ng][c] = rel.p[c] + r[c]; d.geom_axis[h][ng][c] = ax[c]; }
      d.geom_size[h][ng][0] = m->geom_size[3 * g]; d.geom_size[h][ng][1] = m->geom_size[3 * g + 1];
      d.geom_body_invw[h][ng] = m->body_invweight0[2 * b];  // of the geom's own body (a fused body keeps its own value)
      copy_gp(d.geom_gp[h][ng], m, g);
      ng++;
    }
    d.ngeom[h] = ng;
    // colliders the narrowphase does not cover: oriented bounding boxes for the runtime proximity flag
    int ni = 0;
    for (int g = 0; g < m->ngeom; g++) {
      if (m->geom_hand[g] != h || m->geom_collide[g] || !(m->geom_contype[g] | m->geom_conaffinity[g])) continue;
      int b = m->geom_bodyid[g];
      Pose rel;
      int l = body_link(b, &rel);
      if (l < 0) continue;  // welded to the world (a forearm without dofs): its clearance never changes
      if (ni >= RP_NIGN) return bad("too many uncovered hand colliders");
      d.ign_link[h][ni] = l;
      if (m->geom_forearm[g]) {
        if (d.nfore[h] >= 2) return bad("more than two forearm colliders");
        d.fore_ign[h][d.nfore[h]++] = ni;
      }
      double Rb[9], Rg[9], r[3];
      rp::quat2mat(Rb, rel.q);
      rp::mat_vec(r, Rb, m->geom_pos + 3 * g);
      for (int c = 0; c < 3; c++) d.ign_pos[h][ni][c] = rel.p[c] + r[c];
      const double* sz = m->geom_size + 3 * g;
      quat_to_mat(m->geom_quat + 4 * g, Rg);
      for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) d.ign_mat[h][ni][3 * i + k] = Rb[3 * i] * Rg[k] + Rb[3 * i + 1] * Rg[3 + k] + Rb[3 * i + 2] * Rg[6 + k];
      if (m->geom_type[g] == 5) { d.ign_half[h][ni][0] = d.ign_half[h][ni][1] = sz[0]; d.ign_half[h][ni][2] = sz[1]; }  // cylinder: radius, half height
      else for (int c = 0; c < 3; c++) d.ign_half[h][ni][c] = sz[c];                                                    // box / mesh bounding box
      ni++;
    }
    d.nign[h] = ni;
    if (m->nsite != 2 * RP_NSITE) return bad("expected 5 fingertip sites per hand");
    for (int s = 0; s < RP_NSITE; s++) {
      int si = h * RP_NSITE + s, b = m->site_bodyid[si];
      Pose rel;
      int l = body_link(b, &rel);
      if (l < 0 || l >= nl) return bad("fingertip site on a body welded to the world");
      d.site_link[h][s] = l;
      double Rb[9], r[3];
      rp::quat2mat(Rb, rel.q);
      rp::mat_vec(r, Rb, m->site_pos + 3 * si);
      for (int c = 0; c < 3; c++) d.site_pos[h][s][c] = rel.p[c] + r[c];
    }
    for (int l = 0; l < nl; l++) {
      int od = m->hand_obs_dofid[h * nl + l] - base;
      if (od < 0 || od >= nl) return bad("hand_obs_dofid outside the hand's dof range");
      d.obs_dof[h][l] = od;
    }
  }
  // hand geom : hand geom pairs (self and cross hand), mj_collision's filters: contype / conaffinity, same body or same
  // weld body (a jointless body is welded to its parent), parent-child weld bodies unless one is welded to the world,
  // <exclude> pairs.
  if (m->nhand == 2) {
    std::vector<int> weld(m->nbody, 0);
    for (int b = 1; b < m->nbody; b++) weld[b] = m->body_dofnum[b] > 0 ? b : weld[m->body_parentid[b]];
    int np = 0;
    for (int a1 = 0; a1 < 2 * RP_NGH; a1++)
      for (int a2 = a1 + 1; a2 < 2 * RP_NGH; a2++) {
        int h1 = a1 / RP_NGH, i1 = a1 % RP_NGH, h2 = a2 / RP_NGH, i2 = a2 % RP_NGH;
        if (i1 >= d.ngeom[h1] || i2 >= d.ngeom[h2]) continue;
        int g1 = geom_flat[h1][i1], g2 = geom_flat[h2][i2];
        if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
        int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2], w1 = weld[b1], w2 = weld[b2];
        if (b1 == b2 || w1 == w2) continue;
        int wp1 = weld[m->body_parentid[w1]], wp2 = weld[m->body_parentid[w2]];
        if (w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
        bool excl = false;
        for (int e = 0; e < m->nexclude; e++) {
          int e1 = m->exclude_bodies[2 * e], e2 = m->exclude_bodies[2 * e + 1];
          excl = excl || (e1 == b1 && e2 == b2) || (e1 == b2 && e2 == b1);
        }
        if (excl) continue;
        if (g1 > g2 || b1 > b2) return bad("hand geoms must be numbered in body order");
        if (np >= RP_HC_MAXPAIRS) return bad("too many hand geom pairs");
        d.hc_pairs[np++] = (uint16_t)(a1 | (a2 << 8));
      }
    d.hc_npair = np;
    d.hc_enabled = np > 0;
  }
  d.nhandgeom_total = order;
  if (order > 64) return bad("too many hand geoms for the contact sort key");
  for (int c = 0; c < 3; c++) { d.piano_lo[c] = 1e30; d.piano_hi[c] = -1e30; }
  for (int k = 0; k < RP_NK; k++)
    for (int c = 0; c < 3; c++) {
      d.piano_lo[c] = fmin(d.piano_lo[c], d.key_pos[k][c] - d.key_half[k][c]);
      d.piano_hi[c] = fmax(d.piano_hi[c], d.key_pos[k][c] + d.key_half[k][c]);
    }
  if (d.has_base)
    for (int c = 0; c < 3; c++) {
      d.piano_lo[c] = fmin(d.piano_lo[c], d.base_pos[c] - d.base_half[c]);
      d.piano_hi[c] = fmax(d.piano_hi[c], d.base_pos[c] + d.base_half[c]);
    }
  // ---------------- task
  d.T = task->n_steps; d.lookahead = task->n_steps_lookahead;
  d.wrong_press_termination = task->wrong_press_termination;
  d.disable_fingering_reward = task->disable_fingering_reward;
  d.disable_forearm_reward = task->disable_forearm_reward;
  d.ot_fingering = task->disable_fingering_reward && task->ot_fingering_reward && task->kind == 0;
  d.reward_type = task->reward_type;
  d.energy_coef = task->energy_penalty_coef;
  if (task->kind == 0) {
    if (m->nhand != 2) return bad("piano_with_shadow_hands needs two hands");
    d.obs_dim = RP_NK + 1 + 2 * nl + (d.lookahead + 1) * RP_GOAL_DIM + (d.disable_fingering_reward ? 0 : 10);
  } else {
    if (nkey_act != RP_NK) return bad("self_actuated_piano needs key actuators");
    d.obs_dim = RP_NK + 1 + (d.lookahead + 1) * RP_GOAL_DIM;
  }
  d.act_dim = m->nu + 1;
  if (d.T <= 0) return bad("empty song");
  M->goal.assign(task->goal, task->goal + (size_t)d.T * RP_GOAL_DIM);
  M->fingering.assign(task->fingering, task->fingering + (size_t)d.T * RP_NK);
  M->nv = m->nv; M->nu = m->nu; M->nhand = m->nhand; M->obs_dim = d.obs_dim; M->act_dim = d.act_dim;
  M->nl = nl; M->nv_dev = RP_NK + m->nhand * RP_NL;
  return M;
}

extern "C" void rp_model_destroy(rp_model* m) { delete m; }
extern "C" int rp_obs_dim(const rp_model* m) { return m->obs_dim; }
extern "C" int rp_action_dim(const rp_model* m) { return m->act_dim; }
extern "C" int rp_reward_terms(const rp_model* m) { return m->dm.task_kind == 0 ? 5 : 1; }

// rp_debug_fail_alloc(k): the k-th device allocation from now on fails (tests of the failure path); -1 = off
static int g_fail_alloc_after = -1;
extern "C" int rp_debug_fail_alloc(int k) { g_fail_alloc_after = k; return 0; }

template <typename T>
static cudaError_t dalloc(T** p, size_t n) {
  if (g_fail_alloc_after >= 0 && g_fail_alloc_after-- == 0) return cudaErrorMemoryAllocation;  // test hook (rp_debug_fail_alloc)
  cudaError_t e = cudaMalloc((void**)p, sizeof(T) * (n ? n : 1));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, sizeof(T) * (n ? n : 1));
  return e;
}

extern "C" rp_batch* rp_batch_create(const rp_model* model, int nenv, int device, void* cuda_stream) {
  if (!model || nenv <= 0) { fail("bad arguments"); return nullptr; }
  if (device < 0) { fail("device < 0: this library has no CPU path"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) { fail("CUDA device not available"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { fail("cudaSetDevice failed"); return nullptr; }
  rp_batch* b = new rp_batch();
  memset(b, 0, sizeof(*b));
  b->model = model; b->nenv = nenv; b->device = device; b->stream = (cudaStream_t)cuda_stream;
  b->iterations = -1; b->tolerance = -1.0;
  size_t N = nenv, nv = model->nv_dev, nu = model->nu;
  cudaError_t e = cudaSuccess;
  // every allocation is skipped once one has failed (a batch too large for the device must not go on to ask for
  // hundreds of gigabytes of pinned host memory)
#define chk(x) do { if (e == cudaSuccess) e = (x); } while (0)
  chk(cudaMalloc((void**)&b->d_dm, sizeof(RpDevModel)));
  chk(cudaMemcpy(b->d_dm, &model->dm, sizeof(RpDevModel), cudaMemcpyHostToDevice));
  chk(cudaMalloc((void**)&b->d_goal, model->goal.size()));
  chk(cudaMemcpy(b->d_goal, model->goal.data(), model->goal.size(), cudaMemcpyHostToDevice));
  chk(cudaMalloc((void**)&b->d_fingering, model->fingering.size()));
  chk(cudaMemcpy(b->d_fingering, model->fingering.data(), model->fingering.size(), cudaMemcpyHostToDevice));
  chk(dalloc(&b->qpos, N * nv)); chk(dalloc(&b->qvel, N * nv)); chk(dalloc(&b->qacc_ws, N * nv));
  chk(dalloc(&b->ctrl, N * nu)); chk(dalloc(&b->sustain, N)); chk(dalloc(&b->act_force, N * nu));
  chk(dalloc(&b->site_xpos, N * 30)); chk(dalloc(&b->t_idx, N)); chk(dalloc(&b->step_type_state, N));
  chk(dalloc(&b->diag, N * 3)); chk(dalloc(&b->activation, N * RP_NK));
  chk(dalloc(&b->metrics, N * RP_METRICS_STRIDE));
  chk(dalloc(&b->work, N)); chk(dalloc(&b->order, N));
  if (model->dm.ot_fingering) chk(dalloc(&b->ot_scratch, N * RP_OT_SCRATCH_DOUBLES));
  size_t out_doubles = N * (model->obs_dim + 2 + 5) + (N + 7) / 8 + 8;
  chk(dalloc(&b->d_action, N * model->act_dim));
  chk(dalloc(&b->d_out, out_doubles));
  chk(cudaMallocHost((void**)&b->h_action, sizeof(double) * N * model->act_dim));
  chk(cudaMallocHost((void**)&b->h_out, sizeof(double) * out_doubles));
#undef chk
  if (e != cudaSuccess) {  // release whatever was allocated before the failure
    fail(std::string("allocation failed: ") + cudaGetErrorString(e));
    rp_batch_destroy(b);
    return nullptr;
  }
  // step_type_state starts as MID so the first rp_step without rp_reset still steps
  std::vector<int32_t> mid(N, RP_STEP_MID_);
  cudaMemcpy(b->step_type_state, mid.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice);
  return b;
}

extern "C" int rp_batch_destroy(rp_batch* b) {
  if (!b) return 0;
  cudaSetDevice(b->device);
  cudaFree(b->d_dm); cudaFree(b->d_goal); cudaFree(b->d_fingering);
  cudaFree(b->qpos); cudaFree(b->qvel); cudaFree(b->qacc_ws); cudaFree(b->ctrl); cudaFree(b->sustain);
  cudaFree(b->act_force); cudaFree(b->site_xpos); cudaFree(b->t_idx); cudaFree(b->step_type_state);
  cudaFree(b->diag); cudaFree(b->activation); cudaFree(b->metrics); cudaFree(b->work); cudaFree(b->order); cudaFree(b->dbg); cudaFree(b->d_action); cudaFree(b->d_out);
  cudaFree(b->midi_events); cudaFree(b->midi_count); cudaFree(b->ot_scratch);
  cudaFree(b->hand_offset); cudaFree(b->rng_state);
  cudaFreeHost(b->h_action); cudaFreeHost(b->h_out);
  delete b;
  return 0;
}

extern "C" int rp_set_song(rp_batch* b, const uint8_t* goal, const int8_t* fingering, int n_steps) {
  if (!b || !goal || !fingering || n_steps < 1) return fail("rp_set_song: bad arguments");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  uint8_t* ng = nullptr; int8_t* nf = nullptr;
  const size_t gb = (size_t)n_steps * RP_GOAL_DIM, fb = (size_t)n_steps * RP_NK;
  if (cudaMalloc((void**)&ng, gb) != cudaSuccess || cudaMalloc((void**)&nf, fb) != cudaSuccess) {
    cudaFree(ng); cudaFree(nf);
    return fail("rp_set_song: allocation failed");
  }
  CUDA_OK(cudaMemcpy(ng, goal, gb, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(nf, fingering, fb, cudaMemcpyHostToDevice));
  // episode length lives in the batch's device copy of the model constants
  CUDA_OK(cudaMemcpy((char*)b->d_dm + offsetof(RpDevModel, T), &n_steps, sizeof(int), cudaMemcpyHostToDevice));
  cudaFree(b->d_goal); cudaFree(b->d_fingering);
  b->d_goal = ng; b->d_fingering = nf;
  return 0;
}

extern "C" int rp_set_block_envs(rp_batch* b, int block_envs) {
  if (block_envs != 0 && block_envs != 1 && block_envs != RP_EPB) return fail("block_envs must be 0 (auto), 1 or RP_BLOCK_ENVS");
  b->block_envs = block_envs;
  return 0;
}

extern "C" int rp_set_option(rp_batch* b, int iterations, double tolerance) {
  b->iterations = iterations; b->tolerance = tolerance;
  return 0;
}

static RpKernelArgs base_args(rp_batch* b, int mode) {
  RpKernelArgs a;
  memset(&a, 0, sizeof(a));
  a.dm = b->d_dm; a.nenv = b->nenv; a.mode = mode; a.nsub = b->model->dm.nsub;
  a.iterations = b->iterations; a.tolerance = b->tolerance;
  a.qpos = b->qpos; a.qvel = b->qvel; a.qacc_ws = b->qacc_ws; a.ctrl = b->ctrl;
  a.t_idx = b->t_idx; a.step_type_state = b->step_type_state; a.sustain = b->sustain;
  a.goal = b->d_goal; a.fingering = b->d_fingering;
  a.diag = b->diag; a.act_force = b->act_force; a.site_xpos = b->site_xpos; a.activation = b->activation; a.metrics = b->metrics;
  a.midi_events = b->midi_events; a.midi_count = b->midi_count; a.midi_cap = b->midi_cap;
  a.work = b->work; a.order = b->order_valid ? b->order : nullptr;
  a.hand_offset = b->hand_offset; a.rng_state = b->rng_state; a.hand_offset_fresh = b->hand_offset_fresh;
  a.record_cap = b->record_cap;
  a.dbg = b->dbg;
  return a;
}
static int launch(rp_batch* b, const RpKernelArgs& a) {
  CUDA_OK(cudaSetDevice(b->device));
  static const int force_hc = getenv("ROBOPLAY_FORCE_HC") != nullptr;  // measurement hook: the hand-collision instantiation on a model without pairs
  CUDA_OK(rp_launch_step(a, b->model->nhand, b->stream, b->block_envs, b->model->dm.hc_enabled || (force_hc && b->model->nhand == 2)));
  b->launches++;
  return 0;
}

extern "C" int rp_reset(rp_batch* b, const uint8_t* mask_dev, double* obs_dev) {
  RpKernelArgs a = base_args(b, RP_MODE_RESET);
  a.reset_mask = mask_dev; a.obs = obs_dev;
  return launch(b, a);
}

extern "C" int rp_step(rp_batch* b, const double* action_dev, double* obs_dev, double* reward_dev,
                       double* discount_dev, uint8_t* step_type_dev, double* reward_terms_dev) {
  if (!action_dev) return fail("action is required");
  RpKernelArgs a = base_args(b, RP_MODE_STEP);
  a.action = action_dev; a.obs = obs_dev; a.reward = reward_dev; a.discount = discount_dev;
  a.step_type = step_type_dev; a.reward_terms = reward_terms_dev;
  if (launch(b, a)) return 1;
  if (b->model->dm.ot_fingering && (reward_dev || reward_terms_dev)) {
    CUDA_OK(rp_launch_ot_fingering(b->d_dm, b->nenv, b->t_idx, b->qpos, b->model->nv_dev, b->site_xpos, b->d_goal, reward_dev,
                                   reward_terms_dev, b->ot_scratch, b->stream));
    b->launches++;
  }
  // more environments than resident slots: sort the next launch longest-expected-first
  if (b->nenv >= 2048 && b->model->nhand > 0) {
    CUDA_OK(rp_launch_order(b->work, b->order, b->nenv, b->stream));
    b->order_valid = 1;
    b->launches++;
  }
  return 0;
}

// Staged form of the host call: actions are written by the caller into the pinned input buffer and
// results are left in the pinned output buffer (rp_host_buffers), so a step costs one H2D copy, one
// launch, one D2H copy and a stream synchronise — no pageable memcpy, no allocation.
extern "C" int rp_step_staged(rp_batch* b) {
  const rp_model* M = b->model;
  size_t N = b->nenv, od = M->obs_dim, ad = M->act_dim;
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaMemcpyAsync(b->d_action, b->h_action, sizeof(double) * N * ad, cudaMemcpyHostToDevice, b->stream));
  double* d_obs = b->d_out;
  double* d_rew = d_obs + N * od;
  double* d_dis = d_rew + N;
  double* d_terms = d_dis + N;
  uint8_t* d_st = (uint8_t*)(d_terms + N * 5);
  if (rp_step(b, b->d_action, d_obs, d_rew, d_dis, d_st, d_terms)) return 1;
  size_t bytes = sizeof(double) * (N * od + 2 * N + 5 * N) + N;
  CUDA_OK(cudaMemcpyAsync(b->h_out, b->d_out, bytes, cudaMemcpyDeviceToHost, b->stream));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  return 0;
}

// Pinned staging buffers of rp_step_staged / rp_step_host.  out layout (doubles): obs [N,obs_dim],
// reward [N], discount [N], reward_terms [N,5], then step_type as N bytes.
extern "C" int rp_host_buffers(rp_batch* b, double** action_pinned, double** out_pinned) {
  if (action_pinned) *action_pinned = b->h_action;
  if (out_pinned) *out_pinned = b->h_out;
  return 0;
}

extern "C" int rp_step_host(rp_batch* b, const double* action, double* obs, double* reward, double* discount,
                            uint8_t* step_type, double* reward_terms) {
  const rp_model* M = b->model;
  size_t N = b->nenv, od = M->obs_dim, ad = M->act_dim;
  if (action != b->h_action) memcpy(b->h_action, action, sizeof(double) * N * ad);
  if (rp_step_staged(b)) return 1;
  if (obs) memcpy(obs, b->h_out, sizeof(double) * N * od);
  if (reward) memcpy(reward, b->h_out + N * od, sizeof(double) * N);
  if (discount) memcpy(discount, b->h_out + N * od + N, sizeof(double) * N);
  if (reward_terms) memcpy(reward_terms, b->h_out + N * od + 2 * N, sizeof(double) * N * 5);
  if (step_type) memcpy(step_type, (uint8_t*)(b->h_out + N * od + 7 * N), N);
  return 0;
}

extern "C" int rp_physics_step(rp_batch* b, const double* ctrl_dev, int nsub) {
  RpKernelArgs a = base_args(b, RP_MODE_PHYSICS);
  a.ctrl_in = ctrl_dev; a.nsub = nsub;
  return launch(b, a);
}

// Rows of the device state arrays are [keys 88 | hand 0: 26 | hand 1: 26]; the flat model's rows are
// [keys 88 | hand 0: nl | hand 1: nl].  Equal when nl == 26; otherwise three strided copies per array.
static cudaError_t copy_state(const rp_model* M, size_t N, double* dst, const double* src, bool to_host) {
  const size_t nv = M->nv, nvd = M->nv_dev, B = sizeof(double);
  const cudaMemcpyKind kind = to_host ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice;
  if (nv == nvd) return cudaMemcpy(dst, src, B * N * nv, kind);
  const size_t dpitch = (to_host ? nv : nvd) * B, spitch = (to_host ? nvd : nv) * B;
  cudaError_t e = cudaMemcpy2D(dst, dpitch, src, spitch, B * RP_NK, N, kind);
  for (int h = 0; h < M->nhand && e == cudaSuccess; h++) {
    const size_t ho = RP_NK + (size_t)h * M->nl, hd = RP_NK + (size_t)h * RP_NL;
    e = cudaMemcpy2D(dst + (to_host ? ho : hd), dpitch, src + (to_host ? hd : ho), spitch, B * M->nl, N, kind);
  }
  return e;
}
extern "C" int rp_get_state(rp_batch* b, double* qpos, double* qvel, double* qacc_ws, int32_t* t_idx) {
  size_t N = b->nenv;
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  if (qpos) CUDA_OK(copy_state(b->model, N, qpos, b->qpos, true));
  if (qvel) CUDA_OK(copy_state(b->model, N, qvel, b->qvel, true));
  if (qacc_ws) CUDA_OK(copy_state(b->model, N, qacc_ws, b->qacc_ws, true));
  if (t_idx) CUDA_OK(cudaMemcpy(t_idx, b->t_idx, sizeof(int32_t) * N, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int rp_set_state(rp_batch* b, const double* qpos, const double* qvel, const double* qacc_ws, const int32_t* t_idx) {
  size_t N = b->nenv;
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  if (qpos) CUDA_OK(copy_state(b->model, N, b->qpos, qpos, false));
  if (qvel) CUDA_OK(copy_state(b->model, N, b->qvel, qvel, false));
  if (qacc_ws) CUDA_OK(copy_state(b->model, N, b->qacc_ws, qacc_ws, false));
  if (t_idx) CUDA_OK(cudaMemcpy(b->t_idx, t_idx, sizeof(int32_t) * N, cudaMemcpyHostToDevice));
  return 0;
}
// ---- checkpoint: every piece of per-environment state a later step depends on, as one host blob (header + arrays in the
// device layout).  A batch restored from it continues bit for bit (tests/test_gpu_parity.py).
struct CkptHeader { uint32_t magic, version; int32_t nenv, nv_dev, nu, midi_cap, has_offsets, reserved; };
struct CkptPart { void* dev; size_t bytes; };
static std::vector<CkptPart> ckpt_parts(rp_batch* b) {
  const size_t N = b->nenv, nv = b->model->nv_dev, nu = b->model->nu;
  std::vector<CkptPart> p = {
      {b->qpos, N * nv * sizeof(double)}, {b->qvel, N * nv * sizeof(double)}, {b->qacc_ws, N * nv * sizeof(double)},
      {b->ctrl, N * nu * sizeof(double)}, {b->sustain, N * sizeof(double)}, {b->t_idx, N * sizeof(int32_t)},
      {b->step_type_state, N * sizeof(int32_t)}, {b->activation, N * RP_NK}, {b->metrics, N * RP_METRICS_STRIDE * sizeof(double)},
      {b->work, N * sizeof(int32_t)}, {b->order, N * sizeof(int32_t)}};
  if (b->hand_offset) { p.push_back({b->hand_offset, N * sizeof(double)}); p.push_back({b->rng_state, N * sizeof(uint64_t)}); }
  if (b->midi_events) { p.push_back({b->midi_events, N * (size_t)b->midi_cap * sizeof(uint32_t)}); p.push_back({b->midi_count, N * sizeof(int32_t)}); }
  return p;
}
extern "C" size_t rp_checkpoint_bytes(rp_batch* b) {
  size_t n = sizeof(CkptHeader) + sizeof(int32_t);  // + order_valid
  for (const CkptPart& q : ckpt_parts(b)) n += q.bytes;
  return n;
}
extern "C" int rp_checkpoint_save(rp_batch* b, void* host_buf, size_t bytes) {
  if (!host_buf || bytes < rp_checkpoint_bytes(b)) return fail("checkpoint buffer too small (rp_checkpoint_bytes)");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  CkptHeader h = {0x52504b43u, 1u, b->nenv, b->model->nv_dev, b->model->nu, b->midi_events ? b->midi_cap : 0, b->hand_offset ? 1 : 0, 0};
  char* o = (char*)host_buf;
  memcpy(o, &h, sizeof(h)); o += sizeof(h);
  int32_t ov = b->order_valid;
  memcpy(o, &ov, sizeof(ov)); o += sizeof(ov);
  for (const CkptPart& q : ckpt_parts(b)) { CUDA_OK(cudaMemcpy(o, q.dev, q.bytes, cudaMemcpyDeviceToHost)); o += q.bytes; }
  return 0;
}
extern "C" int rp_checkpoint_load(rp_batch* b, const void* host_buf, size_t bytes) {
  if (!host_buf || bytes < sizeof(CkptHeader) + sizeof(int32_t)) return fail("checkpoint blob too small");
  CkptHeader h;
  memcpy(&h, host_buf, sizeof(h));
  if (h.magic != 0x52504b43u || h.version != 1u) return fail("not a roboplay checkpoint (magic / version)");
  if (h.nenv != b->nenv || h.nv_dev != b->model->nv_dev || h.nu != b->model->nu) return fail("checkpoint of a different batch shape or model");
  if ((h.midi_cap != 0) != (b->midi_events != nullptr) || (h.midi_cap && h.midi_cap != b->midi_cap) || (h.has_offsets != 0) != (b->hand_offset != nullptr))
    return fail("checkpoint and batch differ in their optional state (rp_midi_enable / rp_randomize_hand_positions before loading)");
  if (bytes < rp_checkpoint_bytes(b)) return fail("checkpoint blob truncated");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  const char* o = (const char*)host_buf + sizeof(h);
  int32_t ov;
  memcpy(&ov, o, sizeof(ov)); o += sizeof(ov);
  for (const CkptPart& q : ckpt_parts(b)) { CUDA_OK(cudaMemcpy(q.dev, o, q.bytes, cudaMemcpyHostToDevice)); o += q.bytes; }
  b->order_valid = ov;
  return 0;
}

extern "C" int rp_get_diag(rp_batch* b, int32_t* ncon, int32_t* niter, int32_t* flags) {
  size_t N = b->nenv;
  std::vector<int32_t> tmp(N * 3);
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  CUDA_OK(cudaMemcpy(tmp.data(), b->diag, sizeof(int32_t) * N * 3, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < N; i++) {
    if (ncon) ncon[i] = tmp[3 * i];
    if (niter) niter[i] = tmp[3 * i + 1];
    if (flags) flags[i] = tmp[3 * i + 2];
  }
  return 0;
}
extern "C" int rp_get_aux(rp_batch* b, double* actuator_force, double* site_xpos, uint8_t* activation) {
  size_t N = b->nenv;
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  if (actuator_force) CUDA_OK(cudaMemcpy(actuator_force, b->act_force, sizeof(double) * N * b->model->nu, cudaMemcpyDeviceToHost));
  if (site_xpos) CUDA_OK(cudaMemcpy(site_xpos, b->site_xpos, sizeof(double) * N * 30, cudaMemcpyDeviceToHost));
  if (activation) CUDA_OK(cudaMemcpy(activation, b->activation, N * RP_NK, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int rp_get_metrics(rp_batch* b, double* out) {
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  CUDA_OK(cudaMemcpy(out, b->metrics, sizeof(double) * (size_t)b->nenv * RP_METRICS_STRIDE, cudaMemcpyDeviceToHost));
  return 0;
}

// randomize_hand_positions (piano_with_shadow_hands.py:64,491-499): from the next reset on, every episode start shifts
// both hands along y by a draw from U(-0.05, 0.05) m, per environment; the shifts accumulate like the reference's
// (its model is not recompiled between episodes).  The draws come from one splitmix64 stream per environment, seeded
// from `seed` and the environment index — the reference's single numpy RandomState has no batched equivalent.
// enable = 0 turns the feature off and clears the offsets; enable = 2: every episode starts from the unshifted pose (what
// the reference does when asked to recompile its physics every episode).
extern "C" int rp_randomize_hand_positions(rp_batch* b, int enable, uint64_t seed) {
  if (b->model->nhand != 2) return fail("randomize_hand_positions needs the hands task");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  cudaFree(b->hand_offset); cudaFree(b->rng_state);
  b->hand_offset = nullptr; b->rng_state = nullptr;
  b->hand_offset_fresh = enable == 2;
  if (!enable) return 0;
  std::vector<uint64_t> st((size_t)b->nenv);
  for (size_t i = 0; i < st.size(); i++) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    st[i] = z;
  }
  CUDA_OK(dalloc(&b->hand_offset, (size_t)b->nenv));
  CUDA_OK(dalloc(&b->rng_state, (size_t)b->nenv));
  CUDA_OK(cudaMemcpy(b->rng_state, st.data(), sizeof(uint64_t) * st.size(), cudaMemcpyHostToDevice));
  return 0;
}
// current offsets [N] (host buffer); zeros when the feature is off
extern "C" int rp_get_hand_offsets(rp_batch* b, double* out) {
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  if (!b->hand_offset) { memset(out, 0, sizeof(double) * (size_t)b->nenv); return 0; }
  CUDA_OK(cudaMemcpy(out, b->hand_offset, sizeof(double) * (size_t)b->nenv, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int rp_midi_enable(rp_batch* b, int capacity) {
  if (capacity <= 0) return fail("capacity must be positive");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  cudaFree(b->midi_events); cudaFree(b->midi_count);
  b->midi_events = nullptr; b->midi_count = nullptr; b->midi_cap = 0;
  CUDA_OK(dalloc(&b->midi_events, (size_t)b->nenv * capacity));
  CUDA_OK(dalloc(&b->midi_count, (size_t)b->nenv));
  b->midi_cap = capacity;
  return 0;
}
extern "C" int rp_get_midi(rp_batch* b, uint32_t* events, int32_t* counts) {
  if (!b->midi_events) return fail("rp_midi_enable was not called");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  if (events) CUDA_OK(cudaMemcpy(events, b->midi_events, sizeof(uint32_t) * (size_t)b->nenv * b->midi_cap, cudaMemcpyDeviceToHost));
  if (counts) CUDA_OK(cudaMemcpy(counts, b->midi_count, sizeof(int32_t) * (size_t)b->nenv, cudaMemcpyDeviceToHost));
  return 0;
}

// Debug hook (not part of the public header): the per-environment body of rp_ot_fingering_kernel run on the
// host for one environment — lets the CPU tests check the wrapper's indexing without a GPU.
extern "C" double rp_debug_ot_fingering_host(const rp_model* m, int t, const double* qpos, const double* site_xpos) {
  std::vector<double> scratch(RP_OT_SCRATCH_DOUBLES);
  return rp::ot_fingering_env(&m->dm, t, qpos, site_xpos, m->goal.data(), scratch.data());
}

extern "C" int64_t rp_launch_count(const rp_batch* b) { return b->launches; }

// Debug hooks (not part of the public header): stage dump of substep 0 for every env.
extern "C" int rp_debug_set_record_cap(rp_batch* b, int cap) { b->record_cap = cap; return 0; }
// Test hooks (host only): the packed per-link and per-geom tables of one hand, so that the packer — bodies without joints
// fused into their parents' links, merged inertials, padded link slots — can be checked without a GPU.
// links: [RP_NL][28] = parent, depth, jtype, pos 3, quat 4, jpos 3, jaxis 3, mass, ipos 3, iquat 4, inertia 3, armature
// geoms: [RP_NGH][9] = link, type, pos 3, axis 3, radius  (rows beyond the hand's geom count are zero); returns the geom count
extern "C" int rp_debug_link_table(const rp_model* m, int hand, double* out) {
  const RpDevModel& d = m->dm;
  if (hand < 0 || hand >= d.nhand) return -1;
  for (int l = 0; l < RP_NL; l++) {
    double* o = out + 28 * l;
    int k = 0;
    o[k++] = d.link_parent[hand][l]; o[k++] = d.link_depth[hand][l]; o[k++] = d.link_jtype[hand][l];
    for (int c = 0; c < 3; c++) o[k++] = d.link_pos[hand][l][c];
    for (int c = 0; c < 4; c++) o[k++] = d.link_quat[hand][l][c];
    for (int c = 0; c < 3; c++) o[k++] = d.link_jpos[hand][l][c];
    for (int c = 0; c < 3; c++) o[k++] = d.link_jaxis[hand][l][c];
    o[k++] = d.link_mass[hand][l];
    for (int c = 0; c < 3; c++) o[k++] = d.link_ipos[hand][l][c];
    for (int c = 0; c < 4; c++) o[k++] = d.link_iquat[hand][l][c];
    for (int c = 0; c < 3; c++) o[k++] = d.link_inertia[hand][l][c];
    o[k++] = d.link_armature[hand][l];
  }
  return d.nobs[hand];
}
// hand geom pairs that may collide, in launch order: out [npair][2] = hand * RP_NGH + index of the hand's colliding primitive; returns npair
extern "C" int rp_debug_pair_list(const rp_model* m, int32_t* out) {
  const RpDevModel& d = m->dm;
  for (int p = 0; p < d.hc_npair; p++) { out[2 * p] = d.hc_pairs[p] & 0xff; out[2 * p + 1] = d.hc_pairs[p] >> 8; }
  return d.hc_enabled ? d.hc_npair : 0;
}
extern "C" int rp_debug_geom_table(const rp_model* m, int hand, double* out) {
  const RpDevModel& d = m->dm;
  if (hand < 0 || hand >= d.nhand) return -1;
  memset(out, 0, sizeof(double) * RP_NGH * 9);
  for (int g = 0; g < d.ngeom[hand]; g++) {
    double* o = out + 9 * g;
    o[0] = d.geom_link[hand][g]; o[1] = d.geom_type[hand][g];
    for (int c = 0; c < 3; c++) { o[2 + c] = d.geom_pos[hand][g][c]; o[5 + c] = d.geom_axis[hand][g][c]; }
    o[8] = d.geom_size[hand][g][0];
  }
  return d.ngeom[hand];
}
extern "C" int rp_debug_enable(rp_batch* b, int on) {
  CUDA_OK(cudaSetDevice(b->device));
  if (on && !b->dbg) CUDA_OK(dalloc(&b->dbg, (size_t)b->nenv * RP_DBG_STRIDE));
  if (!on && b->dbg) { cudaFree(b->dbg); b->dbg = nullptr; }
  return 0;
}
extern "C" int rp_debug_read(rp_batch* b, double* out) {
  if (!b->dbg) return fail("debug dump not enabled");
  CUDA_OK(cudaSetDevice(b->device));
  CUDA_OK(cudaStreamSynchronize(b->stream));
  CUDA_OK(cudaMemcpy(out, b->dbg, sizeof(double) * (size_t)b->nenv * RP_DBG_STRIDE, cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int rp_debug_stride(void) { return RP_DBG_STRIDE; }
extern "C" int rp_smem_bytes(void) { return (int)rp_kernel_smem_bytes(); }
