// rp_model.h — device-side constant layout of the RoboPianist scene.
//
// The flat mjModel-like arrays (include/roboplay.h) are re-packed on the host into two groups that
// match how the step kernel walks them: 88 single-hinge keys (one tree each, lane-parallel) and up
// to two hands expressed as 26 single-joint links each (lane == link == dof inside a hand).
// A body with several joints (the forearm: two slides) becomes a chain of links, all but the last
// massless, which is kinematically identical to MuJoCo's sequential in-body joints.
#pragma once
#include <stdint.h>

#define RP_NK 88      // keys
#define RP_NL 26      // links (= dofs) per hand
#define RP_NHMAX 2    // hands
#define RP_MAXD 10    // max chain length (depth+1)
#define RP_NGH 20     // colliding primitive geoms per hand (capsules / spheres), max
#define RP_NAH 24     // actuators per hand, max
#define RP_HC_MAXPAIRS 640  // hand geom pairs that may collide with each other (sphere / capsule pairs of both hands: 596)
#define RP_NIGN 32    // hand colliders outside the narrowphase's coverage (per hand), max
#define RP_NSITE 5    // fingertip sites per hand
#ifndef RP_MAXCON
#define RP_MAXCON 21  // contact capacity per environment (shared memory: 584 B each; 3 x 21 rows = one thread per row)
#endif
#define RP_SW_MAXPAIRS 4096  // element updates of one symmetric sweep of a hand matrix (pairs j <= i of every pivot's support)
#define RP_SW_NPAIRS_SMEM 2048  // ... of which the kernel keeps a block-shared copy (Shadow Hand: 1996)
#define RP_NPIANO 89  // piano boxes: index 0 = base, 1 + k = key k (MuJoCo body order)

struct RpGeomParams {  // contact-relevant geom fields
  double solref[2], solimp[5], friction[3], solmix, margin, gap;
  int priority, contype, conaffinity, pad;
};

struct RpDevModel {
  // ---- options ----
  int nhand, nu, nsub, iterations, task_kind, nkey_act, ot_fingering, pad1;
  double h, tol, impratio, solver_scale;  // solver_scale = 1/(meaninertia*max(1,nv))
  double gravity[3];
  // ---- keys ----
  double key_pos[RP_NK][3];     // body origin (world; key bodies hang off the world, identity quat)
  double key_jpos[RP_NK][3];    // hinge anchor, body frame
  double key_axis[RP_NK][3];    // hinge axis (body == world frame)
  double key_ipos[RP_NK][3];    // COM, body frame
  double key_half[RP_NK][3];    // box half sizes (geom at body origin)
  double key_mass[RP_NK], key_M[RP_NK];  // M = dof inertia about the hinge incl. armature (constant)
  double key_stiffness[RP_NK], key_springref[RP_NK], key_damping[RP_NK];
  double key_range[RP_NK][2], key_lim_K[RP_NK], key_lim_B[RP_NK], key_lim_solimp[RP_NK][5];
  double key_dof_invw[RP_NK], key_body_invw[RP_NK];
  double key_act_gain[RP_NK], key_act_range[RP_NK][2];  // self-actuated piano only
  RpGeomParams key_gp[RP_NK];
  // base box (static)
  double base_pos[3], base_half[3];
  RpGeomParams base_gp;
  int has_base, pad2;
  // ---- hands ----
  int link_parent[RP_NHMAX][RP_NL], link_depth[RP_NHMAX][RP_NL], link_subtree[RP_NHMAX][RP_NL];
  int link_anc[RP_NHMAX][RP_NL][RP_MAXD];  // ancestor (or self) at depth d, -1 beyond
  int link_jtype[RP_NHMAX][RP_NL], link_limited[RP_NHMAX][RP_NL];
  int link_nchild_slots[RP_NHMAX][RP_MAXD + 1];  // max #children of any link at depth d
  int link_childrank[RP_NHMAX][RP_NL];           // rank among siblings
  int maxdepth[RP_NHMAX];
  // schedule of the tree-sparse symmetric sweep (M -> -M^-1): pivot k (leaves first) touches the elements (j, i),
  // j <= i, of its support S_k = ancestors(k) + proper subtree(k); entries j * 32 + i, pivot k's run at
  // [sw_off[k], sw_off[k + 1])
  uint16_t sw_pairs[RP_NHMAX][RP_SW_MAXPAIRS];
  int sw_off[RP_NHMAX][RP_NL + 1];
  double link_pos[RP_NHMAX][RP_NL][3], link_quat[RP_NHMAX][RP_NL][4];
  double link_jpos[RP_NHMAX][RP_NL][3], link_jaxis[RP_NHMAX][RP_NL][3];
  double link_ipos[RP_NHMAX][RP_NL][3], link_iquat[RP_NHMAX][RP_NL][4];
  double link_mass[RP_NHMAX][RP_NL], link_inertia[RP_NHMAX][RP_NL][3], link_gravcomp[RP_NHMAX][RP_NL];
  double link_damping[RP_NHMAX][RP_NL], link_armature[RP_NHMAX][RP_NL], link_frictionloss[RP_NHMAX][RP_NL];
  double link_stiffness[RP_NHMAX][RP_NL], link_springref[RP_NHMAX][RP_NL];
  double link_range[RP_NHMAX][RP_NL][2], link_lim_K[RP_NHMAX][RP_NL], link_lim_B[RP_NHMAX][RP_NL];
  double link_lim_solimp[RP_NHMAX][RP_NL][5];
  double link_fr_R[RP_NHMAX][RP_NL], link_fr_B[RP_NHMAX][RP_NL];  // friction-loss row constants
  double link_dof_invw[RP_NHMAX][RP_NL], link_body_invw[RP_NHMAX][RP_NL];
  int dof_act[RP_NHMAX][RP_NL];        // actuator (index within the hand) driving this dof, -1 none
  double dof_act_coef[RP_NHMAX][RP_NL];
  // actuators (per hand, MuJoCo order: XML actuators then forearm slides)
  int nact[RP_NHMAX];
  int act_dof0[RP_NHMAX][RP_NAH], act_dof1[RP_NHMAX][RP_NAH];  // dof1 = -1 for joint transmission
  double act_coef0[RP_NHMAX][RP_NAH], act_coef1[RP_NHMAX][RP_NAH];
  double act_gain[RP_NHMAX][RP_NAH], act_bias[RP_NHMAX][RP_NAH][3];
  double act_ctrlrange[RP_NHMAX][RP_NAH][2], act_forcerange[RP_NHMAX][RP_NAH][2];
  int act_ctrllimited[RP_NHMAX][RP_NAH], act_forcelimited[RP_NHMAX][RP_NAH];
  // colliding hand primitives
  int ngeom[RP_NHMAX];
  int geom_link[RP_NHMAX][RP_NGH], geom_type[RP_NHMAX][RP_NGH], geom_order[RP_NHMAX][RP_NGH];
  double geom_pos[RP_NHMAX][RP_NGH][3], geom_axis[RP_NHMAX][RP_NGH][3];  // capsule axis (local z) in link frame
  double geom_size[RP_NHMAX][RP_NGH][2];
  RpGeomParams geom_gp[RP_NHMAX][RP_NGH];
  double geom_body_invw[RP_NHMAX][RP_NGH];  // body_invweight0 of the geom's own body (differs from its link's when the body is fused into an ancestor)
  int nhandgeom_total, pad3;
  // hand colliders of the reference model the narrowphase does not handle (boxes, cylinders, forearm mesh): oriented
  // bounding boxes, checked every substep against the piano's bounding box -> diag flag 8 when one comes within reach
  int nign[RP_NHMAX];
  int ign_link[RP_NHMAX][RP_NIGN];
  double ign_pos[RP_NHMAX][RP_NIGN][3], ign_mat[RP_NHMAX][RP_NIGN][9], ign_half[RP_NHMAX][RP_NIGN][3];  // oriented bounding box in the link frame
  double piano_lo[3], piano_hi[3];  // world bounding box of base + keys at rest (pressed keys only move down)
  // fingertip sites
  int site_link[RP_NHMAX][RP_NSITE];
  double site_pos[RP_NHMAX][RP_NSITE][3];
  // observation order of hand joints (joints_pos observable)
  int obs_dof[RP_NHMAX][RP_NL];
  int nobs[RP_NHMAX];  // observed (= real) dofs per hand; link slots beyond are inert padding
  // ---- hand : hand collisions (disable_hand_collisions=False): sphere / capsule geom pairs that pass mj_collision's
  // body-pair filters, in MuJoCo's pair order; a | b << 8 with a = hand1 * RP_NGH + geom1 < b = hand2 * RP_NGH + geom2
  int hc_enabled, hc_npair;
  uint16_t hc_pairs[RP_HC_MAXPAIRS];
  int nfore[RP_NHMAX], fore_ign[RP_NHMAX][2];  // the forearm's colliders among ign_* (forearm reward: oriented-box overlap)
  // ---- task ----
  int T, lookahead, wrong_press_termination, disable_fingering_reward, disable_forearm_reward, reward_type;
  int obs_dim, act_dim;
  double energy_coef;
  // white-key lattice for the broadphase
  double lattice_y0, lattice_pitch, lattice_slack;
  // keys come in two geometric classes (white / black): joint anchor, half sizes, hinge axis and margin are shared inside a
  // class (checked on the host: keygeo_ok), only the position differs — the broadphase keeps these in shared memory
  int keygeo_ok, pad4;
  uint8_t key_class[RP_NK];
  double keycls[2][10];   // jpos[3], half[3], axis[3], margin
  int white_of_slot[52];   // key number of the i-th white key
  int black_after[52];     // key number of the black key on the boundary after white i, -1 none
};
