/* oracle.c — CPU restatement (fp64, one environment per orc_data) of the MuJoCo `mj_step`
 * pipeline for the feature subset the RoboPianist scene exercises.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in robopianist_b200/ may import, link or call this file;
 * it is the checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference legs).
 *
 * PARITY UNPINNED: the reference's physics lives in the third-party `mujoco` wheel
 * (setup.py:39 `mujoco>=3.1.1`, not vendored, not installable here) and the reference's own
 * tests pin no numeric physics result.  This file restates MuJoCo's
 * published algorithm ("Computation" chapter + XML reference defaults) stage by stage; the
 * reference call sites it stands in for are
 *   robopianist/suite/__init__.py:87-93        (Environment(..., legacy_step=True))
 *   robopianist/suite/tasks/base.py:28-31,68-70 (0.005 s physics / 0.05 s control)
 *   robopianist/models/piano/piano.py:154-192  (after_substep key state)
 * Stage names below are MuJoCo's (mj_kinematics, mj_comPos, mj_crb, mj_factorM, mj_collision,
 * mj_makeConstraint, mj_projectConstraint, mj_fwdVelocity, mj_fwdActuation, mj_fwdAcceleration,
 * mj_fwdConstraint/mj_solPGS, mj_Euler).  Spatial quantities are expressed at the world origin
 * instead of MuJoCo's subtree-COM frame (same mathematics).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/roboplay.h"

#define MINVAL 1e-15
#define MINIMP 0.0001
#define MAXIMP 0.9999
#define MAXCON 64
#define MAXDEPTH 16

enum { T_FRICTION = 0, T_LIMIT = 1, T_CONTACT = 2 };
enum { G_SPHERE = 2, G_CAPSULE = 3, G_BOX = 6 };
enum { J_SLIDE = 2, J_HINGE = 3 };

typedef struct orc_model {
  rp_flat_model m; /* deep copy, pointers owned */
  void *blocks[96];
  int nblocks;
  int ntree;
  int *tree_adr, *tree_num, *dof_treeid; /* contiguous dof blocks */
  int *body_chain_len, *body_chain;      /* [nbody], [nbody*MAXDEPTH] dof chain root->leaf */
  int nhandgeom, npianogeom, *handgeom, *pianogeom;
  int npair, *pair_g1, *pair_g2; /* hand geom pairs that may collide (mj_collision's body-pair filters), MuJoCo order */
  int maxcon;
  int record_cap; /* > 0: emulate the kernel's contact-record capacity (see prune_to_record_cap) */
} orc_model;

typedef struct orc_contact {
  double dist, pos[3], frame[9], friction[3], solref[2], solimp[5], mu, includemargin;
  int geom1, geom2, body1, body2, efc_address, key; /* geom1 = hand geom, geom2 = piano geom */
} orc_contact;

typedef struct orc_data {
  int nv, nbody, ngeom, nsite, nu;
  double *qpos, *qvel, *qacc, *qacc_warmstart, *ctrl, *qfrc_applied;
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat, *site_xpos;
  double *S;    /* [nv,6] joint motion subspace at world origin: ang(3), lin(3) */
  double *Sdot; /* [nv,6] */
  double *V;    /* [nbody,6] body spatial velocity */
  double *M, *L, *LH; /* dense nv*nv: mass matrix, block Cholesky of M and of M+h*D */
  double *qfrc_bias, *qfrc_passive, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qfrc_constraint;
  double *actuator_length, *actuator_velocity, *actuator_force;
  int ncon, con_overflow;
  orc_contact con[MAXCON];
  int nefc, efc_cap;
  int *efc_type, *efc_id, *efc_t0, *efc_t1; /* trees touched by the row (-1 none) */
  double *efc_J, *efc_B; /* [nefc,nv]; B = Minv J^T */
  double *efc_pos, *efc_margin, *efc_frictionloss, *efc_diagApprox, *efc_R, *efc_D, *efc_K, *efc_Bd,
      *efc_imp, *efc_vel, *efc_aref, *efc_b, *efc_force, *efc_AR;
  int *adj, *adj_n; /* adjacency (rows sharing a tree): adj[i*efc_cap + k] */
  double *scr_MH, *scr_a, *scr_crb, *scr_A, *scr_F, *scr_jar; /* per-step scratch, allocated once: the baseline
      timing runs one instance per host thread and must not serialise on mmap / munmap of 157 KB blocks */
  int solver_niter;
  double solver_improvement;
  long mac_count;
} orc_data;

/* ----------------------------------------------------------------------------------------- */
static void *dupmem(orc_model *o, const void *src, size_t bytes) {
  void *p = malloc(bytes ? bytes : 1);
  if (bytes) memcpy(p, src, bytes);
  o->blocks[o->nblocks++] = p;
  return p;
}
#define DUP(field, count, type) o->m.field = (const type *)dupmem(o, src->field, sizeof(type) * (size_t)(count))

orc_model *orc_model_create(const rp_flat_model *src) {
  orc_model *o = (orc_model *)calloc(1, sizeof(orc_model));
  o->m = *src;
  int nb = src->nbody, nv = src->nv, ng = src->ngeom;
  DUP(body_parentid, nb, int32_t); DUP(body_rootid, nb, int32_t); DUP(body_dofadr, nb, int32_t);
  DUP(body_dofnum, nb, int32_t); DUP(body_pos, nb * 3, double); DUP(body_quat, nb * 4, double);
  DUP(body_ipos, nb * 3, double); DUP(body_iquat, nb * 4, double); DUP(body_mass, nb, double);
  DUP(body_inertia, nb * 3, double); DUP(body_gravcomp, nb, double); DUP(body_invweight0, nb * 2, double);
  DUP(jnt_type, nv, int32_t); DUP(jnt_bodyid, nv, int32_t); DUP(jnt_limited, nv, int32_t);
  DUP(dof_parentid, nv, int32_t); DUP(jnt_pos, nv * 3, double); DUP(jnt_axis, nv * 3, double);
  DUP(jnt_range, nv * 2, double); DUP(jnt_stiffness, nv, double); DUP(jnt_springref, nv, double);
  DUP(jnt_solref, nv * 2, double); DUP(jnt_solimp, nv * 5, double); DUP(jnt_margin, nv, double);
  DUP(dof_damping, nv, double); DUP(dof_armature, nv, double); DUP(dof_frictionloss, nv, double);
  DUP(dof_solref, nv * 2, double); DUP(dof_solimp, nv * 5, double); DUP(dof_invweight0, nv, double);
  DUP(geom_type, ng, int32_t); DUP(geom_bodyid, ng, int32_t); DUP(geom_contype, ng, int32_t);
  DUP(geom_conaffinity, ng, int32_t); DUP(geom_condim, ng, int32_t); DUP(geom_priority, ng, int32_t);
  DUP(geom_key, ng, int32_t); DUP(geom_hand, ng, int32_t); DUP(geom_forearm, ng, int32_t);
  DUP(geom_collide, ng, int32_t); DUP(geom_pos, ng * 3, double); DUP(geom_quat, ng * 4, double);
  DUP(geom_size, ng * 3, double); DUP(geom_friction, ng * 3, double); DUP(geom_solref, ng * 2, double);
  DUP(geom_solimp, ng * 5, double); DUP(geom_solmix, ng, double); DUP(geom_margin, ng, double);
  DUP(geom_gap, ng, double);
  DUP(site_bodyid, src->nsite, int32_t); DUP(site_pos, src->nsite * 3, double);
  DUP(tendon_adr, src->ntendon, int32_t); DUP(tendon_num, src->ntendon, int32_t);
  DUP(wrap_dofid, src->nwrap, int32_t); DUP(wrap_coef, src->nwrap, double);
  DUP(actuator_trntype, src->nu, int32_t); DUP(actuator_trnid, src->nu, int32_t);
  DUP(actuator_ctrllimited, src->nu, int32_t); DUP(actuator_forcelimited, src->nu, int32_t);
  DUP(actuator_gear, src->nu, double); DUP(actuator_gainprm, src->nu, double);
  DUP(actuator_biasprm, src->nu * 3, double); DUP(actuator_ctrlrange, src->nu * 2, double);
  DUP(actuator_forcerange, src->nu * 2, double);
  DUP(key_dofid, src->nkey, int32_t); DUP(key_geomid, src->nkey, int32_t);
  DUP(key_bodyid, src->nkey, int32_t); DUP(key_qmax, src->nkey, double);
  DUP(hand_obs_dofid, src->nhand * src->nobsdof, int32_t);
  /* dof trees: contiguous blocks, a new tree starts where dof_parentid == -1 */
  o->tree_adr = (int *)calloc(nv + 1, sizeof(int));
  o->tree_num = (int *)calloc(nv + 1, sizeof(int));
  o->dof_treeid = (int *)calloc(nv + 1, sizeof(int));
  for (int j = 0; j < nv; j++) {
    if (o->m.dof_parentid[j] < 0) { o->tree_adr[o->ntree] = j; o->ntree++; }
    o->dof_treeid[j] = o->ntree - 1;
    o->tree_num[o->ntree - 1]++;
  }
  o->body_chain_len = (int *)calloc(nb, sizeof(int));
  o->body_chain = (int *)calloc((size_t)nb * MAXDEPTH, sizeof(int));
  for (int b = 1; b < nb; b++) {
    int p = o->m.body_parentid[b], n = o->body_chain_len[p];
    memcpy(o->body_chain + b * MAXDEPTH, o->body_chain + p * MAXDEPTH, sizeof(int) * n);
    for (int k = 0; k < o->m.body_dofnum[b]; k++) o->body_chain[b * MAXDEPTH + n++] = o->m.body_dofadr[b] + k;
    if (n > MAXDEPTH) { fprintf(stderr, "oracle: chain too deep\n"); abort(); }
    o->body_chain_len[b] = n;
  }
  o->handgeom = (int *)calloc(ng + 1, sizeof(int));
  o->pianogeom = (int *)calloc(ng + 1, sizeof(int));
  for (int g = 0; g < ng; g++) {
    if (o->m.geom_hand[g] >= 0 && o->m.geom_collide[g]) o->handgeom[o->nhandgeom++] = g;
    if (o->m.geom_key[g] != -1) o->pianogeom[o->npianogeom++] = g;
  }
  /* hand geom : hand geom pairs (self and cross hand).  mj_collision's filters: contype / conaffinity; same body or
   * same weld body (a body without joints is welded to its parent: body_weldid); parent-child weld bodies unless one
   * of them is welded to the world (filterparent); <exclude> pairs.  Pair order = (body1, body2, geom1, geom2) with
   * body1 < body2, which for geoms numbered in body order is the lexicographic geom order. */
  DUP(exclude_bodies, src->nexclude * 2, int32_t);
  {
    int *weld = (int *)calloc(nb + 1, sizeof(int));
    for (int b = 1; b < nb; b++) weld[b] = o->m.body_dofnum[b] > 0 ? b : weld[o->m.body_parentid[b]];
    o->pair_g1 = (int *)calloc((size_t)o->nhandgeom * o->nhandgeom + 1, sizeof(int));
    o->pair_g2 = (int *)calloc((size_t)o->nhandgeom * o->nhandgeom + 1, sizeof(int));
    for (int i1 = 0; i1 < o->nhandgeom; i1++)
      for (int i2 = i1 + 1; i2 < o->nhandgeom; i2++) {
        int g1 = o->handgeom[i1], g2 = o->handgeom[i2];
        int b1 = o->m.geom_bodyid[g1], b2 = o->m.geom_bodyid[g2];
        if (!((o->m.geom_contype[g1] & o->m.geom_conaffinity[g2]) || (o->m.geom_contype[g2] & o->m.geom_conaffinity[g1]))) continue;
        int w1 = weld[b1], w2 = weld[b2];
        if (b1 == b2 || w1 == w2) continue;
        int wp1 = weld[o->m.body_parentid[w1]], wp2 = weld[o->m.body_parentid[w2]];
        if (w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
        int excl = 0;
        for (int e = 0; e < src->nexclude; e++) {
          int e1 = o->m.exclude_bodies[2 * e], e2 = o->m.exclude_bodies[2 * e + 1];
          if ((e1 == b1 && e2 == b2) || (e1 == b2 && e2 == b1)) excl = 1;
        }
        if (excl) continue;
        if (b1 > b2) { fprintf(stderr, "oracle: hand geoms are not in body order\n"); abort(); }
        o->pair_g1[o->npair] = g1; o->pair_g2[o->npair] = g2; o->npair++;
      }
    free(weld);
  }
  o->maxcon = 32;
  return o;
}

int orc_npair(const orc_model *o) { return o->npair; }
void orc_pair_get(const orc_model *o, int i, int *g1, int *g2) { *g1 = o->pair_g1[i]; *g2 = o->pair_g2[i]; }
void orc_model_destroy(orc_model *o) {
  free(o->pair_g1); free(o->pair_g2);
  for (int i = 0; i < o->nblocks; i++) free(o->blocks[i]);
  free(o->tree_adr); free(o->tree_num); free(o->dof_treeid); free(o->body_chain_len);
  free(o->body_chain); free(o->handgeom); free(o->pianogeom); free(o);
}
/* The CUDA kernel holds a fixed number of contact records (a contact across the hands takes two, every other contact one).  When
 * a step needs more, it keeps the deepest: records ordered by (distance in whole nanometres, contact order), kept while their position is below
 * the capacity, a hand : hand contact as a unit — and flags the environment.  MuJoCo has no such limit; this switch only
 * exists so that the kernel's behaviour beyond its capacity can be tested against the same rule. */
void orc_set_record_cap(orc_model *o, int cap) { o->record_cap = cap; }
void orc_set_option(orc_model *o, int iterations, double tolerance, int maxcon) {
  if (iterations > 0) o->m.iterations = iterations;
  if (tolerance >= 0) o->m.tolerance = tolerance;
  if (maxcon > 0 && maxcon <= MAXCON) o->maxcon = maxcon;
}

#define ALLOC(field, count) d->field = (double *)calloc((size_t)(count) + 1, sizeof(double))
#define ALLOCI(field, count) d->field = (int *)calloc((size_t)(count) + 1, sizeof(int))

orc_data *orc_data_create(const orc_model *o) {
  orc_data *d = (orc_data *)calloc(1, sizeof(orc_data));
  int nv = o->m.nv, nb = o->m.nbody, ng = o->m.ngeom, ns = o->m.nsite, nu = o->m.nu;
  d->nv = nv; d->nbody = nb; d->ngeom = ng; d->nsite = ns; d->nu = nu;
  ALLOC(qpos, nv); ALLOC(qvel, nv); ALLOC(qacc, nv); ALLOC(qacc_warmstart, nv); ALLOC(ctrl, nu);
  ALLOC(qfrc_applied, nv);
  ALLOC(xpos, nb * 3); ALLOC(xquat, nb * 4); ALLOC(xmat, nb * 9); ALLOC(xipos, nb * 3); ALLOC(ximat, nb * 9);
  ALLOC(xanchor, nv * 3); ALLOC(xaxis, nv * 3); ALLOC(geom_xpos, ng * 3); ALLOC(geom_xmat, ng * 9);
  ALLOC(site_xpos, ns * 3); ALLOC(S, nv * 6); ALLOC(Sdot, nv * 6); ALLOC(V, nb * 6);
  ALLOC(M, nv * nv); ALLOC(L, nv * nv); ALLOC(LH, nv * nv);
  ALLOC(qfrc_bias, nv); ALLOC(qfrc_passive, nv); ALLOC(qfrc_actuator, nv); ALLOC(qfrc_smooth, nv);
  ALLOC(qacc_smooth, nv); ALLOC(qfrc_constraint, nv);
  ALLOC(actuator_length, nu); ALLOC(actuator_velocity, nu); ALLOC(actuator_force, nu);
  int cap = 2 * nv + 3 * MAXCON + 8;
  d->efc_cap = cap;
  ALLOCI(efc_type, cap); ALLOCI(efc_id, cap); ALLOCI(efc_t0, cap); ALLOCI(efc_t1, cap);
  ALLOC(efc_J, cap * nv); ALLOC(efc_B, cap * nv);
  ALLOC(efc_pos, cap); ALLOC(efc_margin, cap); ALLOC(efc_frictionloss, cap); ALLOC(efc_diagApprox, cap);
  ALLOC(efc_R, cap); ALLOC(efc_D, cap); ALLOC(efc_K, cap); ALLOC(efc_Bd, cap); ALLOC(efc_imp, cap);
  ALLOC(efc_vel, cap); ALLOC(efc_aref, cap); ALLOC(efc_b, cap); ALLOC(efc_force, cap);
  ALLOC(efc_AR, cap * cap);
  ALLOCI(adj, cap * cap); ALLOCI(adj_n, cap);
  ALLOC(scr_MH, nv * nv); ALLOC(scr_a, nv); ALLOC(scr_crb, nb * 10); ALLOC(scr_A, nb * 6); ALLOC(scr_F, nb * 6); ALLOC(scr_jar, cap);
  return d;
}
void orc_data_destroy(orc_data *d) {
  free(d->qpos); free(d->qvel); free(d->qacc); free(d->qacc_warmstart); free(d->ctrl); free(d->qfrc_applied);
  free(d->xpos); free(d->xquat); free(d->xmat); free(d->xipos); free(d->ximat); free(d->xanchor);
  free(d->xaxis); free(d->geom_xpos); free(d->geom_xmat); free(d->site_xpos); free(d->S); free(d->Sdot);
  free(d->V); free(d->M); free(d->L); free(d->LH); free(d->qfrc_bias); free(d->qfrc_passive);
  free(d->qfrc_actuator); free(d->qfrc_smooth); free(d->qacc_smooth); free(d->qfrc_constraint);
  free(d->actuator_length); free(d->actuator_velocity); free(d->actuator_force);
  free(d->efc_type); free(d->efc_id); free(d->efc_t0); free(d->efc_t1); free(d->efc_J); free(d->efc_B);
  free(d->efc_pos); free(d->efc_margin); free(d->efc_frictionloss); free(d->efc_diagApprox);
  free(d->efc_R); free(d->efc_D); free(d->efc_K); free(d->efc_Bd); free(d->efc_imp); free(d->efc_vel);
  free(d->efc_aref); free(d->efc_b); free(d->efc_force); free(d->efc_AR); free(d->adj); free(d->adj_n);
  free(d->scr_MH); free(d->scr_a); free(d->scr_crb); free(d->scr_A); free(d->scr_F); free(d->scr_jar);
  free(d);
}

/* ---------------------------------- small vector helpers ---------------------------------- */
static inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double norm3(const double *a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double *a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return n; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
static void quat_mul(double *r, const double *a, const double *b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_normalize(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat2mat(double *m, const double *q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void mat_vec(double *r, const double *m, const double *v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void matT_vec(double *r, const double *m, const double *v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  double y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  double z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void rot_vec_quat(double *r, const double *q, const double *v) {
  double m[9];
  quat2mat(m, q);
  mat_vec(r, m, v);
}
static void axis_angle2quat(double *q, const double *axis, double angle) {
  double s = sin(0.5 * angle);
  q[0] = cos(0.5 * angle); q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}

/* ------------------------------------ mj_kinematics ---------------------------------------- */
void orc_kinematics(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  d->xpos[0] = d->xpos[1] = d->xpos[2] = 0;
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    double pos[3], quat[4], tmp[3];
    rot_vec_quat(tmp, d->xquat + 4 * p, m->body_pos + 3 * b);
    for (int k = 0; k < 3; k++) pos[k] = d->xpos[3 * p + k] + tmp[k];
    quat_mul(quat, d->xquat + 4 * p, m->body_quat + 4 * b);
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      int j = m->body_dofadr[b] + k;
      double anchor[3], axis[3];
      rot_vec_quat(tmp, quat, m->jnt_pos + 3 * j);
      for (int c = 0; c < 3; c++) anchor[c] = pos[c] + tmp[c];
      rot_vec_quat(axis, quat, m->jnt_axis + 3 * j);
      double q = d->qpos[j]; /* qpos0 == 0 for every joint of this scene */
      if (m->jnt_type[j] == J_SLIDE) {
        for (int c = 0; c < 3; c++) pos[c] += axis[c] * q;
      } else {
        double qloc[4], qnew[4];
        axis_angle2quat(qloc, m->jnt_axis + 3 * j, q);
        quat_mul(qnew, quat, qloc);
        memcpy(quat, qnew, sizeof(qnew));
        rot_vec_quat(tmp, quat, m->jnt_pos + 3 * j);
        for (int c = 0; c < 3; c++) pos[c] = anchor[c] - tmp[c];
      }
      memcpy(d->xanchor + 3 * j, anchor, sizeof(anchor));
      memcpy(d->xaxis + 3 * j, axis, sizeof(axis));
    }
    quat_normalize(quat);
    memcpy(d->xpos + 3 * b, pos, sizeof(pos));
    memcpy(d->xquat + 4 * b, quat, sizeof(quat));
    quat2mat(d->xmat + 9 * b, quat);
    rot_vec_quat(tmp, quat, m->body_ipos + 3 * b);
    for (int c = 0; c < 3; c++) d->xipos[3 * b + c] = pos[c] + tmp[c];
    double iq[4];
    quat_mul(iq, quat, m->body_iquat + 4 * b);
    quat2mat(d->ximat + 9 * b, iq);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    double tmp[3], q[4];
    mat_vec(tmp, d->xmat + 9 * b, m->geom_pos + 3 * g);
    for (int c = 0; c < 3; c++) d->geom_xpos[3 * g + c] = d->xpos[3 * b + c] + tmp[c];
    quat_mul(q, d->xquat + 4 * b, m->geom_quat + 4 * g);
    quat_normalize(q);
    quat2mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    double tmp[3];
    mat_vec(tmp, d->xmat + 9 * b, m->site_pos + 3 * s);
    for (int c = 0; c < 3; c++) d->site_xpos[3 * s + c] = d->xpos[3 * b + c] + tmp[c];
  }
  /* joint motion subspaces at the world origin (mj_comPos: cdof) */
  for (int j = 0; j < m->nv; j++) {
    double *S = d->S + 6 * j;
    if (m->jnt_type[j] == J_HINGE) {
      memcpy(S, d->xaxis + 3 * j, 3 * sizeof(double));
      cross3(S + 3, d->xanchor + 3 * j, d->xaxis + 3 * j);
    } else {
      S[0] = S[1] = S[2] = 0;
      memcpy(S + 3, d->xaxis + 3 * j, 3 * sizeof(double));
    }
  }
}

/* spatial inertia of body b at the world origin: out = {m, h[3], I[6] (xx,yy,zz,xy,xz,yz)} */
static void body_inertia_world(const orc_model *o, const orc_data *d, int b, double *out) {
  const rp_flat_model *m = &o->m;
  double mass = m->body_mass[b];
  const double *R = d->ximat + 9 * b, *c = d->xipos + 3 * b, *di = m->body_inertia + 3 * b;
  double I[9];
  for (int r = 0; r < 3; r++)
    for (int s = 0; s < 3; s++)
      I[3 * r + s] = R[3 * r] * di[0] * R[3 * s] + R[3 * r + 1] * di[1] * R[3 * s + 1] + R[3 * r + 2] * di[2] * R[3 * s + 2];
  double cc = dot3(c, c);
  out[0] = mass;
  out[1] = mass * c[0]; out[2] = mass * c[1]; out[3] = mass * c[2];
  out[4] = I[0] + mass * (cc - c[0] * c[0]);
  out[5] = I[4] + mass * (cc - c[1] * c[1]);
  out[6] = I[8] + mass * (cc - c[2] * c[2]);
  out[7] = I[1] - mass * c[0] * c[1];
  out[8] = I[2] - mass * c[0] * c[2];
  out[9] = I[5] - mass * c[1] * c[2];
}
/* F = I * V for spatial inertia (10 numbers) and motion vector V = [w; v]; F = [n; f] */
static void inertia_mul(double *F, const double *I, const double *V) {
  const double *w = V, *v = V + 3, *h = I + 1;
  double Iw[3] = {I[4] * w[0] + I[7] * w[1] + I[8] * w[2], I[7] * w[0] + I[5] * w[1] + I[9] * w[2],
                  I[8] * w[0] + I[9] * w[1] + I[6] * w[2]};
  double hxv[3], wxh[3];
  cross3(hxv, h, v);
  cross3(wxh, w, h);
  for (int k = 0; k < 3; k++) { F[k] = Iw[k] + hxv[k]; F[3 + k] = I[0] * v[k] + wxh[k]; }
}
static void cross_motion(double *r, const double *V, const double *S) {
  double a[3], b[3], c[3];
  cross3(a, V, S);
  cross3(b, V, S + 3);
  cross3(c, V + 3, S);
  for (int k = 0; k < 3; k++) { r[k] = a[k]; r[3 + k] = b[k] + c[k]; }
}
static void cross_force(double *r, const double *V, const double *F) {
  double a[3], b[3], c[3];
  cross3(a, V, F);
  cross3(b, V + 3, F + 3);
  cross3(c, V, F + 3);
  for (int k = 0; k < 3; k++) { r[k] = a[k] + b[k]; r[3 + k] = c[k]; }
}
static inline double dot6(const double *a, const double *b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

/* ---------------------------------- mj_crb + mj_factorM ------------------------------------ */
static int chol_block(double *L, const double *A, int ld, int adr, int n) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[(adr + i) * ld + adr + j];
      for (int k = 0; k < j; k++) s -= L[(adr + i) * ld + adr + k] * L[(adr + j) * ld + adr + k];
      if (i == j) {
        if (s <= 0) return -1;
        L[(adr + i) * ld + adr + i] = sqrt(s);
      } else
        L[(adr + i) * ld + adr + j] = s / L[(adr + j) * ld + adr + j];
    }
  return 0;
}
static void chol_solve_block(const double *L, int ld, int adr, int n, double *x) {
  for (int i = 0; i < n; i++) {
    double s = x[adr + i];
    for (int k = 0; k < i; k++) s -= L[(adr + i) * ld + adr + k] * x[adr + k];
    x[adr + i] = s / L[(adr + i) * ld + adr + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[adr + i];
    for (int k = i + 1; k < n; k++) s -= L[(adr + k) * ld + adr + i] * x[adr + k];
    x[adr + i] = s / L[(adr + i) * ld + adr + i];
  }
}
static void solve_M(const orc_model *o, const double *L, int nv, double *x) {
  for (int t = 0; t < o->ntree; t++) chol_solve_block(L, nv, o->tree_adr[t], o->tree_num[t], x);
}

void orc_crb(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  int nv = m->nv, nb = m->nbody;
  double *crb = d->scr_crb;
  memset(crb, 0, sizeof(double) * (size_t)nb * 10);
  for (int b = 1; b < nb; b++) body_inertia_world(o, d, b, crb + 10 * b);
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int k = 0; k < 10; k++) crb[10 * p + k] += crb[10 * b + k];
  }
  memset(d->M, 0, sizeof(double) * nv * nv);
  for (int i = 0; i < nv; i++) {
    double F[6];
    inertia_mul(F, crb + 10 * m->jnt_bodyid[i], d->S + 6 * i);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      double v = dot6(d->S + 6 * j, F);
      d->M[i * nv + j] = v;
      d->M[j * nv + i] = v;
    }
    d->M[i * nv + i] += m->dof_armature[i];
  }
  for (int t = 0; t < o->ntree; t++)
    if (chol_block(d->L, d->M, nv, o->tree_adr[t], o->tree_num[t])) fprintf(stderr, "oracle: M not SPD\n");
}

/* -------------------- mj_comVel + mj_rne (bias) + mj_passive ------------------------------ */
void orc_velocity(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  int nv = m->nv, nb = m->nbody;
  double *A = d->scr_A, *F = d->scr_F;
  memset(A, 0, sizeof(double) * (size_t)nb * 6);
  memset(F, 0, sizeof(double) * (size_t)nb * 6);
  memset(d->V, 0, sizeof(double) * 6 * nb);
  for (int k = 0; k < 3; k++) A[3 + k] = -m->gravity[k];
  for (int b = 1; b < nb; b++) {
    int p = m->body_parentid[b];
    double *V = d->V + 6 * b, *Ab = A + 6 * b;
    memcpy(V, d->V + 6 * p, 6 * sizeof(double));
    memcpy(Ab, A + 6 * p, 6 * sizeof(double));
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      int j = m->body_dofadr[b] + k;
      cross_motion(d->Sdot + 6 * j, V, d->S + 6 * j);
      for (int c = 0; c < 6; c++) {
        Ab[c] += d->Sdot[6 * j + c] * d->qvel[j];
        V[c] += d->S[6 * j + c] * d->qvel[j];
      }
    }
    double I[10], IA[6], IV[6], VxIV[6];
    body_inertia_world(o, d, b, I);
    inertia_mul(IA, I, Ab);
    inertia_mul(IV, I, V);
    cross_force(VxIV, V, IV);
    for (int c = 0; c < 6; c++) F[6 * b + c] = IA[c] + VxIV[c];
  }
  for (int b = nb - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    if (p > 0) for (int c = 0; c < 6; c++) F[6 * p + c] += F[6 * b + c];
  }
  for (int j = 0; j < nv; j++) d->qfrc_bias[j] = dot6(d->S + 6 * j, F + 6 * m->jnt_bodyid[j]);
  /* passive: spring, damper, gravity compensation */
  for (int j = 0; j < nv; j++)
    d->qfrc_passive[j] = -m->jnt_stiffness[j] * (d->qpos[j] - m->jnt_springref[j]) - m->dof_damping[j] * d->qvel[j];
  for (int b = 1; b < nb; b++) {
    double gc = m->body_gravcomp[b];
    if (gc == 0) continue;
    double f[3], Fgc[6];
    for (int k = 0; k < 3; k++) f[k] = -m->gravity[k] * m->body_mass[b] * gc;
    cross3(Fgc, d->xipos + 3 * b, f);
    memcpy(Fgc + 3, f, sizeof(f));
    for (int c = 0; c < o->body_chain_len[b]; c++) {
      int j = o->body_chain[b * MAXDEPTH + c];
      d->qfrc_passive[j] += dot6(d->S + 6 * j, Fgc);
    }
  }
  /* transmission: lengths and velocities (mj_transmission, mj_fwdVelocity) */
  for (int a = 0; a < m->nu; a++) {
    double len = 0, vel = 0;
    if (m->actuator_trntype[a] == 0) {
      int j = m->actuator_trnid[a];
      len = d->qpos[j]; vel = d->qvel[j];
    } else {
      int t = m->actuator_trnid[a];
      for (int w = 0; w < m->tendon_num[t]; w++) {
        int j = m->wrap_dofid[m->tendon_adr[t] + w];
        len += m->wrap_coef[m->tendon_adr[t] + w] * d->qpos[j];
        vel += m->wrap_coef[m->tendon_adr[t] + w] * d->qvel[j];
      }
    }
    d->actuator_length[a] = m->actuator_gear[a] * len;
    d->actuator_velocity[a] = m->actuator_gear[a] * vel;
  }
}

/* ------------------------------------ mj_collision ----------------------------------------- */
/* Signed distance from sphere (centre c in box frame, radius r) to a box with half sizes s.
 * Returns dist; n = unit normal in box frame pointing from the sphere towards the box;
 * p = contact point (mid-way between the surfaces) in box frame. */
static double sphere_box(const double *c, double r, const double *s, double *n, double *p) {
  double cl[3], dlt[3];
  int inside = 1;
  for (int k = 0; k < 3; k++) {
    cl[k] = c[k] < -s[k] ? -s[k] : (c[k] > s[k] ? s[k] : c[k]);
    dlt[k] = c[k] - cl[k];
    if (dlt[k] != 0) inside = 0;
  }
  double dist;
  if (!inside) {
    double len = norm3(dlt);
    dist = len - r;
    for (int k = 0; k < 3; k++) n[k] = -dlt[k] / len;
    for (int k = 0; k < 3; k++) p[k] = cl[k] - n[k] * 0.5 * dist;
  } else {
    /* centre inside the box: push out through the nearest face */
    int best = 0;
    double bd = 1e300;
    for (int k = 0; k < 3; k++) {
      double dk = s[k] - fabs(c[k]);
      if (dk < bd) { bd = dk; best = k; }
    }
    double sg = c[best] >= 0 ? 1.0 : -1.0;
    n[0] = n[1] = n[2] = 0;
    n[best] = -sg;
    dist = -bd - r;
    for (int k = 0; k < 3; k++) p[k] = c[k];
    p[best] = sg * s[best];
    for (int k = 0; k < 3; k++) p[k] -= n[k] * 0.5 * dist;
  }
  return dist;
}
/* squared distance from point to box (0 inside) */
static double point_box_d2(const double *c, const double *s) {
  double d2 = 0;
  for (int k = 0; k < 3; k++) {
    double e = fabs(c[k]) - s[k];
    if (e > 0) d2 += e * e;
  }
  return d2;
}
/* parameter t in [0,1] minimising the distance from a + t*(b-a) to the box.  The squared
 * distance is convex piecewise quadratic in t with breakpoints where a coordinate crosses a
 * face plane; each piece is minimised in closed form.  Ties keep the smallest t. */
static double segment_box_closest_t(const double *a, const double *b, const double *s) {
  double bp[8];
  int nbp = 0;
  bp[nbp++] = 0; bp[nbp++] = 1;
  double dir[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  for (int k = 0; k < 3; k++) {
    if (fabs(dir[k]) < MINVAL) continue;
    double t1 = (-s[k] - a[k]) / dir[k], t2 = (s[k] - a[k]) / dir[k];
    if (t1 > 0 && t1 < 1) bp[nbp++] = t1;
    if (t2 > 0 && t2 < 1) bp[nbp++] = t2;
  }
  for (int i = 1; i < nbp; i++) { /* insertion sort */
    double v = bp[i];
    int j = i - 1;
    while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; j--; }
    bp[j + 1] = v;
  }
  double best_t = 0, best_d = 1e300;
  for (int i = 0; i + 1 < nbp; i++) {
    double t0 = bp[i], t1 = bp[i + 1];
    if (t1 - t0 < 1e-14) continue;
    double tm = 0.5 * (t0 + t1), qa = 0, qb = 0; /* d2(t) = qa t^2 + qb t + c on this piece */
    for (int k = 0; k < 3; k++) {
      double pm = a[k] + tm * dir[k];
      if (pm > s[k]) { qa += dir[k] * dir[k]; qb += 2 * dir[k] * (a[k] - s[k]); }
      else if (pm < -s[k]) { qa += dir[k] * dir[k]; qb += 2 * dir[k] * (a[k] + s[k]); }
    }
    double tc = t0;
    if (qa > MINVAL) { tc = -qb / (2 * qa); tc = tc < t0 ? t0 : (tc > t1 ? t1 : tc); }
    double cand[3] = {t0, tc, t1};
    for (int c = 0; c < 3; c++) {
      double p[3] = {a[0] + cand[c] * dir[0], a[1] + cand[c] * dir[1], a[2] + cand[c] * dir[2]};
      double d2 = point_box_d2(p, s);
      if (d2 < best_d - 1e-18) { best_d = d2; best_t = cand[c]; }
    }
  }
  if (nbp == 2 && best_d == 1e300) best_t = 0;
  return best_t;
}
static void make_frame(double *frame) {
  /* mju_makeFrame: frame[0:3] = normal (unit); builds tangents */
  double *x = frame, *y = frame + 3, *z = frame + 6;
  y[0] = y[1] = y[2] = 0;
  if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
  double t = dot3(x, y);
  for (int k = 0; k < 3; k++) y[k] -= t * x[k];
  normalize3(y);
  cross3(z, x, y);
}
static double mix_solparams(const rp_flat_model *m, int g1, int g2, double *solref, double *solimp, double *friction) {
  /* mj_contactParam: priority equal in this scene -> solmix weighted average, max friction */
  double mix;
  int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  if (p1 != p2) mix = p1 > p2 ? 1.0 : 0.0;
  else {
    double s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
    if (s1 >= MINVAL && s2 >= MINVAL) mix = s1 / (s1 + s2);
    else if (s1 < MINVAL && s2 < MINVAL) mix = 0.5;
    else mix = s1 < MINVAL ? 0.0 : 1.0;
  }
  const double *r1 = m->geom_solref + 2 * g1, *r2 = m->geom_solref + 2 * g2;
  if (r1[0] > 0 && r2[0] > 0) for (int k = 0; k < 2; k++) solref[k] = mix * r1[k] + (1 - mix) * r2[k];
  else for (int k = 0; k < 2; k++) solref[k] = r1[k] < r2[k] ? r1[k] : r2[k];
  for (int k = 0; k < 5; k++) solimp[k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
  for (int k = 0; k < 3; k++) {
    double f1 = m->geom_friction[3 * g1 + k], f2 = m->geom_friction[3 * g2 + k];
    friction[k] = p1 == p2 ? (f1 > f2 ? f1 : f2) : (p1 > p2 ? f1 : f2);
  }
  return mix;
}
static void add_contact(const orc_model *o, orc_data *d, int gh, int gp, double dist, const double *n_box,
                        const double *p_box, double margin, double gap) {
  const rp_flat_model *m = &o->m;
  if (!(dist < margin)) return;
  if (d->ncon >= o->maxcon) { d->con_overflow = 1; return; }
  orc_contact *c = &d->con[d->ncon++];
  const double *R = d->geom_xmat + 9 * gp, *bp = d->geom_xpos + 3 * gp;
  double tmp[3];
  mat_vec(c->frame, R, n_box);
  normalize3(c->frame);
  make_frame(c->frame);
  mat_vec(tmp, R, p_box);
  for (int k = 0; k < 3; k++) c->pos[k] = bp[k] + tmp[k];
  c->dist = dist;
  c->includemargin = margin - gap;
  c->geom1 = gh; c->geom2 = gp;
  c->body1 = m->geom_bodyid[gh]; c->body2 = m->geom_bodyid[gp];
  c->key = m->geom_key[gp];
  mix_solparams(m, gh, gp, c->solref, c->solimp, c->friction);
  c->mu = c->friction[0];
  c->efc_address = -1;
}
/* Hand capsule / sphere against piano box.  Capsule: contacts at the segment end points and,
 * when it is interior, at the segment point closest to the box (our own scheme; MuJoCo's
 * mjc_CapsuleBox is not restated verbatim — see DESIGN.md "narrowphase").  Order: t=0, t=1,
 * interior. */
static void collide_pair(const orc_model *o, orc_data *d, int gh, int gp) {
  const rp_flat_model *m = &o->m;
  const double *R = d->geom_xmat + 9 * gp, *bp = d->geom_xpos + 3 * gp, *s = m->geom_size + 3 * gp;
  double margin = m->geom_margin[gh] > m->geom_margin[gp] ? m->geom_margin[gh] : m->geom_margin[gp];
  double gap = m->geom_gap[gh] > m->geom_gap[gp] ? m->geom_gap[gh] : m->geom_gap[gp];
  double r = m->geom_size[3 * gh];
  double rel[3], c[3], n[3], p[3];
  for (int k = 0; k < 3; k++) rel[k] = d->geom_xpos[3 * gh + k] - bp[k];
  matT_vec(c, R, rel);
  if (m->geom_type[gh] == G_SPHERE) {
    /* bounding test */
    if (point_box_d2(c, s) > (r + margin) * (r + margin)) return;
    double dist = sphere_box(c, r, s, n, p);
    add_contact(o, d, gh, gp, dist, n, p, margin, gap);
    return;
  }
  double half = m->geom_size[3 * gh + 1];
  double axw[3] = {d->geom_xmat[9 * gh + 2], d->geom_xmat[9 * gh + 5], d->geom_xmat[9 * gh + 8]}, ax[3];
  matT_vec(ax, R, axw);
  double bound = r + half + margin;
  if (point_box_d2(c, s) > bound * bound) return;
  double a[3], b[3];
  for (int k = 0; k < 3; k++) { a[k] = c[k] - half * ax[k]; b[k] = c[k] + half * ax[k]; }
  double dist = sphere_box(a, r, s, n, p);
  add_contact(o, d, gh, gp, dist, n, p, margin, gap);
  dist = sphere_box(b, r, s, n, p);
  add_contact(o, d, gh, gp, dist, n, p, margin, gap);
  double t = segment_box_closest_t(a, b, s);
  if (t > 1e-9 && t < 1 - 1e-9) {
    double q[3] = {a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1]), a[2] + t * (b[2] - a[2])};
    dist = sphere_box(q, r, s, n, p);
    add_contact(o, d, gh, gp, dist, n, p, margin, gap);
  }
}
/* ---- hand geom : hand geom narrowphase: MuJoCo's analytic routines for spheres and capsules
 * (engine_collision_primitive.c: mjraw_SphereSphere, mjc_SphereCapsule, mjraw_CapsuleCapsule), restated.
 * A "sphere" below is (centre, radius); the z axes are only used for the coincident-centre normal. */
static int raw_sphere_sphere(double margin, const double *p1, const double *z1, double r1, const double *p2, const double *z2,
                             double r2, double *dist, double *normal, double *pos) {
  double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  double cd2 = dot3(dif, dif), mind = margin + r1 + r2;
  if (cd2 > mind * mind) return 0;
  double len = sqrt(cd2);
  if (len < MINVAL) { dif[0] = 1; dif[1] = 0; dif[2] = 0; len = 0; } /* mju_normalize3 */
  else for (int k = 0; k < 3; k++) dif[k] /= len;
  *dist = len - r1 - r2;
  for (int k = 0; k < 3; k++) normal[k] = dif[k];
  if (len < MINVAL) { /* coincident centres: cross product of the z axes, or x */
    cross3(normal, z1, z2);
    if (normalize3(normal) < MINVAL) { normal[0] = 1; normal[1] = 0; normal[2] = 0; }
  }
  for (int k = 0; k < 3; k++) pos[k] = p1[k] + normal[k] * (r1 + 0.5 * *dist);
  return 1;
}
static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
/* returns the number of contacts (0..2); outputs [2] */
static int raw_capsule_capsule(double margin, const double *p1, const double *ax1, double r1, double h1, const double *p2,
                               const double *ax2, double r2, double h2, double *dist, double *normal, double *pos) {
  double dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  double ma = dot3(ax1, ax1), mb = -dot3(ax1, ax2), mc = dot3(ax2, ax2);
  double u = -dot3(ax1, dif), v = dot3(ax2, dif);
  double det = ma * mc - mb * mb;
  double v1[3], v2[3];
  if (fabs(det) >= MINVAL) {
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > h1) { x1 = h1; x2 = (v - mb * h1) / mc; }
    else if (x1 < -h1) { x1 = -h1; x2 = (v + mb * h1) / mc; }
    if (x2 > h2) { x2 = h2; x1 = clipd((u - mb * h2) / ma, -h1, h1); }
    else if (x2 < -h2) { x2 = -h2; x1 = clipd((u + mb * h2) / ma, -h1, h1); }
    for (int k = 0; k < 3; k++) { v1[k] = p1[k] + ax1[k] * x1; v2[k] = p2[k] + ax2[k] * x2; }
    return raw_sphere_sphere(margin, v1, ax1, r1, v2, ax2, r2, dist, normal, pos);
  }
  /* parallel axes: both ends of capsule 1 against segment 2, then both ends of capsule 2 against segment 1, at most 2 */
  int n = 0;
  for (int pass = 0; pass < 4 && n < 2; pass++) {
    double x1, x2;
    if (pass == 0) { x1 = h1; x2 = clipd((v - mb * h1) / mc, -h2, h2); }
    else if (pass == 1) { x1 = -h1; x2 = clipd((v + mb * h1) / mc, -h2, h2); }
    else if (pass == 2) { x2 = h2; x1 = clipd((u - mb * h2) / ma, -h1, h1); }
    else { x2 = -h2; x1 = clipd((u + mb * h2) / ma, -h1, h1); }
    for (int k = 0; k < 3; k++) { v1[k] = p1[k] + ax1[k] * x1; v2[k] = p2[k] + ax2[k] * x2; }
    n += raw_sphere_sphere(margin, v1, ax1, r1, v2, ax2, r2, dist + n, normal + 3 * n, pos + 3 * n);
  }
  return n;
}
static void add_pair_contact(const orc_model *o, orc_data *d, int g1, int g2, double dist, const double *normal,
                             const double *pos, double margin, double gap) {
  const rp_flat_model *m = &o->m;
  if (d->ncon >= o->maxcon) { d->con_overflow = 1; return; }
  orc_contact *c = &d->con[d->ncon++];
  for (int k = 0; k < 3; k++) { c->frame[k] = normal[k]; c->pos[k] = pos[k]; }
  make_frame(c->frame);
  c->dist = dist;
  c->includemargin = margin - gap;
  c->geom1 = g1; c->geom2 = g2;
  c->body1 = m->geom_bodyid[g1]; c->body2 = m->geom_bodyid[g2];
  c->key = -1;
  mix_solparams(m, g1, g2, c->solref, c->solimp, c->friction);
  c->mu = c->friction[0];
  c->efc_address = -1;
}
/* one hand geom pair; mj_collideGeoms orders the two geoms by type first (sphere before capsule) */
static void collide_hand_pair(const orc_model *o, orc_data *d, int g1, int g2) {
  const rp_flat_model *m = &o->m;
  if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
  double margin = m->geom_margin[g1] > m->geom_margin[g2] ? m->geom_margin[g1] : m->geom_margin[g2];
  double gap = m->geom_gap[g1] > m->geom_gap[g2] ? m->geom_gap[g1] : m->geom_gap[g2];
  const double *p1 = d->geom_xpos + 3 * g1, *p2 = d->geom_xpos + 3 * g2;
  double z1[3] = {d->geom_xmat[9 * g1 + 2], d->geom_xmat[9 * g1 + 5], d->geom_xmat[9 * g1 + 8]};
  double z2[3] = {d->geom_xmat[9 * g2 + 2], d->geom_xmat[9 * g2 + 5], d->geom_xmat[9 * g2 + 8]};
  double r1 = m->geom_size[3 * g1], r2 = m->geom_size[3 * g2];
  double h1 = m->geom_type[g1] == G_CAPSULE ? m->geom_size[3 * g1 + 1] : 0.0, h2 = m->geom_type[g2] == G_CAPSULE ? m->geom_size[3 * g2 + 1] : 0.0;
  /* bounding spheres (mj_collideGeoms' first filter) */
  {
    double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, bound = r1 + h1 + r2 + h2 + margin;
    if (dot3(dif, dif) > bound * bound) return;
  }
  double dist[2], normal[6], pos[6];
  int n;
  if (m->geom_type[g1] == G_SPHERE && m->geom_type[g2] == G_SPHERE) {
    n = raw_sphere_sphere(margin, p1, z1, r1, p2, z2, r2, dist, normal, pos);
  } else if (m->geom_type[g1] == G_SPHERE) { /* mjc_SphereCapsule: nearest point of the segment, then sphere : sphere */
    double vec[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    double x = clipd(dot3(z2, vec), -h2, h2);
    double q[3] = {p2[0] + z2[0] * x, p2[1] + z2[1] * x, p2[2] + z2[2] * x};
    n = raw_sphere_sphere(margin, p1, z1, r1, q, z2, r2, dist, normal, pos);
  } else {
    n = raw_capsule_capsule(margin, p1, z1, r1, h1, p2, z2, r2, h2, dist, normal, pos);
  }
  for (int i = 0; i < n; i++)
    if (dist[i] < margin) add_pair_contact(o, d, g1, g2, dist[i], normal + 3 * i, pos + 3 * i, margin, gap);
}
void orc_collision(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  d->ncon = 0;
  d->con_overflow = 0;
  /* MuJoCo orders contacts by body pair (lower body id first): piano bodies precede hand bodies,
   * so the outer loop is over piano geoms, the inner one over hand geoms. */
  for (int ip = 0; ip < o->npianogeom; ip++) {
    int gp = o->pianogeom[ip];
    for (int ih = 0; ih < o->nhandgeom; ih++) {
      int gh = o->handgeom[ih];
      int ok = (m->geom_contype[gh] & m->geom_conaffinity[gp]) || (m->geom_contype[gp] & m->geom_conaffinity[gh]);
      if (!ok) continue;
      collide_pair(o, d, gh, gp);
    }
  }
  /* hand bodies follow the piano bodies: their pairs come last */
  for (int i = 0; i < o->npair; i++) collide_hand_pair(o, d, o->pair_g1[i], o->pair_g2[i]);
  if (o->record_cap > 0) {
    const rp_flat_model *mm = &o->m;
    int nrec = 0;
    /* records: a contact across the hands takes two (one per body), every other contact one */
#define ORC_CROSS(c_) (mm->geom_hand[d->con[c_].geom2] >= 0 && mm->geom_hand[d->con[c_].geom1] >= 0 && mm->geom_hand[d->con[c_].geom1] != mm->geom_hand[d->con[c_].geom2])
    for (int c = 0; c < d->ncon; c++) nrec += ORC_CROSS(c) ? 2 : 1;
    if (nrec > o->record_cap) {
      int order[MAXCON], keep[MAXCON];
      for (int c = 0; c < d->ncon; c++) order[c] = c;
      for (int a = 1; a < d->ncon; a++) { /* insertion sort by (dist, index): stable */
        int x = order[a], b = a - 1;
        while (b >= 0 && floor(d->con[order[b]].dist * 1e9) > floor(d->con[x].dist * 1e9)) { order[b + 1] = order[b]; b--; } /* whole nanometres */
        order[b + 1] = x;
      }
      int pos = 0;
      for (int a = 0; a < d->ncon; a++) {
        int c = order[a], pair = ORC_CROSS(c);
        keep[c] = (pair ? pos + 1 : pos) < o->record_cap;
        pos += pair ? 2 : 1;
      }
      int n = 0;
      for (int c = 0; c < d->ncon; c++) if (keep[c]) d->con[n++] = d->con[c];
      d->ncon = n;
      d->con_overflow = 1;
    }
  }
}

/* --------------------------------- mj_makeConstraint --------------------------------------- */
static void point_jac_row(const orc_model *o, const orc_data *d, int body, const double *point, const double *dir,
                          double sign, double *Jrow) {
  for (int c = 0; c < o->body_chain_len[body]; c++) {
    int j = o->body_chain[body * MAXDEPTH + c];
    const double *S = d->S + 6 * j;
    double wxp[3];
    cross3(wxp, S, point);
    Jrow[j] += sign * (dir[0] * (S[3] + wxp[0]) + dir[1] * (S[4] + wxp[1]) + dir[2] * (S[5] + wxp[2]));
  }
}
static double impedance(const double *solimp_in, double pos, double margin) {
  double si[5];
  memcpy(si, solimp_in, sizeof(si));
  si[0] = si[0] < MINIMP ? MINIMP : (si[0] > MAXIMP ? MAXIMP : si[0]);
  si[1] = si[1] < MINIMP ? MINIMP : (si[1] > MAXIMP ? MAXIMP : si[1]);
  si[2] = si[2] < 0 ? 0 : si[2];
  si[3] = si[3] < MINIMP ? MINIMP : (si[3] > MAXIMP ? MAXIMP : si[3]);
  si[4] = si[4] < 1 ? 1 : si[4];
  if (si[0] == si[1] || si[2] <= MINVAL) return 0.5 * (si[0] + si[1]);
  double x = (pos - margin) / si[2];
  if (x < 0) x = -x;
  if (x >= 1) return si[1];
  if (x <= 0) return si[0];
  double y;
  if (si[4] == 1) y = x;
  else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1);
  else y = 1 - pow(1 - x, si[4]) / pow(1 - si[3], si[4] - 1);
  return si[0] + y * (si[1] - si[0]);
}
static int add_row(orc_data *d, int nv, int type, int id) {
  int i = d->nefc++;
  memset(d->efc_J + (size_t)i * nv, 0, sizeof(double) * nv);
  d->efc_type[i] = type; d->efc_id[i] = id;
  d->efc_pos[i] = 0; d->efc_margin[i] = 0; d->efc_frictionloss[i] = 0;
  d->efc_t0[i] = d->efc_t1[i] = -1;
  return i;
}
void orc_make_constraint(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  int nv = m->nv;
  d->nefc = 0;
  /* friction loss */
  for (int j = 0; j < nv; j++) {
    if (m->dof_frictionloss[j] <= 0) continue;
    int i = add_row(d, nv, T_FRICTION, j);
    d->efc_J[(size_t)i * nv + j] = 1;
    d->efc_frictionloss[i] = m->dof_frictionloss[j];
    d->efc_t0[i] = o->dof_treeid[j];
  }
  /* joint limits */
  for (int j = 0; j < nv; j++) {
    if (!m->jnt_limited[j]) continue;
    double margin = m->jnt_margin[j];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - d->qpos[j]);
      if (dist < margin) {
        int i = add_row(d, nv, T_LIMIT, j);
        d->efc_J[(size_t)i * nv + j] = -side;
        d->efc_pos[i] = dist;
        d->efc_margin[i] = margin;
        d->efc_t0[i] = o->dof_treeid[j];
      }
    }
  }
  /* contacts, elliptic cones, condim 3: rows (normal, tangent1, tangent2); J = jac(body2)-jac(body1) */
  for (int c = 0; c < d->ncon; c++) {
    orc_contact *con = &d->con[c];
    con->efc_address = d->nefc;
    for (int r = 0; r < 3; r++) {
      int i = add_row(d, nv, T_CONTACT, c);
      double *J = d->efc_J + (size_t)i * nv;
      point_jac_row(o, d, con->body2, con->pos, con->frame + 3 * r, 1.0, J);
      point_jac_row(o, d, con->body1, con->pos, con->frame + 3 * r, -1.0, J);
      if (r == 0) { d->efc_pos[i] = con->dist; d->efc_margin[i] = con->includemargin; }
      int n1 = o->body_chain_len[con->body1], n2 = o->body_chain_len[con->body2];
      d->efc_t0[i] = n1 ? o->dof_treeid[o->body_chain[con->body1 * MAXDEPTH]] : -1;
      d->efc_t1[i] = n2 ? o->dof_treeid[o->body_chain[con->body2 * MAXDEPTH]] : -1;
    }
  }
  /* impedance, reference acceleration (mj_makeImpedance + mj_referenceConstraint) */
  int nefc = d->nefc;
  for (int i = 0; i < nefc; i++) {
    const double *solref, *solimp;
    double sr[2];
    int type = d->efc_type[i], id = d->efc_id[i], fric_row = 0;
    if (type == T_FRICTION) { solref = m->dof_solref + 2 * id; solimp = m->dof_solimp + 5 * id; d->efc_diagApprox[i] = m->dof_invweight0[id]; fric_row = 1; }
    else if (type == T_LIMIT) { solref = m->jnt_solref + 2 * id; solimp = m->jnt_solimp + 5 * id; d->efc_diagApprox[i] = m->dof_invweight0[id]; }
    else {
      orc_contact *con = &d->con[id];
      solref = con->solref; solimp = con->solimp;
      d->efc_diagApprox[i] = m->body_invweight0[2 * con->body1] + m->body_invweight0[2 * con->body2];
      fric_row = i > con->efc_address;
    }
    sr[0] = solref[0]; sr[1] = solref[1];
    if (sr[0] > 0 && sr[0] < 2 * m->timestep) sr[0] = 2 * m->timestep; /* refsafe */
    double imp = impedance(solimp, d->efc_pos[i], d->efc_margin[i]);
    double dmax = solimp[1] < MINIMP ? MINIMP : (solimp[1] > MAXIMP ? MAXIMP : solimp[1]);
    double K, B;
    if (sr[0] > 0) {
      double den = dmax * dmax * sr[0] * sr[0] * sr[1] * sr[1];
      K = 1 / (den > MINVAL ? den : MINVAL);
      den = dmax * sr[0];
      B = 2 / (den > MINVAL ? den : MINVAL);
    } else {
      K = -sr[0] / (dmax * dmax > MINVAL ? dmax * dmax : MINVAL);
      B = -sr[1] / (dmax > MINVAL ? dmax : MINVAL);
    }
    if (fric_row) K = 0;
    d->efc_K[i] = K; d->efc_Bd[i] = B; d->efc_imp[i] = imp;
    double R = (1 - imp) * d->efc_diagApprox[i] / imp;
    d->efc_R[i] = R > MINVAL ? R : MINVAL;
  }
  /* elliptic friction rows: R[t] = R[n]/impratio, scaled by mu^2/friction^2 */
  for (int c = 0; c < d->ncon; c++) {
    orc_contact *con = &d->con[c];
    int i = con->efc_address;
    double ir = m->impratio > MINVAL ? m->impratio : MINVAL;
    d->efc_R[i + 1] = d->efc_R[i] / ir;
    /* R[i+2] = R[i+1]*mu^2/friction_t2^2; condim 3 uses the geom sliding friction for both
     * tangents (mjContact.friction[0] == friction[1]), so the factor is 1 */
    d->efc_R[i + 2] = d->efc_R[i + 1];
    /* friction coefficient of the regularised cone (mj_makeImpedance: contact.mu = friction[0]*sqrt(R[i+1]/R[i])):
     * with R_t = R_n/impratio the primal map of mj_constraintUpdate works in the space where the cone is regular;
     * the dual solver (PGS) keeps using contact.friction */
    con->mu = con->friction[0] * sqrt(d->efc_R[i + 1] / d->efc_R[i]);
  }
  for (int i = 0; i < nefc; i++) {
    d->efc_D[i] = 1 / d->efc_R[i];
    double vel = 0;
    const double *J = d->efc_J + (size_t)i * nv;
    for (int j = 0; j < nv; j++) vel += J[j] * d->qvel[j];
    d->efc_vel[i] = vel;
    d->efc_aref[i] = -d->efc_Bd[i] * vel - d->efc_K[i] * d->efc_imp[i] * (d->efc_pos[i] - d->efc_margin[i]);
  }
}

/* ------------------------------- mj_projectConstraint -------------------------------------- */
static int rows_share_tree(const orc_data *d, int i, int k) {
  int a0 = d->efc_t0[i], a1 = d->efc_t1[i], b0 = d->efc_t0[k], b1 = d->efc_t1[k];
  return (a0 >= 0 && (a0 == b0 || a0 == b1)) || (a1 >= 0 && (a1 == b0 || a1 == b1));
}
void orc_project_constraint(const orc_model *o, orc_data *d) {
  int nv = o->m.nv, nefc = d->nefc, cap = d->efc_cap;
  for (int i = 0; i < nefc; i++) {
    double *B = d->efc_B + (size_t)i * nv;
    memset(B, 0, sizeof(double) * nv);
    int ts[2] = {d->efc_t0[i], d->efc_t1[i]};
    for (int q = 0; q < 2; q++) {
      int t = ts[q];
      if (t < 0 || (q == 1 && t == ts[0])) continue;
      int adr = o->tree_adr[t], n = o->tree_num[t];
      memcpy(B + adr, d->efc_J + (size_t)i * nv + adr, sizeof(double) * n);
      chol_solve_block(d->L, nv, adr, n, B);
    }
  }
  for (int i = 0; i < nefc; i++) {
    d->adj_n[i] = 0;
    for (int k = 0; k < nefc; k++) {
      double v = 0;
      if (rows_share_tree(d, i, k)) {
        int ts[2] = {d->efc_t0[k], d->efc_t1[k]};
        for (int q = 0; q < 2; q++) {
          int t = ts[q];
          if (t < 0 || (q == 1 && t == ts[0])) continue;
          int adr = o->tree_adr[t], n = o->tree_num[t];
          const double *J = d->efc_J + (size_t)k * nv + adr, *B = d->efc_B + (size_t)i * nv + adr;
          for (int j = 0; j < n; j++) v += J[j] * B[j];
          d->mac_count += n;
        }
        d->adj[(size_t)i * cap + d->adj_n[i]++] = k;
      }
      if (i == k) v += d->efc_R[i];
      d->efc_AR[(size_t)i * cap + k] = v;
    }
  }
}

/* ------------------------------------ actuation -------------------------------------------- */
void orc_actuation(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  int nv = m->nv;
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int a = 0; a < m->nu; a++) {
    double ctrl = d->ctrl[a];
    if (m->actuator_ctrllimited[a]) {
      double lo = m->actuator_ctrlrange[2 * a], hi = m->actuator_ctrlrange[2 * a + 1];
      ctrl = ctrl < lo ? lo : (ctrl > hi ? hi : ctrl);
    }
    const double *bp = m->actuator_biasprm + 3 * a;
    double force = m->actuator_gainprm[a] * ctrl + bp[0] + bp[1] * d->actuator_length[a] + bp[2] * d->actuator_velocity[a];
    if (m->actuator_forcelimited[a]) {
      double lo = m->actuator_forcerange[2 * a], hi = m->actuator_forcerange[2 * a + 1];
      force = force < lo ? lo : (force > hi ? hi : force);
    }
    d->actuator_force[a] = force;
    double g = m->actuator_gear[a];
    if (m->actuator_trntype[a] == 0) d->qfrc_actuator[m->actuator_trnid[a]] += g * force;
    else {
      int t = m->actuator_trnid[a];
      for (int w = 0; w < m->tendon_num[t]; w++)
        d->qfrc_actuator[m->wrap_dofid[m->tendon_adr[t] + w]] += g * m->wrap_coef[m->tendon_adr[t] + w] * force;
    }
  }
}

/* -------------------------------- constraint solver (PGS) ---------------------------------- */
/* mju_QCQP2 */
static int qcqp2(double *res, const double *Ain, const double *bin, const double *dd, double r) {
  double b1 = bin[0] * dd[0], b2 = bin[1] * dd[1];
  double A11 = Ain[0] * dd[0] * dd[0], A22 = Ain[3] * dd[1] * dd[1], A12 = Ain[1] * dd[0] * dd[1];
  double la = 0, v1 = 0, v2 = 0;
  for (int iter = 0; iter < 20; iter++) {
    double det = (A11 + la) * (A22 + la) - A12 * A12;
    if (det < 1e-10) { res[0] = 0; res[1] = 0; return 0; }
    double detinv = 1 / det, P11 = (A22 + la) * detinv, P22 = (A11 + la) * detinv, P12 = -A12 * detinv;
    v1 = -P11 * b1 - P12 * b2;
    v2 = -P12 * b1 - P22 * b2;
    double val = v1 * v1 + v2 * v2 - r * r;
    if (val < 1e-10) break;
    double deriv = -2 * (P11 * v1 * v1 + 2 * P12 * v1 * v2 + P22 * v2 * v2);
    double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  res[0] = v1 * dd[0];
  res[1] = v2 * dd[1];
  return la != 0;
}
/* primal force mapping used to turn qacc_warmstart into efc_force (mj_constraintUpdate) */
static void constraint_update(const orc_model *o, orc_data *d, const double *jar) {
  int nefc = d->nefc;
  for (int i = 0; i < nefc; i++) {
    int type = d->efc_type[i];
    if (type == T_FRICTION) {
      double f = d->efc_frictionloss[i], R = d->efc_R[i];
      if (jar[i] <= -R * f) d->efc_force[i] = f;
      else if (jar[i] >= R * f) d->efc_force[i] = -f;
      else d->efc_force[i] = -d->efc_D[i] * jar[i];
    } else if (type == T_LIMIT) {
      d->efc_force[i] = jar[i] < 0 ? -d->efc_D[i] * jar[i] : 0;
    } else {
      orc_contact *con = &d->con[d->efc_id[i]];
      if (i != con->efc_address) continue;
      double mu = con->mu, fr[2] = {con->friction[0], con->friction[0]};
      double U0 = jar[i] * mu, U1 = jar[i + 1] * fr[0], U2 = jar[i + 2] * fr[1];
      double N = U0, T = sqrt(U1 * U1 + U2 * U2);
      if ((N >= mu * T) || (T <= 0 && N >= 0)) {
        d->efc_force[i] = d->efc_force[i + 1] = d->efc_force[i + 2] = 0;
      } else if ((mu * N + T <= 0) || (T <= 0 && N < 0)) {
        for (int r = 0; r < 3; r++) d->efc_force[i + r] = -d->efc_D[i + r] * jar[i + r];
      } else {
        double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu));
        double NmT = N - mu * T;
        d->efc_force[i] = -Dm * NmT * mu;
        d->efc_force[i + 1] = -d->efc_force[i] / T * U1 * fr[0];
        d->efc_force[i + 2] = -d->efc_force[i] / T * U2 * fr[1];
      }
    }
  }
}
static double row_residual(orc_data *d, int i) {
  int cap = d->efc_cap;
  double res = d->efc_b[i];
  const double *AR = d->efc_AR + (size_t)i * cap;
  const int *adj = d->adj + (size_t)i * cap;
  int n = d->adj_n[i];
  for (int q = 0; q < n; q++) res += AR[adj[q]] * d->efc_force[adj[q]];
  d->mac_count += n;
  return res;
}
/* orc_solve = orc_solve_init (b, warm start, cost test) + orc_solve_iterate (the PGS sweeps); split so that the
 * tests can read the solver's starting point (efc_b, warm-start efc_force) before the sweeps overwrite it. */
void orc_solve_init(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  int nv = m->nv, nefc = d->nefc;
  double *jar = d->scr_jar;
  memset(jar, 0, sizeof(double) * ((size_t)nefc + 1));
  /* b = J*qacc_smooth - aref */
  for (int i = 0; i < nefc; i++) {
    const double *J = d->efc_J + (size_t)i * nv;
    double s = 0;
    for (int j = 0; j < nv; j++) s += J[j] * d->qacc_smooth[j];
    d->efc_b[i] = s - d->efc_aref[i];
  }
  /* warm start (mj_fwdConstraint): force from qacc_warmstart, kept only if its dual cost < 0 */
  for (int i = 0; i < nefc; i++) {
    const double *J = d->efc_J + (size_t)i * nv;
    double s = 0;
    for (int j = 0; j < nv; j++) s += J[j] * d->qacc_warmstart[j];
    jar[i] = s - d->efc_aref[i];
  }
  constraint_update(o, d, jar);
  {
    double cost = 0;
    for (int i = 0; i < nefc; i++) {
      double res = row_residual(d, i) - d->efc_b[i];
      cost += 0.5 * d->efc_force[i] * res + d->efc_force[i] * d->efc_b[i];
    }
    if (cost > 0) memset(d->efc_force, 0, sizeof(double) * nefc);
  }
}
void orc_solve_iterate(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  int nv = m->nv, nefc = d->nefc, cap = d->efc_cap;
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  int iter = 0;
  d->solver_improvement = 0;
  while (iter < m->iterations) {
    double improvement = 0;
    for (int i = 0; i < nefc;) {
      int type = d->efc_type[i];
      int dim = type == T_CONTACT ? 3 : 1;
      double res[3], old[3];
      for (int r = 0; r < dim; r++) { res[r] = row_residual(d, i + r); old[r] = d->efc_force[i + r]; }
      if (dim == 1) {
        double Aii = d->efc_AR[(size_t)i * cap + i];
        double f = d->efc_force[i] - res[0] / Aii;
        if (type == T_FRICTION) {
          double fl = d->efc_frictionloss[i];
          f = f < -fl ? -fl : (f > fl ? fl : f);
        } else if (f < 0) f = 0;
        d->efc_force[i] = f;
        double delta = f - old[0];
        double change = 0.5 * delta * delta * Aii + delta * res[0];
        if (change > 1e-10) { d->efc_force[i] = old[0]; change = 0; }
        improvement -= change;
      } else {
        orc_contact *con = &d->con[d->efc_id[i]];
        double fr[2] = {con->friction[0], con->friction[0]};
        double A[9];
        for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) A[3 * r + s] = d->efc_AR[(size_t)(i + r) * cap + i + s];
        double *f = d->efc_force + i;
        if (f[0] < MINVAL) {
          f[0] -= res[0] / A[0];
          if (f[0] < 0) f[0] = 0;
          f[1] = f[2] = 0;
        } else {
          double v[3] = {old[0], old[1], old[2]}, v1[3];
          for (int r = 0; r < 3; r++) v1[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
          double denom = dot3(v, v1);
          if (denom >= MINVAL) {
            double x = -dot3(v, res) / denom;
            if (f[0] + x * v[0] < 0) x = -v[0] / v[0];
            for (int r = 0; r < 3; r++) f[r] += x * v[r];
          }
        }
        if (f[0] >= MINVAL) {
          /* friction subproblem with the normal force fixed */
          double Ac[4] = {A[4], A[5], A[7], A[8]}, bc[2], vv[2];
          for (int r = 1; r < 3; r++) {
            bc[r - 1] = res[r];
            for (int s = 1; s < 3; s++) bc[r - 1] -= A[3 * r + s] * old[s];
            bc[r - 1] += A[3 * r] * (f[0] - old[0]);
          }
          int flg = qcqp2(vv, Ac, bc, fr, f[0]);
          if (flg) { /* rescale onto the cone exactly */
            double s = 0;
            for (int r = 0; r < 2; r++) s += vv[r] * vv[r] / (fr[r] * fr[r]);
            s = sqrt(f[0] * f[0] / (s > MINVAL ? s : MINVAL));
            vv[0] *= s; vv[1] *= s;
          }
          f[1] = vv[0]; f[2] = vv[1];
        }
        double dl[3] = {f[0] - old[0], f[1] - old[1], f[2] - old[2]}, Ad[3];
        for (int r = 0; r < 3; r++) Ad[r] = A[3 * r] * dl[0] + A[3 * r + 1] * dl[1] + A[3 * r + 2] * dl[2];
        double change = 0.5 * dot3(dl, Ad) + dot3(dl, res);
        if (change > 1e-10) { f[0] = old[0]; f[1] = old[1]; f[2] = old[2]; change = 0; }
        improvement -= change;
      }
      i += dim;
    }
    improvement *= scale;
    iter++;
    d->solver_improvement = improvement;
    if (improvement < m->tolerance) break;
  }
  d->solver_niter = iter;
  /* dual -> primal: qfrc_constraint = J^T f; qacc = qacc_smooth + Minv qfrc_constraint */
  memset(d->qfrc_constraint, 0, sizeof(double) * nv);
  for (int i = 0; i < nefc; i++) {
    const double *J = d->efc_J + (size_t)i * nv;
    double f = d->efc_force[i];
    if (f == 0) continue;
    for (int j = 0; j < nv; j++) d->qfrc_constraint[j] += J[j] * f;
  }
  memcpy(d->qacc, d->qfrc_constraint, sizeof(double) * nv);
  solve_M(o, d->L, nv, d->qacc);
  for (int j = 0; j < nv; j++) d->qacc[j] += d->qacc_smooth[j];
}
void orc_solve(const orc_model *o, orc_data *d) {
  orc_solve_init(o, d);
  orc_solve_iterate(o, d);
}

/* ---------------------------------------- steps -------------------------------------------- */
/* position + velocity dependent stages (mj_step1 = mj_fwdPosition + mj_fwdVelocity) */
void orc_step1(const orc_model *o, orc_data *d) {
  orc_kinematics(o, d);
  orc_crb(o, d);
  orc_collision(o, d);
  orc_make_constraint(o, d);
  orc_project_constraint(o, d);
  orc_velocity(o, d);
}
/* mj_step2 = mj_fwdActuation + mj_fwdAcceleration + mj_fwdConstraint + mj_Euler */
/* mj_fwdActuation + mj_fwdAcceleration: qfrc_smooth, qacc_smooth */
void orc_smooth(const orc_model *o, orc_data *d) {
  int nv = o->m.nv;
  orc_actuation(o, d);
  for (int j = 0; j < nv; j++) {
    d->qfrc_smooth[j] = d->qfrc_passive[j] - d->qfrc_bias[j] + d->qfrc_actuator[j] + d->qfrc_applied[j];
    d->qacc_smooth[j] = d->qfrc_smooth[j];
  }
  solve_M(o, d->L, nv, d->qacc_smooth);
}
void orc_step2(const orc_model *o, orc_data *d) {
  const rp_flat_model *m = &o->m;
  int nv = m->nv;
  double h = m->timestep;
  orc_smooth(o, d);
  orc_solve(o, d);
  /* mj_Euler with implicit joint damping: (M + h*diag(damping)) a = qfrc_smooth + qfrc_constraint */
  double *a = d->scr_a, *MH = d->scr_MH;
  memcpy(MH, d->M, sizeof(double) * nv * nv);
  for (int j = 0; j < nv; j++) { MH[j * nv + j] += h * m->dof_damping[j]; a[j] = d->qfrc_smooth[j] + d->qfrc_constraint[j]; }
  for (int t = 0; t < o->ntree; t++) chol_block(d->LH, MH, nv, o->tree_adr[t], o->tree_num[t]);
  solve_M(o, d->LH, nv, a);
  for (int j = 0; j < nv; j++) {
    d->qvel[j] += h * a[j];
    d->qpos[j] += h * d->qvel[j];
    d->qacc_warmstart[j] = d->qacc[j];
  }
}
void orc_reset(const orc_model *o, orc_data *d) {
  int nv = o->m.nv;
  memset(d->qpos, 0, sizeof(double) * nv); memset(d->qvel, 0, sizeof(double) * nv);
  memset(d->qacc, 0, sizeof(double) * nv); memset(d->qacc_warmstart, 0, sizeof(double) * nv);
  memset(d->ctrl, 0, sizeof(double) * o->m.nu); memset(d->qfrc_applied, 0, sizeof(double) * nv);
  memset(d->actuator_force, 0, sizeof(double) * o->m.nu);
  orc_step1(o, d);
}
/* One control step with dm_control's legacy stepping: nsub x (mj_step2; mj_step1). The data must
 * hold a valid step1 (after orc_reset / a previous call). */
void orc_control_step(const orc_model *o, orc_data *d, const double *ctrl, int nsub) {
  memcpy(d->ctrl, ctrl, sizeof(double) * o->m.nu);
  for (int s = 0; s < nsub; s++) { orc_step2(o, d); orc_step1(o, d); }
}

/* threaded rollout for CPU-baseline timing: nenv independent envs, nsteps control steps each,
 * ctrl [nsteps, nenv, nu] (or [nsteps, nu] broadcast if per_env == 0); nthreads POSIX threads pull
 * envs from a shared counter. */
#include <pthread.h>
typedef struct {
  const orc_model *o;
  int nenv, nsteps, per_env;
  const double *ctrl;
  double *qpos_out;
  int next;
  pthread_mutex_t mu;
} rollout_job;
static void *rollout_worker(void *arg) {
  rollout_job *job = (rollout_job *)arg;
  const orc_model *o = job->o;
  int nu = o->m.nu, nv = o->m.nv, nsub = o->m.nsubstep;
  orc_data *d = orc_data_create(o);
  for (;;) {
    pthread_mutex_lock(&job->mu);
    int e = job->next++;
    pthread_mutex_unlock(&job->mu);
    if (e >= job->nenv) break;
    orc_reset(o, d);
    for (int t = 0; t < job->nsteps; t++) {
      const double *c = job->ctrl + ((size_t)t * (job->per_env ? job->nenv : 1) + (job->per_env ? e : 0)) * nu;
      orc_control_step(o, d, c, nsub);
    }
    if (job->qpos_out) memcpy(job->qpos_out + (size_t)e * nv, d->qpos, sizeof(double) * nv);
  }
  orc_data_destroy(d);
  return NULL;
}
int orc_rollout(const orc_model *o, int nenv, int nsteps, const double *ctrl, int per_env, double *qpos_out,
                int nthreads) {
  rollout_job job = {o, nenv, nsteps, per_env, ctrl, qpos_out, 0};
  pthread_mutex_init(&job.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, rollout_worker, &job);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  pthread_mutex_destroy(&job.mu);
  return 0;
}

/* Persistent pool for the CPU-baseline timing: one orc_data workspace per host thread, allocated (and touched) once,
 * so that a timed call does no allocation at all — the r01 timing created 128 workspaces of ~3 MB per call and measured
 * page faults and mmap contention as much as the solver (3.4k vs 20.9k env-steps/s on two identical boxes).  The
 * environments' states live in a caller-owned array [nenv][3 nv + nu] (qpos, qvel, qacc_warmstart, ctrl); a thread
 * loads a state, recomputes the position / velocity stage (orc_step1, what a running environment already holds),
 * runs `nsteps` control steps and stores the state back, so a rollout can be continued call after call. */
typedef struct orc_pool {
  const orc_model *o;
  int nthreads;
  orc_data **d;
} orc_pool;
typedef struct {
  orc_pool *pool;
  int nenv, nsteps, per_env, reset, tid;
  const double *ctrl;
  double *state;
  int *next;
  pthread_mutex_t *mu;
  long macs;
} pool_job;
orc_pool *orc_pool_create(const orc_model *o, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  orc_pool *p = (orc_pool *)calloc(1, sizeof(orc_pool));
  p->o = o; p->nthreads = nthreads;
  p->d = (orc_data **)calloc(nthreads, sizeof(orc_data *));
  for (int i = 0; i < nthreads; i++) { p->d[i] = orc_data_create(o); orc_reset(o, p->d[i]); }
  return p;
}
void orc_pool_destroy(orc_pool *p) {
  for (int i = 0; i < p->nthreads; i++) orc_data_destroy(p->d[i]);
  free(p->d); free(p);
}
static void *pool_worker(void *arg) {
  pool_job *job = (pool_job *)arg;
  const orc_model *o = job->pool->o;
  orc_data *d = job->pool->d[job->tid];
  int nv = o->m.nv, nu = o->m.nu, nsub = o->m.nsubstep, stride = 3 * nv + nu;
  long mac0 = d->mac_count;
  for (;;) {
    pthread_mutex_lock(job->mu);
    int e = (*job->next)++;
    pthread_mutex_unlock(job->mu);
    if (e >= job->nenv) break;
    double *st = job->state + (size_t)e * stride;
    if (job->reset) orc_reset(o, d);
    else {
      memcpy(d->qpos, st, sizeof(double) * nv); memcpy(d->qvel, st + nv, sizeof(double) * nv);
      memcpy(d->qacc_warmstart, st + 2 * nv, sizeof(double) * nv); memcpy(d->ctrl, st + 3 * nv, sizeof(double) * nu);
      memset(d->qfrc_applied, 0, sizeof(double) * nv);
      orc_step1(o, d);
    }
    for (int t = 0; t < job->nsteps; t++) {
      const double *c = job->ctrl + ((size_t)t * (job->per_env ? job->nenv : 1) + (job->per_env ? e : 0)) * nu;
      orc_control_step(o, d, c, nsub);
    }
    memcpy(st, d->qpos, sizeof(double) * nv); memcpy(st + nv, d->qvel, sizeof(double) * nv);
    memcpy(st + 2 * nv, d->qacc_warmstart, sizeof(double) * nv); memcpy(st + 3 * nv, d->ctrl, sizeof(double) * nu);
  }
  job->macs = d->mac_count - mac0;
  return NULL;
}
/* returns the multiply-adds the solver counted during the call (projection + Gauss-Seidel residuals) */
long orc_pool_rollout(orc_pool *p, int nenv, int nsteps, const double *ctrl, int per_env, double *state, int reset, int nthreads) {
  if (nthreads < 1 || nthreads > p->nthreads) nthreads = p->nthreads;
  pthread_mutex_t mu;
  pthread_mutex_init(&mu, NULL);
  int next = 0;
  pool_job *jobs = (pool_job *)calloc(nthreads, sizeof(pool_job));
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  for (int i = 0; i < nthreads; i++) {
    pool_job j = {p, nenv, nsteps, per_env, reset, i, ctrl, state, &next, &mu, 0};
    jobs[i] = j;
    pthread_create(&th[i], NULL, pool_worker, &jobs[i]);
  }
  long macs = 0;
  for (int i = 0; i < nthreads; i++) { pthread_join(th[i], NULL); macs += jobs[i].macs; }
  pthread_mutex_destroy(&mu);
  free(jobs); free(th);
  return macs;
}

/* accessors for ctypes */
double *orc_ptr(orc_data *d, const char *name) {
#define F(n) if (!strcmp(name, #n)) return d->n;
  F(qpos) F(qvel) F(qacc) F(qacc_warmstart) F(ctrl) F(qfrc_applied) F(xpos) F(xquat) F(xmat) F(xipos) F(ximat)
  F(xanchor) F(xaxis) F(geom_xpos) F(geom_xmat) F(site_xpos) F(S) F(M) F(qfrc_bias) F(qfrc_passive)
  F(qfrc_actuator) F(qfrc_smooth) F(qacc_smooth) F(qfrc_constraint) F(actuator_length) F(actuator_velocity)
  F(actuator_force) F(efc_J) F(efc_pos) F(efc_R) F(efc_D) F(efc_aref) F(efc_b) F(efc_force) F(efc_AR)
  F(efc_vel) F(efc_imp) F(efc_K) F(efc_Bd) F(efc_diagApprox) F(efc_frictionloss)
#undef F
  return NULL;
}
int orc_int(orc_data *d, const char *name) {
  if (!strcmp(name, "ncon")) return d->ncon;
  if (!strcmp(name, "nefc")) return d->nefc;
  if (!strcmp(name, "efc_cap")) return d->efc_cap;
  if (!strcmp(name, "solver_niter")) return d->solver_niter;
  if (!strcmp(name, "con_overflow")) return d->con_overflow;
  return -1;
}
int *orc_iptr(orc_data *d, const char *name) {
  if (!strcmp(name, "efc_type")) return d->efc_type;
  if (!strcmp(name, "efc_id")) return d->efc_id;
  return NULL;
}
/* contact i -> out[0..] = dist, pos(3), frame(9), geom1, geom2, key, mu (regularised cone), friction[0] */
void orc_contact_get(orc_data *d, int i, double *out) {
  orc_contact *c = &d->con[i];
  out[0] = c->dist;
  memcpy(out + 1, c->pos, 3 * sizeof(double));
  memcpy(out + 4, c->frame, 9 * sizeof(double));
  out[13] = c->geom1; out[14] = c->geom2; out[15] = c->key; out[16] = c->mu; out[17] = c->friction[0];
}
long orc_mac_count(orc_data *d) { return d->mac_count; }
