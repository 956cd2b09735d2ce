// rp_kernels.cu — sm_100a step kernel: one warp per environment, fp64.
//
// Replaces the reference's per-control-step host loop
//   before_step -> 10 x physics.step() -> after_step -> observables -> reward -> termination
// (robopianist/suite/tasks/piano_with_shadow_hands.py:176-220, tasks/base.py:28-31; the physics is
// MuJoCo's mj_step, restated stage by stage) with a single launch.  Lanes map to the 26 links of a
// hand (both hands in lock-step, two independent dependency chains per warp) or to keys
// (lane l owns keys l, l+32, l+64).  Shared memory holds the per-hand inverse inertia, the joint
// motion subspaces and the contact blocks; everything else lives in registers.
#include <cuda_runtime.h>
#include <stdint.h>

#include "rp_kernels.h"
#include "rp_math.cuh"

#define FULL 0xffffffffu
#define NL RP_NL
#define NK RP_NK
#define MAXD RP_MAXD
#define MAXCON RP_MAXCON

namespace {

struct Contact {  // lives in shared memory
  double Jc[3][MAXD];  // hand-chain Jacobian entries by depth (already signed: body2 - body1)
  double Jk[3];        // key dof entry
  double A[9];         // 3x3 block of AR (R on the diagonal)
  double R[3], b[3], f[3], aref[3];
  double mu, invMk;
  int hand, link, key, dep;  // key = -1: piano base (static)
  signed char chain[16];     // dof (lane) at depth d of the hand chain
};
struct CKey {  // limit row of a key that is coupled to a hand through a contact
  double J, R, b, AR, ARi, f, invM;
  int key, active;
};
struct RawContact {  // narrowphase output before sorting
  double dist, pos[3], n[3];
  int hand, geom, key, sub;
  unsigned sortkey, pad;
};

#define RS 27  // row stride of the hand matrices (odd: lane-per-row accesses are bank-conflict free)
struct Smem {
  double A[RP_NHMAX][NL * RS];  // per hand: M, then -M^-1 (sweep), later -(M + h*D)^-1; symmetric, row-major
  double S[RP_NHMAX][NL][6];
  double qk[NK], vk[NK], vak[NK];
  double wbuf[RP_NHMAX][32];
  double Mrow[RP_NHMAX][NL][MAXD];   // mass-matrix entries M[link][ancestor at depth d] (kept for the Euler solve)
  signed char anc[RP_NHMAX][NL][MAXD];  // ancestor-or-self at depth d (copy of the model table, for the sweep)
  signed char dep[RP_NHMAX][32], sub[RP_NHMAX][32];  // link depth / subtree size
  double actf[RP_NHMAX][RP_NAH];
  double vh[RP_NHMAX][NL];  // hand joint velocities (for chain walks)
  double kfrc[NK], kqs[NK], kfc[NK];  // per key: qfrc_smooth, qacc_smooth, constraint force J^T f
  double stash[RP_NHMAX][NL][2];      // qfrc_smooth, qacc_smooth of the hand dofs (parked during PGS)
  CKey ck[MAXCON];                    // limit rows of keys that also carry contacts
  signed char keyslot[NK];            // index into ck[] (-1: key not in contact)
  int nck, pad2;
  union {
    double acc[RP_NHMAX][NL][16];  // upward accumulation scratch: F (6) + composite inertia (10)
    struct {
      Contact con[MAXCON];
      RawContact raw[MAXCON];
    } c;
  } u;
  int ncon, nraw, overflow, pad;
};

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ void rotq(double* r, const double* q, const double* v) {
  double m[9];
  rp::quat2mat(m, q);
  rp::mat_vec(r, m, v);
}

// rotation of vector v about unit axis by angle (Rodrigues)
__device__ __forceinline__ void rot_axis(double* r, const double* axis, double s, double c, const double* v) {
  double axv[3];
  rp::cross3(axv, axis, v);
  double ad = rp::dot3(axis, v) * (1 - c);
  for (int k = 0; k < 3; k++) r[k] = v[k] * c + axv[k] * s + axis[k] * ad;
}

// ------------------------------------------------------------------------------------------------
// key helpers (one key, scalar)
// ------------------------------------------------------------------------------------------------
struct KeyRow {       // active limit row of a key (at most one side can be active)
  double J, R, b, aref, ARinv, AR, f;
  int active;
};

}  // namespace

// In-place symmetric sweep of the NH hand matrices in sm.A: afterwards A = -M^-1 (callers fold the
// sign into their multiply-adds).  Lane j owns row j (== column j; the sweep keeps the matrix
// symmetric).  Pivots run leaves first (k = NL-1 .. 0): at that point row k is non-zero only at the
// ancestors of k and inside its subtree, so a pivot touches depth_k + subtree_k rows instead of 26
// (290 row updates per hand instead of 676), and the loops stay rolled: the code is a few hundred
// bytes, not the 230 KB of a register-resident unrolled elimination (instruction-cache bound).
template <int NH>
__device__ __noinline__ void sweep_inverse(Smem& sm, int lane) {
#pragma unroll 1
  for (int h = 0; h < NH; h++) {
    double* A = sm.A[h];
    double* myrow = A + (lane < NL ? lane : 0) * RS;
#pragma unroll 1
    for (int k = NL - 1; k >= 0; k--) {
      const double* rowk = A + k * RS;
      const double dinv = 1.0 / rowk[k];
      const bool upd = lane < NL && lane != k;  // lane k's own row is rewritten after the updates
      const double t = upd ? rowk[lane] * dinv : 0.0;
      const int depk = sm.dep[h][k], endk = k + sm.sub[h][k];
      if (upd) {
        for (int d = 0; d < depk; d++) {
          const int i = sm.anc[h][k][d];
          myrow[i] -= rowk[i] * t;
        }
#pragma unroll 2
        for (int i = k + 1; i < endk; i++) myrow[i] -= rowk[i] * t;
      }
      __syncwarp();
      if (upd) { myrow[k] = t; A[k * RS + lane] = t; }
      if (lane == k) A[k * RS + k] = -dinv;
      __syncwarp();
    }
  }
}

// lane == dof values of hand 0 / hand 1 -> half-warp mapping (both shuffles run on the full warp)
__device__ __forceinline__ double to_new(double x0, double x1, int src, int hh) {
  const double v0 = __shfl_sync(FULL, x0, src), v1 = __shfl_sync(FULL, x1, src);
  return hh ? v1 : v0;
}

// Scalar constraint rows of the two hands in the half-warp mapping (see the PGS section).
struct HandRows { double f[2], b[2], R[2], AR[2], ARi[2], J[2], lo[2], hi[2]; };

// One Gauss-Seidel pass over the scalar rows `anyrow` (bit i: dof i has a row in either hand).
// Ah: this lane's hand matrix (A = -M^-1), l0/l1: the dofs owned by the lane.  Returns the lane's
// share of the cost improvement.  Every lane evaluates the candidate update of its own row; only
// the owner of row i commits at step i and broadcasts J*delta inside its half-warp.
__device__ __forceinline__ double gs_rows(const double* __restrict__ Ah, int lane, int l0, int l1, unsigned anyrow,
                                          HandRows& r, double (&va)[2]) {
  double impr = 0;
  const int hb = lane & 16;
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const double base = r.b[s] + r.R[s] * r.f[s];  // f only changes at the lane's own commit
    const int i0 = 16 * s, i1 = s ? NL : 16, own = s ? l1 : l0;
#pragma unroll 1
    for (int i = i0; i < i1; i++) {
      if (!(anyrow >> i & 1)) continue;
      const double m0 = Ah[i * RS + l0], m1 = Ah[i * RS + l1];
      const double res = base + r.J[s] * va[s];
      const double x = r.f[s] - res * r.ARi[s];
      double fx = x < r.lo[s] ? r.lo[s] : x;
      fx = x > r.hi[s] ? r.hi[s] : fx;
      const double dl = fx - r.f[s];
      const double dbc = __shfl_sync(FULL, r.J[s] * dl, hb | (i & 15));
      va[0] -= m0 * dbc;
      va[1] -= m1 * dbc;
      if (own == i) { impr -= dl * (0.5 * dl * r.AR[s] + res); r.f[s] = fx; }
    }
  }
  return impr;
}

// The kernel proper is large; it is written as one function so that lane-owned state stays in
// registers across the stages of a substep.
template <int NH>
__global__ void __launch_bounds__(32) rp_step_kernel(RpKernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const RpDevModel* __restrict__ dm = a.dm;
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));  // opaque: keeps the compiler from re-reading S2R in loops
  const int env = blockIdx.x;
  if (env >= a.nenv) return;
  const int nvtot = NK + NH * NL;
  const int nu = dm->nu;
  double* __restrict__ gq = a.qpos + (size_t)env * nvtot;
  double* __restrict__ gv = a.qvel + (size_t)env * nvtot;
  double* __restrict__ gw = a.qacc_ws + (size_t)env * nvtot;
  double* __restrict__ gctrl = a.ctrl + (size_t)env * nu;
  const bool hl = lane < NL;  // lane owns a hand link

  // ------------------------------------------------------------------ task prologue (full step)
  int t_idx = 0, prev_type = RP_STEP_MID_;
  bool do_reset = false;
  if (a.mode == RP_MODE_STEP) {
    t_idx = a.t_idx[env];
    prev_type = a.step_type_state[env];
    do_reset = (prev_type == RP_STEP_LAST_);
  } else if (a.mode == RP_MODE_RESET) {
    do_reset = a.reset_mask == nullptr ? true : (a.reset_mask[env] != 0);
    if (!do_reset) return;
  }
  if (do_reset) {
    for (int j = lane; j < nvtot; j += 32) { gq[j] = 0; gv[j] = 0; gw[j] = 0; }
    for (int j = lane; j < nu; j += 32) gctrl[j] = 0;
    if (lane == 0) { a.t_idx[env] = 0; a.sustain[env] = 0; }
    t_idx = 0;
    __syncwarp();
  }
  const int act_dim = dm->act_dim;
  if (a.mode == RP_MODE_STEP && !do_reset) {
    // before_step: ctrl <- action[:-1] verbatim (clamping happens in the actuator model), sustain <- action[-1]
    const double* act = a.action + (size_t)env * act_dim;
    for (int j = lane; j < nu; j += 32) gctrl[j] = act[j];
    if (lane == 0) a.sustain[env] = act[act_dim - 1];
    __syncwarp();
  } else if (a.mode == RP_MODE_PHYSICS && a.ctrl_in != nullptr) {
    const double* c = a.ctrl_in + (size_t)env * nu;
    for (int j = lane; j < nu; j += 32) gctrl[j] = c[j];
    __syncwarp();
  }

  // ------------------------------------------------------------------ load state
  double qk[3], vk[3], wk[3];  // keys lane, lane+32, lane+64
#pragma unroll
  for (int s = 0; s < 3; s++) {
    int k = lane + 32 * s;
    bool ok = k < NK;
    qk[s] = ok ? gq[k] : 0.0;
    vk[s] = ok ? gv[k] : 0.0;
    wk[s] = ok ? gw[k] : 0.0;
  }
  double qh[NH > 0 ? NH : 1], vh[NH > 0 ? NH : 1], wh[NH > 0 ? NH : 1];
#pragma unroll
  for (int h = 0; h < NH; h++) {
    qh[h] = hl ? gq[NK + h * NL + lane] : 0.0;
    vh[h] = hl ? gv[NK + h * NL + lane] : 0.0;
    wh[h] = hl ? gw[NK + h * NL + lane] : 0.0;
  }
  // per-lane link constants
  int lpar[NH > 0 ? NH : 1], ldep[NH > 0 ? NH : 1], ljt[NH > 0 ? NH : 1];
#pragma unroll
  for (int h = 0; h < NH; h++) {
    lpar[h] = hl ? dm->link_parent[h][lane] : -1;
    ldep[h] = hl ? dm->link_depth[h][lane] : -1;
    ljt[h] = hl ? dm->link_jtype[h][lane] : 0;
  }
  int lsub[NH > 0 ? NH : 1], ldc[NH > 0 ? NH : 1];  // subtree size, depth clamped to >= 0
#pragma unroll
  for (int h = 0; h < NH; h++) { lsub[h] = hl ? dm->link_subtree[h][lane] : 0; ldc[h] = ldep[h] > 0 ? ldep[h] : 0; }
  const double hstep = dm->h;
  const double g3[3] = {dm->gravity[0], dm->gravity[1], dm->gravity[2]};
  const int nsub = (a.mode == RP_MODE_RESET || do_reset) ? 0 : a.nsub;
  double actforce[NH > 0 ? NH : 1];  // lane = actuator index within the hand
#pragma unroll
  for (int h = 0; h < NH; h++) actforce[h] = 0;
  double key_actforce[3] = {0, 0, 0};
  int niter_last = 0, ncon_last = 0, flags = 0;

  // topology tables into shared memory (read in the inner loops of the sweep / contact rows)
  if (NH > 0) {
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (hl) {
        for (int d = 0; d < MAXD; d++) sm.anc[h][lane][d] = (signed char)dm->link_anc[h][lane][d];
      }
      sm.dep[h][lane] = (signed char)ldep[h];
      sm.sub[h][lane] = (signed char)lsub[h];
    }
    __syncwarp();
  }
  // link world poses (kept for the task epilogue)
  double xp[NH > 0 ? NH : 1][3], xq[NH > 0 ? NH : 1][4];
#ifdef RP_PROFILE
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#define TICK(i) { long long tn = clock64(); tacc[i] += tn - tlast; tlast = tn; }
#else
#define TICK(i)
#endif

  // =============================================================================================
  for (int sub = 0; sub <= nsub; sub++) {
    const bool last_pass = (sub == nsub);  // kinematics-only pass for the task epilogue / reset
    if (last_pass && a.mode == RP_MODE_PHYSICS) break;
    // ---------------------------------------------------------------- hand kinematics (by level)
#pragma unroll
    for (int h = 0; h < NH; h++) {
      xp[h][0] = xp[h][1] = xp[h][2] = 0;
      xq[h][0] = 1; xq[h][1] = xq[h][2] = xq[h][3] = 0;
    }
    double xanchor[NH > 0 ? NH : 1][3], xaxis[NH > 0 ? NH : 1][3];
    if (NH > 0) {
      int maxd = max(dm->maxdepth[0], dm->maxdepth[NH > 1 ? 1 : 0]);
      for (int d = 0; d <= maxd; d++) {
#pragma unroll
        for (int h = 0; h < NH; h++) {
          int src = lpar[h] >= 0 ? lpar[h] : lane;
          double pp[3], pq[4];
#pragma unroll
          for (int k = 0; k < 3; k++) pp[k] = shfl(xp[h][k], src);
#pragma unroll
          for (int k = 0; k < 4; k++) pq[k] = shfl(xq[h][k], src);
          if (hl && ldep[h] == d) {
            if (lpar[h] < 0) { pp[0] = pp[1] = pp[2] = 0; pq[0] = 1; pq[1] = pq[2] = pq[3] = 0; }
            double lp[3] = {dm->link_pos[h][lane][0], dm->link_pos[h][lane][1], dm->link_pos[h][lane][2]};
            double lq[4] = {dm->link_quat[h][lane][0], dm->link_quat[h][lane][1], dm->link_quat[h][lane][2], dm->link_quat[h][lane][3]};
            double jp[3] = {dm->link_jpos[h][lane][0], dm->link_jpos[h][lane][1], dm->link_jpos[h][lane][2]};
            double ja[3] = {dm->link_jaxis[h][lane][0], dm->link_jaxis[h][lane][1], dm->link_jaxis[h][lane][2]};
            double pos[3], quat[4], tmp[3];
            rotq(tmp, pq, lp);
            for (int k = 0; k < 3; k++) pos[k] = pp[k] + tmp[k];
            rp::quat_mul(quat, pq, lq);
            rotq(tmp, quat, jp);
            for (int k = 0; k < 3; k++) xanchor[h][k] = pos[k] + tmp[k];
            rotq(xaxis[h], quat, ja);
            if (ljt[h] == 2) {
              for (int k = 0; k < 3; k++) pos[k] += xaxis[h][k] * qh[h];
            } else {
              double sn, cs;
              sincos(0.5 * qh[h], &sn, &cs);
              double ql[4] = {cs, ja[0] * sn, ja[1] * sn, ja[2] * sn}, qn[4];
              rp::quat_mul(qn, quat, ql);
              for (int k = 0; k < 4; k++) quat[k] = qn[k];
              rotq(tmp, quat, jp);
              for (int k = 0; k < 3; k++) pos[k] = xanchor[h][k] - tmp[k];
            }
            rp::quat_normalize(quat);
            for (int k = 0; k < 3; k++) xp[h][k] = pos[k];
            for (int k = 0; k < 4; k++) xq[h][k] = quat[k];
          }
        }
      }
    }
    TICK(0)
    if (last_pass) break;

    // ------------------------------------------------------------------ publish S, velocities
    double Sown[NH > 0 ? NH : 1][6];
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (hl) {
        if (ljt[h] == 3) {
          for (int k = 0; k < 3; k++) Sown[h][k] = xaxis[h][k];
          rp::cross3(&Sown[h][3], xanchor[h], xaxis[h]);
        } else {
          for (int k = 0; k < 3; k++) { Sown[h][k] = 0; Sown[h][3 + k] = xaxis[h][k]; }
        }
        for (int k = 0; k < 6; k++) sm.S[h][lane][k] = Sown[h][k];
        sm.vh[h][lane] = vh[h];
      }
    }
#pragma unroll
    for (int s = 0; s < 3; s++) {
      int k = lane + 32 * s;
      if (k < NK) { sm.qk[k] = qk[s]; sm.vk[k] = vk[s]; sm.vak[k] = 0; }
    }
    __syncwarp();

    // --------------------------------------------------- RNE forward (chain walk), inertia, F
    double qfrc_bias[NH > 0 ? NH : 1];
#pragma unroll
    for (int h = 0; h < NH; h++) {
      qfrc_bias[h] = 0;
      if (hl) {
        double V[6] = {0, 0, 0, 0, 0, 0}, A[6] = {0, 0, 0, 0, 0, 0};
        for (int d = 0; d <= ldep[h]; d++) {
          int j = sm.anc[h][lane][d];
          double Sj[6], Sd[6];
          for (int k = 0; k < 6; k++) Sj[k] = sm.S[h][j][k];
          double qd = sm.vh[h][j];
          rp::cross_motion(Sd, V, Sj);
          for (int k = 0; k < 6; k++) { A[k] += Sd[k] * qd; V[k] += Sj[k] * qd; }
        }
        double mass = dm->link_mass[h][lane];
        double xm[9], tmp[3], c[3], iq[4], R[9], I[10];
        rp::quat2mat(xm, xq[h]);
        double ip[3] = {dm->link_ipos[h][lane][0], dm->link_ipos[h][lane][1], dm->link_ipos[h][lane][2]};
        rp::mat_vec(tmp, xm, ip);
        for (int k = 0; k < 3; k++) c[k] = xp[h][k] + tmp[k];
        double liq[4] = {dm->link_iquat[h][lane][0], dm->link_iquat[h][lane][1], dm->link_iquat[h][lane][2], dm->link_iquat[h][lane][3]};
        rp::quat_mul(iq, xq[h], liq);
        rp::quat2mat(R, iq);
        double di[3] = {dm->link_inertia[h][lane][0], dm->link_inertia[h][lane][1], dm->link_inertia[h][lane][2]};
        rp::body_inertia_world(I, mass, di, R, c);
        // F = I*A + V x* (I*V) + (1-gravcomp) * I*[0;-g]
        double IA[6], IV[6], VxIV[6];
        rp::inertia_mul(IA, I, A);
        rp::inertia_mul(IV, I, V);
        rp::cross_force(VxIV, V, IV);
        double gc = 1.0 - dm->link_gravcomp[h][lane];
        double Ag[6] = {0, 0, 0, -g3[0] * gc, -g3[1] * gc, -g3[2] * gc}, IAg[6];
        rp::inertia_mul(IAg, I, Ag);
        for (int k = 0; k < 6; k++) sm.u.acc[h][lane][k] = IA[k] + VxIV[k] + IAg[k];
        for (int k = 0; k < 10; k++) sm.u.acc[h][lane][6 + k] = I[k];
      }
    }
    __syncwarp();
    // upward accumulation by level
    if (NH > 0) {
      int maxd = max(dm->maxdepth[0], dm->maxdepth[NH > 1 ? 1 : 0]);
      for (int d = maxd; d >= 1; d--) {
        int nslots = max(dm->link_nchild_slots[0][d], dm->link_nchild_slots[NH > 1 ? 1 : 0][d]);
        for (int r = 0; r < nslots; r++) {
#pragma unroll
          for (int h = 0; h < NH; h++) {
            if (hl && ldep[h] == d && dm->link_childrank[h][lane] == r) {
              int p = lpar[h];
              for (int k = 0; k < 16; k++) sm.u.acc[h][p][k] += sm.u.acc[h][lane][k];
            }
          }
          __syncwarp();
        }
      }
    }
    // bias force, mass-matrix rows
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (hl) {
        double F[6], Ic[10], fI[6];
        for (int k = 0; k < 6; k++) F[k] = sm.u.acc[h][lane][k];
        for (int k = 0; k < 10; k++) Ic[k] = sm.u.acc[h][lane][6 + k];
        qfrc_bias[h] = rp::dot6(Sown[h], F);
        rp::inertia_mul(fI, Ic, Sown[h]);
        for (int d = 0; d <= ldep[h]; d++) {
          const int j = sm.anc[h][lane][d];
          double Sj[6];
          for (int k = 0; k < 6; k++) Sj[k] = sm.S[h][j][k];
          double m = rp::dot6(Sj, fI);
          if (d == ldep[h]) m += dm->link_armature[h][lane];
          sm.Mrow[h][lane][d] = m;
        }
      }
    }
    // scatter M into the dense symmetric work matrix
    if (NH > 0) {
      double* Aflat = &sm.A[0][0];
      for (int idx = lane; idx < NH * NL * RS; idx += 32) Aflat[idx] = 0;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (hl) {
        for (int d = 0; d <= ldep[h]; d++) {
          const int j = sm.anc[h][lane][d];
          const double m = sm.Mrow[h][lane][d];
          sm.A[h][lane * RS + j] = m;
          sm.A[h][j * RS + lane] = m;
        }
      }
    }
    __syncwarp();
    if (a.dbg && sub == 0 && NH > 0) {
      double* dbg = a.dbg + (size_t)env * RP_DBG_STRIDE;
      for (int i = 0; i < NL; i++) if (hl) dbg[RP_DBG_M + i * NL + lane] = sm.A[0][i * RS + lane];
    }
    TICK(1)
    // ---------------------------------------------------------------- invert M (symmetric sweep): A = -M^-1
    if (NH > 0) sweep_inverse<NH>(sm, lane);
    if (a.dbg && sub == 0 && NH > 0) {
      double* dbg = a.dbg + (size_t)env * RP_DBG_STRIDE;
      for (int i = 0; i < NL; i++) if (hl) dbg[RP_DBG_MINV + i * NL + lane] = -sm.A[0][i * RS + lane];
    }

    TICK(2)
    // ------------------------------------------------------------------ actuation (lane = actuator)
#pragma unroll
    for (int h = 0; h < NH; h++) {
      bool isact = lane < dm->nact[h];
      int d0 = isact ? dm->act_dof0[h][lane] : 0, d1 = isact ? dm->act_dof1[h][lane] : -1;
      double q0 = shfl(qh[h], d0), v0 = shfl(vh[h], d0);
      double q1 = shfl(qh[h], d1 >= 0 ? d1 : 0), v1 = shfl(vh[h], d1 >= 0 ? d1 : 0);
      double force = 0;
      if (isact) {
        double c0 = dm->act_coef0[h][lane], c1 = d1 >= 0 ? dm->act_coef1[h][lane] : 0.0;
        double len = c0 * q0 + c1 * q1, vel = c0 * v0 + c1 * v1;
        double ctrl = gctrl[h * dm->nact[0] + lane];
        if (dm->act_ctrllimited[h][lane]) ctrl = fmin(fmax(ctrl, dm->act_ctrlrange[h][lane][0]), dm->act_ctrlrange[h][lane][1]);
        force = dm->act_gain[h][lane] * ctrl + dm->act_bias[h][lane][0] + dm->act_bias[h][lane][1] * len + dm->act_bias[h][lane][2] * vel;
        if (dm->act_forcelimited[h][lane]) force = fmin(fmax(force, dm->act_forcerange[h][lane][0]), dm->act_forcerange[h][lane][1]);
        sm.actf[h][lane] = force;
      }
      actforce[h] = force;
    }
    __syncwarp();
    // ---------------------------------------------------- smooth forces and accelerations (hands)
    double qfrc_smooth[NH > 0 ? NH : 1], qacc_smooth[NH > 0 ? NH : 1];
#pragma unroll
    for (int h = 0; h < NH; h++) {
      qfrc_smooth[h] = 0;
      if (hl) {
        double passive = -dm->link_stiffness[h][lane] * (qh[h] - dm->link_springref[h][lane]) - dm->link_damping[h][lane] * vh[h];
        int ai = dm->dof_act[h][lane];
        double qa = ai >= 0 ? dm->dof_act_coef[h][lane] * sm.actf[h][ai] : 0.0;
        qfrc_smooth[h] = passive - qfrc_bias[h] + qa;
        sm.wbuf[h][lane] = qfrc_smooth[h];
      }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < NH; h++) {
      double s = 0;
      if (hl) for (int k = 0; k < NL; k++) s -= sm.A[h][k * RS + lane] * sm.wbuf[h][k];
      qacc_smooth[h] = s;
    }
    __syncwarp();
    // ------------------------------------------------------------------ keys: smooth dynamics + limit rows
    double ks_qfrc[3], ks_qacc[3], kM[3];
    KeyRow krow[3];
#pragma unroll
    for (int s = 0; s < 3; s++) {
      int k = lane + 32 * s;
      krow[s].active = 0; krow[s].f = 0; krow[s].J = 0; krow[s].R = 0; krow[s].b = 0; krow[s].ARinv = 0; krow[s].AR = 1; krow[s].aref = 0;
      ks_qfrc[s] = 0; ks_qacc[s] = 0; kM[s] = 1; key_actforce[s] = 0;
      if (k < NK) {
        double ax[3] = {dm->key_axis[k][0], dm->key_axis[k][1], dm->key_axis[k][2]};
        double r0[3] = {dm->key_ipos[k][0] - dm->key_jpos[k][0], dm->key_ipos[k][1] - dm->key_jpos[k][1], dm->key_ipos[k][2] - dm->key_jpos[k][2]};
        double sn, cs, r[3], mg[3], t[3];
        sincos(qk[s], &sn, &cs);
        rot_axis(r, ax, sn, cs, r0);
        double mass = dm->key_mass[k];
        for (int c = 0; c < 3; c++) mg[c] = mass * g3[c];
        rp::cross3(t, r, mg);
        double bias = -rp::dot3(ax, t);
        double passive = -dm->key_stiffness[k] * (qk[s] - dm->key_springref[k]) - dm->key_damping[k] * vk[s];
        double actq = 0;
        if (dm->nkey_act) {
          double ctrl = fmin(fmax(gctrl[k], dm->key_act_range[k][0]), dm->key_act_range[k][1]);
          actq = dm->key_act_gain[k] * ctrl;
          key_actforce[s] = actq;
        }
        kM[s] = dm->key_M[k];
        ks_qfrc[s] = passive - bias + actq;
        ks_qacc[s] = ks_qfrc[s] / kM[s];
        sm.kfrc[k] = ks_qfrc[s];
        sm.kqs[k] = ks_qacc[s];
        sm.keyslot[k] = -1;
        double lo = dm->key_range[k][0], hi = dm->key_range[k][1];
        double dlo = qk[s] - lo, dhi = hi - qk[s];
        double J = 0, pos = 0;
        if (dlo < 0) { J = 1; pos = dlo; krow[s].active = 1; }
        else if (dhi < 0) { J = -1; pos = dhi; krow[s].active = 1; }
        if (krow[s].active) {
          double solimp[5];
          for (int c = 0; c < 5; c++) solimp[c] = dm->key_lim_solimp[k][c];
          double imp = rp::impedance(solimp, pos, 0.0);
          double R = fmax(RP_MINVAL, (1 - imp) * dm->key_dof_invw[k] / imp);
          double aref = -dm->key_lim_B[k] * (J * vk[s]) - dm->key_lim_K[k] * imp * pos;
          krow[s].J = J; krow[s].R = R; krow[s].aref = aref;
          krow[s].b = J * ks_qacc[s] - aref;
          krow[s].AR = 1.0 / kM[s] + R;
          krow[s].ARinv = 1.0 / krow[s].AR;
        }
      }
    }
    // ------------------------------------------------------------------ hand friction + limit rows
    double fr_R[NH > 0 ? NH : 1], fr_b[NH > 0 ? NH : 1], fr_AR[NH > 0 ? NH : 1], fr_f[NH > 0 ? NH : 1], fr_loss[NH > 0 ? NH : 1], fr_aref[NH > 0 ? NH : 1];
    double lm_J[NH > 0 ? NH : 1], lm_R[NH > 0 ? NH : 1], lm_b[NH > 0 ? NH : 1], lm_AR[NH > 0 ? NH : 1], lm_f[NH > 0 ? NH : 1], lm_aref[NH > 0 ? NH : 1];
    unsigned lm_mask[NH > 0 ? NH : 1], fr_mask[NH > 0 ? NH : 1];
#pragma unroll
    for (int h = 0; h < NH; h++) {
      fr_R[h] = 1; fr_b[h] = 0; fr_AR[h] = 1; fr_f[h] = 0; fr_loss[h] = 0; fr_aref[h] = 0;
      lm_J[h] = 0; lm_R[h] = 1; lm_b[h] = 0; lm_AR[h] = 1; lm_f[h] = 0; lm_aref[h] = 0;
      bool fr_on = false, lm_on = false;
      if (hl) {
        double mii = -sm.A[h][lane * RS + lane];
        fr_loss[h] = dm->link_frictionloss[h][lane];
        fr_on = fr_loss[h] > 0;
        if (fr_on) {
          fr_R[h] = dm->link_fr_R[h][lane];
          fr_aref[h] = -dm->link_fr_B[h][lane] * vh[h];
          fr_b[h] = qacc_smooth[h] - fr_aref[h];
          fr_AR[h] = mii + fr_R[h];
        }
        if (dm->link_limited[h][lane]) {
          double lo = dm->link_range[h][lane][0], hi = dm->link_range[h][lane][1];
          double dlo = qh[h] - lo, dhi = hi - qh[h], pos = 0;
          if (dlo < 0) { lm_J[h] = 1; pos = dlo; lm_on = true; }
          else if (dhi < 0) { lm_J[h] = -1; pos = dhi; lm_on = true; }
          if (lm_on) {
            double solimp[5];
            for (int c = 0; c < 5; c++) solimp[c] = dm->link_lim_solimp[h][lane][c];
            double imp = rp::impedance(solimp, pos, 0.0);
            lm_R[h] = fmax(RP_MINVAL, (1 - imp) * dm->link_dof_invw[h][lane] / imp);
            lm_aref[h] = -dm->link_lim_B[h][lane] * (lm_J[h] * vh[h]) - dm->link_lim_K[h][lane] * imp * pos;
            lm_b[h] = lm_J[h] * qacc_smooth[h] - lm_aref[h];
            lm_AR[h] = mii + lm_R[h];
          }
        }
      }
      fr_mask[h] = __ballot_sync(FULL, fr_on);
      lm_mask[h] = __ballot_sync(FULL, lm_on);
    }

    TICK(3)
    // ------------------------------------------------------------------ collision (lane = hand geom)
    __syncwarp();
    if (lane == 0) { sm.nraw = 0; sm.ncon = 0; sm.overflow = 0; }
    __syncwarp();
    if (NH > 0) {
#pragma unroll
      for (int h = 0; h < NH; h++) {
        bool isg = lane < dm->ngeom[h];
        int gl = isg ? dm->geom_link[h][lane] : 0;
        double lp[3], lq[4];
#pragma unroll
        for (int k = 0; k < 3; k++) lp[k] = shfl(xp[h][k], gl);
#pragma unroll
        for (int k = 0; k < 4; k++) lq[k] = shfl(xq[h][k], gl);
        double gc[3] = {0, 0, 0}, gax[3] = {0, 0, 1}, rad = 0, half = 0, gmargin = 0, rb = 0;
        int i0 = 0, i1 = -1, ncand = 0, gtype = 2, gorder = 0;
        if (isg) {
          double xm[9], tmp[3];
          rp::quat2mat(xm, lq);
          gtype = dm->geom_type[h][lane];
          gorder = dm->geom_order[h][lane];
          double gpos[3] = {dm->geom_pos[h][lane][0], dm->geom_pos[h][lane][1], dm->geom_pos[h][lane][2]};
          double gaxl[3] = {dm->geom_axis[h][lane][0], dm->geom_axis[h][lane][1], dm->geom_axis[h][lane][2]};
          rp::mat_vec(tmp, xm, gpos);
          for (int k = 0; k < 3; k++) gc[k] = lp[k] + tmp[k];
          rp::mat_vec(gax, xm, gaxl);
          rad = dm->geom_size[h][lane][0];
          half = gtype == 3 ? dm->geom_size[h][lane][1] : 0.0;
          gmargin = dm->geom_gp[h][lane].margin;
          rb = rad + half;
          // candidate piano boxes: base + keys whose y-extent overlaps
          i0 = (int)floor((gc[1] - rb - 0.02 - dm->lattice_y0) / dm->lattice_pitch);
          i1 = (int)ceil((gc[1] + rb + 0.02 - dm->lattice_y0) / dm->lattice_pitch);
          i0 = max(i0, 0); i1 = min(i1, 51);
          ncand = (i1 >= i0) ? 2 * (i1 - i0 + 1) + 1 : 0;
        }
        // structured, warp-uniform trip count (no continue/break: keeps the warp reconvergent)
        const int ncand_max = __reduce_max_sync(FULL, isg ? ncand : 0);
        for (int ci = -1; ci < ncand_max; ci++) {
          // candidate -> piano box: ci = -1 base; 2k / 2k+1 = white / following black of slot i0+k;
          // last = the black key before the first slot
          int key = -2;
          if (isg && ci < ncand) {
            if (ci < 0) key = dm->has_base ? -1 : -2;
            else if (ci == ncand - 1) key = i0 > 0 ? dm->black_after[i0 - 1] : -1, key = key < 0 ? -2 : key;
            else {
              int slot = i0 + (ci >> 1);
              key = (ci & 1) ? dm->black_after[slot] : dm->white_of_slot[slot];
              key = key < 0 ? -2 : key;
            }
          }
          if (key > -2) {
            // box pose
            double bpos[3], bhalf[3], R[9];
            double margin;
            if (key < 0) {
              for (int k = 0; k < 3; k++) { bpos[k] = dm->base_pos[k]; bhalf[k] = dm->base_half[k]; }
              R[0] = R[4] = R[8] = 1; R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0;
              margin = fmax(gmargin, dm->base_gp.margin);
            } else {
              double ax[3] = {dm->key_axis[key][0], dm->key_axis[key][1], dm->key_axis[key][2]};
              double sn, cs;
              sincos(sm.qk[key], &sn, &cs);
              // rotation matrix about axis: columns = rotated basis vectors
              double ex[3] = {1, 0, 0}, ey[3] = {0, 1, 0}, ez[3] = {0, 0, 1}, cx[3], cy[3], cz[3];
              rot_axis(cx, ax, sn, cs, ex); rot_axis(cy, ax, sn, cs, ey); rot_axis(cz, ax, sn, cs, ez);
              R[0] = cx[0]; R[3] = cx[1]; R[6] = cx[2];
              R[1] = cy[0]; R[4] = cy[1]; R[7] = cy[2];
              R[2] = cz[0]; R[5] = cz[1]; R[8] = cz[2];
              double jp[3] = {dm->key_jpos[key][0], dm->key_jpos[key][1], dm->key_jpos[key][2]}, rj[3];
              rp::mat_vec(rj, R, jp);
              for (int k = 0; k < 3; k++) { bpos[k] = dm->key_pos[key][k] + jp[k] - rj[k]; bhalf[k] = dm->key_half[key][k]; }
              margin = fmax(gmargin, dm->key_gp[key].margin);
            }
            double rel[3] = {gc[0] - bpos[0], gc[1] - bpos[1], gc[2] - bpos[2]}, c[3];
            rp::matT_vec(c, R, rel);
            double bound = rb + margin;
            if (rp::point_box_d2(c, bhalf) <= bound * bound) {
              // narrowphase: sphere, or capsule end points + interior closest point
              int ntest = 1;
              double pts[3][3];
              if (gtype == 2) {
                for (int k = 0; k < 3; k++) { pts[0][k] = c[k]; pts[1][k] = c[k]; pts[2][k] = c[k]; }
              } else {
                double ax[3];
                rp::matT_vec(ax, R, gax);
                for (int k = 0; k < 3; k++) { pts[0][k] = c[k] - half * ax[k]; pts[1][k] = c[k] + half * ax[k]; pts[2][k] = c[k]; }
                ntest = 2;
                double t = rp::segment_box_closest_t(pts[0], pts[1], bhalf);
                if (t > 1e-9 && t < 1 - 1e-9) {
                  for (int k = 0; k < 3; k++) pts[2][k] = pts[0][k] + t * (pts[1][k] - pts[0][k]);
                  ntest = 3;
                }
              }
#pragma unroll
              for (int ti = 0; ti < 3; ti++) {
                double n[3], p[3];
                double dist = rp::sphere_box(pts[ti], rad, bhalf, n, p);
                if (ti < ntest && dist < margin) {
                  int slot = atomicAdd(&sm.nraw, 1);
                  if (slot < MAXCON) {
                    RawContact& rc = sm.u.c.raw[slot];
                    double nw[3], pw[3];
                    rp::mat_vec(nw, R, n);
                    rp::mat_vec(pw, R, p);
                    rc.dist = dist;
                    for (int k = 0; k < 3; k++) { rc.n[k] = nw[k]; rc.pos[k] = bpos[k] + pw[k]; }
                    rc.hand = h; rc.geom = lane; rc.key = key; rc.sub = ti;
                    // MuJoCo order: piano geom (base, then keys) major, hand geom (model order) minor
                    rc.sortkey = (unsigned)(((key + 1) * 64 + gorder) * 4 + ti);
                  } else {
                    sm.overflow = 1;
                  }
                }
              }
            }
          }
        }
        __syncwarp();
      }
    }
    __syncwarp();
    TICK(4)
    int nraw = min(sm.nraw, MAXCON);
    if (sm.nraw > MAXCON) flags |= 1;
    // NOTE: on overflow the kept subset depends on atomic order; the flag marks the env invalid.
    // rank-sort into con[] and build the constraint blocks (lane c < nraw handles contact c's geometry)
    int myrank = -1;
    if (lane < nraw) {
      unsigned mykey = sm.u.c.raw[lane].sortkey;
      myrank = 0;
      for (int j = 0; j < nraw; j++) myrank += (sm.u.c.raw[j].sortkey < mykey) ? 1 : 0;
    }
    __syncwarp();
    const int ncon = nraw;
    ncon_last = ncon;
    // contact geometry is first moved to registers (raw[] aliases nothing in con[], both live in the union)
    double c_dist = 0, c_pos[3] = {0, 0, 0}, c_frame[9];
    int c_hand = 0, c_geom = 0, c_key = -1;
    if (lane < ncon) {
      RawContact rc = sm.u.c.raw[lane];
      c_dist = rc.dist;
      for (int k = 0; k < 3; k++) { c_pos[k] = rc.pos[k]; c_frame[k] = rc.n[k]; }
      double nn = sqrt(rp::dot3(c_frame, c_frame));
      for (int k = 0; k < 3; k++) c_frame[k] /= nn;
      rp::make_frame(c_frame);
      c_hand = rc.hand; c_geom = rc.geom; c_key = rc.key;
    }
    __syncwarp();
    if (lane < ncon) {
      Contact& cn = sm.u.c.con[myrank];
      int h = c_hand, link = dm->geom_link[h][c_geom];
      cn.hand = h; cn.link = link; cn.key = c_key;
      // mixed parameters
      const RpGeomParams& g1 = dm->geom_gp[h][c_geom];
      const RpGeomParams& g2 = c_key < 0 ? dm->base_gp : dm->key_gp[c_key];
      double solref[2], solimp[5], mu;
      rp::mix_params(g1.solref, g1.solimp, g1.friction, g1.solmix, g1.priority, g2.solref, g2.solimp, g2.friction,
                     g2.solmix, g2.priority, solref, solimp, &mu);
      double margin = fmax(g1.margin, g2.margin), gap = fmax(g1.gap, g2.gap), incl = margin - gap;
      cn.mu = mu;
      // Jacobian: rows = frame directions, body2 (key) minus body1 (hand link)
      int dep = sm.dep[h][link];
      cn.dep = dep;
      for (int d = 0; d < 16; d++) cn.chain[d] = (signed char)(d <= dep ? sm.anc[h][link][d] : 0);
      for (int r = 0; r < 3; r++) {
        const double* dir = c_frame + 3 * r;
        for (int d = 0; d < MAXD; d++) {
          double v = 0;
          if (d <= dep) {
            int j = sm.anc[h][link][d];
            const double* S = sm.S[h][j];
            double wxp[3];
            rp::cross3(wxp, S, c_pos);
            v = -(dir[0] * (S[3] + wxp[0]) + dir[1] * (S[4] + wxp[1]) + dir[2] * (S[5] + wxp[2]));
          }
          cn.Jc[r][d] = v;
        }
        double jk = 0;
        if (c_key >= 0) {
          double ax[3] = {dm->key_axis[c_key][0], dm->key_axis[c_key][1], dm->key_axis[c_key][2]};
          double an[3] = {dm->key_pos[c_key][0] + dm->key_jpos[c_key][0], dm->key_pos[c_key][1] + dm->key_jpos[c_key][1],
                          dm->key_pos[c_key][2] + dm->key_jpos[c_key][2]};
          double lin[3], wxp[3];
          rp::cross3(lin, an, ax);  // S_lin = anchor x axis
          rp::cross3(wxp, ax, c_pos);
          jk = dir[0] * (lin[0] + wxp[0]) + dir[1] * (lin[1] + wxp[1]) + dir[2] * (lin[2] + wxp[2]);
        }
        cn.Jk[r] = jk;
      }
      cn.invMk = c_key >= 0 ? 1.0 / dm->key_M[c_key] : 0.0;
      // impedance / regularisation
      double imp = rp::impedance(solimp, c_dist, incl);
      double K, B;
      rp::ref_KB(solref, solimp[1], hstep, &K, &B);
      double diag = dm->link_body_invw[h][link] + (c_key >= 0 ? dm->key_body_invw[c_key] : 0.0);
      double Rn = fmax(RP_MINVAL, (1 - imp) * diag / imp);
      cn.R[0] = Rn;
      cn.R[1] = Rn / fmax(RP_MINVAL, dm->impratio);
      cn.R[2] = cn.R[1];
      // velocities, aref, b, block A
      double vel[3], jqs[3];
      for (int r = 0; r < 3; r++) {
        double sv = 0;
        for (int d = 0; d <= dep; d++) sv += cn.Jc[r][d] * sm.vh[h][sm.anc[h][link][d]];
        if (c_key >= 0) sv += cn.Jk[r] * sm.vk[c_key];
        vel[r] = sv;
      }
      cn.aref[0] = -B * vel[0] - K * imp * (c_dist - incl);
      cn.aref[1] = -B * vel[1];
      cn.aref[2] = -B * vel[2];
      (void)jqs;
      // A = Jc Minv_chain Jc^T + Jk Jk^T / Mk + R
      double u[3][MAXD];
      for (int r = 0; r < 3; r++)
        for (int d = 0; d <= dep; d++) {
          int jd = sm.anc[h][link][d];
          double s2 = 0;
          for (int e = 0; e <= dep; e++) s2 -= sm.A[h][jd * RS + sm.anc[h][link][e]] * cn.Jc[r][e];
          u[r][d] = s2;
        }
      for (int r = 0; r < 3; r++)
        for (int s2 = 0; s2 < 3; s2++) {
          double v = 0;
          for (int d = 0; d <= dep; d++) v += cn.Jc[s2][d] * u[r][d];
          v += cn.Jk[r] * cn.Jk[s2] * cn.invMk;
          if (r == s2) v += cn.R[r];
          cn.A[3 * s2 + r] = v;
        }
      cn.f[0] = cn.f[1] = cn.f[2] = 0;
    }
    __syncwarp();
    // b = J qacc_smooth - aref needs qacc_smooth of chain dofs: publish via wbuf
#pragma unroll
    for (int h = 0; h < NH; h++) if (hl) sm.wbuf[h][lane] = qacc_smooth[h];
#pragma unroll
    for (int s = 0; s < 3; s++) { int k = lane + 32 * s; if (k < NK) sm.vak[k] = ks_qacc[s]; }
    __syncwarp();
    if (lane < ncon) {
      Contact& cn = sm.u.c.con[lane];
      int h = cn.hand, link = cn.link, dep = sm.dep[h][link];
      for (int r = 0; r < 3; r++) {
        double s2 = 0;
        for (int d = 0; d <= dep; d++) s2 += cn.Jc[r][d] * sm.wbuf[h][sm.anc[h][link][d]];
        if (cn.key >= 0) s2 += cn.Jk[r] * sm.vak[cn.key];
        cn.b[r] = s2 - cn.aref[r];
      }
    }
    __syncwarp();

    TICK(5)
    // =============================================================== coupled keys
    // keys touched by a contact get their limit row solved inside the Gauss-Seidel sweep; all other
    // key rows are independent 1x1 problems that are finished in closed form below.
    {
      bool firstc = false;
      int mykey = -1;
      if (lane < ncon) {
        mykey = sm.u.c.con[lane].key;
        firstc = mykey >= 0 && (lane == 0 || sm.u.c.con[lane - 1].key != mykey);
      }
      unsigned fmask = __ballot_sync(FULL, firstc);
      if (firstc) sm.keyslot[mykey] = (signed char)__popc(fmask & ((1u << lane) - 1u));
      if (lane == 0) sm.nck = __popc(fmask);
    }
    __syncwarp();
    const int nck = sm.nck;

    // =============================================================== warm start
    // forces from qacc_warmstart through the primal map; kept only if the dual cost is negative
#pragma unroll
    for (int h = 0; h < NH; h++) if (hl) sm.wbuf[h][lane] = wh[h];
#pragma unroll
    for (int s = 0; s < 3; s++) { int k = lane + 32 * s; if (k < NK) sm.vak[k] = wk[s]; }
    __syncwarp();
    if (lane < ncon) {
      Contact& cn = sm.u.c.con[lane];
      int h = cn.hand, link = cn.link, dep = sm.dep[h][link];
      double jar[3], D[3];
      for (int r = 0; r < 3; r++) {
        double s2 = 0;
        for (int d = 0; d <= dep; d++) s2 += cn.Jc[r][d] * sm.wbuf[h][sm.anc[h][link][d]];
        if (cn.key >= 0) s2 += cn.Jk[r] * sm.vak[cn.key];
        jar[r] = s2 - cn.aref[r];
        D[r] = 1.0 / cn.R[r];
      }
      rp::elliptic_primal_force(jar, D, cn.mu, cn.f);
    }
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (hl && (fr_mask[h] >> lane & 1)) {
        double jar = wh[h] - fr_aref[h], fl = fr_loss[h], R = fr_R[h];
        fr_f[h] = jar <= -R * fl ? fl : (jar >= R * fl ? -fl : -jar / R);
      }
      if (hl && (lm_mask[h] >> lane & 1)) {
        double jar = lm_J[h] * wh[h] - lm_aref[h];
        lm_f[h] = jar < 0 ? -jar / lm_R[h] : 0.0;
      }
    }
#pragma unroll
    for (int s = 0; s < 3; s++) {
      if (krow[s].active) {
        double jar = krow[s].J * wk[s] - krow[s].aref;
        krow[s].f = jar < 0 ? -jar / krow[s].R : 0.0;
      }
    }
    __syncwarp();
    // va = Minv J^T f ; cost = 0.5 f'(J va) + 0.5 sum R f^2 + f'b
    double va[NH > 0 ? NH : 1];
    double improvement0 = 0;  // contribution of the closed-form key rows to the first sweep
    {
      double cost = 0;
      double wq[NH > 0 ? NH : 1];
#pragma unroll
      for (int h = 0; h < NH; h++) {
        wq[h] = 0;
        if (hl) {
          wq[h] = fr_f[h] + lm_J[h] * lm_f[h];
          cost += 0.5 * fr_R[h] * fr_f[h] * fr_f[h] + fr_f[h] * fr_b[h] + 0.5 * lm_R[h] * lm_f[h] * lm_f[h] + lm_f[h] * lm_b[h];
          for (int c = 0; c < ncon; c++) {
            const Contact& cn = sm.u.c.con[c];
            if (cn.hand == h) {
              int link = cn.link;
              bool anc = lane <= link && link < lane + sm.sub[h][lane];
              if (anc) { int d = ldep[h]; wq[h] += cn.Jc[0][d] * cn.f[0] + cn.Jc[1][d] * cn.f[1] + cn.Jc[2][d] * cn.f[2]; }
            }
          }
          sm.wbuf[h][lane] = wq[h];
        }
      }
      double wqk[3];
#pragma unroll
      for (int s = 0; s < 3; s++) {
        int k = lane + 32 * s;
        wqk[s] = 0;
        if (k < NK) {
          wqk[s] = krow[s].J * krow[s].f;
          cost += 0.5 * krow[s].R * krow[s].f * krow[s].f + krow[s].f * krow[s].b;
          if (sm.keyslot[k] >= 0)
            for (int c = 0; c < ncon; c++) {
              const Contact& cn = sm.u.c.con[c];
              if (cn.key == k) wqk[s] += cn.Jk[0] * cn.f[0] + cn.Jk[1] * cn.f[1] + cn.Jk[2] * cn.f[2];
            }
          cost += 0.5 * wqk[s] * wqk[s] / kM[s];
        }
      }
      if (lane < ncon) {
        const Contact& cn = sm.u.c.con[lane];
        for (int r = 0; r < 3; r++) cost += 0.5 * cn.R[r] * cn.f[r] * cn.f[r] + cn.f[r] * cn.b[r];
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < NH; h++) {
        double s2 = 0;
        if (hl) for (int k = 0; k < NL; k++) s2 -= sm.A[h][k * RS + lane] * sm.wbuf[h][k];
        va[h] = s2;
        cost += 0.5 * wq[h] * s2;
      }
      cost = warp_sum(cost);
      const bool keep = !(cost > 0);
      if (!keep) {
#pragma unroll
        for (int h = 0; h < NH; h++) { fr_f[h] = 0; lm_f[h] = 0; va[h] = 0; }
        if (lane < ncon) { Contact& cn = sm.u.c.con[lane]; cn.f[0] = cn.f[1] = cn.f[2] = 0; }
      }
      if (a.dbg && sub == 0 && lane == 0) a.dbg[(size_t)env * RP_DBG_STRIDE + RP_DBG_SCALARS + 3] = cost;
      // key rows: uncoupled ones are solved exactly now (one Gauss-Seidel update of a 1x1 system is
      // its solution); coupled ones move to the ck[] list with their warm-start force
#pragma unroll
      for (int s = 0; s < 3; s++) {
        int k = lane + 32 * s;
        if (k < NK) {
          double f = keep ? krow[s].f : 0.0;
          int slot = sm.keyslot[k];
          if (slot < 0) {
            if (krow[s].active) {
              double res = krow[s].b + krow[s].AR * f;
              double fn = f - res * krow[s].ARinv;
              if (fn < 0) fn = 0;
              double dl = fn - f;
              improvement0 -= 0.5 * dl * dl * krow[s].AR + dl * res;
              f = fn;
            }
            sm.kfc[k] = krow[s].J * f;
            sm.vak[k] = krow[s].J * f / kM[s];
          } else {
            CKey& ck = sm.ck[slot];
            ck.key = k; ck.active = krow[s].active;
            ck.J = krow[s].J; ck.R = krow[s].R; ck.b = krow[s].b; ck.AR = krow[s].AR; ck.ARi = krow[s].ARinv;
            ck.f = f; ck.invM = 1.0 / kM[s];
            sm.vak[k] = (keep ? wqk[s] : 0.0) / kM[s];
          }
        }
      }
    }
    // park what is only needed after the sweep
#pragma unroll
    for (int h = 0; h < NH; h++) if (hl) { sm.stash[h][lane][0] = qfrc_smooth[h]; sm.stash[h][lane][1] = qacc_smooth[h]; }
    __syncwarp();

    TICK(6)
    // =============================================================== PGS
    // Row order inside a sweep follows MuJoCo: friction-loss dofs, joint limits (keys, then hand
    // joints), contacts.  Rows that do not interact (the two hands, untouched keys) are advanced
    // together.  Scalar hand rows (friction loss, limits) run in a half-warp-per-hand mapping: lanes
    // 0-15 serve hand 0, lanes 16-31 hand 1, each lane owning dofs l and l+16 of its hand, so one
    // instruction stream sweeps both hands (gs_rows).  Contacts still use lane == dof.
    int iter = 0;
    const int maxiter = a.iterations > 0 ? a.iterations : dm->iterations;
    const double tol = a.tolerance >= 0 ? a.tolerance : dm->tol;
    const unsigned anyrow_fr = (NH > 0 ? fr_mask[0] : 0u) | (NH > 1 ? fr_mask[NH > 1 ? 1 : 0] : 0u);
    const unsigned anyrow_lm = (NH > 0 ? lm_mask[0] : 0u) | (NH > 1 ? lm_mask[NH > 1 ? 1 : 0] : 0u);
    const int hh = lane >> 4, l0 = lane & 15, l1 = (l0 + 16 < NL) ? l0 + 16 : l0;
    const bool slot1 = l0 + 16 < NL;
    HandRows FR, LM;
    double van[2] = {0, 0};  // va in the half-warp mapping
    if (NH > 1) {
      const int H1 = NH > 1 ? 1 : 0;
#define TO_NEW(x, s) to_new(x[0], x[H1], (s) ? l1 : l0, hh)
#pragma unroll
      for (int s = 0; s < 2; s++) {
        const bool ok = s == 0 || slot1;
        // all shuffles first (full warp), selects afterwards
        const double fl = TO_NEW(fr_loss, s), far = TO_NEW(fr_AR, s), ff = TO_NEW(fr_f, s);
        const double lar = TO_NEW(lm_AR, s), lf = TO_NEW(lm_f, s), lj = TO_NEW(lm_J, s), vv = TO_NEW(va, s);
        FR.b[s] = TO_NEW(fr_b, s); FR.R[s] = TO_NEW(fr_R, s);
        LM.b[s] = TO_NEW(lm_b, s); LM.R[s] = TO_NEW(lm_R, s);
        const bool fon = ok && ((hh ? fr_mask[H1] : fr_mask[0]) >> (s ? l1 : l0) & 1);
        const bool lon = ok && ((hh ? lm_mask[H1] : lm_mask[0]) >> (s ? l1 : l0) & 1);
        FR.f[s] = ok ? ff : 0.0;
        FR.AR[s] = far; FR.ARi[s] = fon ? 1.0 / far : 0.0; FR.J[s] = 1.0;
        FR.hi[s] = fon ? fl : 0.0; FR.lo[s] = -FR.hi[s];
        LM.f[s] = lon ? lf : 0.0;
        LM.AR[s] = lar; LM.ARi[s] = lon ? 1.0 / lar : 0.0; LM.J[s] = lon ? lj : 0.0;
        LM.lo[s] = 0.0; LM.hi[s] = lon ? __longlong_as_double(0x7ff0000000000000LL) : 0.0;
        van[s] = ok ? vv : 0.0;
      }
    }
    const double* Ah = sm.A[NH > 1 ? hh : 0];
    while (iter < maxiter) {
      double improvement = iter == 0 ? improvement0 : 0.0;
      // ---- friction-loss rows, dof order
      if (NH > 1) improvement += gs_rows(Ah, lane, l0, l1, anyrow_fr, FR, van);
      TICK(9)
      // ---- limit rows of keys that carry contacts (uniform scalar work, values in shared memory)
      for (int c = 0; c < nck; c++) {
        CKey& ck = sm.ck[c];
        double vak = sm.vak[ck.key], f = ck.f;
        double res = ck.b + ck.R * f + ck.J * vak;
        double fn = f - res * ck.ARi;
        fn = (fn < 0 || !ck.active) ? 0.0 : fn;
        double dl = fn - f;
        __syncwarp();
        improvement -= lane == 0 ? dl * (0.5 * dl * ck.AR + res) : 0.0;
        ck.f = fn;                                    // every lane stores the same value
        sm.vak[ck.key] = vak + ck.J * dl * ck.invM;
        __syncwarp();
      }
      // ---- hand limit rows
      if (NH > 1 && anyrow_lm) improvement += gs_rows(Ah, lane, l0, l1, anyrow_lm, LM, van);
      TICK(10)
      if (NH > 1 && ncon > 0) {
        // half-warp mapping -> lane == dof for the contact rows
        const int H1 = NH > 1 ? 1 : 0;
        const double a0 = shfl(van[0], l0), a1 = shfl(van[1], l0);        // hand 0: dofs 0-15 / 16-31
        const double b0 = shfl(van[0], 16 + l0), b1 = shfl(van[1], 16 + l0);
        va[0] = lane < 16 ? a0 : a1;
        va[H1] = lane < 16 ? b0 : b1;
      }
      // ---- contacts, sorted order
      for (int c = 0; c < ncon; c++) {
        Contact& cn = sm.u.c.con[c];
        const int h = cn.hand, link = cn.link, key = cn.key, dep = cn.dep;
        const bool h1 = NH > 1 && h == 1;
        const double vsel = h1 ? va[NH > 1 ? 1 : 0] : va[0];
        const int mysub = h1 ? lsub[NH > 1 ? 1 : 0] : lsub[0];
        const int dme = h1 ? ldc[NH > 1 ? 1 : 0] : ldc[0];
        const bool anc = lane <= link && link < lane + mysub;
        const double j0 = anc ? cn.Jc[0][dme] : 0.0, j1 = anc ? cn.Jc[1][dme] : 0.0, j2 = anc ? cn.Jc[2][dme] : 0.0;
        double part[3] = {j0 * vsel, j1 * vsel, j2 * vsel};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int r = 0; r < 3; r++) part[r] += __shfl_xor_sync(FULL, part[r], o);
        }
        double res[3], f[3];
        const double vak = sm.vak[key >= 0 ? key : 0];
        const double jkm = key >= 0 ? 1.0 : 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) {
          f[r] = cn.f[r];
          res[r] = cn.b[r] + cn.R[r] * f[r] + part[r] + jkm * cn.Jk[r] * vak;
        }
        double A[9];
#pragma unroll
        for (int k = 0; k < 9; k++) A[k] = cn.A[k];
        const double fold[3] = {f[0], f[1], f[2]};
        const double change = rp::pgs_elliptic_update(A, res, f, cn.mu);
        const double df[3] = {f[0] - fold[0], f[1] - fold[1], f[2] - fold[2]};
        __syncwarp();
        improvement -= lane == 0 ? change : 0.0;
        cn.f[0] = f[0]; cn.f[1] = f[1]; cn.f[2] = f[2];  // identical on every lane
        if (key >= 0) sm.vak[key] = vak + (cn.Jk[0] * df[0] + cn.Jk[1] * df[1] + cn.Jk[2] * df[2]) * cn.invMk;
        // hand side: w = Jc^T df on the chain, va += Minv[:,chain] w
        const double w = j0 * df[0] + j1 * df[1] + j2 * df[2];
        double acc = 0;
        for (int d = 0; d <= dep; d++) {
          const int j = cn.chain[d];
          const double wj = shfl(w, j);
          acc -= sm.A[h][j * RS + lane] * wj;
        }
        if (h1) va[NH > 1 ? 1 : 0] += acc; else va[0] += acc;
        __syncwarp();
      }
      if (NH > 1 && ncon > 0) {
        const int H1 = NH > 1 ? 1 : 0;
#pragma unroll
        for (int s = 0; s < 2; s++) { const double vv = TO_NEW(va, s); van[s] = (s == 0 || slot1) ? vv : 0.0; }
        (void)H1;
      }
      TICK(11)
      improvement = warp_sum(improvement) * dm->solver_scale;
      iter++;
      if (improvement < tol) break;
    }
    if (NH > 1) {
      // back to lane == dof
      const int H1 = NH > 1 ? 1 : 0;
#define TO_OLD(x, dst0, dst1) { const double a0_ = shfl(x[0], l0), a1_ = shfl(x[1], l0), b0_ = shfl(x[0], 16 + l0), b1_ = shfl(x[1], 16 + l0); \
                               dst0 = lane < 16 ? a0_ : a1_; dst1 = lane < 16 ? b0_ : b1_; }
      TO_OLD(van, va[0], va[H1]);
      TO_OLD(FR.f, fr_f[0], fr_f[H1]);
      TO_OLD(LM.f, lm_f[0], lm_f[H1]);
      if (!hl) { va[0] = 0; va[H1] = 0; fr_f[0] = 0; fr_f[H1] = 0; lm_f[0] = 0; lm_f[H1] = 0; }
    }
    niter_last = iter;

    TICK(7)
    // =============================================================== dual -> primal, Euler
    // qacc = qacc_smooth + va ; qfrc_constraint = J^T f
    double qacc[NH > 0 ? NH : 1], qfrc_c[NH > 0 ? NH : 1];
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (hl) { qfrc_smooth[h] = sm.stash[h][lane][0]; qacc_smooth[h] = sm.stash[h][lane][1]; }
      qacc[h] = qacc_smooth[h] + va[h];
      qfrc_c[h] = 0;
      if (hl) {
        qfrc_c[h] = fr_f[h] + lm_J[h] * lm_f[h];
        for (int c = 0; c < ncon; c++) {
          const Contact& cn = sm.u.c.con[c];
          if (cn.hand == h) {
            int link = cn.link;
            bool anc = lane <= link && link < lane + sm.sub[h][lane];
            if (anc) { int d = ldep[h]; qfrc_c[h] += cn.Jc[0][d] * cn.f[0] + cn.Jc[1][d] * cn.f[1] + cn.Jc[2][d] * cn.f[2]; }
          }
        }
      }
    }
    // constraint force on coupled keys: limit row + contacts, accumulated in contact order
    if (lane == 0) {
      for (int c = 0; c < nck; c++) sm.kfc[sm.ck[c].key] = sm.ck[c].J * sm.ck[c].f;
      for (int c = 0; c < ncon; c++) {
        const Contact& cn = sm.u.c.con[c];
        if (cn.key >= 0) sm.kfc[cn.key] += cn.Jk[0] * cn.f[0] + cn.Jk[1] * cn.f[1] + cn.Jk[2] * cn.f[2];
      }
    }
    __syncwarp();
    double kq_acc[3], kq_frc[3], kq_smooth[3];
#pragma unroll
    for (int s = 0; s < 3; s++) {
      int k = lane + 32 * s;
      kq_acc[s] = 0; kq_frc[s] = 0; kq_smooth[s] = 0;
      if (k < NK) {
        kq_frc[s] = sm.kfc[k];
        kq_smooth[s] = sm.kfrc[k];
        kq_acc[s] = sm.kqs[k] + sm.vak[k];
      }
    }
    if (a.dbg && sub == 0) {
      double* dbg = a.dbg + (size_t)env * RP_DBG_STRIDE;
#pragma unroll
      for (int s = 0; s < 3; s++) {
        int k = lane + 32 * s;
        if (k < NK) { dbg[RP_DBG_QACC_SMOOTH + k] = sm.kqs[k]; dbg[RP_DBG_QACC + k] = kq_acc[s]; dbg[RP_DBG_QFRC_C + k] = kq_frc[s]; dbg[RP_DBG_BIAS + k] = 0; }
      }
#pragma unroll
      for (int h = 0; h < NH; h++) if (hl) {
        dbg[RP_DBG_QACC_SMOOTH + NK + h * NL + lane] = qacc_smooth[h];
        dbg[RP_DBG_QACC + NK + h * NL + lane] = qacc[h];
        dbg[RP_DBG_QFRC_C + NK + h * NL + lane] = qfrc_c[h];
        dbg[RP_DBG_BIAS + NK + h * NL + lane] = qfrc_bias[h];
      }
      if (lane == 0) { dbg[RP_DBG_SCALARS + 0] = ncon; dbg[RP_DBG_SCALARS + 1] = iter; dbg[RP_DBG_SCALARS + 2] = sm.overflow; }
      if (lane < ncon) {
        const Contact& cn = sm.u.c.con[lane];
        double* o = dbg + RP_DBG_CONTACTS + lane * 16;
        o[0] = cn.key; o[1] = cn.hand; o[2] = cn.link; o[3] = cn.f[0]; o[4] = cn.f[1]; o[5] = cn.f[2];
        o[6] = cn.b[0]; o[7] = cn.b[1]; o[8] = cn.b[2]; o[9] = cn.R[0]; o[10] = cn.A[0]; o[11] = cn.A[4]; o[12] = cn.A[1];
        o[13] = cn.aref[0]; o[14] = cn.mu; o[15] = cn.Jk[0];
      }
    }
    __syncwarp();
    // Euler with implicit joint damping: (M + h*D) a = qfrc_smooth + qfrc_constraint
    double aeul[NH > 0 ? NH : 1];
    if (NH > 0) {
      double* Aflat = &sm.A[0][0];
      for (int idx = lane; idx < NH * NL * RS; idx += 32) Aflat[idx] = 0;
      __syncwarp();
#pragma unroll
      for (int h = 0; h < NH; h++) {
        if (hl) {
          for (int d = 0; d <= ldep[h]; d++) {
            const int j = sm.anc[h][lane][d];
            double v = sm.Mrow[h][lane][d];
            if (d == ldep[h]) v += hstep * dm->link_damping[h][lane];
            sm.A[h][lane * RS + j] = v;
            sm.A[h][j * RS + lane] = v;
          }
          sm.wbuf[h][lane] = qfrc_smooth[h] + qfrc_c[h];  // right-hand side
        }
      }
      __syncwarp();
      // a = (M + h*D)^-1 rhs: sweep (A = -inverse), then one row-times-vector per lane
      sweep_inverse<NH>(sm, lane);
#pragma unroll
      for (int h = 0; h < NH; h++) {
        double acc = 0;
        if (hl) for (int i = 0; i < NL; i++) acc -= sm.A[h][lane * RS + i] * sm.wbuf[h][i];
        aeul[h] = acc;
      }
      __syncwarp();
    }
    TICK(8)
    // integrate
#pragma unroll
    for (int h = 0; h < NH; h++) {
      if (hl) {
        vh[h] += hstep * aeul[h];
        qh[h] += hstep * vh[h];
        wh[h] = qacc[h];
      }
    }
#pragma unroll
    for (int s = 0; s < 3; s++) {
      int k = lane + 32 * s;
      if (k < NK) {
        double ae = (kq_smooth[s] + kq_frc[s]) / (dm->key_M[k] + hstep * dm->key_damping[k]);
        vk[s] += hstep * ae;
        qk[s] += hstep * vk[s];
        wk[s] = kq_acc[s];
      }
    }
    __syncwarp();
  }  // substeps

  // ------------------------------------------------------------------ store state
  {
    bool bad = false;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      int k = lane + 32 * s;
      if (k < NK) { gq[k] = qk[s]; gv[k] = vk[s]; gw[k] = wk[s]; bad |= !isfinite(qk[s]) || !isfinite(vk[s]); }
    }
#pragma unroll
    for (int h = 0; h < NH; h++)
      if (hl) { gq[NK + h * NL + lane] = qh[h]; gv[NK + h * NL + lane] = vh[h]; gw[NK + h * NL + lane] = wh[h]; bad |= !isfinite(qh[h]) || !isfinite(vh[h]); }
    if (__any_sync(FULL, bad)) flags |= 2;
    if (lane == 0 && a.diag) { a.diag[3 * env + 0] = ncon_last; a.diag[3 * env + 1] = niter_last; a.diag[3 * env + 2] = flags; }
    if (a.act_force) {
#pragma unroll
      for (int h = 0; h < NH; h++) if (lane < dm->nact[h]) a.act_force[(size_t)env * nu + h * dm->nact[0] + lane] = actforce[h];
      if (dm->nkey_act) {
#pragma unroll
        for (int s = 0; s < 3; s++) { int k = lane + 32 * s; if (k < NK) a.act_force[(size_t)env * nu + k] = key_actforce[s]; }
      }
    }
  }
#ifdef RP_PROFILE
  if (a.dbg && lane == 0) for (int i = 0; i < 12; i++) a.dbg[(size_t)env * RP_DBG_STRIDE + RP_DBG_SCALARS + 4 + i] = (double)tacc[i];
#endif
  if (a.mode == RP_MODE_PHYSICS) return;

  // =============================================================================================
  // task epilogue: after_step, observables, reward, termination  (piano_with_shadow_hands.py:188-412,
  // self_actuated_piano.py:143-243) — uses the poses of the final kinematics pass.
  // =============================================================================================
  const int T = dm->T, L = dm->lookahead;
  const double sustain_state = a.sustain[env];
  const bool first = do_reset || a.mode == RP_MODE_RESET;
  int t_new = first ? 0 : t_idx + 1;
  // key state
  double knorm[3];
  bool kact[3];
#pragma unroll
  for (int s = 0; s < 3; s++) {
    int k = lane + 32 * s;
    knorm[s] = 0; kact[s] = false;
    if (k < NK) {
      double hi = dm->key_range[k][1];
      if (dm->task_kind == 1) {
        double mid = 0.5 * (dm->key_act_range[k][0] + dm->key_act_range[k][1]);
        kact[s] = gctrl[k] >= mid;
        knorm[s] = 0;  // state observable stays at its initial value in the self-actuated task
      } else {
        double st = fmin(fmax(qk[s], dm->key_range[k][0]), hi);
        knorm[s] = st / hi;
        kact[s] = fabs(st - hi) <= 0.00872665;
      }
      if (a.activation) a.activation[(size_t)env * NK + k] = kact[s] ? 1 : 0;
    }
  }
  const bool sustain_act = sustain_state >= 0.5;
  double reward = 0, terms[5] = {0, 0, 0, 0, 0}, discount = 1.0;
  int step_type = first ? RP_STEP_FIRST_ : RP_STEP_MID_;
  // fingertip sites (lane s < 5: site s of each hand)
  double site[NH > 0 ? NH : 1][3];
#pragma unroll
  for (int h = 0; h < NH; h++) {
    bool iss = lane < RP_NSITE;
    int sl = iss ? dm->site_link[h][lane] : 0;
    double lp[3], lq[4];
#pragma unroll
    for (int k = 0; k < 3; k++) lp[k] = shfl(xp[h][k], sl);
#pragma unroll
    for (int k = 0; k < 4; k++) lq[k] = shfl(xq[h][k], sl);
    double tmp[3], spos[3] = {dm->site_pos[h][iss ? lane : 0][0], dm->site_pos[h][iss ? lane : 0][1], dm->site_pos[h][iss ? lane : 0][2]};
    rotq(tmp, lq, spos);
    for (int k = 0; k < 3; k++) site[h][k] = lp[k] + tmp[k];
    if (iss && a.site_xpos) for (int k = 0; k < 3; k++) a.site_xpos[((size_t)env * 2 * RP_NSITE + h * RP_NSITE + lane) * 3 + k] = site[h][k];
  }
  if (!first) {
    const uint8_t* gcur = a.goal + (size_t)min(t_idx, T - 1) * RP_GOAL_DIM_;  // goal emitted with the previous observation
    bool should_terminate = (t_new - 1) == T - 1;
    // ---- key-press / activation statistics
    double sum_on = 0; int n_on = 0; bool false_pos = false;
    double sq = 0, xent = 0;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      int k = lane + 32 * s;
      if (k < NK) {
        int g = gcur[k];
        if (g) {
          double x = (double)g - knorm[s];  // goal - state/qmax, bounds (0, 0.05), margin 0.5, gaussian
          double v = 1.0;
          if (x < 0 || x > 0.05) { double dd = (x < 0 ? -x : x - 0.05) / 0.5; double sc = sqrt(-2 * log(0.1)); v = exp(-0.5 * (dd * sc) * (dd * sc)); }
          sum_on += v; n_on++;
        } else if (kact[s]) false_pos = true;
        double df = (kact[s] ? 1.0 : 0.0) - (double)g;
        sq += df * df;
        double pr = kact[s] ? 1.0 : 0.0;  // negative_binary_cross_entropy, eps 1e-6 (self_actuated_piano.py:38-47)
        xent += g ? log(pr + 1e-6) : log(1 - pr + 1e-6);
      }
    }
    sum_on = warp_sum(sum_on);
    n_on = __reduce_add_sync(FULL, n_on);
    false_pos = __any_sync(FULL, false_pos);
    sq = warp_sum(sq); xent = warp_sum(xent);
    int gs = gcur[NK];
    if (dm->task_kind == 0) {
      double r_key = (n_on > 0 ? 0.5 * sum_on / n_on : 0.0) + 0.5 * (false_pos ? 0.0 : 1.0);
      // sustain: tolerance(goal - sustain_activation)
      double xs = (double)gs - (sustain_act ? 1.0 : 0.0), r_sus = 1.0;
      if (xs < 0 || xs > 0.05) { double dd = (xs < 0 ? -xs : xs - 0.05) / 0.5; double sc = sqrt(-2 * log(0.1)); r_sus = exp(-0.5 * (dd * sc) * (dd * sc)); }
      // energy: |actuator force (pre-integration)| * |actuator velocity (post-integration)|
      double pw = 0;
#pragma unroll
      for (int h = 0; h < NH; h++) {
        bool isact = lane < dm->nact[h];
        int d0 = isact ? dm->act_dof0[h][lane] : 0, d1 = isact ? dm->act_dof1[h][lane] : -1;
        double v0 = shfl(vh[h], d0), v1 = shfl(vh[h], d1 >= 0 ? d1 : 0);
        if (isact) {
          double vel = dm->act_coef0[h][lane] * v0 + (d1 >= 0 ? dm->act_coef1[h][lane] * v1 : 0.0);
          pw += fabs(actforce[h]) * fabs(vel);
        }
      }
      // per-hand sums in hand order, as the reference accumulates them
      double r_en = -dm->energy_coef * warp_sum(pw);
      // fingering: mean tolerance of fingertip-to-key distances for (key, finger) pairs due now
      double r_fing = 0;
      if (!dm->disable_fingering_reward) {
        const int8_t* fcur = a.fingering + (size_t)min(t_idx, T - 1) * NK;
        double fsum = 0; int fn = 0;
        for (int k = 0; k < NK; k++) {
          int fg = fcur[k];
          if (fg < 0 || !gcur[k]) continue;
          int hh = fg < 5 ? 0 : 1, fi = fg < 5 ? fg : fg - 5;
          // key geom centre in the world: body pos rotated about the hinge
          double ax[3] = {dm->key_axis[k][0], dm->key_axis[k][1], dm->key_axis[k][2]};
          double sn, cs, r0[3] = {-dm->key_jpos[k][0], -dm->key_jpos[k][1], -dm->key_jpos[k][2]}, rr[3];
          double qkk = shfl(k < 32 ? qk[0] : (k < 64 ? qk[1] : qk[2]), k & 31);
          sincos(qkk, &sn, &cs);
          rot_axis(rr, ax, sn, cs, r0);
          double kp[3];
          for (int c = 0; c < 3; c++) kp[c] = dm->key_pos[k][c] + dm->key_jpos[k][c] + rr[c];
          kp[2] += 0.5 * dm->key_half[k][2];
          kp[0] += 0.35 * dm->key_half[k][0];
          double sx = shfl(NH > 1 && hh == 1 ? site[NH > 1 ? 1 : 0][0] : site[0][0], fi);
          double sy = shfl(NH > 1 && hh == 1 ? site[NH > 1 ? 1 : 0][1] : site[0][1], fi);
          double sz = shfl(NH > 1 && hh == 1 ? site[NH > 1 ? 1 : 0][2] : site[0][2], fi);
          double dd = sqrt((kp[0] - sx) * (kp[0] - sx) + (kp[1] - sy) * (kp[1] - sy) + (kp[2] - sz) * (kp[2] - sz));
          double v = 1.0;
          if (dd > 0.01) { double e = (dd - 0.01) / 0.1; double sc = sqrt(-2 * log(0.1)); v = exp(-0.5 * (e * sc) * (e * sc)); }
          fsum += v; fn++;
        }
        r_fing = fn > 0 ? fsum / fn : 0.0;
      }
      double r_fore = dm->disable_forearm_reward ? 0.0 : 0.5;  // hands never collide with each other here
      terms[0] = r_key; terms[1] = r_sus; terms[2] = r_en; terms[3] = r_fing; terms[4] = r_fore;
      reward = r_key + r_sus + r_en;
      if (!dm->disable_fingering_reward) reward += r_fing;
      if (!dm->disable_forearm_reward) reward += r_fore;
      bool failure = false_pos;
      if (should_terminate) step_type = RP_STEP_LAST_;
      else if (dm->wrong_press_termination && failure) { discount = 0.0; step_type = RP_STEP_LAST_; }
    } else {
      // self-actuated: reward on [activation, sustain_activation] vs goal
      double dfs = (sustain_act ? 1.0 : 0.0) - (double)gs;
      double prs = sustain_act ? 1.0 : 0.0;
      if (dm->reward_type == 0) reward = -sqrt(sq + dfs * dfs);
      else reward = xent + (gs ? log(prs + 1e-6) : log(1 - prs + 1e-6));
      terms[0] = reward;
      if (should_terminate) step_type = RP_STEP_LAST_;
    }
  }
  // ---- observation (order: piano/state, piano/sustain_state, rh joints_pos, lh joints_pos, goal, fingering)
  if (a.obs) {
    double* o = a.obs + (size_t)env * dm->obs_dim;
    int off = 0;
    if (dm->task_kind == 0) {
#pragma unroll
      for (int s = 0; s < 3; s++) { int k = lane + 32 * s; if (k < NK) o[k] = knorm[s]; }
      off = NK;
      if (lane == 0) o[off] = sustain_state;
      off += 1;
#pragma unroll
      for (int h = 0; h < NH; h++) {
        // joints_pos in the reference's joint order
        int src = hl ? dm->obs_dof[h][lane] : 0;
        double v = shfl(qh[h], src);
        if (hl) o[off + h * NL + lane] = v;
      }
      off += NH * NL;
    } else {
      // self-actuated piano observables: piano/activation (88), piano/sustain_activation (1)
#pragma unroll
      for (int s = 0; s < 3; s++) { int k = lane + 32 * s; if (k < NK) o[k] = kact[s] ? 1.0 : 0.0; }
      off = NK;
      if (lane == 0) o[off] = sustain_act ? 1.0 : 0.0;
      off += 1;
    }
    // goal lookahead window starting at the new t_idx; kept from the previous observation when t == T
    int tg = t_new;
    bool frozen = (tg >= T);
    for (int i = 0; i <= L; i++) {
      for (int c = lane; c < RP_GOAL_DIM_; c += 32) {
        double v;
        if (frozen) { int t2 = (T - 1) + i; v = t2 < T ? (double)a.goal[(size_t)t2 * RP_GOAL_DIM_ + c] : 0.0; }
        else { int t2 = tg + i; v = t2 < T ? (double)a.goal[(size_t)t2 * RP_GOAL_DIM_ + c] : 0.0; }
        o[off + i * RP_GOAL_DIM_ + c] = v;
      }
    }
    off += (L + 1) * RP_GOAL_DIM_;
    if (dm->task_kind == 0 && !dm->disable_fingering_reward) {
      int tf = frozen ? T - 1 : tg;
      if (lane < 10) {
        double v = 0;
        const int8_t* fc = a.fingering + (size_t)tf * NK;
        for (int k = 0; k < NK; k++) if (fc[k] == lane && a.goal[(size_t)tf * RP_GOAL_DIM_ + k]) v = 1.0;
        o[off + lane] = v;
      }
    }
  }
  if (lane == 0) {
    a.t_idx[env] = t_new;
    a.step_type_state[env] = step_type;
    if (a.reward) a.reward[env] = first ? 0.0 : reward;
    if (a.discount) a.discount[env] = first ? 1.0 : discount;
    if (a.step_type) a.step_type[env] = (uint8_t)step_type;
    if (a.reward_terms) for (int k = 0; k < 5; k++) a.reward_terms[(size_t)env * 5 + k] = first ? 0.0 : terms[k];
  }
}

// ------------------------------------------------------------------------------------------------
size_t rp_kernel_smem_bytes() { return sizeof(Smem); }

cudaError_t rp_launch_step(const RpKernelArgs& a, int nhand, cudaStream_t stream) {
  size_t smem = sizeof(Smem);
  static bool configured[2] = {false, false};
  if (nhand == 2) {
    if (!configured[0]) {
      cudaError_t e = cudaFuncSetAttribute(rp_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured[0] = true;
    }
    rp_step_kernel<2><<<a.nenv, 32, smem, stream>>>(a);
  } else {
    if (!configured[1]) {
      cudaError_t e = cudaFuncSetAttribute(rp_step_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured[1] = true;
    }
    rp_step_kernel<0><<<a.nenv, 32, smem, stream>>>(a);
  }
  return cudaGetLastError();
}
