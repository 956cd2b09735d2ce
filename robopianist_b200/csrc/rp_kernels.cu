// rp_kernels.cu — sm_100a step kernel: 64 threads per environment, RP_BLOCK_ENVS environments per block, fp64.
//
// Replaces the reference's per-control-step host loop
//   before_step -> 10 x physics.step() -> after_step -> observables -> reward -> termination
// (robopianist/suite/tasks/piano_with_shadow_hands.py:176-220, tasks/base.py:28-31; the physics is
// MuJoCo's mj_step, restated stage by stage) with a single launch.
//
// Mapping: rp_step_kernel<2, EPB, HC> runs two warps per environment, warp h owning hand h (lane == link ==
// dof of that hand, up to 26 of 32 lanes; hands with fewer dofs are padded with inert links); the 88 keys are
// spread over the 64 threads (thread t owns keys t and t+64).  The two hands only interact through keys they both
// touch — and, in the HC instantiation (hand geoms colliding with each other), through cross-hand contacts — so each
// warp runs its own instruction stream — kinematics, RNE/CRB, M^-1, constraint rows, Gauss-Seidel — and the block
// synchronises where key / contact data crosses (a handful of barriers per substep plus one per
// PGS iteration for the convergence test).  rp_step_kernel<1, 1, false> is the keys-only variant
// (self_actuated_piano): one warp, lane l owns keys l, l+32, l+64.
// Shared memory holds the per-hand inverse inertia, motion subspaces, contact blocks and key
// arrays; everything lane-owned lives in registers across the stages of a substep.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/roboplay.h"
#include "rp_kernels.h"
#include "rp_math.cuh"

#define FULL 0xffffffffu
#define NL RP_NL
#define NK RP_NK
#define MAXD RP_MAXD
#define MAXCON RP_MAXCON
#define RAWCAP 40  // raw narrowphase entries of the hand-collision instantiation (more than MAXCON: the deepest MAXCON records are kept)
#define RS 27  // row stride of the hand matrices (odd: lane-per-row accesses are bank-conflict free)
#define H1 (RP_NHMAX - 1)
// tuning knobs (DESIGN.md §5 lists what was measured, including the variants that were removed again)
#ifndef RP_CH_UNROLL
#define RP_CH_UNROLL 1
#endif
#ifndef RP_EPB_MINBLOCKS
#define RP_EPB_MINBLOCKS (RP_EPB <= 6 ? 6 / RP_EPB : 1)  // resident blocks per SM the register allocation has to allow (six environments per SM)
#endif
#ifndef RP_PHASE_EVERY
#define RP_PHASE_EVERY 1  // phase barrier at the top of every RP_PHASE_EVERY-th substep (RP_EPB > 1)
#endif
#ifndef RP_PHASE_GROUP
#define RP_PHASE_GROUP RP_EPB  // environments per phase group: the groups of a block are phase-locked separately (the last group takes the remainder)
#endif
#ifndef RP_NU
#define RP_NU 8  // contacts whose M^-1 Jc^T is kept in shared memory (the others walk the kinematic chain)
#endif
#ifndef RP_MINBLOCKS
#define RP_MINBLOCKS 6
#endif

namespace {

constexpr int kChUnroll = RP_CH_UNROLL;

struct Contact {  // lives in shared memory
  double Jc[3][MAXD];  // hand-chain Jacobian entries by depth (already signed: body2 - body1)
  double Jk[3];        // key dof entry
  double A[9];         // 3x3 block of AR (R on the diagonal)
  double R[3], b[3], f[3];
  double mu, invMk;
  rp::EllInv inv;            // sweep-invariant part of the elliptic block update (1/A00, friction block at lambda = 0, ...)
  int hand, link, key, dep;  // key = -1: piano base (static); -2 / -3: the two records of a cross-hand contact (the block lives in the
                             // first, the second carries the other body's Jacobian); -4: a contact between two bodies of one hand, one
                             // record whose "chain" lists the dofs below the bodies' common ancestor (those above cancel exactly);
                             // dep: chain length - 1, bit 8: friction block regular (qcqp ok)
  signed char chain[16];     // dof (lane) at depth d of the hand chain
};
struct CKey {  // limit row of a key that is coupled to a hand through a contact
  double J, R, b, AR, ARi, f, invM;
  int key, active;
};
struct RawContact {  // narrowphase output before sorting
  double dist, pos[3], n[3];
  int hand, geom, key, sub;  // hand : hand contacts: across the hands two entries, key = -2 (geom 1's side) / -3 (geom 2's side); within a hand
                             // one entry, key = -4; sub = partner's hand * 32 + geom
  unsigned sortkey, pad;
};

struct Smem {
  double A[RP_NHMAX][NL * RS];  // per hand: M, then -M^-1 (sweep), later -(M + h*D)^-1; symmetric, row-major
  union {
    struct {  // per-substep scratch, dead once the contact rows are built
      double S[RP_NHMAX][NL][6];
      double vk[NK];
      double ksn[NK], kcs[NK];          // sin / cos of the key angles of this substep
      double wsb[RP_NHMAX][32];         // qacc_warmstart of the hand dofs (read by the contact rows)
      double vh[RP_NHMAX][NL];          // hand joint velocities (for chain walks)
    } pre;
    double U[RP_NU > 0 ? RP_NU : 1][3][NL];  // Gauss-Seidel: M^-1 Jc^T of the first RP_NU contacts, lane-major
  } v;
  double vak[NK];
  double wbuf[RP_NHMAX][32];
  double actf[RP_NHMAX][RP_NAH];
  double kfc[NK];                     // per key: constraint force J^T f (qacc_warmstart before the sweep)
  double Mrow[RP_NHMAX][160];         // mass-matrix entries M[link][ancestor at depth d], packed by link (kept for the Euler solve)
  double ctrl[NK];                    // actuator controls of this control step
  double site[RP_NHMAX][RP_NSITE][3]; // fingertip sites (task epilogue)
  double red[2][RP_NHMAX][10];        // block reductions (double buffered)
  CKey ck[MAXCON];                    // limit rows of keys that also carry contacts
  signed char anc[RP_NHMAX][NL][MAXD];  // ancestor-or-self at depth d
  signed char dep[RP_NHMAX][32], sub[RP_NHMAX][32];  // link depth / subtree size
  unsigned char moff[RP_NHMAX][32];   // offset of link k's packed mass-matrix row in Mrow
  signed char keyslot[NK];            // index into ck[] (-1: key not in contact)
  signed char hlist[RP_NHMAX][MAXCON];  // contacts of hand h, in sorted order
  signed char ckown[MAXCON];          // hand that sweeps the key row ck[i]
  signed char perm[MAXCON];           // sorted position -> raw[] slot
  int nhc[RP_NHMAX];
  int nck, shared_key;                // shared_key: some key carries contacts of both hands
  union {
    double acc[RP_NHMAX][NL][16];  // upward accumulation scratch: F (6) + composite inertia (10)
    Contact con[MAXCON];           // constraint blocks of the contacts (sorted)
    RawContact raw[MAXCON];        // narrowphase output; moved to registers before con[] is written
    struct {                       // hand : hand collisions: the raw list plus the world-frame primitives of both hands
      RawContact raw[RAWCAP];
      double gw[RP_NHMAX][RP_NGH][8];  // centre, axis, radius, half length
    } col;
  } u;
  int ncon, nraw, overflow, pad;
  int near_uncovered, npairc;  // near_uncovered: some collider outside the narrowphase's coverage came within reach of the piano (this launch); npairc: hand : hand contacts of this substep (each takes two records)
};

constexpr size_t kSmemStride = (sizeof(Smem) + 15) / 16 * 16;  // per-environment shared memory (RP_EPB > 1: slot stride)
// block-shared copy of the sweep schedule (the two hands have the same tree: one copy, checked on the host),
// placed behind the environment slots of the hands kernels
struct SweepSched { int off[32]; uint16_t pairs[RP_SW_NPAIRS_SMEM]; };
// block-shared broadphase tables (behind the schedule): per-key position, per-class geometry, the white-key lattice.
// In global memory every candidate cost two dependent round trips to L2 (slot -> key -> pose).
struct KeyGeo {
  double pos[NK][3];
  double cls[2][10];  // jpos[3], half[3], axis[3], margin
  unsigned char cls_of[NK];
  signed char white_of_slot[52], black_after[52];
};
constexpr size_t kSchedBytes = (sizeof(SweepSched) + 15) / 16 * 16 + (sizeof(KeyGeo) + 15) / 16 * 16;
constexpr size_t kGeoOffset = (sizeof(SweepSched) + 15) / 16 * 16;

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// Barrier over the threads of one environment.  With several environments per thread block (RP_EPB > 1) each
// environment owns a named barrier (ids 1..EPB, 64 threads); barrier 0 (__syncthreads) is the block-wide
// phase barrier.
template <int NW, int EPB>
__device__ __forceinline__ void env_sync(int bar_id) {
  if (NW == 1) __syncwarp();
  else if (EPB == 1) __syncthreads();
  else { __syncwarp(); asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); }
}
// the two warps of one environment, on the environment's named barrier (also when the block holds a single
// environment: pgs_solve is one function shared by every launch shape, which keeps their results bit-identical)
__device__ __forceinline__ void hand_pair_sync(int bar_id) {
  __syncwarp();
  asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
}
// Phase barrier: the environments of one phase group (RP_PHASE_GROUP consecutive slots of the block; the last group
// takes the remainder) start the substep together.  One group: barrier 0; several: named barriers EPB + 1 + group.
// A function of its own (not inlined): every thread of the block — stepping, finished or without an environment — waits
// on the same barrier instruction.
template <int EPB>
__device__ __noinline__ void phase_sync(int slot) {
  constexpr int PG = RP_PHASE_GROUP, NG = EPB / (PG > 0 ? PG : 1);
  if (NG <= 1) { __syncthreads(); return; }
  const int g = slot / PG < NG - 1 ? slot / PG : NG - 1;
  const int gsize = g == NG - 1 ? EPB - PG * (NG - 1) : PG;
  __syncwarp();
  asm volatile("bar.sync %0, %1;" ::"r"(EPB + 1 + g), "r"(64 * gsize) : "memory");
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
// dm_control.utils.rewards.tolerance, sigmoid="gaussian", value_at_margin=0.1
__device__ __noinline__ double tolerance_gauss(double x, double lo, double hi, double margin) {
  if (x >= lo && x <= hi) return 1.0;
  const double dd = (x < lo ? lo - x : x - hi) / margin;
  const double sc = sqrt(-2 * log(0.1));
  return exp(-0.5 * (dd * sc) * (dd * sc));
}

// precision / recall / F-score ratio with sklearn's zero_division=1
__device__ __forceinline__ double prf_ratio(double num, double den) { return den > 0 ? num / den : 1.0; }

struct KeyRow {  // active limit row of a key (at most one side can be active)
  double J, R, b, aref, ARinv, AR, f;
  int active;
};

// In-place symmetric sweep of one hand matrix: afterwards A = -M^-1 (callers fold the sign into their
// multiply-adds).  Pivots run leaves first (k = NL-1 .. 0): at that point row k is non-zero only at the ancestors
// of k and inside its subtree (its support S_k), so the rank-one update of pivot k touches |S_k|^2 elements instead
// of 25^2.  The element updates of one pivot are independent of each other: they are spread over the 32 lanes from a
// precomputed schedule (RpDevModel::sw_pairs, one entry per element j <= i, mirrored on store), 2-20 independent
// load / multiply-add / store groups per lane and pivot instead of one lane walking a row element by element
// through shared memory (r01: 16 % of the step's warp time, 1100 cycles per pivot).
__device__ __noinline__ void sweep_inverse(double* A, const SweepSched* __restrict__ sched, int lane) {
  const uint16_t* const pairs = sched->pairs;
  const int* const off = sched->off;
  int pend = off[NL];
#pragma unroll 1
  for (int k = NL - 1; k >= 0; k--) {
    const double* rowk = A + k * RS;
    const int pbeg = off[k];
    const double dinv = rp::rcp(rowk[k]);
#pragma unroll 2
    for (int p = pbeg + lane; p < pend; p += 32) {
      const unsigned e = pairs[p];
      const int j = e >> 5, i = e & 31;
      const double t = rowk[j] * dinv;
      const double v = fma(-rowk[i], t, A[j * RS + i]);
      A[j * RS + i] = v;
      A[i * RS + j] = v;
    }
    pend = pbeg;
    __syncwarp();
    if (lane < NL) {  // row / column k (zero outside the support: the scaled value is zero there as well)
      const double t = rowk[lane] * dinv;
      const double w = lane == k ? -dinv : t;
      A[lane * RS + k] = w;
      if (lane != k) A[k * RS + lane] = w;
    }
    __syncwarp();
  }
}

#ifdef RP_PROFILE
#define PTICK(i) { const long long tn = clock64(); tacc[i] += tn - *tlast; *tlast = tn; }
#else
#define PTICK(i)
#endif

struct PgsLane { double fr_f, lm_f, va; };

// Projected Gauss-Seidel over the rows of one hand (one warp; lane == dof), MuJoCo's row order inside a sweep:
// friction-loss dofs, joint limits (keys, then hand joints), contacts.  The two hands are independent islands
// unless they touch the same key, so each warp sweeps the rows of its own hand and the rows of a key are handled
// by the warp of the hand touching it.  If some key carries contacts of both hands (shared_key) the key rows go to
// warp 0 and the contact rows are stepped in global order with an environment barrier per row.  One environment
// barrier per sweep for MuJoCo's global convergence test.  A function of its own (not inlined, not a template: one
// body for every launch shape) so that the hot loop gets a register allocation that does not have to carry the
// lane state of the whole step.
// Cost changes are accumulated warp-uniformly in row order (the committing lane broadcasts its term), the order
// mj_solPGS adds them in.  Returns the number of sweeps; forces / va through io.
__device__ __noinline__ int pgs_solve(Smem* smp, const int h, const int bar_id, const int ncon, const int maxiter, const double tol,
                                      const double scale, const double improvement0, const double fr_R, const double fr_b,
                                      const double fr_AR, const double fr_hi, const double lm_J, const double lm_R,
                                      const double lm_b, const double lm_AR, const unsigned lm_mask, const int lsub,
                                      const int ldc, PgsLane* io
#ifdef RP_PROFILE
                                      , long long* tacc, long long* tlast
#endif
) {
  Smem& sm = *smp;
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  const bool hl = lane < NL;
  const int lc = hl ? lane : 0;
  const double* const Ah = sm.A[h];
  const double* const pcol = Ah + lc;
  double fr_f = io->fr_f, lm_f = io->lm_f, va = io->va;
  const int nck = sm.nck, shared_key = sm.shared_key, nmine = sm.nhc[h];
  const double fr_ARi = 1.0 / fr_AR, lm_ARi = ((lm_mask >> lane) & 1u) ? 1.0 / lm_AR : 0.0;
  const double fr_lo = -fr_hi;
  // limit rows of the coupled keys this warp sweeps; an inactive row keeps its zero force (it never moves anything)
  unsigned ckmask = 0;
  for (int c = 0; c < nck; c++) {
    const bool mine = shared_key ? (h == 0) : (sm.ckown[c] == h);
    if (mine && sm.ck[c].active) ckmask |= 1u << c;
  }
  const signed char* const hlist = sm.hlist[h];
  Contact* const con = sm.u.con;
  const int ntrip = shared_key ? ncon : nmine;
  const double imp0 = warp_sum(improvement0);
  int iter = 0;
  while (iter < maxiter) {
    double improvement = iter == 0 ? imp0 : 0.0;
    // ---- friction-loss rows, dof order (row i of the classic sweep: res = b + R f + va_i,
    // f' = clamp(f - res / AR, +-frictionloss), va += M^-1[:, i] (f' - f), cost change (f' - f)(0.5 (f' - f) AR + res)).
    // Every lane evaluates the update of its own row against the current va.  Few changing rows (most rows sit at
    // +-frictionloss while the hand moves): a ballot finds the first row at or after the cursor whose force changes,
    // that row commits and broadcasts, the search resumes behind it — the rows skipped in between would have
    // committed a change of exactly zero at their turn, so the iterates are those of the row-by-row sweep.  (A plain
    // row-by-row loop for sweeps in which most rows change — a hand at rest, nothing saturated — was measured and is
    // slower at every switch-over threshold: 18.2 ms per step always on, 14.83 never, r02g; fetching the column of the
    // most likely next row one commit ahead changed nothing, r02p; a fixed-order sweep that broadcasts the unclamped change
    // first and redoes the rare clamped row — a shorter dependency chain on paper — is slower as well, r02x.)
    {
      const double fbase = fr_b + fr_R * fr_f;  // f only changes at the lane's own commit
      double res = fbase + va;
      double x = fr_f - res * fr_ARi;
      double fx = x < fr_lo ? fr_lo : x;
      fx = x > fr_hi ? fr_hi : fx;
      unsigned m = __ballot_sync(FULL, hl && fx != fr_f);
      while (m != 0u) {
        const int i = __ffs(m) - 1;
        const double col = pcol[i * RS];
        const double dl = fx - fr_f;
        const double dbc = shfl(dl, i);  // (the cost term is computed in the shadow of this shuffle and broadcast off the va chain)
        const double imp = dl * (0.5 * dl * fr_AR + res);
        va -= col * dbc;
        if (lane == i) fr_f = fx;
        res = fbase + va;
        x = fr_f - res * fr_ARi;
        fx = x < fr_lo ? fr_lo : x;
        fx = x > fr_hi ? fr_hi : fx;
        improvement -= shfl(imp, i);
        m = __ballot_sync(FULL, hl && lane > i && fx != fr_f);
      }
    }
    PTICK(8)
    // ---- limit rows of keys that carry contacts (values in shared memory, every lane computes the same)
    for (unsigned mm = ckmask; mm; mm &= mm - 1) {
      CKey& ck = sm.ck[__ffs(mm) - 1];
      const int key = ck.key;
      const double vak = sm.vak[key], f = ck.f;
      const double res = ck.b + ck.R * f + ck.J * vak;
      double fn = f - res * ck.ARi;
      fn = fn < 0 ? 0.0 : fn;
      const double dl = fn - f;
      __syncwarp();
      improvement -= dl * (0.5 * dl * ck.AR + res);
      ck.f = fn;  // every lane stores the same value
      sm.vak[key] = vak + ck.J * dl * ck.invM;
      __syncwarp();
    }
    if (shared_key) hand_pair_sync(bar_id);
    // ---- hand limit rows
    {
      const double lbase = lm_b + lm_R * lm_f;
      for (unsigned m2 = lm_mask; m2; m2 &= m2 - 1) {
        const int i = __ffs(m2) - 1;
        const double res = lbase + lm_J * va;
        const double x = lm_f - res * lm_ARi;
        const double fn = x < 0 ? 0.0 : x;
        const double dl = fn - lm_f;
        const double col = Ah[i * RS + lc];
        const double dbc = shfl(lm_J * dl, i);
        const double imp = dl * (0.5 * dl * lm_AR + res);
        va -= col * dbc;
        improvement -= shfl(imp, i);
        if (lane == i) lm_f = fn;
      }
    }
    PTICK(9)
    // ---- contacts of this hand, sorted order.  The per-lane Jacobian entries of the next contact are fetched while
    // the current one is being updated; the block of the current one arrives in the shadow of the J*v reduction.
    if (ntrip > 0) {
      int c = shared_key ? 0 : hlist[0];
      double j0, j1, j2;
      {
        const Contact& q = con[c];
        const int link = q.link;
        const bool anc = lane <= link && link < lane + lsub;
        j0 = anc ? q.Jc[0][ldc] : 0.0; j1 = anc ? q.Jc[1][ldc] : 0.0; j2 = anc ? q.Jc[2][ldc] : 0.0;
      }
      for (int ci = 0; ci < ntrip; ci++) {
        Contact& cn = con[c];
        const int cnext = ci + 1 < ntrip ? (shared_key ? ci + 1 : (int)hlist[ci + 1]) : c;
        double n0, n1, n2;
        {
          const Contact& q = con[cnext];
          const int link = q.link;
          const bool anc = lane <= link && link < lane + lsub;
          n0 = anc ? q.Jc[0][ldc] : 0.0; n1 = anc ? q.Jc[1][ldc] : 0.0; n2 = anc ? q.Jc[2][ldc] : 0.0;
        }
        if (!shared_key || cn.hand == h) {
          double p0 = j0 * va, p1 = j1 * va, p2 = j2 * va;
          const int key = cn.key, depq = cn.dep;
          double Ab[9], fold[3], rbase[3];
          // The J*va reduction (five shuffle stages, each followed by adds that wait ~25 cycles for the shuffles) interleaved by hand
          // with the block's loads and the ray denominator: the compiler keeps shuffle sequences compact and neither hoists
          // independent work into their latency nor overlaps the two chains (measured: the loads and the ray's dependent
          // multiply-adds were issued first, the reduction after them).  Same operations as pgs_elliptic_ray.
          double s0 = __shfl_xor_sync(FULL, p0, 16), s1 = __shfl_xor_sync(FULL, p1, 16), s2 = __shfl_xor_sync(FULL, p2, 16);
#pragma unroll
          for (int k = 0; k < 9; k++) Ab[k] = cn.A[k];
#pragma unroll
          for (int r = 0; r < 3; r++) fold[r] = cn.f[r];
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 8); s1 = __shfl_xor_sync(FULL, p1, 8); s2 = __shfl_xor_sync(FULL, p2, 8);
          const double cmu = cn.mu;
          const double vak = sm.vak[key >= 0 ? key : 0], jkm0 = key >= 0 ? 1.0 : 0.0;
          const double jk0 = cn.Jk[0], jk1 = cn.Jk[1], jk2 = cn.Jk[2], invMk = cn.invMk;
          const double cb0 = cn.b[0], cb1 = cn.b[1], cb2 = cn.b[2], cR0 = cn.R[0], cR1 = cn.R[1], cR2 = cn.R[2];
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 4); s1 = __shfl_xor_sync(FULL, p1, 4); s2 = __shfl_xor_sync(FULL, p2, 4);
          double rv[3];
#pragma unroll
          for (int r = 0; r < 3; r++) rv[r] = Ab[3 * r] * fold[0] + Ab[3 * r + 1] * fold[1] + Ab[3 * r + 2] * fold[2];
          const rp::EllInv inv = cn.inv;
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 2); s1 = __shfl_xor_sync(FULL, p1, 2); s2 = __shfl_xor_sync(FULL, p2, 2);
          const double denom = rp::dot3(fold, rv);
          const int den_ok = denom >= RP_MINVAL;
          const double rden = rp::rcp(denom);
          rbase[0] = cb0 + cR0 * fold[0] + jkm0 * jk0 * vak; rbase[1] = cb1 + cR1 * fold[1] + jkm0 * jk1 * vak; rbase[2] = cb2 + cR2 * fold[2] + jkm0 * jk2 * vak;
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 1); s1 = __shfl_xor_sync(FULL, p1, 1); s2 = __shfl_xor_sync(FULL, p2, 1);
#if RP_NU > 0
          double u0 = 0, u1 = 0, u2 = 0;
          if (c < RP_NU) { const double* Uc = &sm.v.U[c][0][lc]; u0 = Uc[0]; u1 = Uc[NL]; u2 = Uc[2 * NL]; }
#endif
          p0 += s0; p1 += s1; p2 += s2;
          PTICK(16)
          const double res[3] = {rbase[0] + p0, rbase[1] + p1, rbase[2] + p2};
          double f[3] = {fold[0], fold[1], fold[2]};
          const double change = rp::pgs_elliptic_apply(inv, depq >> 8, rden, den_ok, Ab, res, f, cmu);
          const double df0 = f[0] - fold[0], df1 = f[1] - fold[1], df2 = f[2] - fold[2];
          PTICK(17)
          __syncwarp();
          improvement -= change;
          cn.f[0] = f[0]; cn.f[1] = f[1]; cn.f[2] = f[2];  // identical on every lane
          if (df0 != 0.0 || df1 != 0.0 || df2 != 0.0) {  // warp-uniform; an unchanged force moves nothing
            if (key >= 0) sm.vak[key] = vak + (jk0 * df0 + jk1 * df1 + jk2 * df2) * invMk;
#if RP_NU > 0
            if (c < RP_NU) {
              va += fma(u2, df2, fma(u1, df1, u0 * df0));
            } else
#endif
            {
              // hand side: w = Jc^T df on the chain, va += Minv[:,chain] w
              const double w = j0 * df0 + j1 * df1 + j2 * df2;
              const int dep = depq & 0xff;
              double acc = 0;
#pragma unroll 1
              for (int d = 0; d <= dep; d++) {
                const int j = cn.chain[d];
                const double wj = shfl(w, j);
                acc -= Ah[j * RS + lc] * wj;
              }
              va += acc;
            }
          }
          __syncwarp();
          PTICK(18)
        }
        if (shared_key) hand_pair_sync(bar_id);
        c = cnext; j0 = n0; j1 = n1; j2 = n2;
      }
    }
    PTICK(10)
    if (lane == 0) sm.red[iter & 1][h][1] = improvement;
    hand_pair_sync(bar_id);
    PTICK(11)
    improvement = (sm.red[iter & 1][0][1] + sm.red[iter & 1][H1][1]) * scale;
    iter++;
    if (improvement < tol) break;
  }
  io->fr_f = fr_f; io->lm_f = lm_f; io->va = va;
  return iter;
}

// Jacobian entries of the lane's dof in a contact record whose chain is any list of dofs (hand-collision instantiation: the
// chain of a same-hand contact is not a root path, so the lane's depth does not index it)
__device__ __forceinline__ void lane_jac(const Contact& cn, int lane, int lsub, int ldc, double& j0, double& j1, double& j2) {
  j0 = 0; j1 = 0; j2 = 0;
  if (cn.key != -4) {  // root path: the lane's depth indexes it
    const int link = cn.link;
    if (lane <= link && link < lane + lsub) { j0 = cn.Jc[0][ldc]; j1 = cn.Jc[1][ldc]; j2 = cn.Jc[2][ldc]; }
    return;
  }
  const int dep = cn.dep & 0xff;
  for (int d = 0; d <= dep; d++)
    if (cn.chain[d] == lane) { j0 = cn.Jc[0][d]; j1 = cn.Jc[1][d]; j2 = cn.Jc[2][d]; }
}

// The solver of models whose hand geoms may collide with each other (disable_hand_collisions=False).  Same sweep as
// pgs_solve — friction rows, limit rows of coupled keys, hand limit rows, contacts in sorted order (this copy loads
// the Jacobian entries on demand instead of a contact ahead).  A contact between two bodies of one hand is one record over
// the dofs below their common ancestor (key = -4), swept like any other by the warp of that hand.  A cross-hand contact
// occupies two adjacent records: the first (key = -2) holds the block and the Jacobian of geom 1's body, the second
// (key = -3) the Jacobian of geom 2's body; both mirror the force; sm.shared_key is set, the rows are stepped in global order, each warp reduces its own half of J * va, the
// halves are exchanged through shared memory, both warps evaluate the (identical) block update and apply it to their own
// dofs; the warp of geom 1 accounts for the cost change.
__device__ __noinline__ int pgs_solve_hc(Smem* smp, const int h, const int bar_id, const int ncon, const int maxiter, const double tol,
                                         const double scale, const double improvement0, const double fr_R, const double fr_b,
                                         const double fr_AR, const double fr_hi, const double lm_J, const double lm_R,
                                         const double lm_b, const double lm_AR, const unsigned lm_mask, const int lsub,
                                         const int ldc, PgsLane* io) {
  Smem& sm = *smp;
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  const bool hl = lane < NL;
  const int lc = hl ? lane : 0;
  const double* const Ah = sm.A[h];
  const double* const pcol = Ah + lc;
  double fr_f = io->fr_f, lm_f = io->lm_f, va = io->va;
  const int nck = sm.nck, shared_key = sm.shared_key, nmine = sm.nhc[h];
  const double fr_ARi = 1.0 / fr_AR, lm_ARi = ((lm_mask >> lane) & 1u) ? 1.0 / lm_AR : 0.0;
  const double fr_lo = -fr_hi;
  unsigned ckmask = 0;
  for (int c = 0; c < nck; c++) {
    const bool mine = shared_key ? (h == 0) : (sm.ckown[c] == h);
    if (mine && sm.ck[c].active) ckmask |= 1u << c;
  }
  const signed char* const hlist = sm.hlist[h];
  Contact* const con = sm.u.con;
  const int ntrip = shared_key ? ncon : nmine;
  const double imp0 = warp_sum(improvement0);
  int iter = 0;
  while (iter < maxiter) {
    double improvement = iter == 0 ? imp0 : 0.0;
    {  // friction-loss rows (see pgs_solve)
      const double fbase = fr_b + fr_R * fr_f;
      double res = fbase + va;
      double x = fr_f - res * fr_ARi;
      double fx = x < fr_lo ? fr_lo : x;
      fx = x > fr_hi ? fr_hi : fx;
      unsigned m = __ballot_sync(FULL, hl && fx != fr_f);
      while (m != 0u) {
        const int i = __ffs(m) - 1;
        const double col = pcol[i * RS];
        const double dl = fx - fr_f;
        const double dbc = shfl(dl, i);  // (the cost term is computed in the shadow of this shuffle and broadcast off the va chain)
        const double imp = dl * (0.5 * dl * fr_AR + res);
        va -= col * dbc;
        if (lane == i) fr_f = fx;
        res = fbase + va;
        x = fr_f - res * fr_ARi;
        fx = x < fr_lo ? fr_lo : x;
        fx = x > fr_hi ? fr_hi : fx;
        improvement -= shfl(imp, i);
        m = __ballot_sync(FULL, hl && lane > i && fx != fr_f);
      }
    }
    for (unsigned mm = ckmask; mm; mm &= mm - 1) {  // limit rows of keys that carry contacts
      CKey& ck = sm.ck[__ffs(mm) - 1];
      const int key = ck.key;
      const double vak = sm.vak[key], f = ck.f;
      const double res = ck.b + ck.R * f + ck.J * vak;
      double fn = f - res * ck.ARi;
      fn = fn < 0 ? 0.0 : fn;
      const double dl = fn - f;
      __syncwarp();
      improvement -= dl * (0.5 * dl * ck.AR + res);
      ck.f = fn;
      sm.vak[key] = vak + ck.J * dl * ck.invM;
      __syncwarp();
    }
    if (shared_key) hand_pair_sync(bar_id);
    {  // hand limit rows
      const double lbase = lm_b + lm_R * lm_f;
      for (unsigned m2 = lm_mask; m2; m2 &= m2 - 1) {
        const int i = __ffs(m2) - 1;
        const double res = lbase + lm_J * va;
        const double x = lm_f - res * lm_ARi;
        const double fn = x < 0 ? 0.0 : x;
        const double dl = fn - lm_f;
        const double col = Ah[i * RS + lc];
        const double dbc = shfl(lm_J * dl, i);
        const double imp = dl * (0.5 * dl * lm_AR + res);
        va -= col * dbc;
        improvement -= shfl(imp, i);
        if (lane == i) lm_f = fn;
      }
    }
    // contacts, sorted order
    for (int ci = 0; ci < ntrip; ci++) {
      const int c = shared_key ? ci : (int)hlist[ci];
      Contact& cn = con[c];
      const int kind = cn.key;  // >= -1: hand : piano; -2 / -3: first / second record of a hand : hand contact
      if (kind != -3) {         // the second record is handled together with the first
        const bool pair = kind == -2;
        Contact& cs = con[pair ? c + 1 : c];
        const bool mineP = cn.hand == h, mineS = pair && cs.hand == h;
        const bool cross = pair && cs.hand != cn.hand;
        if (mineP || mineS) {  // warp-uniform
          double j0 = 0, j1 = 0, j2 = 0, k0 = 0, k1 = 0, k2 = 0;
          if (mineP) lane_jac(cn, lane, lsub, ldc, j0, j1, j2);
          if (mineS) lane_jac(cs, lane, lsub, ldc, k0, k1, k2);
          double p0 = (j0 + k0) * va, p1 = (j1 + k1) * va, p2 = (j2 + k2) * va;
          const int key = cn.key, depq = cn.dep;
          double Ab[9], fold[3], rbase[3];
          // reduction interleaved with the block's loads and the ray denominator (see pgs_solve)
          double s0 = __shfl_xor_sync(FULL, p0, 16), s1 = __shfl_xor_sync(FULL, p1, 16), s2 = __shfl_xor_sync(FULL, p2, 16);
#pragma unroll
          for (int k = 0; k < 9; k++) Ab[k] = cn.A[k];
#pragma unroll
          for (int r = 0; r < 3; r++) fold[r] = cn.f[r];
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 8); s1 = __shfl_xor_sync(FULL, p1, 8); s2 = __shfl_xor_sync(FULL, p2, 8);
          const double cmu = cn.mu;
          const double vak = sm.vak[key >= 0 ? key : 0], jkm0 = key >= 0 ? 1.0 : 0.0;
          const double jk0 = cn.Jk[0], jk1 = cn.Jk[1], jk2 = cn.Jk[2], invMk = cn.invMk;
          const double cb0 = cn.b[0], cb1 = cn.b[1], cb2 = cn.b[2], cR0 = cn.R[0], cR1 = cn.R[1], cR2 = cn.R[2];
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 4); s1 = __shfl_xor_sync(FULL, p1, 4); s2 = __shfl_xor_sync(FULL, p2, 4);
          double rv[3];
#pragma unroll
          for (int r = 0; r < 3; r++) rv[r] = Ab[3 * r] * fold[0] + Ab[3 * r + 1] * fold[1] + Ab[3 * r + 2] * fold[2];
          const rp::EllInv inv = cn.inv;
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 2); s1 = __shfl_xor_sync(FULL, p1, 2); s2 = __shfl_xor_sync(FULL, p2, 2);
          const double denom = rp::dot3(fold, rv);
          const int den_ok = denom >= RP_MINVAL;
          const double rden = rp::rcp(denom);
          rbase[0] = cb0 + cR0 * fold[0] + jkm0 * jk0 * vak; rbase[1] = cb1 + cR1 * fold[1] + jkm0 * jk1 * vak; rbase[2] = cb2 + cR2 * fold[2] + jkm0 * jk2 * vak;
          p0 += s0; p1 += s1; p2 += s2;
          s0 = __shfl_xor_sync(FULL, p0, 1); s1 = __shfl_xor_sync(FULL, p1, 1); s2 = __shfl_xor_sync(FULL, p2, 1);
          p0 += s0; p1 += s1; p2 += s2;
          if (cross) {  // both warps are here: exchange the halves of J * va (the block and the old force are already read)
            if (lane == 0) { double* r = sm.red[0][h]; r[2] = p0; r[3] = p1; r[4] = p2; }
            hand_pair_sync(bar_id);
            p0 = sm.red[0][0][2] + sm.red[0][H1][2]; p1 = sm.red[0][0][3] + sm.red[0][H1][3]; p2 = sm.red[0][0][4] + sm.red[0][H1][4];
          }
          const double res[3] = {rbase[0] + p0, rbase[1] + p1, rbase[2] + p2};
          double f[3] = {fold[0], fold[1], fold[2]};
          const double change = rp::pgs_elliptic_apply(inv, depq >> 8, rden, den_ok, Ab, res, f, cmu);
          const double df0 = f[0] - fold[0], df1 = f[1] - fold[1], df2 = f[2] - fold[2];
          __syncwarp();
          if (mineP) improvement -= change;
          cn.f[0] = f[0]; cn.f[1] = f[1]; cn.f[2] = f[2];  // identical on every lane (and in both warps of a cross-hand pair)
          if (pair) { cs.f[0] = f[0]; cs.f[1] = f[1]; cs.f[2] = f[2]; }
          if (df0 != 0.0 || df1 != 0.0 || df2 != 0.0) {
            if (key >= 0) sm.vak[key] = vak + (jk0 * df0 + jk1 * df1 + jk2 * df2) * invMk;
            // hand side, per record of this hand: w = Jc^T df on the chain, va += Minv[:, chain] w
            // (the first RP_NU records have M^-1 Jc^T in shared memory, lane-major: three loads instead of the chain walk)
            if (mineP) {
#if RP_NU > 0
              if (c < RP_NU) {
                const double* Uc = &sm.v.U[c][0][lc];
                va += fma(Uc[2 * NL], df2, fma(Uc[NL], df1, Uc[0] * df0));
              } else
#endif
              {
                const double w = j0 * df0 + j1 * df1 + j2 * df2;
                const int dep = depq & 0xff;
                double acc = 0;
#pragma unroll 1
                for (int d = 0; d <= dep; d++) { const int j = cn.chain[d]; acc -= Ah[j * RS + lc] * shfl(w, j); }
                va += acc;
              }
            }
            if (mineS) {
#if RP_NU > 0
              if (c + 1 < RP_NU) {
                const double* Uc = &sm.v.U[c + 1][0][lc];
                va += fma(Uc[2 * NL], df2, fma(Uc[NL], df1, Uc[0] * df0));
              } else
#endif
              {
                const double w = k0 * df0 + k1 * df1 + k2 * df2;
                const int dep = cs.dep & 0xff;
                double acc = 0;
#pragma unroll 1
                for (int d = 0; d <= dep; d++) { const int j = cs.chain[d]; acc -= Ah[j * RS + lc] * shfl(w, j); }
                va += acc;
              }
            }
          }
          __syncwarp();
        }
      }
      if (shared_key) hand_pair_sync(bar_id);
    }
    if (lane == 0) sm.red[iter & 1][h][1] = improvement;
    hand_pair_sync(bar_id);
    improvement = (sm.red[iter & 1][0][1] + sm.red[iter & 1][H1][1]) * scale;
    iter++;
    if (improvement < tol) break;
  }
  io->fr_f = fr_f; io->lm_f = lm_f; io->va = va;
  return iter;
}

}  // namespace

// The kernel proper is large; it is written as one function so that lane-owned state stays in
// registers across the stages of a substep.
// EPB environments share a thread block (EPB x NW warps).  EPB > 1 keeps the environments of a block in step:
// a block-wide barrier at the top of every substep re-aligns their warps, so that the warps resident on an SM
// run the same stage (the same instruction lines) at the same time instead of each streaming its own
// position of a 370 KB program through the instruction caches.
// HC: the hand geoms may collide with each other (disable_hand_collisions=False): an instantiation of its own, so that the
// default configuration's code is unchanged.
template <int NW, int EPB, bool HC>
__global__ void __launch_bounds__(32 * NW * EPB, EPB > 1 ? RP_EPB_MINBLOCKS : (NW == 2 ? RP_MINBLOCKS : 8)) rp_step_kernel(RpKernelArgs a) {
  constexpr bool HANDS = NW == 2;
  static_assert(!HC || HANDS, "hand : hand collisions need hands");
  constexpr int KT = 32 * NW;               // threads per environment
  constexpr int KS = (NK + KT - 1) / KT;    // keys per thread
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int slot = EPB > 1 ? (int)threadIdx.x / KT : 0;  // environment slot inside the block
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw + (size_t)slot * kSmemStride);
  const int bar_id = slot + 1;
  const RpDevModel* __restrict__ dm = a.dm;
  const SweepSched* const sched = reinterpret_cast<const SweepSched*>(smem_raw + (size_t)EPB * kSmemStride);
  const KeyGeo* const geo = reinterpret_cast<const KeyGeo*>(smem_raw + (size_t)EPB * kSmemStride + kGeoOffset);
  if (HANDS) {  // every thread of the block, before any early exit
    SweepSched* w = reinterpret_cast<SweepSched*>(smem_raw + (size_t)EPB * kSmemStride);
    const int np = dm->sw_off[0][NL];
    for (int i = (int)threadIdx.x; i <= NL; i += KT * EPB) w->off[i] = dm->sw_off[0][i];
    for (int i = (int)threadIdx.x; i < np; i += KT * EPB) w->pairs[i] = dm->sw_pairs[0][i];
    KeyGeo* g = reinterpret_cast<KeyGeo*>(smem_raw + (size_t)EPB * kSmemStride + kGeoOffset);
    for (int i = (int)threadIdx.x; i < NK * 3; i += KT * EPB) g->pos[i / 3][i % 3] = dm->key_pos[i / 3][i % 3];
    for (int i = (int)threadIdx.x; i < 20; i += KT * EPB) g->cls[i / 10][i % 10] = dm->keycls[i / 10][i % 10];
    for (int i = (int)threadIdx.x; i < NK; i += KT * EPB) g->cls_of[i] = dm->key_class[i];
    for (int i = (int)threadIdx.x; i < 52; i += KT * EPB) { g->white_of_slot[i] = (signed char)dm->white_of_slot[i]; g->black_after[i] = (signed char)dm->black_after[i]; }
    __syncthreads();
  }
  const int tid = (int)threadIdx.x - slot * KT;
  int lane;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));  // opaque: keeps the compiler from re-reading S2R in loops
  // warp <-> hand; swizzled by block parity so that the warps of the (usually busier) right hand do not
  // all land on the same two warp schedulers of an SM
  const int h = HANDS ? ((tid >> 5) ^ (EPB > 1 ? ((slot >> 1) & 1) : (int)(blockIdx.x & 1))) : 0;
  // block-wide phase barriers every thread of the block has to pass: one per iteration of the substep loop
  const int nsub_loop = a.mode == RP_MODE_RESET ? 0 : a.nsub;
  const int eidx = (int)blockIdx.x * EPB + slot;
  if (eidx >= a.nenv) {
    if (EPB > 1) for (int s2 = 0; s2 <= nsub_loop; s2++) if (s2 % RP_PHASE_EVERY == 0) phase_sync<EPB>(slot);
    return;
  }
  // longest-expected-first launch order (rp_order_kernel) shortens the tail of the last wave
  const int env = a.order ? a.order[eidx] : eidx;
  const int nvtot = NK + (HANDS ? 2 * NL : 0);
  const int nu = dm->nu;
  double* __restrict__ gq = a.qpos + (size_t)env * nvtot;
  double* __restrict__ gv = a.qvel + (size_t)env * nvtot;
  double* __restrict__ gw = a.qacc_ws + (size_t)env * nvtot;
  double* __restrict__ gctrl = a.ctrl + (size_t)env * nu;
  const bool hl = HANDS && lane < NL;  // lane owns a hand link
  const int lc = hl ? lane : 0;        // column index that is always inside the hand matrix

  // ------------------------------------------------------------------ task prologue (full step)
  int t_idx = 0;
  bool do_reset = false;
  if (a.mode == RP_MODE_STEP) {
    t_idx = a.t_idx[env];
    do_reset = (a.step_type_state[env] == RP_STEP_LAST_);
  } else if (a.mode == RP_MODE_RESET) {
    do_reset = a.reset_mask == nullptr ? true : (a.reset_mask[env] != 0);
    if (!do_reset) {
      if (EPB > 1) for (int s2 = 0; s2 <= nsub_loop; s2++) if (s2 % RP_PHASE_EVERY == 0) phase_sync<EPB>(slot);
      return;
    }
  }
  const int act_dim = dm->act_dim;
  double sustain_prev = 0;  // thread 0: sustain state before this step's action (MIDI sustain events)
  // controls of this step -> shared memory (and the persistent ctrl state)
  if (do_reset) {
    for (int j = tid; j < nu; j += KT) { sm.ctrl[j] = 0; gctrl[j] = 0; }
    t_idx = 0;
  } else if (a.mode == RP_MODE_STEP) {
    // before_step: ctrl <- action[:-1] verbatim (clamping happens in the actuator model), sustain <- action[-1]
    const double* act = a.action + (size_t)env * act_dim;
    for (int j = tid; j < nu; j += KT) { const double c = act[j]; sm.ctrl[j] = c; gctrl[j] = c; }
    if (tid == 0) {
      sustain_prev = a.sustain[env];
      a.sustain[env] = act[act_dim - 1];
    }
  } else if (a.mode == RP_MODE_PHYSICS && a.ctrl_in != nullptr) {
    const double* c = a.ctrl_in + (size_t)env * nu;
    for (int j = tid; j < nu; j += KT) { const double cv = c[j]; sm.ctrl[j] = cv; gctrl[j] = cv; }
  } else {
    for (int j = tid; j < nu; j += KT) sm.ctrl[j] = gctrl[j];
  }
  if (do_reset && tid == 0) { a.t_idx[env] = 0; a.sustain[env] = 0; }
  // randomize_hand_positions (piano_with_shadow_hands.py:491-499): at every episode start both hands' attachment
  // frames are shifted along y by one draw from U(-0.05, 0.05).  The reference shifts the frames of a model it does
  // not recompile between episodes (suite/__init__.py:56,91), so the shifts accumulate over episodes; so do these.
  // Every thread derives the same value from the environment's generator state; thread 0 stores the new state.
  double hand_off = 0.0;
  if (HANDS && a.hand_offset != nullptr) {
    hand_off = a.hand_offset[env];
    if (do_reset) {
      uint64_t z = a.rng_state[env] + 0x9E3779B97F4A7C15ull;  // splitmix64
      const uint64_t st = z;
      z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
      z ^= z >> 31;
      hand_off = (a.hand_offset_fresh ? 0.0 : hand_off) + (-0.05 + 0.1 * ((double)(z >> 11) * (1.0 / 9007199254740992.0)));
      env_sync<NW, EPB>(bar_id);  // every thread has read the old state
      if (tid == 0) { a.rng_state[env] = st; a.hand_offset[env] = hand_off; }
    }
  }
  // ---- MIDI event stream (models/piano/midi_module.py:47-98): activation changes per physics substep.
  // The previous activation is the one stored by the last step's epilogue; an episode start clears it.
  const bool midi_on = a.midi_events != nullptr && a.mode != RP_MODE_PHYSICS;
  bool midi_prev[KS];
#pragma unroll
  for (int s = 0; s < KS; s++) {
    const int k = tid + KT * s;
    midi_prev[s] = midi_on && !do_reset && k < NK && a.activation[(size_t)env * NK + k] != 0;
  }
  if (midi_on && tid == 0) {
    if (do_reset) a.midi_count[env] = 0;
    else if (a.mode == RP_MODE_STEP) {
      // the pedal state only changes with the action, so its event belongs to the first substep
      const bool was = sustain_prev >= 0.5, is = a.sustain[env] >= 0.5;
      if (was != is) {
        const int slot = atomicAdd(&a.midi_count[env], 1);
        if (slot < a.midi_cap) a.midi_events[(size_t)env * a.midi_cap + slot] = RP_MIDI_WORD(t_idx * a.nsub + 1, is ? 2 : 3, 0);
      }
    }
  }

  // ------------------------------------------------------------------ load state
  double qk[KS], vk[KS], wk[KS];  // keys tid, tid+KT, ...
#pragma unroll
  for (int s = 0; s < KS; s++) {
    const int k = tid + KT * s;
    const bool ok = k < NK && !do_reset;
    qk[s] = ok ? gq[k] : 0.0;
    vk[s] = ok ? gv[k] : 0.0;
    wk[s] = ok ? gw[k] : 0.0;
  }
  double qh = 0, vh = 0, wh = 0;
  if (hl && !do_reset) {
    qh = gq[NK + h * NL + lane];
    vh = gv[NK + h * NL + lane];
    wh = gw[NK + h * NL + lane];
  }
  // per-lane link constants
  const int lpar = hl ? dm->link_parent[h][lane] : -1;
  const int ldep = hl ? dm->link_depth[h][lane] : -1;
  const int ljt = hl ? dm->link_jtype[h][lane] : 0;
  const int lsub = hl ? dm->link_subtree[h][lane] : 0;
  int flags_init_overflow = 0;
  const int ldc = ldep > 0 ? ldep : 0;
  const int maxd = HANDS ? dm->maxdepth[h] : 0;
  int moff = 0;  // offset of this link's packed mass-matrix row: exclusive prefix sum of depth + 1
  if (HANDS) {
    int v = hl ? ldep + 1 : 0, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += t; }
    moff = incl - v;
    sm.moff[h][lane] = (unsigned char)moff;
    if (__shfl_sync(FULL, incl, 31) > 160) flags_init_overflow = 1;
  }
  // topology tables into shared memory (read in the inner loops of the sweep / contact rows)
  if (HANDS) {
    if (hl) for (int d = 0; d < MAXD; d++) sm.anc[h][lane][d] = (signed char)dm->link_anc[h][lane][d];
    sm.dep[h][lane] = (signed char)ldep;
    sm.sub[h][lane] = (signed char)lsub;
  }
  const double hstep = dm->h;
  const double g3[3] = {dm->gravity[0], dm->gravity[1], dm->gravity[2]};
  const int nsub = (a.mode == RP_MODE_RESET || do_reset) ? 0 : a.nsub;
  double actforce = 0;            // lane = actuator index within the hand
  double key_actforce[KS];
#pragma unroll
  for (int s = 0; s < KS; s++) key_actforce[s] = 0;
  int niter_last = 0, ncon_last = 0, flags = flags_init_overflow ? 4 : 0;
  long long busy = 0;  // cycles this environment spent in its own substeps (phase-barrier waits excluded): predictor of the next step's duration
  const int maxiter = a.iterations > 0 ? a.iterations : dm->iterations;
  const double tol = a.tolerance >= 0 ? a.tolerance : dm->tol;

#ifdef RP_PROFILE
  long long tacc[RP_DBG_NPROF] = {0}, tlast = clock64();
#define TICK(i) { const long long tn = clock64(); tacc[i] += tn - tlast; tlast = tn; }
#else
#define TICK(i)
#endif
  // link world pose (kept for the task epilogue)
  double xp[3] = {0, 0, 0}, xq[4] = {1, 0, 0, 0};
  if (tid == 0) sm.near_uncovered = 0;
  env_sync<NW, EPB>(bar_id);

  // =============================================================================================
  for (int sub = 0; sub <= (EPB > 1 ? nsub_loop : nsub); sub++) {
    if (EPB > 1) {
      if (sub % RP_PHASE_EVERY == 0) phase_sync<EPB>(slot);  // phase barrier: the environments of the group start the substep together
      if (sub > nsub) continue;  // this environment is done (episode reset: no physics), only keeps the barrier count
    }
    const long long t_sub = clock64();
    const bool last_pass = (sub == nsub);  // kinematics-only pass for the task epilogue / reset
    if (last_pass && a.mode == RP_MODE_PHYSICS) break;
    // ---------------------------------------------------------------- hand kinematics (by level)
    double xanchor[3] = {0, 0, 0}, xaxis[3] = {0, 0, 1};
    if (HANDS) {
      xp[0] = xp[1] = xp[2] = 0;
      xq[0] = 1; xq[1] = xq[2] = xq[3] = 0;
      // half-angle sine / cosine of the lane's own hinge, once (the level loop below runs its whole body
      // for every level: anything lane-local that is hoisted out of it is executed once instead of 9 times)
      double hsn = 0, hcs = 1;
      if (hl && ljt != 2) sincos(0.5 * qh, &hsn, &hcs);
      // the lane's link constants, fetched once in front of the level loop: inside it they were a fresh round trip to
      // L2 at every level (the model constants live in global memory and the L1 left over by 225 KB of shared memory
      // does not keep them)
      double lp[3] = {0, 0, 0}, lq[4] = {1, 0, 0, 0}, jp[3] = {0, 0, 0}, ja[3] = {0, 0, 1};
      if (hl) {
#pragma unroll
        for (int k = 0; k < 3; k++) { lp[k] = dm->link_pos[h][lane][k]; jp[k] = dm->link_jpos[h][lane][k]; ja[k] = dm->link_jaxis[h][lane][k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) lq[k] = dm->link_quat[h][lane][k];
      }
      for (int d = 0; d <= maxd; d++) {
        const int src = lpar >= 0 ? lpar : lane;
        double pp[3], pq[4];
#pragma unroll
        for (int k = 0; k < 3; k++) pp[k] = shfl(xp[k], src);
#pragma unroll
        for (int k = 0; k < 4; k++) pq[k] = shfl(xq[k], src);
        if (hl && ldep == d) {
          if (lpar < 0) { pp[0] = pp[1] = pp[2] = 0; pq[0] = 1; pq[1] = pq[2] = pq[3] = 0; }
          double pos[3], quat[4], tmp[3];
          rotq(tmp, pq, lp);
          for (int k = 0; k < 3; k++) pos[k] = pp[k] + tmp[k];
          if (lpar < 0) pos[1] += hand_off;
          rp::quat_mul(quat, pq, lq);
          rotq(tmp, quat, jp);
          for (int k = 0; k < 3; k++) xanchor[k] = pos[k] + tmp[k];
          rotq(xaxis, quat, ja);
          if (ljt == 2) {
            for (int k = 0; k < 3; k++) pos[k] += xaxis[k] * qh;
          } else {
            const double sn = hsn, cs = hcs;
            double ql[4] = {cs, ja[0] * sn, ja[1] * sn, ja[2] * sn}, qn[4];
            rp::quat_mul(qn, quat, ql);
            for (int k = 0; k < 4; k++) quat[k] = qn[k];
            rotq(tmp, quat, jp);
            for (int k = 0; k < 3; k++) pos[k] = xanchor[k] - tmp[k];
          }
          rp::quat_normalize(quat);
          for (int k = 0; k < 3; k++) xp[k] = pos[k];
          for (int k = 0; k < 4; k++) xq[k] = quat[k];
        }
      }
    }
    // key angles (and their sin / cos) of this substep
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int k = tid + KT * s;
      if (k < NK) {
        double sn, cs;
        sincos(qk[s], &sn, &cs);
        sm.v.pre.vk[k] = vk[s]; sm.v.pre.ksn[k] = sn; sm.v.pre.kcs[k] = cs;
      }
    }
    if (last_pass) {
      if (EPB == 1) break;
      continue;  // an episode-reset environment among stepping ones
    }

    TICK(0)
    // ------------------------------------------------------------------ publish S, velocities
    double Sown[6] = {0, 0, 0, 0, 0, 0};
    if (hl) {
      if (ljt == 3) {
        for (int k = 0; k < 3; k++) Sown[k] = xaxis[k];
        rp::cross3(&Sown[3], xanchor, xaxis);
      } else {
        for (int k = 0; k < 3; k++) { Sown[k] = 0; Sown[3 + k] = xaxis[k]; }
      }
      for (int k = 0; k < 6; k++) sm.v.pre.S[h][lane][k] = Sown[k];
      sm.v.pre.vh[h][lane] = vh;
    }
    __syncwarp();

    // --------------------------------------------------- RNE forward (chain walk), inertia, F
    double qfrc_bias = 0;
    if (hl) {
      double V[6] = {0, 0, 0, 0, 0, 0}, Ac[6] = {0, 0, 0, 0, 0, 0};
      for (int d = 0; d <= ldep; d++) {
        const int j = sm.anc[h][lane][d];
        double Sj[6], Sd[6];
        for (int k = 0; k < 6; k++) Sj[k] = sm.v.pre.S[h][j][k];
        const double qd = sm.v.pre.vh[h][j];
        rp::cross_motion(Sd, V, Sj);
        for (int k = 0; k < 6; k++) { Ac[k] += Sd[k] * qd; V[k] += Sj[k] * qd; }
      }
      const double mass = dm->link_mass[h][lane];
      double xm[9], tmp[3], c[3], iq[4], R[9], I[10];
      rp::quat2mat(xm, xq);
      double ip[3] = {dm->link_ipos[h][lane][0], dm->link_ipos[h][lane][1], dm->link_ipos[h][lane][2]};
      rp::mat_vec(tmp, xm, ip);
      for (int k = 0; k < 3; k++) c[k] = xp[k] + tmp[k];
      double liq[4] = {dm->link_iquat[h][lane][0], dm->link_iquat[h][lane][1], dm->link_iquat[h][lane][2], dm->link_iquat[h][lane][3]};
      rp::quat_mul(iq, xq, liq);
      rp::quat2mat(R, iq);
      double di[3] = {dm->link_inertia[h][lane][0], dm->link_inertia[h][lane][1], dm->link_inertia[h][lane][2]};
      rp::body_inertia_world(I, mass, di, R, c);
      // F = I*A + V x* (I*V) + (1-gravcomp) * I*[0;-g]
      double IA[6], IV[6], VxIV[6];
      rp::inertia_mul(IA, I, Ac);
      rp::inertia_mul(IV, I, V);
      rp::cross_force(VxIV, V, IV);
      const double gc = 1.0 - dm->link_gravcomp[h][lane];
      double Ag[6] = {0, 0, 0, -g3[0] * gc, -g3[1] * gc, -g3[2] * gc}, IAg[6];
      rp::inertia_mul(IAg, I, Ag);
      for (int k = 0; k < 6; k++) sm.u.acc[h][lane][k] = IA[k] + VxIV[k] + IAg[k];
      for (int k = 0; k < 10; k++) sm.u.acc[h][lane][6 + k] = I[k];
    }
    __syncwarp();
    // upward accumulation by level
    if (HANDS) {
      const int crank = hl ? dm->link_childrank[h][lane] : -1;
      for (int d = maxd; d >= 1; d--) {
        const int nslots = dm->link_nchild_slots[h][d];
        for (int r = 0; r < nslots; r++) {
          if (hl && ldep == d && crank == r) {
            for (int k = 0; k < 16; k++) sm.u.acc[h][lpar][k] += sm.u.acc[h][lane][k];
          }
          __syncwarp();
        }
      }
    }
    // bias force, mass-matrix rows
    if (hl) {
      double F[6], Ic[10], fI[6];
      for (int k = 0; k < 6; k++) F[k] = sm.u.acc[h][lane][k];
      for (int k = 0; k < 10; k++) Ic[k] = sm.u.acc[h][lane][6 + k];
      qfrc_bias = rp::dot6(Sown, F);
      rp::inertia_mul(fI, Ic, Sown);
      for (int d = 0; d <= ldep; d++) {
        const int j = sm.anc[h][lane][d];
        double Sj[6];
        for (int k = 0; k < 6; k++) Sj[k] = sm.v.pre.S[h][j][k];
        double m = rp::dot6(Sj, fI);
        if (d == ldep) m += dm->link_armature[h][lane];
        sm.Mrow[h][moff + d] = m;
      }
    }
    // scatter M into the dense symmetric work matrix
    if (HANDS) {
      double* Aw = sm.A[h];
      for (int idx = lane; idx < NL * RS; idx += 32) Aw[idx] = 0;
      __syncwarp();
      if (hl) {
        for (int d = 0; d <= ldep; d++) {
          const int j = sm.anc[h][lane][d];
          const double m = sm.Mrow[h][moff + d];
          Aw[lane * RS + j] = m;
          Aw[j * RS + lane] = m;
        }
      }
      __syncwarp();
      if (a.dbg && sub == 0 && h == 0) {
        double* dbg = a.dbg + (size_t)env * RP_DBG_STRIDE;
        for (int i = 0; i < NL; i++) if (hl) dbg[RP_DBG_M + i * NL + lane] = Aw[i * RS + lane];
      }
    TICK(1)
      // ---------------------------------------------------------------- invert M (symmetric sweep): A = -M^-1
      sweep_inverse(Aw, sched, lane);
      if (a.dbg && sub == 0 && h == 0) {
        double* dbg = a.dbg + (size_t)env * RP_DBG_STRIDE;
        for (int i = 0; i < NL; i++) if (hl) dbg[RP_DBG_MINV + i * NL + lane] = -Aw[i * RS + lane];
      }
    }
    const double* Ah = sm.A[h];

    TICK(2)
    // ------------------------------------------------------------------ actuation (lane = actuator)
    if (HANDS) {
      const bool isact = lane < dm->nact[h];
      const int d0 = isact ? dm->act_dof0[h][lane] : 0, d1 = isact ? dm->act_dof1[h][lane] : -1;
      const double q0 = shfl(qh, d0), v0 = shfl(vh, d0);
      const double q1 = shfl(qh, d1 >= 0 ? d1 : 0), v1 = shfl(vh, d1 >= 0 ? d1 : 0);
      double force = 0;
      if (isact) {
        const double c0 = dm->act_coef0[h][lane], c1 = d1 >= 0 ? dm->act_coef1[h][lane] : 0.0;
        const double len = c0 * q0 + c1 * q1, vel = c0 * v0 + c1 * v1;
        double ctrl = sm.ctrl[h * dm->nact[0] + lane];
        if (dm->act_ctrllimited[h][lane]) ctrl = fmin(fmax(ctrl, dm->act_ctrlrange[h][lane][0]), dm->act_ctrlrange[h][lane][1]);
        force = dm->act_gain[h][lane] * ctrl + dm->act_bias[h][lane][0] + dm->act_bias[h][lane][1] * len + dm->act_bias[h][lane][2] * vel;
        if (dm->act_forcelimited[h][lane]) force = fmin(fmax(force, dm->act_forcerange[h][lane][0]), dm->act_forcerange[h][lane][1]);
        sm.actf[h][lane] = force;
      }
      actforce = force;
      __syncwarp();
    }
    // ---------------------------------------------------- smooth forces and accelerations (hands)
    double qfrc_smooth = 0, qacc_smooth = 0;
    if (HANDS) {
      if (hl) {
        const double passive = -dm->link_stiffness[h][lane] * (qh - dm->link_springref[h][lane]) - dm->link_damping[h][lane] * vh;
        const int ai = dm->dof_act[h][lane];
        const double qa = ai >= 0 ? dm->dof_act_coef[h][lane] * sm.actf[h][ai] : 0.0;
        qfrc_smooth = passive - qfrc_bias + qa;
        sm.wbuf[h][lane] = qfrc_smooth;
      }
      __syncwarp();
      if (hl) {
        double s = 0;
        for (int k = 0; k < NL; k++) s -= Ah[k * RS + lane] * sm.wbuf[h][k];
        qacc_smooth = s;
      }
      __syncwarp();
    }
    // ------------------------------------------------------------------ keys: smooth dynamics + limit rows
    double ks_qacc[KS], ks_qfrc[KS], kM[KS];
    KeyRow krow[KS];
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int k = tid + KT * s;
      krow[s].active = 0; krow[s].f = 0; krow[s].J = 0; krow[s].R = 0; krow[s].b = 0; krow[s].ARinv = 0; krow[s].AR = 1; krow[s].aref = 0;
      ks_qacc[s] = 0; ks_qfrc[s] = 0; kM[s] = 1; key_actforce[s] = 0;
      if (k < NK) {
        double ax[3] = {dm->key_axis[k][0], dm->key_axis[k][1], dm->key_axis[k][2]};
        double r0[3] = {dm->key_ipos[k][0] - dm->key_jpos[k][0], dm->key_ipos[k][1] - dm->key_jpos[k][1], dm->key_ipos[k][2] - dm->key_jpos[k][2]};
        double r[3], mg[3], t[3];
        rot_axis(r, ax, sm.v.pre.ksn[k], sm.v.pre.kcs[k], r0);
        const double mass = dm->key_mass[k];
        for (int c = 0; c < 3; c++) mg[c] = mass * g3[c];
        rp::cross3(t, r, mg);
        const double bias = -rp::dot3(ax, t);
        const double passive = -dm->key_stiffness[k] * (qk[s] - dm->key_springref[k]) - dm->key_damping[k] * vk[s];
        double actq = 0;
        if (dm->nkey_act) {
          const double ctrl = fmin(fmax(sm.ctrl[k], dm->key_act_range[k][0]), dm->key_act_range[k][1]);
          actq = dm->key_act_gain[k] * ctrl;
          key_actforce[s] = actq;
        }
        kM[s] = dm->key_M[k];
        const double kq = passive - bias + actq;
        ks_qacc[s] = kq / kM[s];
        ks_qfrc[s] = kq;
        sm.keyslot[k] = -1;
        const double lo = dm->key_range[k][0], hi = dm->key_range[k][1];
        const double dlo = qk[s] - lo, dhi = hi - qk[s];
        double J = 0, pos = 0;
        if (dlo < 0) { J = 1; pos = dlo; krow[s].active = 1; }
        else if (dhi < 0) { J = -1; pos = dhi; krow[s].active = 1; }
        if (krow[s].active) {
          double solimp[5];
          for (int c = 0; c < 5; c++) solimp[c] = dm->key_lim_solimp[k][c];
          const double imp = rp::impedance(solimp, pos, 0.0);
          const double R = fmax(RP_MINVAL, (1 - imp) * dm->key_dof_invw[k] / imp);
          const double aref = -dm->key_lim_B[k] * (J * vk[s]) - dm->key_lim_K[k] * imp * pos;
          krow[s].J = J; krow[s].R = R; krow[s].aref = aref;
          krow[s].b = J * ks_qacc[s] - aref;
          krow[s].AR = 1.0 / kM[s] + R;
          krow[s].ARinv = 1.0 / krow[s].AR;
        }
      }
    }
    // ------------------------------------------------------------------ hand friction + limit rows
    double fr_R = 1, fr_b = 0, fr_AR = 1, fr_f = 0, fr_loss = 0, fr_aref = 0;
    double lm_J = 0, lm_R = 1, lm_b = 0, lm_AR = 1, lm_f = 0, lm_aref = 0;
    bool fr_on = false, lm_on = false;
    if (hl) {
      const double mii = -Ah[lane * RS + lane];
      const double loss = dm->link_frictionloss[h][lane];
      fr_on = loss > 0;
      if (fr_on) {
        fr_loss = loss;
        fr_R = dm->link_fr_R[h][lane];
        fr_aref = -dm->link_fr_B[h][lane] * vh;
        fr_b = qacc_smooth - fr_aref;
        fr_AR = mii + fr_R;
      }
      if (dm->link_limited[h][lane]) {
        const double lo = dm->link_range[h][lane][0], hi = dm->link_range[h][lane][1];
        const double dlo = qh - lo, dhi = hi - qh;
        double pos = 0;
        if (dlo < 0) { lm_J = 1; pos = dlo; lm_on = true; }
        else if (dhi < 0) { lm_J = -1; pos = dhi; lm_on = true; }
        if (lm_on) {
          double solimp[5];
          for (int c = 0; c < 5; c++) solimp[c] = dm->link_lim_solimp[h][lane][c];
          const double imp = rp::impedance(solimp, pos, 0.0);
          lm_R = fmax(RP_MINVAL, (1 - imp) * dm->link_dof_invw[h][lane] / imp);
          lm_aref = -dm->link_lim_B[h][lane] * (lm_J * vh) - dm->link_lim_K[h][lane] * imp * pos;
          lm_b = lm_J * qacc_smooth - lm_aref;
          lm_AR = mii + lm_R;
        }
      }
    }
    const unsigned lm_mask = HANDS ? __ballot_sync(FULL, lm_on) : 0u;

    TICK(3)
    // ------------------------------------------------------------------ collision (lane = geom of this warp's hand)
    if (tid == 0) { sm.nraw = 0; sm.ncon = 0; sm.overflow = 0; sm.nck = 0; sm.shared_key = 0; sm.nhc[0] = 0; sm.nhc[H1] = 0; sm.npairc = 0; }
    env_sync<NW, EPB>(bar_id);  // key arrays published; acc scratch of both hands dead (it aliases the contact lists)
    TICK(4)
    if (HANDS) {
      const bool isg = lane < dm->ngeom[h];
      const int gl = isg ? dm->geom_link[h][lane] : 0;
      double lp[3], lq[4];
#pragma unroll
      for (int k = 0; k < 3; k++) lp[k] = shfl(xp[k], gl);
#pragma unroll
      for (int k = 0; k < 4; k++) lq[k] = shfl(xq[k], gl);
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
        // white keys whose y extent (unchanged by the rotation about their y-parallel hinge: lattice_slack is their half
        // width plus the largest margin; 0.02 m when some hinge is not y-parallel) overlaps the geom's bounding sphere
        i0 = (int)floor((gc[1] - rb - dm->lattice_slack - dm->lattice_y0) / dm->lattice_pitch);
        i1 = (int)ceil((gc[1] + rb + dm->lattice_slack - dm->lattice_y0) / dm->lattice_pitch);
        i0 = max(i0, 0); i1 = min(i1, 51);
        ncand = (i1 >= i0) ? 2 * (i1 - i0 + 1) + 1 : 0;
      }
      // Two passes.  Broadphase: every lane walks the candidate boxes of its geom (warp-uniform trip count) and only tests
      // the bounding sphere against the box in the box's frame; the pairs that pass go to a per-hand list (wbuf is
      // free between the smooth-acceleration product and the publication of qacc_smooth).  Narrowphase: one lane per
      // listed pair — the expensive part (segment / box closest point, up to three sphere / box tests) runs once per
      // 32 pairs instead of once per candidate round in which some lane passed.
      // Box frame.  Base: axis aligned at base_pos.  Key: the geom sits at the body origin with the body's orientation,
      // the body is turned by q about the hinge axis a through anchor = key_pos + jpos, so a world point p maps to
      // Rot(a, -q)(p - anchor) + jpos — one Rodrigues rotation per vector; contacts map back with Rot(a, +q).
      unsigned short* plist = reinterpret_cast<unsigned short*>(sm.wbuf[h]);  // 128 entries: geom lane | (key + 1) << 8
      int npair = 0;
      const int ncand_max = __reduce_max_sync(FULL, isg ? ncand : 0);
      const unsigned ltmask = (1u << lane) - 1u;
      auto box_frame = [&](int key, const double* gcw, double* c, double* bhalf, double* kax, double* jp, double* anchor,
                           double& sn, double& cs) -> double {  // returns the box's margin
        if (key < 0) {
          for (int k = 0; k < 3; k++) { anchor[k] = dm->base_pos[k]; bhalf[k] = dm->base_half[k]; c[k] = gcw[k] - anchor[k]; kax[k] = k == 2; jp[k] = 0; }
          sn = 0; cs = 1;
          return dm->base_gp.margin;
        }
        sn = sm.v.pre.ksn[key]; cs = sm.v.pre.kcs[key];
        double rel[3];
        // class geometry from the block-shared table (or per key from the model when the classes are not uniform)
        const double* kc = dm->keygeo_ok ? geo->cls[geo->cls_of[key]] : nullptr;
        for (int k = 0; k < 3; k++) {
          kax[k] = kc ? kc[6 + k] : dm->key_axis[key][k]; jp[k] = kc ? kc[k] : dm->key_jpos[key][k];
          bhalf[k] = kc ? kc[3 + k] : dm->key_half[key][k];
          anchor[k] = geo->pos[key][k] + jp[k];
          rel[k] = gcw[k] - anchor[k];
        }
        rot_axis(c, kax, -sn, cs, rel);
        for (int k = 0; k < 3; k++) c[k] += jp[k];
        return kc ? kc[9] : dm->key_gp[key].margin;
      };
      for (int ci = -1; ci < ncand_max; ci++) {
        // candidate -> piano box: ci = -1 base; 2k / 2k+1 = white / following black of slot i0+k;
        // last = the black key before the first slot
        int key = -2;
        if (isg && ci < ncand) {
          if (ci < 0) key = dm->has_base ? -1 : -2;
          else if (ci == ncand - 1) key = i0 > 0 ? geo->black_after[i0 - 1] : -1, key = key < 0 ? -2 : key;
          else {
            const int slot = i0 + (ci >> 1);
            key = (ci & 1) ? geo->black_after[slot] : geo->white_of_slot[slot];
            key = key < 0 ? -2 : key;
          }
        }
        bool pass = false;
        if (key > -2) {
          double c[3], bhalf[3], kax[3], jp[3], anchor[3], sn, cs;
          const double margin = fmax(gmargin, box_frame(key, gc, c, bhalf, kax, jp, anchor, sn, cs));
          const double bound = rb + margin;
          pass = rp::point_box_d2(c, bhalf) <= bound * bound;
        }
        const unsigned pm = __ballot_sync(FULL, pass);
        if (pass) {
          const int slot = npair + __popc(pm & ltmask);
          if (slot < 128) plist[slot] = (unsigned short)(lane | ((key + 1) << 8));
        }
        npair += __popc(pm);
      }
      if (npair > 128) { npair = 128; sm.overflow = 1; }  // more close pairs than the list holds: reported like a contact overflow
      __syncwarp();
      for (int p0 = 0; p0 < npair; p0 += 32) {
        const bool act = p0 + lane < npair;
        const unsigned e = act ? plist[p0 + lane] : 0u;
        const int g = e & 0xff, key = (int)(e >> 8) - 1;
        // the pair's geom lives in lane g
        double pgc[3], pgax[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { pgc[k] = shfl(gc[k], g); pgax[k] = shfl(gax[k], g); }
        const double prad = shfl(rad, g), phalf = shfl(half, g), pgmargin = shfl(gmargin, g);
        const int pgtype = __shfl_sync(FULL, gtype, g), pgorder = __shfl_sync(FULL, gorder, g);
        if (act) {
          double c[3], bhalf[3], kax[3], jp[3], anchor[3], sn, cs;
          const double margin = fmax(pgmargin, box_frame(key, pgc, c, bhalf, kax, jp, anchor, sn, cs));
          const bool iskey = key >= 0;
          // narrowphase: sphere, or capsule end points + interior closest point
          int ntest = 1;
          double pts[3][3];
          if (pgtype == 2) {
            for (int k = 0; k < 3; k++) { pts[0][k] = c[k]; pts[1][k] = c[k]; pts[2][k] = c[k]; }
          } else {
            double ax[3] = {pgax[0], pgax[1], pgax[2]};
            if (iskey) rot_axis(ax, kax, -sn, cs, pgax);
            for (int k = 0; k < 3; k++) { pts[0][k] = c[k] - phalf * ax[k]; pts[1][k] = c[k] + phalf * ax[k]; pts[2][k] = c[k]; }
            ntest = 2;
            const double t = rp::segment_box_closest_t(pts[0], pts[1], bhalf);
            if (t > 1e-9 && t < 1 - 1e-9) {
              for (int k = 0; k < 3; k++) pts[2][k] = pts[0][k] + t * (pts[1][k] - pts[0][k]);
              ntest = 3;
            }
          }
#pragma unroll 1
          for (int ti = 0; ti < ntest; ti++) {
            double n[3], p[3];
            const double dist = rp::sphere_box(pts[ti], prad, bhalf, n, p);
            if (dist < margin) {
              const int slot = atomicAdd(&sm.nraw, 1);
              if (slot < (HC ? RAWCAP : MAXCON)) {
                RawContact& rc = sm.u.col.raw[slot];  // (same storage as u.raw)
                double nw[3] = {n[0], n[1], n[2]}, pw[3] = {p[0], p[1], p[2]};
                if (iskey) {
                  double pr[3] = {p[0] - jp[0], p[1] - jp[1], p[2] - jp[2]};
                  rot_axis(nw, kax, sn, cs, n);
                  rot_axis(pw, kax, sn, cs, pr);
                }
                rc.dist = dist;
                for (int k = 0; k < 3; k++) { rc.n[k] = nw[k]; rc.pos[k] = anchor[k] + pw[k]; }
                rc.hand = h; rc.geom = g; rc.key = key; rc.sub = ti;
                // MuJoCo order: piano geom (base, then keys) major, hand geom (model order) minor
                rc.sortkey = (unsigned)(((key + 1) * 64 + pgorder) * 4 + ti);
              } else {
                sm.overflow = 1;
              }
            }
          }
        }
      }
      __syncwarp();
      if (HC && isg) {  // world-frame primitive for the hand : hand pass
        double* w = sm.u.col.gw[h][lane];
        for (int k = 0; k < 3; k++) { w[k] = gc[k]; w[3 + k] = gax[k]; }
        w[6] = rad; w[7] = half;
      }
    }
    if (HC && dm->hc_enabled) {
      // ---- hand geom : hand geom pairs (self and cross hand; host-filtered list in MuJoCo's pair order), one thread
      // per pair: bounding spheres, then the analytic sphere / capsule routines.  A contact across the hands takes two
      // adjacent raw entries (one per body), a contact within a hand one; all sorted behind every hand : piano contact.
      env_sync<NW, EPB>(bar_id);
      const int npairs = dm->hc_npair;
      for (int p = tid; p < npairs; p += KT) {
        const unsigned e = dm->hc_pairs[p];
        int a1 = e & 0xff, a2 = e >> 8;
        const double* w1 = &sm.u.col.gw[0][0][0] + 8 * a1;
        const double* w2 = &sm.u.col.gw[0][0][0] + 8 * a2;
        if (w1[7] > 0 && w2[7] == 0) { const double* t = w1; w1 = w2; w2 = t; const int ta = a1; a1 = a2; a2 = ta; }  // sphere before capsule
        const int h1 = a1 / RP_NGH, g1 = a1 - h1 * RP_NGH, h2 = a2 / RP_NGH, g2 = a2 - h2 * RP_NGH;
        const double margin = fmax(dm->geom_gp[h1][g1].margin, dm->geom_gp[h2][g2].margin);
        const double dx = w2[0] - w1[0], dy = w2[1] - w1[1], dz = w2[2] - w1[2], bound = w1[6] + w1[7] + w2[6] + w2[7] + margin;
        if (dx * dx + dy * dy + dz * dz > bound * bound) continue;
        double dist[2], nrm[6], pos[6];
        const int nc = rp::sphere_capsule_pair(margin, w1, w1 + 3, w1[6], w1[7], w2, w2 + 3, w2[6], w2[7], dist, nrm, pos);
        for (int i = 0; i < nc; i++) {
          if (!(dist[i] < margin)) continue;
          const int cross = h1 != h2;
          const int slot = atomicAdd(&sm.nraw, 1 + cross);
          if (slot + cross < RAWCAP) {
            if (cross) atomicAdd(&sm.npairc, 1);
            for (int side = 0; side <= cross; side++) {
              RawContact& rc = sm.u.col.raw[slot + side];
              rc.dist = dist[i];
              for (int k = 0; k < 3; k++) { rc.n[k] = nrm[3 * i + k]; rc.pos[k] = pos[3 * i + k]; }
              rc.hand = side ? h2 : h1; rc.geom = side ? g2 : g1; rc.key = cross ? -2 - side : -4;
              rc.sub = side ? h1 * 32 + g1 : h2 * 32 + g2;
              rc.sortkey = 32768u + (unsigned)((p * 2 + i) * 2 + side);
            }
          } else {
            sm.overflow = 1;
            if (slot < RAWCAP) {  // the odd slot left over: an inert record (the environment is flagged as overflowed)
              RawContact& rc = sm.u.col.raw[slot];
              rc.dist = 1.0; rc.n[0] = 0; rc.n[1] = 0; rc.n[2] = 1;
              for (int k = 0; k < 3; k++) rc.pos[k] = w1[k];
              rc.hand = h1; rc.geom = g1; rc.key = -1; rc.sub = 0; rc.sortkey = 32768u + (unsigned)((p * 2 + i) * 2);
            }
          }
        }
      }
    }
    if (HANDS) {
      // colliders outside the narrowphase's coverage (palm / knuckle / wrist boxes and cylinders, forearm): if the
      // bounding box of one comes within reach of the piano the step is flagged (diag flag 8) — they are never
      // dropped silently.  (They stay 5 cm clear of the piano on the recorded replay and cannot reach it while the joints
      // respect their limits: see the tests.)
      const bool isi = lane < dm->nign[h];
      const int il = isi ? dm->ign_link[h][lane] : 0;
      double lp[3], lq[4];
#pragma unroll
      for (int k = 0; k < 3; k++) lp[k] = shfl(xp[k], il);
#pragma unroll
      for (int k = 0; k < 4; k++) lq[k] = shfl(xq[k], il);
      if (isi) {
        double tmp[3], gp[3] = {dm->ign_pos[h][lane][0], dm->ign_pos[h][lane][1], dm->ign_pos[h][lane][2]};
        double Rl[9];
        rp::quat2mat(Rl, lq);
        rp::mat_vec(tmp, Rl, gp);
        bool near = true;
        for (int k = 0; k < 3; k++) {
          double e = 0;  // half extent of the box along world axis k: sum_j |(Rl Rg)_kj| half_j
          for (int j = 0; j < 3; j++) {
            const double rkj = Rl[3 * k] * dm->ign_mat[h][lane][j] + Rl[3 * k + 1] * dm->ign_mat[h][lane][3 + j] + Rl[3 * k + 2] * dm->ign_mat[h][lane][6 + j];
            e += fabs(rkj) * dm->ign_half[h][lane][j];
          }
          near = near && (lp[k] + tmp[k] + e > dm->piano_lo[k]) && (lp[k] + tmp[k] - e < dm->piano_hi[k]);
        }
        if (near) atomicOr(&sm.near_uncovered, 1);
      }
    }
    TICK(5)
    // publish what the contact rows read from the other warp: qacc_smooth / qacc_warmstart of the hand
    // dofs (wbuf / wsb) and of the keys (vak / kfc; kfc is free until the end of the sweep)
    if (hl) { sm.wbuf[h][lane] = qacc_smooth; sm.v.pre.wsb[h][lane] = wh; }
#pragma unroll
    for (int s = 0; s < KS; s++) { const int k = tid + KT * s; if (k < NK) { sm.vak[k] = ks_qacc[s]; sm.kfc[k] = wk[s]; } }
    env_sync<NW, EPB>(bar_id);
    TICK(4)
    const int nraw = min(sm.nraw, HC ? RAWCAP : MAXCON);
    if (sm.nraw > (HC ? RAWCAP : MAXCON) || sm.overflow) flags |= 1;
    // NOTE: beyond the raw capacity the kept subset depends on atomic order; the flag marks the env invalid.
    int ncon = nraw, npairc = HC ? sm.npairc : 0;
    // Hand-collision instantiation: more raw records than contact records (hands flailing over their whole joint ranges can
    // need 30).  The deepest are kept — ordered by (distance in nm, sort key), a hand : hand contact's two records as a unit — the
    // shallow ones dropped, and the environment is flagged (bit 0) all the same.  Deterministic, unlike the atomics' order.
    const int reccap = (HC && a.record_cap > 0 && a.record_cap < MAXCON) ? a.record_cap : MAXCON;
    const bool select = HC && nraw > reccap;
    signed char* const keepf = &sm.hlist[0][0];  // RAWCAP <= 2 * MAXCON bytes, free until the contact lists are built
    static_assert(!HC || (RAWCAP <= RP_NHMAX * MAXCON && RAWCAP <= KT), "keep flags live in hlist; one thread per raw record");
    if (select) {
      flags |= 1;
      if (tid < nraw) {
        // depth in whole nanometres: geometrically equal distances (two capsules of one finger meeting at a joint, seen
        // from a neighbouring finger) differ in their last bits, and must not be ordered by them
        const RawContact& me = sm.u.col.raw[tid];
        const long long myq = (long long)floor(me.dist * 1e9);
        int rank = 0;
        for (int j = 0; j < nraw; j++) {
          const RawContact& o = sm.u.col.raw[j];
          const long long oq = (long long)floor(o.dist * 1e9);
          rank += (oq < myq || (oq == myq && o.sortkey < me.sortkey)) ? 1 : 0;
        }
        keepf[tid] = (signed char)((me.key == -2 ? rank + 1 : rank) < reccap ? (me.key == -2 ? 2 : 1) : 0);  // 2: kept first record of a cross-hand pair
      }
      env_sync<NW, EPB>(bar_id);
      ncon = 0; npairc = 0;
      for (int j = 0; j < nraw; j++) { ncon += keepf[j] != 0; npairc += keepf[j] == 2; }
    }
    ncon_last = HC ? ncon - npairc : ncon;  // a hand : hand contact takes two records
    // ---- contact rows (rank-sorted into con[]): one thread per (contact, frame row), spread over both warps
    // (RP_MAXCON = 21 contacts = 63 rows)
    static_assert(!HANDS || 3 * MAXCON <= KT, "one thread per contact row");
    if (HANDS && ncon > 0) {
      if (tid < nraw && (!select || keepf[tid])) {
        const unsigned mykey = sm.u.col.raw[tid].sortkey;
        int rank = 0;
        for (int j = 0; j < nraw; j++) rank += (sm.u.col.raw[j].sortkey < mykey && (!select || keepf[j])) ? 1 : 0;
        sm.perm[rank] = (signed char)tid;
      }
      env_sync<NW, EPB>(bar_id);
      const bool rowt = tid < 3 * ncon;
      const int c = rowt ? tid / 3 : 0, r = tid - 3 * c;
      // contact geometry to registers: raw[] and con[] share the same shared memory
      double c_dist = 0, c_pos[3] = {0, 0, 0}, c_frame[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      int ch = 0, c_geom = 0, c_key = -1, c_sub = 0;
      if (rowt) {
        const RawContact rc = sm.u.col.raw[sm.perm[c]];
        c_sub = rc.sub;
        c_dist = rc.dist;
        for (int k = 0; k < 3; k++) { c_pos[k] = rc.pos[k]; c_frame[k] = rc.n[k]; }
        const double nn = sqrt(rp::dot3(c_frame, c_frame));
        for (int k = 0; k < 3; k++) c_frame[k] /= nn;
        rp::make_frame(c_frame);
        ch = rc.hand; c_geom = rc.geom; c_key = rc.key;
      }
      env_sync<NW, EPB>(bar_id);
      double u[MAXD], arefr = 0, mu_row = 0;
      double hc_sv = 0, hc_B = 0, hc_K = 0;  // hand : hand contacts: this side's row velocity, damping and stiffness term
      const bool pairc = HC && c_key < -1;   // record of a hand : hand contact (-2 / -3: cross-hand, geom 1's / geom 2's body; -4: within a hand)
      const bool crossc = HC && (c_key == -2 || c_key == -3);
      int link = 0, dep = 0;
      if (rowt) {
        Contact& cn = sm.u.con[c];
        link = dm->geom_link[ch][c_geom];
        dep = sm.dep[ch][link];
        const double* dir = c_frame + 3 * r;
        const double* Ac = sm.A[ch];
        // chain of the contact's link (dof at depth d), once, into registers: the loops below index shared memory
        // through it and are fully unrolled (an index load in front of every element made this stage the longest
        // serial chain of the contact rows: depth^2 dependent load pairs per row)
        int cidx[MAXD];
        unsigned plus = (HC && c_key == -3) ? 0xffffffffu : 0u;  // chain entries that enter with a plus sign (body 2 of a hand : hand contact)
        if (HC && c_key == -4) {
          // two bodies of one hand: the dofs of their common ancestors move the contact point of both bodies alike — those
          // Jacobian entries cancel exactly — so the record's chain lists body 1's dofs below the branch point (minus), then
          // body 2's (plus): at most 5 + 5 for two fingers of the Shadow Hand
          const int link2 = dm->geom_link[ch][c_sub & 31], dep2 = sm.dep[ch][link2];
          int b = 0;
          while (b <= dep && b <= dep2 && sm.anc[ch][link][b] == sm.anc[ch][link2][b]) b++;
          const int n1 = dep - b + 1, n = min(n1 + dep2 - b + 1, MAXD);
          if (n1 + dep2 - b + 1 > MAXD) sm.overflow = 1;  // deeper than the record holds (not with this hand model): flagged
#pragma unroll
          for (int d = 0; d < MAXD; d++) {
            const int dd = d < n ? d : n - 1;
            const bool first = dd < n1;
            cidx[d] = sm.anc[ch][first ? link : link2][first ? b + dd : b + dd - n1];
            if (!first) plus |= 1u << d;
          }
          dep = n - 1;
        } else {
#pragma unroll
          for (int d = 0; d < MAXD; d++) cidx[d] = sm.anc[ch][link][d <= dep ? d : dep];
        }
        // Jacobian row: frame direction r, body2 (key) minus body1 (hand link)
        double jrow[MAXD];
#pragma unroll
        for (int d = 0; d < MAXD; d++) {
          const double* S = sm.v.pre.S[ch][cidx[d]];
          double wxp[3];
          rp::cross3(wxp, S, c_pos);
          double v = d <= dep ? -(dir[0] * (S[3] + wxp[0]) + dir[1] * (S[4] + wxp[1]) + dir[2] * (S[5] + wxp[2])) : 0.0;
          if (HC && ((plus >> d) & 1u)) v = -v;
          jrow[d] = v;
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
        // mixed parameters, impedance / regularisation (evaluated by each of the three row threads)
        const RpGeomParams& g1 = dm->geom_gp[ch][c_geom];
        const RpGeomParams& g2 = pairc ? dm->geom_gp[c_sub >> 5][c_sub & 31] : (c_key < 0 ? dm->base_gp : dm->key_gp[c_key]);
        double solref[2], solimp[5], mu;
        rp::mix_params(g1.solref, g1.solimp, g1.friction, g1.solmix, g1.priority, g2.solref, g2.solimp, g2.friction,
                       g2.solmix, g2.priority, solref, solimp, &mu);
        const double margin = fmax(g1.margin, g2.margin), gap = fmax(g1.gap, g2.gap), incl = margin - gap;
        const double imp = rp::impedance(solimp, c_dist, incl);
        double K, B;
        rp::ref_KB(solref, solimp[1], hstep, &K, &B);
        const double diag = dm->geom_body_invw[ch][c_geom] + (pairc ? dm->geom_body_invw[c_sub >> 5][c_sub & 31] : (c_key >= 0 ? dm->key_body_invw[c_key] : 0.0));
        const double Rn = fmax(RP_MINVAL, (1 - imp) * diag / imp);
        cn.R[r] = r == 0 ? Rn : Rn / fmax(RP_MINVAL, dm->impratio);
        mu_row = mu;
        // row velocity -> reference acceleration
        double sv = 0;
#pragma unroll
        for (int d = 0; d < MAXD; d++) sv += jrow[d] * sm.v.pre.vh[ch][cidx[d]];  // jrow is zero beyond the chain
        if (c_key >= 0) sv += jk * sm.v.pre.vk[c_key];
        arefr = r == 0 ? -B * sv - K * imp * (c_dist - incl) : -B * sv;
        // u = Minv_chain Jc_r^T (kept in registers for the A block)
#pragma unroll
        for (int d = 0; d < MAXD; d++) {
          const double* rowd = Ac + cidx[d] * RS;
          double s2 = 0;
#pragma unroll
          for (int e = 0; e < MAXD; e++) s2 -= rowd[cidx[e]] * jrow[e];
          u[d] = s2;
        }
        // b = J qacc_smooth - aref; jar = J qacc_warmstart - aref (parked in f[] for the primal map)
        {
          double s2 = 0, s3 = 0;
#pragma unroll
          for (int d = 0; d < MAXD; d++) {
            s2 += jrow[d] * sm.wbuf[ch][cidx[d]];
            s3 += jrow[d] * sm.v.pre.wsb[ch][cidx[d]];
          }
          if (c_key >= 0) { s2 += jk * sm.vak[c_key]; s3 += jk * sm.kfc[c_key]; }
          if (crossc) {  // partial sums of this body's chain; the first record combines both after the barrier
            cn.b[r] = s2; cn.f[r] = s3;
            if (c_key == -3) cn.R[r] = sv;
            hc_sv = sv; hc_B = B; hc_K = r == 0 ? K * imp * (c_dist - incl) : 0.0;
          } else {
            cn.b[r] = s2 - arefr;
            cn.f[r] = s3 - arefr;
          }
        }
        if (r == 0) {
          cn.hand = ch; cn.link = link; cn.key = c_key; cn.mu = mu;
          cn.invMk = c_key >= 0 ? 1.0 / dm->key_M[c_key] : 0.0;
#pragma unroll
          for (int d = 0; d < 16; d++) cn.chain[d] = (signed char)(d < MAXD && d <= dep ? cidx[d < MAXD ? d : 0] : 0);
        }
      }
      env_sync<NW, EPB>(bar_id);
      if (rowt && !crossc) {
        Contact& cn = sm.u.con[c];
        // A[s2][r] = Jc_s2 Minv Jc_r^T + Jk_s2 Jk_r / Mk + R (diagonal)
        for (int s2 = 0; s2 < 3; s2++) {
          double v = 0;
          for (int d = 0; d <= dep; d++) v += cn.Jc[s2][d] * u[d];
          v += cn.Jk[r] * cn.Jk[s2] * cn.invMk;
          if (r == s2) v += cn.R[r];
          cn.A[3 * s2 + r] = v;
        }
      }
      if (HC && rowt && c_key == -2) {
        // cross-hand contact, first record (P) with its partner (S = c + 1).  Reference acceleration from the row velocity
        // of both bodies; b and the warm-start jar from both partial sums; the block from J = [J_P J_S]:
        // A[s2][r] = J_s2 Minv J_r^T with Minv block diagonal per hand — the cross terms P x S only when both bodies belong
        // to the same hand.
        Contact& cn = sm.u.con[c];
        Contact& cs = sm.u.con[c + 1];
        const double sv = hc_sv + cs.R[r];
        const double ar = -hc_B * sv - hc_K;
        cn.b[r] = cn.b[r] + cs.b[r] - ar;
        cn.f[r] = cn.f[r] + cs.f[r] - ar;
        const bool same = cs.hand == ch;
        const double* AP = sm.A[ch];
        const double* AS = sm.A[cs.hand];
        double uP[MAXD], uS[MAXD];
        for (int d = 0; d < MAXD; d++) {
          const int iP = cn.chain[d], iS = cs.chain[d];
          double aP = u[d], aS = 0;
          for (int e = 0; e < MAXD; e++) {
            const int eP = cn.chain[e], eS = cs.chain[e];
            const double jP = cn.Jc[r][e], jS = cs.Jc[r][e];  // zero beyond the chains
            aS -= AS[iS * RS + eS] * jS;
            if (same) { aP -= AP[iP * RS + eS] * jS; aS -= AP[iS * RS + eP] * jP; }
          }
          uP[d] = aP; uS[d] = aS;
        }
        for (int s2 = 0; s2 < 3; s2++) {
          double v = 0;
          for (int d = 0; d < MAXD; d++) v += cn.Jc[s2][d] * uP[d] + cs.Jc[s2][d] * uS[d];
          if (r == s2) v += cn.R[r];
          cn.A[3 * s2 + r] = v;
        }
        cs.R[r] = 0; cs.b[r] = 0;  // the second record carries no row of its own
      }
      env_sync<NW, EPB>(bar_id);
      if (rowt && r == 0 && !(HC && c_key == -3)) {
        Contact& cn = sm.u.con[c];
        double jar[3] = {cn.f[0], cn.f[1], cn.f[2]}, D[3] = {1.0 / cn.R[0], 1.0 / cn.R[1], 1.0 / cn.R[2]};
        rp::elliptic_primal_force(jar, D, cn.mu * sqrt(cn.R[1] / cn.R[0]), cn.mu, cn.f);
      }
      if (HC && rowt && r == 1 && c_key == -3) sm.u.con[c].dep = dep;
      if (rowt && r == 1 && !(HC && c_key == -3)) {  // what the Gauss-Seidel update of this block needs in every sweep but that never changes
        Contact& cn = sm.u.con[c];
        double Ab[9];
        for (int k = 0; k < 9; k++) Ab[k] = cn.A[k];
        rp::EllInv q;
        const int qok = rp::pgs_elliptic_invariants(q, Ab, mu_row);
        cn.inv = q;
        cn.dep = dep | (qok << 8);
      }
      env_sync<NW, EPB>(bar_id);
    }
    if (HANDS && h == 0 && ncon > 0) {
      // coupled keys, per-hand contact lists, cross-hand key sharing (sorted order: same key => adjacent)
      {
        bool firstc = false, sharedc = false;
        int mykey = -1, myhand = -1;
        if (lane < ncon) {
          mykey = sm.u.con[lane].key;
          myhand = sm.u.con[lane].hand;
          const bool samekey = lane > 0 && sm.u.con[lane - 1].key == mykey;
          firstc = mykey >= 0 && !samekey;
          sharedc = mykey >= 0 && samekey && sm.u.con[lane - 1].hand != myhand;
        }
        bool crossc = false;
        if (HC && lane < ncon && mykey == -3) {  // second record of a hand : hand contact: mirrors the warm-start force
          Contact& cs = sm.u.con[lane];
          const Contact& cp = sm.u.con[lane - 1];
          cs.f[0] = cp.f[0]; cs.f[1] = cp.f[1]; cs.f[2] = cp.f[2];
          crossc = cp.hand != myhand;
        }
        const unsigned xmask = HC ? __ballot_sync(FULL, crossc) : 0u;  // a cross-hand pair couples the two warps like a shared key
        const unsigned fmask = __ballot_sync(FULL, firstc);
        const unsigned smask = __ballot_sync(FULL, sharedc);
        const unsigned h0mask = __ballot_sync(FULL, myhand == 0), h1mask = __ballot_sync(FULL, myhand == 1);
        const unsigned lt = (1u << lane) - 1u;
        if (firstc) { const int slot = __popc(fmask & lt); sm.keyslot[mykey] = (signed char)slot; sm.ckown[slot] = (signed char)myhand; }
        if (myhand == 0) sm.hlist[0][__popc(h0mask & lt)] = (signed char)lane;
        if (myhand == 1) sm.hlist[H1][__popc(h1mask & lt)] = (signed char)lane;
        if (lane == 0) { sm.nck = __popc(fmask); sm.shared_key = smask != 0 || xmask != 0; sm.nhc[0] = __popc(h0mask); sm.nhc[H1] = __popc(h1mask); }
      }
    }
    TICK(6)
    // =============================================================== warm start
    // forces from qacc_warmstart through the primal map; kept only if the dual cost is negative
    if (hl && fr_on) {
      const double jar = wh - fr_aref, fl = fr_loss, R = fr_R;
      fr_f = jar <= -R * fl ? fl : (jar >= R * fl ? -fl : -jar / R);
    }
    if (hl && lm_on) {
      const double jar = lm_J * wh - lm_aref;
      lm_f = jar < 0 ? -jar / lm_R : 0.0;
    }
#pragma unroll
    for (int s = 0; s < KS; s++) {
      if (krow[s].active) {
        const double jar = krow[s].J * wk[s] - krow[s].aref;
        krow[s].f = jar < 0 ? -jar / krow[s].R : 0.0;
      }
    }
    env_sync<NW, EPB>(bar_id);
    TICK(4)
    const int nck = sm.nck;
    const int shared_key = sm.shared_key;
    const int nmine = HANDS ? sm.nhc[h] : 0;
#if RP_NU > 0
    // U_c = M^-1 Jc^T (three 26-vectors per contact, lane == dof) for the first RP_NU contacts: turns the
    // per-update walk along the kinematic chain (index load -> shuffle -> load -> FMA per chain dof, ~115
    // instructions) into three loads and three FMAs.  Lives in the scratch the contact rows no longer need.
    if (HANDS) {
      for (int ci = 0; ci < nmine; ci++) {
        const int c = sm.hlist[h][ci];
        if (c < RP_NU) {  // warp-uniform
          const Contact& cn = sm.u.con[c];
          const int dep = cn.dep & 0xff;
          double u0 = 0, u1 = 0, u2 = 0;
#pragma unroll
          for (int d = 0; d < MAXD; d++) {  // Jc is zero beyond the chain, chain[] repeats dof 0 there
            const double m = Ah[cn.chain[d] * RS + lc];
            u0 -= m * cn.Jc[0][d]; u1 -= m * cn.Jc[1][d]; u2 -= m * cn.Jc[2][d];
          }
          if (hl) { sm.v.U[c][0][lane] = u0; sm.v.U[c][1][lane] = u1; sm.v.U[c][2][lane] = u2; }
        }
      }
      __syncwarp();
    }
#endif
    // va = Minv J^T f ; cost = 0.5 f'(J va) + 0.5 sum R f^2 + f'b
    double va = 0;
    double improvement0 = 0;  // contribution of the closed-form key rows to the first sweep
    {
      double cost = 0;
      double wq = 0;
      if (hl) {
        wq = fr_f + lm_J * lm_f;
        cost += 0.5 * fr_R * fr_f * fr_f + fr_f * fr_b + 0.5 * lm_R * lm_f * lm_f + lm_f * lm_b;
        for (int ci = 0; ci < nmine; ci++) {
          const Contact& cn = sm.u.con[sm.hlist[h][ci]];
          const int link = cn.link;
          if (HC) { double j0, j1, j2; lane_jac(cn, lane, lsub, ldc, j0, j1, j2); wq += j0 * cn.f[0] + j1 * cn.f[1] + j2 * cn.f[2]; }
          else if (lane <= link && link < lane + lsub) wq += cn.Jc[0][ldc] * cn.f[0] + cn.Jc[1][ldc] * cn.f[1] + cn.Jc[2][ldc] * cn.f[2];
        }
        sm.wbuf[h][lane] = wq;
      }
      double wqk[KS];
#pragma unroll
      for (int s = 0; s < KS; s++) {
        const int k = tid + KT * s;
        wqk[s] = 0;
        if (k < NK) {
          wqk[s] = krow[s].J * krow[s].f;
          cost += 0.5 * krow[s].R * krow[s].f * krow[s].f + krow[s].f * krow[s].b;
          if (sm.keyslot[k] >= 0)
            for (int c = 0; c < ncon; c++) {
              const Contact& cn = sm.u.con[c];
              if (cn.key == k) wqk[s] += cn.Jk[0] * cn.f[0] + cn.Jk[1] * cn.f[1] + cn.Jk[2] * cn.f[2];
            }
          cost += 0.5 * wqk[s] * wqk[s] / kM[s];
        }
      }
      if (HANDS && h == 0 && lane < ncon) {
        const Contact& cn = sm.u.con[lane];
        for (int r = 0; r < 3; r++) cost += 0.5 * cn.R[r] * cn.f[r] * cn.f[r] + cn.f[r] * cn.b[r];
      }
      __syncwarp();
      if (hl) {
        double s2 = 0;
        for (int k = 0; k < NL; k++) s2 -= Ah[k * RS + lane] * sm.wbuf[h][k];
        va = s2;
        cost += 0.5 * wq * s2;
      }
      cost = warp_sum(cost);
      if (HANDS) {
        if (lane == 0) sm.red[0][h][0] = cost;
        env_sync<NW, EPB>(bar_id);
        cost = sm.red[0][0][0] + sm.red[0][H1][0];
      }
      const bool keep = !(cost > 0);
      if (!keep) {
        fr_f = 0; lm_f = 0; va = 0;
        if (HANDS && h == 0 && lane < ncon) { Contact& cn = sm.u.con[lane]; cn.f[0] = cn.f[1] = cn.f[2] = 0; }
      }
      if (a.dbg && sub == 0 && tid == 0) a.dbg[(size_t)env * RP_DBG_STRIDE + RP_DBG_SCALARS + 3] = cost;
      // key rows: uncoupled ones are solved exactly now (one Gauss-Seidel update of a 1x1 system is
      // its solution); coupled ones move to the ck[] list with their warm-start force
#pragma unroll
      for (int s = 0; s < KS; s++) {
        const int k = tid + KT * s;
        if (k < NK) {
          double f = keep ? krow[s].f : 0.0;
          const int slot = sm.keyslot[k];
          if (slot < 0) {
            if (krow[s].active) {
              const double res = krow[s].b + krow[s].AR * f;
              double fn = f - res * krow[s].ARinv;
              if (fn < 0) fn = 0;
              const double dl = fn - f;
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
    env_sync<NW, EPB>(bar_id);

    TICK(7)
    // =============================================================== PGS
    // Row order inside a sweep follows MuJoCo: friction-loss dofs, joint limits (keys, then hand
    // joints), contacts.  The two hands are independent islands unless they touch the same key, so
    // each warp sweeps the rows of its own hand; the rows of a key are handled by the warp of the
    // hand touching it.  If some key carries contacts of both hands (shared_key) the key rows go to
    // warp 0 and the contact rows are stepped in global order with a block barrier per row.
    int iter = 0;
    if (HANDS) {
      PgsLane io = {fr_f, lm_f, va};
      if (HC) iter = pgs_solve_hc(&sm, h, bar_id, ncon, maxiter, tol, dm->solver_scale, improvement0, fr_R, fr_b, fr_AR,
                                  fr_on ? fr_loss : 0.0, lm_J, lm_R, lm_b, lm_AR, lm_mask, lsub, ldc, &io);
      else iter = pgs_solve(&sm, h, bar_id, ncon, maxiter, tol, dm->solver_scale, improvement0, fr_R, fr_b, fr_AR,
                            fr_on ? fr_loss : 0.0, lm_J, lm_R, lm_b, lm_AR, lm_mask, lsub, ldc, &io
#ifdef RP_PROFILE
                            , tacc, &tlast
#endif
      );
      fr_f = io.fr_f; lm_f = io.lm_f; va = io.va;
    } else {
      // keys only: every limit row is a 1 x 1 island solved in closed form above; the sweeps only run the
      // convergence test (the first one sees the improvement of the closed-form updates)
      while (iter < maxiter) {
        double improvement = warp_sum(iter == 0 ? improvement0 : 0.0) * dm->solver_scale;
        iter++;
        if (improvement < tol) break;
      }
    }
    niter_last = iter;

    TICK(12)
    // =============================================================== dual -> primal, Euler
    // qacc = qacc_smooth + va ; qfrc_constraint = J^T f
    const double qacc = qacc_smooth + va;
    double qfrc_c = 0;
    if (hl) {
      qfrc_c = fr_f + lm_J * lm_f;
      for (int ci = 0; ci < nmine; ci++) {
        const Contact& cn = sm.u.con[sm.hlist[h][ci]];
        const int link = cn.link;
        if (HC) { double j0, j1, j2; lane_jac(cn, lane, lsub, ldc, j0, j1, j2); qfrc_c += j0 * cn.f[0] + j1 * cn.f[1] + j2 * cn.f[2]; }
        else if (lane <= link && link < lane + lsub) qfrc_c += cn.Jc[0][ldc] * cn.f[0] + cn.Jc[1][ldc] * cn.f[1] + cn.Jc[2][ldc] * cn.f[2];
      }
    }
    // constraint force on coupled keys: limit row + contacts, accumulated in contact order
    // (all rows are final: every warp passed the barrier of the last convergence test)
    if (tid == 0) {
      for (int c = 0; c < nck; c++) sm.kfc[sm.ck[c].key] = sm.ck[c].J * sm.ck[c].f;
      for (int c = 0; c < ncon; c++) {
        const Contact& cn = sm.u.con[c];
        if (cn.key >= 0) sm.kfc[cn.key] += cn.Jk[0] * cn.f[0] + cn.Jk[1] * cn.f[1] + cn.Jk[2] * cn.f[2];
      }
    }
    env_sync<NW, EPB>(bar_id);
    double kq_acc[KS], kq_frc[KS], kq_smooth[KS];
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int k = tid + KT * s;
      kq_acc[s] = 0; kq_frc[s] = 0; kq_smooth[s] = 0;
      if (k < NK) {
        kq_frc[s] = sm.kfc[k];
        kq_smooth[s] = ks_qfrc[s];
        kq_acc[s] = ks_qacc[s] + sm.vak[k];
      }
    }
    if (a.dbg && sub == 0) {
      double* dbg = a.dbg + (size_t)env * RP_DBG_STRIDE;
#pragma unroll
      for (int s = 0; s < KS; s++) {
        const int k = tid + KT * s;
        if (k < NK) { dbg[RP_DBG_QACC_SMOOTH + k] = ks_qacc[s]; dbg[RP_DBG_QACC + k] = kq_acc[s]; dbg[RP_DBG_QFRC_C + k] = kq_frc[s]; dbg[RP_DBG_BIAS + k] = 0; }
      }
      if (hl) {
        dbg[RP_DBG_QACC_SMOOTH + NK + h * NL + lane] = qacc_smooth;
        dbg[RP_DBG_QACC + NK + h * NL + lane] = qacc;
        dbg[RP_DBG_QFRC_C + NK + h * NL + lane] = qfrc_c;
        dbg[RP_DBG_BIAS + NK + h * NL + lane] = qfrc_bias;
      }
      if (tid == 0) { dbg[RP_DBG_SCALARS + 0] = ncon; dbg[RP_DBG_SCALARS + 1] = iter; dbg[RP_DBG_SCALARS + 2] = sm.overflow; }
      if (tid < ncon) {
        const Contact& cn = sm.u.con[tid];
        double* o = dbg + RP_DBG_CONTACTS + tid * 16;
        o[0] = cn.key; o[1] = cn.hand; o[2] = cn.link; o[3] = cn.f[0]; o[4] = cn.f[1]; o[5] = cn.f[2];
        o[6] = cn.b[0]; o[7] = cn.b[1]; o[8] = cn.b[2]; o[9] = cn.R[0]; o[10] = cn.A[0]; o[11] = cn.A[4]; o[12] = cn.A[1];
        o[13] = 0; o[14] = cn.mu; o[15] = cn.Jk[0];
      }
    }
    TICK(13)
    // Euler with implicit joint damping: (M + h*D) a = qfrc_smooth + qfrc_constraint
    double aeul = 0;
    if (HANDS) {
      double* Aw = sm.A[h];
      __syncwarp();
      for (int idx = lane; idx < NL * RS; idx += 32) Aw[idx] = 0;
      __syncwarp();
      if (hl) {
        for (int d = 0; d <= ldep; d++) {
          const int j = sm.anc[h][lane][d];
          double v = sm.Mrow[h][moff + d];
          if (d == ldep) v += hstep * dm->link_damping[h][lane];
          Aw[lane * RS + j] = v;
          Aw[j * RS + lane] = v;
        }
        sm.wbuf[h][lane] = qfrc_smooth + qfrc_c;  // right-hand side
      }
      __syncwarp();
      // a = (M + h*D)^-1 rhs: sweep (A = -inverse), then one row-times-vector per lane
      sweep_inverse(Aw, sched, lane);
      TICK(14)
      if (hl) {
        double acc = 0;
        for (int i = 0; i < NL; i++) acc -= Aw[lane * RS + i] * sm.wbuf[h][i];
        aeul = acc;
      }
    }
    // integrate
    if (hl) {
      vh += hstep * aeul;
      qh += hstep * vh;
      wh = qacc;
    }
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int k = tid + KT * s;
      if (k < NK) {
        const double ae = (kq_smooth[s] + kq_frc[s]) / (dm->key_M[k] + hstep * dm->key_damping[k]);
        vk[s] += hstep * ae;
        qk[s] += hstep * vk[s];
        wk[s] = kq_acc[s];
      }
    }
    if (midi_on) {  // Piano.after_substep -> MidiModule.after_substep: activation XOR previous activation
#pragma unroll
      for (int s = 0; s < KS; s++) {
        const int k = tid + KT * s;
        if (k < NK) {
          bool act;
          if (dm->task_kind == 1) {
            act = sm.ctrl[k] >= 0.5 * (dm->key_act_range[k][0] + dm->key_act_range[k][1]);
          } else {
            const double hi = dm->key_range[k][1];
            act = fabs(fmin(fmax(qk[s], dm->key_range[k][0]), hi) - hi) <= 0.00872665;
          }
          if (act != midi_prev[s]) {
            const int slot = atomicAdd(&a.midi_count[env], 1);
            if (slot < a.midi_cap)
              a.midi_events[(size_t)env * a.midi_cap + slot] = RP_MIDI_WORD(t_idx * a.nsub + sub + 1, act ? 0 : 1, k);
            midi_prev[s] = act;
          }
        }
      }
    }
    env_sync<NW, EPB>(bar_id);
    busy += clock64() - t_sub;
    TICK(15)
  }  // substeps

  // ------------------------------------------------------------------ store state
  {
    bool bad = false;
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int k = tid + KT * s;
      if (k < NK) { gq[k] = qk[s]; gv[k] = vk[s]; gw[k] = wk[s]; bad |= !isfinite(qk[s]) || !isfinite(vk[s]); }
    }
    if (hl) { gq[NK + h * NL + lane] = qh; gv[NK + h * NL + lane] = vh; gw[NK + h * NL + lane] = wh; bad |= !isfinite(qh) || !isfinite(vh); }
    int anybad;
    if (EPB == 1) {
      anybad = __syncthreads_or(bad ? 1 : 0);  // also publishes the key sin/cos of the final pass
    } else {
      if (tid == 0) sm.pad = 0;
      env_sync<NW, EPB>(bar_id);
      if (bad) atomicOr(&sm.pad, 1);
      env_sync<NW, EPB>(bar_id);
      anybad = sm.pad;
    }
    if (anybad) flags |= 2;
    if (tid == 0 && a.diag) { a.diag[3 * env + 0] = ncon_last; a.diag[3 * env + 1] = niter_last; a.diag[3 * env + 2] = flags | (sm.near_uncovered ? 8 : 0); }
    if (tid == 0 && a.work) a.work[env] = (int)(busy >> 12);
    //  // rp_order_kernel buckets by work >> 4: 64 K cycles per bucket
    if (a.act_force) {
      if (HANDS && lane < dm->nact[h]) a.act_force[(size_t)env * nu + h * dm->nact[0] + lane] = actforce;
      if (dm->nkey_act) {
#pragma unroll
        for (int s = 0; s < KS; s++) { const int k = tid + KT * s; if (k < NK) a.act_force[(size_t)env * nu + k] = key_actforce[s]; }
      }
    }
  }
#ifdef RP_PROFILE
  TICK(19)
  if (a.dbg && lane == 0) for (int i = 0; i < RP_DBG_NPROF; i++) a.dbg[(size_t)env * RP_DBG_STRIDE + RP_DBG_PROFILE + h * RP_DBG_NPROF + i] = (double)tacc[i];
#endif
  if (a.mode == RP_MODE_PHYSICS) return;

  // =============================================================================================
  // task epilogue: after_step, observables, reward, termination  (piano_with_shadow_hands.py:188-412,
  // self_actuated_piano.py:143-243) — uses the poses of the final kinematics pass.
  // =============================================================================================
  const int T = dm->T, L = dm->lookahead;
  const double sustain_state = a.sustain[env];
  const bool first = do_reset || a.mode == RP_MODE_RESET;
  const int t_new = first ? 0 : t_idx + 1;
  // key state
  double knorm[KS];
  bool kact[KS];
#pragma unroll
  for (int s = 0; s < KS; s++) {
    const int k = tid + KT * s;
    knorm[s] = 0; kact[s] = false;
    if (k < NK) {
      const double hi = dm->key_range[k][1];
      if (dm->task_kind == 1) {
        const double mid = 0.5 * (dm->key_act_range[k][0] + dm->key_act_range[k][1]);
        kact[s] = sm.ctrl[k] >= mid;
        knorm[s] = 0;  // state observable stays at its initial value in the self-actuated task
      } else {
        const double st = fmin(fmax(qk[s], dm->key_range[k][0]), hi);
        knorm[s] = st / hi;
        kact[s] = fabs(st - hi) <= 0.00872665;
      }
      if (a.activation) a.activation[(size_t)env * NK + k] = kact[s] ? 1 : 0;
    }
  }
  const bool sustain_act = sustain_state >= 0.5;
  // fingertip sites (lane s < 5 of warp h: site s of hand h)
  if (HANDS) {
    const bool iss = lane < RP_NSITE;
    const int sl = iss ? dm->site_link[h][lane] : 0;
    double lp[3], lq[4];
#pragma unroll
    for (int k = 0; k < 3; k++) lp[k] = shfl(xp[k], sl);
#pragma unroll
    for (int k = 0; k < 4; k++) lq[k] = shfl(xq[k], sl);
    double tmp[3], spos[3] = {dm->site_pos[h][iss ? lane : 0][0], dm->site_pos[h][iss ? lane : 0][1], dm->site_pos[h][iss ? lane : 0][2]};
    rotq(tmp, lq, spos);
    if (iss) {
      for (int k = 0; k < 3; k++) {
        const double v = lp[k] + tmp[k];
        sm.site[h][lane][k] = v;
        if (a.site_xpos) a.site_xpos[((size_t)env * 2 * RP_NSITE + h * RP_NSITE + lane) * 3 + k] = v;
      }
    }
  }
  // forearm reward with hand : hand collisions enabled (piano_with_shadow_hands.py:251-259: has_collision between the two
  // root bodies' geoms).  The forearm colliders are a box and a mesh, outside the narrowphase's coverage: the test is an
  // oriented-box overlap of the box / the mesh's bounding box (separating axes) — an approximation of MuJoCo's convex
  // contact test, conservative for the mesh.
  bool forearms_touch = false;
  if (HC) {
    const int nf = dm->nfore[h];
    const bool isf = lane < nf;
    const int fi = isf ? dm->fore_ign[h][lane] : 0;
    const int il = dm->ign_link[h][fi];
    double lp[3], lq[4];
#pragma unroll
    for (int k = 0; k < 3; k++) lp[k] = shfl(xp[k], il);
#pragma unroll
    for (int k = 0; k < 4; k++) lq[k] = shfl(xq[k], il);
    if (isf) {
      double Rl[9], tmp[3], gp[3] = {dm->ign_pos[h][fi][0], dm->ign_pos[h][fi][1], dm->ign_pos[h][fi][2]};
      rp::quat2mat(Rl, lq);
      rp::mat_vec(tmp, Rl, gp);
      double* o = &sm.wbuf[h][15 * lane];  // centre 3, rotation 9, half sizes 3
      for (int k = 0; k < 3; k++) { o[k] = lp[k] + tmp[k]; o[12 + k] = dm->ign_half[h][fi][k]; }
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[3 + 3 * i + j] = Rl[3 * i] * dm->ign_mat[h][fi][j] + Rl[3 * i + 1] * dm->ign_mat[h][fi][3 + j] + Rl[3 * i + 2] * dm->ign_mat[h][fi][6 + j];
    }
    if (tid == 0) sm.nraw = 0;
    env_sync<NW, EPB>(bar_id);
    if (tid < dm->nfore[0] * dm->nfore[H1]) {
      const int i = tid / dm->nfore[H1], j = tid - i * dm->nfore[H1];
      const double *A_ = &sm.wbuf[0][15 * i], *B_ = &sm.wbuf[H1][15 * j];
      if (rp::obb_overlap(A_, A_ + 3, A_ + 12, B_, B_ + 3, B_ + 12)) atomicOr(&sm.nraw, 1);
    }
    env_sync<NW, EPB>(bar_id);
    forearms_touch = sm.nraw != 0;
  }
  double reward = 0, terms[5] = {0, 0, 0, 0, 0}, discount = 1.0;
  int step_type = first ? RP_STEP_FIRST_ : RP_STEP_MID_;
  if (!first) {
    const uint8_t* gcur = a.goal + (size_t)min(t_idx, T - 1) * RP_GOAL_DIM_;  // goal emitted with the previous observation
    const bool should_terminate = (t_new - 1) == T - 1;
    // ---- key-press / activation statistics (block reduction through red[0])
    double sum_on = 0, n_on = 0, false_pos = 0, true_pos = 0, sq = 0, xent = 0, pw = 0;  // false_pos / true_pos: counts
#pragma unroll
    for (int s = 0; s < KS; s++) {
      const int k = tid + KT * s;
      if (k < NK) {
        const int g = gcur[k];
        if (g) {
          sum_on += tolerance_gauss((double)g - knorm[s], 0.0, 0.05, 0.5);  // goal - state/qmax
          n_on += 1.0;
          if (kact[s]) true_pos += 1.0;
        } else if (kact[s]) false_pos += 1.0;
        const double df = (kact[s] ? 1.0 : 0.0) - (double)g;
        sq += df * df;
        const double pr = kact[s] ? 1.0 : 0.0;  // negative_binary_cross_entropy, eps 1e-6 (self_actuated_piano.py:38-47)
        xent += g ? log(pr + 1e-6) : log(1 - pr + 1e-6);
      }
    }
    // energy: |actuator force (pre-integration)| * |actuator velocity (post-integration)|
    if (HANDS) {
      const bool isact = lane < dm->nact[h];
      const int d0 = isact ? dm->act_dof0[h][lane] : 0, d1 = isact ? dm->act_dof1[h][lane] : -1;
      const double v0 = shfl(vh, d0), v1 = shfl(vh, d1 >= 0 ? d1 : 0);
      if (isact) {
        const double vel = dm->act_coef0[h][lane] * v0 + (d1 >= 0 ? dm->act_coef1[h][lane] * v1 : 0.0);
        pw = fabs(actforce) * fabs(vel);
      }
    }
    sum_on = warp_sum(sum_on); n_on = warp_sum(n_on); false_pos = warp_sum(false_pos); true_pos = warp_sum(true_pos);
    sq = warp_sum(sq); xent = warp_sum(xent); pw = warp_sum(pw);
    if (HANDS) {
      if (lane == 0) { double* r = sm.red[0][h]; r[2] = sum_on; r[3] = n_on; r[4] = false_pos; r[5] = sq; r[6] = xent; r[7] = pw; r[8] = true_pos; }
      env_sync<NW, EPB>(bar_id);  // also publishes the sites
      const double* r0 = sm.red[0][0]; const double* r1 = sm.red[0][H1];
      sum_on = r0[2] + r1[2]; n_on = r0[3] + r1[3]; false_pos = r0[4] + r1[4]; sq = r0[5] + r1[5]; xent = r0[6] + r1[6]; pw = r0[7] + r1[7]; true_pos = r0[8] + r1[8];
    }
    const int gs = gcur[NK];
    // musical metrics of this step (wrappers/evaluation.py:114-177): activation vs. the goal of this step
    if (a.metrics && tid == 0) {
      double* mt = a.metrics + (size_t)env * RP_METRICS_STRIDE;
      const double tp = true_pos, fp = false_pos, fn = n_on - true_pos;
      const double stp = (gs && sustain_act) ? 1.0 : 0.0, sfp = (!gs && sustain_act) ? 1.0 : 0.0, sfn = (gs && !sustain_act) ? 1.0 : 0.0;
      mt[0] += prf_ratio(tp, tp + fp); mt[1] += prf_ratio(tp, tp + fn); mt[2] += prf_ratio(2 * tp, 2 * tp + fp + fn);
      mt[3] += prf_ratio(stp, stp + sfp); mt[4] += prf_ratio(stp, stp + sfn); mt[5] += prf_ratio(2 * stp, 2 * stp + sfp + sfn);
      mt[6] += 1.0;
    }
    if (dm->task_kind == 0) {
      const double r_key = (n_on > 0 ? 0.5 * sum_on / n_on : 0.0) + 0.5 * (false_pos > 0 ? 0.0 : 1.0);
      // sustain: tolerance(goal - sustain_activation)
      const double r_sus = tolerance_gauss((double)gs - (sustain_act ? 1.0 : 0.0), 0.0, 0.05, 0.5);
      const double r_en = -dm->energy_coef * pw;
      // fingering: mean tolerance of fingertip-to-key distances for (key, finger) pairs due now
      double r_fing = 0;
      if (!dm->disable_fingering_reward) {
        const int8_t* fcur = a.fingering + (size_t)min(t_idx, T - 1) * NK;
        double fsum = 0; int fn = 0;
        for (int k = 0; k < NK; k++) {
          const int fg = fcur[k];
          if (fg < 0 || !gcur[k]) continue;
          const int hh = fg < 5 ? 0 : H1, fi = fg < 5 ? fg : fg - 5;
          // key geom centre in the world: body pos rotated about the hinge (angles of the final pass)
          double ax[3] = {dm->key_axis[k][0], dm->key_axis[k][1], dm->key_axis[k][2]};
          double r0[3] = {-dm->key_jpos[k][0], -dm->key_jpos[k][1], -dm->key_jpos[k][2]}, rr[3];
          rot_axis(rr, ax, sm.v.pre.ksn[k], sm.v.pre.kcs[k], r0);
          double kp[3];
          for (int c = 0; c < 3; c++) kp[c] = dm->key_pos[k][c] + dm->key_jpos[k][c] + rr[c];
          kp[2] += 0.5 * dm->key_half[k][2];
          kp[0] += 0.35 * dm->key_half[k][0];
          const double sx = sm.site[hh][fi][0], sy = sm.site[hh][fi][1], sz = sm.site[hh][fi][2];
          const double dd = sqrt((kp[0] - sx) * (kp[0] - sx) + (kp[1] - sy) * (kp[1] - sy) + (kp[2] - sz) * (kp[2] - sz));
          fsum += tolerance_gauss(dd, 0.0, 0.01, 0.1);
          fn++;
        }
        r_fing = fn > 0 ? fsum / fn : 0.0;
      }
      const double r_fore = dm->disable_forearm_reward ? 0.0 : ((HC && forearms_touch) ? 0.0 : 0.5);  // without hand : hand collisions the forearms never touch
      terms[0] = r_key; terms[1] = r_sus; terms[2] = r_en; terms[3] = r_fing; terms[4] = r_fore;
      reward = r_key + r_sus + r_en;
      if (!dm->disable_fingering_reward) reward += r_fing;
      if (!dm->disable_forearm_reward) reward += r_fore;
      const bool failure = false_pos > 0;
      if (should_terminate) step_type = RP_STEP_LAST_;
      else if (dm->wrong_press_termination && failure) { discount = 0.0; step_type = RP_STEP_LAST_; }
    } else {
      // self-actuated: reward on [activation, sustain_activation] vs goal
      const double dfs = (sustain_act ? 1.0 : 0.0) - (double)gs;
      const double prs = sustain_act ? 1.0 : 0.0;
      if (dm->reward_type == 0) reward = -sqrt(sq + dfs * dfs);
      else reward = xent + (gs ? log(prs + 1e-6) : log(1 - prs + 1e-6));
      terms[0] = reward;
      if (should_terminate) step_type = RP_STEP_LAST_;
    }
  }
  // A diverged environment (non-finite qpos / qvel: where MuJoCo raises mjWARN_BADQPOS / BADQACC and dm_control a
  // PhysicsError) ends its episode instead: LAST with discount 0 and reward 0, diag flag 2 — and the next step
  // auto-resets it like any other finished episode.  Its last observation may hold non-finite values.
  if (!first && (flags & 2)) {
    step_type = RP_STEP_LAST_; discount = 0.0; reward = 0.0;
    for (int k = 0; k < 5; k++) terms[k] = 0.0;
  }
  // ---- observation (order: piano/state, piano/sustain_state, rh joints_pos, lh joints_pos, goal, fingering)
  if (a.obs) {
    double* o = a.obs + (size_t)env * dm->obs_dim;
    int off = 0;
    if (dm->task_kind == 0) {
#pragma unroll
      for (int s = 0; s < KS; s++) { const int k = tid + KT * s; if (k < NK) o[k] = knorm[s]; }
      off = NK;
      if (tid == 0) o[off] = sustain_state;
      off += 1;
      if (HANDS) {
        // joints_pos in the reference's joint order
        const int nob = dm->nobs[h];
        const int src = lane < nob ? dm->obs_dof[h][lane] : 0;
        const double v = shfl(qh, src);
        if (lane < nob) o[off + h * dm->nobs[0] + lane] = v;
        off += dm->nobs[0] + dm->nobs[H1];
      }
    } else {
      // self-actuated piano observables: piano/activation (88), piano/sustain_activation (1)
#pragma unroll
      for (int s = 0; s < KS; s++) { const int k = tid + KT * s; if (k < NK) o[k] = kact[s] ? 1.0 : 0.0; }
      off = NK;
      if (tid == 0) o[off] = sustain_act ? 1.0 : 0.0;
      off += 1;
    }
    // goal lookahead window starting at the new t_idx; kept from the previous observation when t == T
    const int tg = t_new;
    const bool frozen = (tg >= T);
    for (int idx = tid; idx < (L + 1) * RP_GOAL_DIM_; idx += KT) {
      const int i = idx / RP_GOAL_DIM_, c = idx - i * RP_GOAL_DIM_;
      const int t2 = (frozen ? T - 1 : tg) + i;
      o[off + idx] = t2 < T ? (double)a.goal[(size_t)t2 * RP_GOAL_DIM_ + c] : 0.0;
    }
    off += (L + 1) * RP_GOAL_DIM_;
    if (dm->task_kind == 0 && !dm->disable_fingering_reward) {
      const int tf = frozen ? T - 1 : tg;
      if (tid < 10) {
        double v = 0;
        const int8_t* fc = a.fingering + (size_t)tf * NK;
        for (int k = 0; k < NK; k++) if (fc[k] == tid && a.goal[(size_t)tf * RP_GOAL_DIM_ + k]) v = 1.0;
        o[off + tid] = v;
      }
    }
  }
  if (tid == 0 && a.metrics) {
    double* mt = a.metrics + (size_t)env * RP_METRICS_STRIDE;
    if (first) {
      for (int k = 0; k < 7; k++) mt[k] = 0;  // a new episode starts
    } else if (step_type == RP_STEP_LAST_ && mt[6] > 0) {
      for (int k = 0; k < 6; k++) mt[8 + k] = mt[k] / mt[6];
      mt[7] += 1.0;
    }
  }
  if (tid == 0) {
    a.t_idx[env] = t_new;
    a.step_type_state[env] = step_type;
    if (a.reward) a.reward[env] = first ? 0.0 : reward;
    if (a.discount) a.discount[env] = first ? 1.0 : discount;
    if (a.step_type) a.step_type[env] = (uint8_t)step_type;
    if (a.reward_terms) for (int k = 0; k < 5; k++) a.reward_terms[(size_t)env * 5 + k] = first ? 0.0 : terms[k];
  }
}

// Launch order for the next step: environments sorted by decreasing duration of their last step (bucket sort,
// one block).  Durations vary 4x between environments (PGS sweeps x contacts); starting the long ones
// first shortens the tail of the last wave.  Order inside a bucket is arbitrary (results do not depend
// on which block runs which environment).
__global__ void __launch_bounds__(1024) rp_order_kernel(const int32_t* __restrict__ work, int32_t* __restrict__ order, int n) {
  __shared__ int hist[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[255 - min(255, work[i] >> 4)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int b = 0; b < 256; b++) { const int c = hist[b]; hist[b] = run; run += c; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) order[atomicAdd(&hist[255 - min(255, work[i] >> 4)], 1)] = i;
}

cudaError_t rp_launch_order(const int32_t* work, int32_t* order, int n, cudaStream_t stream) {
  rp_order_kernel<<<1, 1024, 0, stream>>>(work, order, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
size_t rp_kernel_smem_bytes() { return kSmemStride * RP_EPB + kSchedBytes; }

// the opt-in to more than 48 KB of dynamic shared memory is per device
static int current_device_slot() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < 64 ? dev : 0;
}

template <int EPB, bool HC>
static cudaError_t launch_hands(const RpKernelArgs& a, cudaStream_t stream) {
  static bool configured[64] = {};
  const int dev = current_device_slot();
  const size_t smem = kSmemStride * EPB + kSchedBytes;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(rp_step_kernel<2, EPB, HC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  rp_step_kernel<2, EPB, HC><<<(a.nenv + EPB - 1) / EPB, 64 * EPB, smem, stream>>>(a);
  return cudaGetLastError();
}

// block_envs: environments per thread block for the hands kernel (1 or RP_EPB; 0 = choose by batch size: the
// phase-locked RP_EPB-environment blocks need at least two full waves of blocks to pay off).
// hand_collisions: the model's hand geoms may collide with each other (RpDevModel::hc_enabled): the HC instantiation.
cudaError_t rp_launch_step(const RpKernelArgs& a, int nhand, cudaStream_t stream, int block_envs, int hand_collisions) {
  if (nhand == 2) {
    int epb = block_envs;
    if (epb <= 0) {
      int dev = 0, sms = 148;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      epb = a.nenv >= 2 * sms * 6 ? RP_EPB : 1;  // two waves at six environments per SM
    }
    if (epb != 1 && epb != RP_EPB) return cudaErrorInvalidValue;
    if (hand_collisions) return epb == 1 ? launch_hands<1, true>(a, stream) : launch_hands<RP_EPB, true>(a, stream);
    return epb == 1 ? launch_hands<1, false>(a, stream) : launch_hands<RP_EPB, false>(a, stream);
  }
  static bool configured[64] = {};
  const int dev = current_device_slot();
  const size_t smem = kSmemStride;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(rp_step_kernel<1, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  rp_step_kernel<1, 1, false><<<a.nenv, 32, smem, stream>>>(a);
  return cudaGetLastError();
}
