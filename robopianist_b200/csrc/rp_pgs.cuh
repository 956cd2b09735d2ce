// rp_pgs.cuh — the Gauss-Seidel constraint solve over a structure-of-arrays constraint problem, one
// environment per caller (one thread on the device): groundwork for the split step pipeline described in
// DESIGN.md §5 "next steps" (setup kernel -> this solver -> finish kernel).  NOT on the product path yet:
// the fused step kernel (rp_kernels.cu) carries its own copy of this solve, spread over two warps.
//
// What it restates: MuJoCo's mj_solPGS with elliptic cones as the reference's mj_step runs it (the tests pin
// it to the dense restatement of the CPU oracle (oracle/)), on the representation the step kernel builds
// instead of a dense efc_J / efc_AR: per hand M^-1 (26 x 26), scalar rows for the 52 friction-loss dofs and the active joint
// limits, the limit rows of keys that carry contacts, and per contact the 3 x 3 block, the chain Jacobian
// and U = M^-1 Jc^T.  Keys without contacts are 1 x 1 islands: their limit row is solved exactly by the
// producer of the problem (one update, then it never moves again) and enters only through `improvement0`.
// Row order inside a sweep: friction-loss rows (hand 0 dofs, hand 1 dofs), key limit rows, hand limit rows,
// contacts in MuJoCo's geom-pair order — the hands are separate islands unless a key couples them, so this
// is MuJoCo's global order up to commuting islands.
//
// Layout: P[element * stride + env], element-major and environment-minor, so that a warp whose lanes are 32
// consecutive environments reads every element with one coalesced 256-byte transaction.
#pragma once
#include "rp_math.cuh"
#include "rp_model.h"

namespace rp {

enum PgsLayout : int {
  PG_NCON = 0,          // number of contacts
  PG_NCK = 1,           // number of coupled key rows
  PG_IMPROVEMENT0 = 2,  // cost decrease of the closed-form key rows, added to the first sweep
  PG_NITER = 3,         // out: sweeps used
  PG_HDR = 8,
  PG_MINV = PG_HDR,                          // [2][NL][NL]  M^-1 of each hand (symmetric)
  PG_FR = PG_MINV + 2 * RP_NL * RP_NL,       // [2*NL][5]    friction-loss rows: R, b, AR, 1/AR, bound
  PG_FR_F = PG_FR + 2 * RP_NL * 5,           // [2*NL]       ... force (in: warm start, out: solution)
  PG_LM = PG_FR_F + 2 * RP_NL,               // [2*NL][5]    joint-limit rows: J (+-1, 0 = inactive), R, b, AR, 1/AR
  PG_LM_F = PG_LM + 2 * RP_NL * 5,           // [2*NL]       ... force
  PG_VA = PG_LM_F + 2 * RP_NL,               // [2*NL]       M^-1 J^T f of the hand dofs (in: of the warm start)
  PG_CK = PG_VA + 2 * RP_NL,                 // [MAXCON][8]  coupled key rows: key, active, J, R, b, AR, 1/AR, 1/M
  PG_CK_F = PG_CK + RP_MAXCON * 8,           // [MAXCON]     ... force
  PG_CK_VA = PG_CK_F + RP_MAXCON,            // [MAXCON]     ... (J^T f) / M of that key: limit row + its contacts
  PG_CON = PG_CK_VA + RP_MAXCON,             // [MAXCON][PG_CON_STRIDE]
  // per contact (offsets inside its record)
  PC_HAND = 0, PC_CK = 1 /* slot in the coupled-key list, -1: piano base */, PC_DEP = 2, PC_MU = 3, PC_INVMK = 4,
  PC_JK = 5 /*[3]*/, PC_A = 8 /*[9]*/, PC_R = 17 /*[3]*/, PC_B = 20 /*[3]*/, PC_F = 23 /*[3] in/out*/,
  PC_CHAIN = 26 /*[MAXD] dof (0..NL-1) at depth d*/, PC_JC = PC_CHAIN + RP_MAXD /*[3][MAXD]*/,
  PC_U = PC_JC + 3 * RP_MAXD /*[3][NL]*/, PG_CON_STRIDE = PC_U + 3 * RP_NL,
  PG_SIZE = PG_CON + RP_MAXCON * PG_CON_STRIDE  // doubles per environment
};

struct PgsView {
  double* P;
  size_t stride;  // number of environments in the buffer
  int env;
  RP_HD double& at(int e) const { return P[(size_t)e * stride + env]; }
  RP_HD double ld(int e) const { return P[(size_t)e * stride + env]; }  // element the solve never writes
};

// One environment's solve.  View: anything with `double& at(int element) const` (PgsView reads the global
// buffer; the kernel's staged view keeps the scalar rows in shared memory).  va: 2*NL doubles of scratch
// (va[i * va_stride]); returns the number of sweeps.
template <class View>
RP_HD int pgs_solve_env(const View p, int maxiter, double tol, double solver_scale, double* va, int va_stride) {
  const int NL = RP_NL;
  const int ncon = (int)p.at(PG_NCON), nck = (int)p.at(PG_NCK);
  for (int i = 0; i < 2 * NL; i++) va[i * va_stride] = p.at(PG_VA + i);
  int iter = 0;
  while (iter < maxiter) {
    double improvement = iter == 0 ? p.at(PG_IMPROVEMENT0) : 0.0;
    // ---- friction-loss rows
    for (int i = 0; i < 2 * NL; i++) {
      const int r = PG_FR + 5 * i;
      const double R = p.ld(r), b = p.ld(r + 1), AR = p.ld(r + 2), ARi = p.ld(r + 3), hi = p.ld(r + 4);
      const double f = p.at(PG_FR_F + i);
      const double res = b + R * f + va[i * va_stride];
      const double x = f - res * ARi;
      double fx = x < -hi ? -hi : x;
      fx = x > hi ? hi : fx;
      if (fx != f) {
        const double dl = fx - f;
        const int h = i / NL, col = i - h * NL;
        for (int j = 0; j < NL; j++) va[(h * NL + j) * va_stride] += p.ld(PG_MINV + (h * NL + col) * NL + j) * dl;
        improvement -= dl * (0.5 * dl * AR + res);
        p.at(PG_FR_F + i) = fx;
      }
    }
    // ---- limit rows of keys that carry contacts
    for (int c = 0; c < nck; c++) {
      const int r = PG_CK + 8 * c;
      const double active = p.ld(r + 1), J = p.ld(r + 2), R = p.ld(r + 3), b = p.ld(r + 4), AR = p.ld(r + 5),
                   ARi = p.ld(r + 6), invM = p.ld(r + 7);
      const double f = p.at(PG_CK_F + c), vak = p.at(PG_CK_VA + c);
      const double res = b + R * f + J * vak;
      double fn = f - res * ARi;
      fn = (fn < 0 || active == 0.0) ? 0.0 : fn;
      const double dl = fn - f;
      improvement -= dl * (0.5 * dl * AR + res);
      p.at(PG_CK_F + c) = fn;
      p.at(PG_CK_VA + c) = vak + J * dl * invM;
    }
    // ---- hand joint-limit rows
    for (int i = 0; i < 2 * NL; i++) {
      const int r = PG_LM + 5 * i;
      const double J = p.ld(r);
      if (J == 0.0) continue;
      const double R = p.ld(r + 1), b = p.ld(r + 2), AR = p.ld(r + 3), ARi = p.ld(r + 4);
      const double f = p.at(PG_LM_F + i);
      const double res = b + R * f + J * va[i * va_stride];
      const double x = f - res * ARi;
      const double fn = x < 0 ? 0.0 : x;
      const double dl = fn - f;
      if (dl != 0.0) {
        const int h = i / NL, col = i - h * NL;
        const double w = J * dl;
        for (int j = 0; j < NL; j++) va[(h * NL + j) * va_stride] += p.ld(PG_MINV + (h * NL + col) * NL + j) * w;
      }
      improvement -= dl * (0.5 * dl * AR + res);
      p.at(PG_LM_F + i) = fn;
    }
    // ---- contacts.  The record (read-only but for the force) is loaded up front through ld(): on the device
    // these are independent read-only loads that go out back to back (one memory latency per contact); read
    // through at() they would be ordered behind the force stores of the previous contact one by one.
    for (int c = 0; c < ncon; c++) {
      const int q = PG_CON + c * PG_CON_STRIDE;
      const int h = (int)p.ld(q + PC_HAND), ck = (int)p.ld(q + PC_CK), dep = (int)p.ld(q + PC_DEP);
      const double mu = p.ld(q + PC_MU), invMk = p.ld(q + PC_INVMK);
      double A[9], Rr[3], br[3], jk[3], res[3], f[3], old[3], jc[3][RP_MAXD];
      int chain[RP_MAXD];
      for (int k = 0; k < 9; k++) A[k] = p.ld(q + PC_A + k);
      for (int r = 0; r < 3; r++) { Rr[r] = p.ld(q + PC_R + r); br[r] = p.ld(q + PC_B + r); jk[r] = p.ld(q + PC_JK + r); }
      for (int d = 0; d < RP_MAXD; d++) {
        chain[d] = (int)p.ld(q + PC_CHAIN + d);  // zero-padded past the chain (as are the Jacobian entries)
        for (int r = 0; r < 3; r++) jc[r][d] = p.ld(q + PC_JC + r * RP_MAXD + d);
      }
      const double vak = ck >= 0 ? p.at(PG_CK_VA + ck) : 0.0;
      for (int r = 0; r < 3; r++) {
        f[r] = old[r] = p.at(q + PC_F + r);
        double s = 0;
        for (int d = 0; d <= dep; d++) s += jc[r][d] * va[(h * NL + chain[d]) * va_stride];
        res[r] = br[r] + Rr[r] * f[r] + s + (ck >= 0 ? jk[r] * vak : 0.0);
      }
      const double change = pgs_elliptic_update(A, res, f, mu);
      improvement -= change;
      const double df[3] = {f[0] - old[0], f[1] - old[1], f[2] - old[2]};
      for (int r = 0; r < 3; r++) p.at(q + PC_F + r) = f[r];
      if (df[0] != 0.0 || df[1] != 0.0 || df[2] != 0.0) {
        if (ck >= 0) p.at(PG_CK_VA + ck) = vak + (jk[0] * df[0] + jk[1] * df[1] + jk[2] * df[2]) * invMk;
        for (int j = 0; j < NL; j++)
          va[(h * NL + j) * va_stride] += fma(p.ld(q + PC_U + 2 * NL + j), df[2], fma(p.ld(q + PC_U + NL + j), df[1], p.ld(q + PC_U + j) * df[0]));
      }
    }
    improvement *= solver_scale;
    iter++;
    if (improvement < tol) break;
  }
  for (int i = 0; i < 2 * NL; i++) p.at(PG_VA + i) = va[i * va_stride];
  p.at(PG_NITER) = (double)iter;
  return iter;
}

}  // namespace rp
