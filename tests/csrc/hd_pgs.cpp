// Host build of csrc/rp_pgs.cuh for tests/test_pgs_soa.py (g++ -shared): the SoA Gauss-Seidel solve, one environment.
#include <cstddef>
#include "rp_pgs.cuh"
extern "C" int hd_pgs_size(void) { return rp::PG_SIZE; }
extern "C" int hd_pgs_offset(int which) {
  const int t[] = {rp::PG_MINV, rp::PG_FR, rp::PG_FR_F, rp::PG_LM, rp::PG_LM_F, rp::PG_VA, rp::PG_CK, rp::PG_CK_F, rp::PG_CK_VA,
                   rp::PG_CON, rp::PG_CON_STRIDE, rp::PC_HAND, rp::PC_CK, rp::PC_DEP, rp::PC_MU, rp::PC_INVMK, rp::PC_JK, rp::PC_A,
                   rp::PC_R, rp::PC_B, rp::PC_F, rp::PC_CHAIN, rp::PC_JC, rp::PC_U, RP_MAXD, RP_MAXCON, RP_NL};
  return t[which];
}
extern "C" int hd_pgs_solve(double* P, size_t stride, int env, int maxiter, double tol, double scale) {
  double va[2 * RP_NL];
  rp::PgsView v{P, stride, env};
  return rp::pgs_solve_env(v, maxiter, tol, scale, va, 1);
}
