// Host emulation of ldl_inverse (csrc/rp_kernels.cu, RP_MINV_LDL: experimental) — lanes as an outer loop, one loop nest per
// barrier phase — so that its index logic (packed rows, subtree / ancestor gathers, folded 1/D) is checked against inv(M).
#include <cstdio>
#include <cmath>
#include <cstring>
#define NL 26
#define MAXD 10
#define RS 27
extern "C" void ldl_host(double* A, const double* Mrow, const double* extra, const unsigned char* moff, const signed char* anc, const signed char* dep, const signed char* sub) {
  double F[160 + 32]; double* Dinv = F + 160;
  for (int lane = 0; lane < NL; lane++) { int md = dep[lane], mo = moff[lane]; for (int d = 0; d < md; d++) F[mo + d] = Mrow[mo + d]; F[mo + md] = Mrow[mo + md] + extra[lane]; }
  for (int k = NL - 1; k >= 0; k--) {
    int depk = dep[k]; const double* rk = F + moff[k]; double dinv = 1.0 / rk[depk];
    double tt[NL]; 
    if (depk > 0) {
      for (int lane = 0; lane < NL; lane++) { bool isanc = lane < k && k < lane + sub[lane]; tt[lane] = 0; if (isanc) { int md = dep[lane], mo = moff[lane]; double t = rk[md] * dinv; tt[lane] = t; double r[MAXD], m[MAXD]; for (int d = 0; d < MAXD; d++) { int dd = d <= md ? d : md; r[d] = rk[dd]; m[d] = F[mo + dd]; } for (int d = 0; d < MAXD; d++) if (d <= md) F[mo + d] = m[d] - r[d] * t; } }
      for (int lane = 0; lane < NL; lane++) { bool isanc = lane < k && k < lane + sub[lane]; if (isanc) F[moff[k] + dep[lane]] = tt[lane]; }
    }
    Dinv[k] = dinv;
  }
  for (int lane = 0; lane < NL; lane++) {
    double* x = A + lane;
    for (int i = 0; i < NL; i++) x[i * RS] = i == lane ? -1.0 : 0.0;
    for (int i = NL - 1; i >= 0; i--) { int di = dep[i], end = i + sub[i]; double acc = x[i * RS]; for (int k = i + 1; k < end; k++) acc -= F[moff[k] + di] * x[k * RS]; x[i * RS] = acc; }
    for (int k = 0; k < NL; k++) { int dk = dep[k]; const double* lk = F + moff[k]; double acc = x[k * RS] * Dinv[k]; for (int d = 0; d < dk; d++) acc -= lk[d] * x[anc[k * MAXD + d] * RS]; x[k * RS] = acc; }
  }
}
