#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "rp_math.cuh"
static double rnd() { return rand() / (double)RAND_MAX; }
int main() {
  srand(1); int bad = 0, nz = 0, slide = 0;
  for (int t = 0; t < 2000000; t++) {
    double L[9] = {0}, A[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j <= i; j++) L[3*i+j] = (i == j ? 0.3 + rnd() : rnd() - 0.5);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += L[3*i+k]*L[3*j+k]; A[3*i+j] = s; }
    double mu = 0.2 + rnd();
    double f[3] = {0, 0, 0};
    int mode = t % 4;
    if (mode >= 1) { f[0] = rnd() * 2; double th = rnd() * 6.28, rr = (mode == 3 ? 1.0 : rnd()) * mu * f[0]; f[1] = rr * cos(th); f[2] = rr * sin(th); }
    double res[3] = {4 * (rnd() - 0.5), 4 * (rnd() - 0.5), 4 * (rnd() - 0.5)};
    double f1[3] = {f[0], f[1], f[2]}, f2[3] = {f[0], f[1], f[2]};
    double c1 = rp::pgs_elliptic_update(A, res, f1, mu);
    rp::EllPre p; rp::pgs_elliptic_pre(p, A, f2, mu);
    double c2 = rp::pgs_elliptic_post(p, A, res, f2, mu);
    double f3[3] = {f[0], f[1], f[2]};
    rp::EllInv q; int qok = rp::pgs_elliptic_invariants(q, A, mu); int dok; double rden = rp::pgs_elliptic_ray(A, f3, &dok);
    double c3 = rp::pgs_elliptic_apply(q, qok, rden, dok, A, res, f3, mu);
    if (memcmp(f1, f3, sizeof f1) || memcmp(&c1, &c3, 8)) { if (bad < 5) printf("mismatch3 t=%d %g %g | %g %g %g | %g %g %g\n", t, c1, c3, f1[0], f1[1], f1[2], f3[0], f3[1], f3[2]); bad++; }
    if (memcmp(f1, f2, sizeof f1) || memcmp(&c1, &c2, 8)) { if (bad < 5) printf("mismatch t=%d %g %g | %g %g %g | %g %g %g\n", t, c1, c2, f1[0], f1[1], f1[2], f2[0], f2[1], f2[2]); bad++; }
    if (f1[0] > 0) nz++;
    if (f1[0] > 0 && fabs(f1[1]*f1[1] + f1[2]*f1[2] - mu*mu*f1[0]*f1[0]) < 1e-9) slide++;
  }
  printf("bad %d nonzero %d sliding %d\n", bad, nz, slide);
  return bad != 0;
}
