// Host build of csrc/rp_math.cuh: hand geom pair narrowphase and the oriented-box overlap test, driven by stdin.
// line "P margin p1(3) z1(3) r1 h1 p2(3) z2(3) r2 h2" -> "n dist normal(3) pos(3) [dist normal pos]"
// line "B ca(3) Ra(9) ha(3) cb(3) Rb(9) hb(3)"       -> "0|1"
#include <cstdio>
#include <cmath>
#include "rp_math.cuh"
int main() {
  char kind;
  while (scanf(" %c", &kind) == 1) {
    if (kind == 'P') {
      double m, p1[3], z1[3], r1, h1, p2[3], z2[3], r2, h2;
      scanf("%lf", &m);
      for (int k = 0; k < 3; k++) scanf("%lf", &p1[k]);
      for (int k = 0; k < 3; k++) scanf("%lf", &z1[k]);
      scanf("%lf %lf", &r1, &h1);
      for (int k = 0; k < 3; k++) scanf("%lf", &p2[k]);
      for (int k = 0; k < 3; k++) scanf("%lf", &z2[k]);
      scanf("%lf %lf", &r2, &h2);
      double dist[2], nrm[6], pos[6];
      int n = rp::sphere_capsule_pair(m, p1, z1, r1, h1, p2, z2, r2, h2, dist, nrm, pos);
      printf("%d", n);
      for (int i = 0; i < n; i++) printf(" %.17g %.17g %.17g %.17g %.17g %.17g %.17g", dist[i], nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
      printf("\n");
    } else {
      double v[30];
      for (int k = 0; k < 30; k++) scanf("%lf", &v[k]);
      printf("%d\n", rp::obb_overlap(v, v + 3, v + 12, v + 15, v + 18, v + 27) ? 1 : 0);
    }
  }
  return 0;
}
