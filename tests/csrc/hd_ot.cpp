// Host build of the assignment solver and the optimal-transport fingering reward of csrc/rp_math.cuh (tests/test_hd_math.py).
#include "rp_math.cuh"
extern "C" void hd_lsa(int nr, int nc, const double* cost, int* col4row) {
  double wd[10 + 2 * 128]; int wi[5 * 128];
  rp::lsa_rect(nr, nc, cost, col4row, wd, wi);
}
extern "C" double hd_ot_reward(const double* fingertip, const double* keypos, int K) {
  double wd[10 * 88 + 10 + 2 * 88]; int wi[4 * 88 + 20];
  return rp::ot_fingering_reward(fingertip, keypos, K, wd, wi);
}
