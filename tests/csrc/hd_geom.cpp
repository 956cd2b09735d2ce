// Host build of the step kernel's geometry / constraint-parameter helpers (csrc/rp_math.cuh) for tests/test_hd_math.py.
#include "rp_math.cuh"
extern "C" {
double hd_sphere_box(const double* c, double r, const double* s, double* n, double* p) { return rp::sphere_box(c, r, s, n, p); }
double hd_segment_box_closest_t(const double* a, const double* b, const double* s) { return rp::segment_box_closest_t(a, b, s); }
double hd_point_box_d2(const double* c, const double* s) { return rp::point_box_d2(c, s); }
void hd_make_frame(double* frame) { rp::make_frame(frame); }
double hd_impedance(const double* solimp, double pos, double margin) { return rp::impedance(solimp, pos, margin); }
void hd_ref_KB(const double* solref, double dmax, double h, double* K, double* B) { rp::ref_KB(solref, dmax, h, K, B); }
void hd_elliptic_primal_force(const double* jar, const double* D, double mu, double fr, double* f) { rp::elliptic_primal_force(jar, D, mu, fr, f); }
}
