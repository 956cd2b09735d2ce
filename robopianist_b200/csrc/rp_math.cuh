// rp_math.cuh — scalar building blocks of the step kernel (host+device so they can be unit-tested on the
// CPU: tests/test_hd_math.py builds tests/csrc/hd_*.cpp with g++ and checks them against brute force,
// MuJoCo's documented formulas and each other).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define RP_HD __host__ __device__ __forceinline__
#else
#define RP_HD inline
#endif

#define RP_MINVAL 1e-15
#define RP_MINIMP 0.0001
#define RP_MAXIMP 0.9999

namespace rp {

// 1/x: on the device MUFU.RCP64H seed + two Newton steps (about 1 ulp, no special-case slow path;
// the callers guard against zero/denormal denominators); plain division on the host.
RP_HD double rcp(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

RP_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
RP_HD void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
RP_HD void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
RP_HD void quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < RP_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double inv = 1.0 / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
RP_HD void quat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
RP_HD void mat_vec(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
RP_HD void matT_vec(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  double y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  double z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// spatial algebra at the world origin: motion V=[w;v], force F=[n;f], inertia I={m,h[3],Ixx,Iyy,Izz,Ixy,Ixz,Iyz}
RP_HD void inertia_mul(double* F, const double* I, const double* V) {
  const double *w = V, *v = V + 3, *h = I + 1;
  double Iw0 = I[4] * w[0] + I[7] * w[1] + I[8] * w[2];
  double Iw1 = I[7] * w[0] + I[5] * w[1] + I[9] * w[2];
  double Iw2 = I[8] * w[0] + I[9] * w[1] + I[6] * w[2];
  double hxv[3], wxh[3];
  cross3(hxv, h, v);
  cross3(wxh, w, h);
  F[0] = Iw0 + hxv[0]; F[1] = Iw1 + hxv[1]; F[2] = Iw2 + hxv[2];
  F[3] = I[0] * v[0] + wxh[0]; F[4] = I[0] * v[1] + wxh[1]; F[5] = I[0] * v[2] + wxh[2];
}
RP_HD void cross_motion(double* r, const double* V, const double* S) {
  double a[3], b[3], c[3];
  cross3(a, V, S);
  cross3(b, V, S + 3);
  cross3(c, V + 3, S);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
RP_HD void cross_force(double* r, const double* V, const double* F) {
  double a[3], b[3], c[3];
  cross3(a, V, F);
  cross3(b, V + 3, F + 3);
  cross3(c, V, F + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
RP_HD double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
// body inertia {mass, principal inertia di} with COM c (world) and inertial orientation R (world) -> at origin
RP_HD void body_inertia_world(double* out, double mass, const double* di, const double* R, const double* c) {
  double I00 = R[0] * di[0] * R[0] + R[1] * di[1] * R[1] + R[2] * di[2] * R[2];
  double I11 = R[3] * di[0] * R[3] + R[4] * di[1] * R[4] + R[5] * di[2] * R[5];
  double I22 = R[6] * di[0] * R[6] + R[7] * di[1] * R[7] + R[8] * di[2] * R[8];
  double I01 = R[0] * di[0] * R[3] + R[1] * di[1] * R[4] + R[2] * di[2] * R[5];
  double I02 = R[0] * di[0] * R[6] + R[1] * di[1] * R[7] + R[2] * di[2] * R[8];
  double I12 = R[3] * di[0] * R[6] + R[4] * di[1] * R[7] + R[5] * di[2] * R[8];
  double cc = dot3(c, c);
  out[0] = mass;
  out[1] = mass * c[0]; out[2] = mass * c[1]; out[3] = mass * c[2];
  out[4] = I00 + mass * (cc - c[0] * c[0]);
  out[5] = I11 + mass * (cc - c[1] * c[1]);
  out[6] = I22 + mass * (cc - c[2] * c[2]);
  out[7] = I01 - mass * c[0] * c[1];
  out[8] = I02 - mass * c[0] * c[2];
  out[9] = I12 - mass * c[1] * c[2];
}

// solimp -> impedance d(r) (MuJoCo getimpedance), with the usual clamps
RP_HD double impedance(const double* solimp, double pos, double margin) {
  double d0 = fmin(fmax(solimp[0], RP_MINIMP), RP_MAXIMP), d1 = fmin(fmax(solimp[1], RP_MINIMP), RP_MAXIMP);
  double width = fmax(solimp[2], 0.0), mid = fmin(fmax(solimp[3], RP_MINIMP), RP_MAXIMP), power = fmax(solimp[4], 1.0);
  if (d0 == d1 || width <= RP_MINVAL) return 0.5 * (d0 + d1);
  double x = fabs((pos - margin) / width);
  if (x >= 1) return d1;
  if (x <= 0) return d0;
  double y;
  if (power == 1) y = x;
  else if (power == 2) y = (x <= mid) ? x * x / mid : 1 - (1 - x) * (1 - x) / (1 - mid);
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return d0 + y * (d1 - d0);
}
// solref (+dmax, timestep) -> stiffness K and damping B of the reference acceleration
RP_HD void ref_KB(const double* solref, double dmax_in, double timestep, double* K, double* B) {
  double tc = solref[0], dr = solref[1];
  double dmax = fmin(fmax(dmax_in, RP_MINIMP), RP_MAXIMP);
  if (tc > 0 && tc < 2 * timestep) tc = 2 * timestep;  // refsafe
  if (tc > 0) {
    *K = 1.0 / fmax(RP_MINVAL, dmax * dmax * tc * tc * dr * dr);
    *B = 2.0 / fmax(RP_MINVAL, dmax * tc);
  } else {
    *K = -tc / fmax(RP_MINVAL, dmax * dmax);
    *B = -dr / fmax(RP_MINVAL, dmax);
  }
}

// sphere (centre c in box frame, radius r) vs box half sizes s. Returns signed distance; n = unit
// normal (box frame) from the sphere towards the box; p = point mid-way between the surfaces.
RP_HD double sphere_box(const double* c, double r, const double* s, double* n, double* p) {
  double cl[3], dl[3];
  bool inside = true;
  for (int k = 0; k < 3; k++) {
    cl[k] = c[k] < -s[k] ? -s[k] : (c[k] > s[k] ? s[k] : c[k]);
    dl[k] = c[k] - cl[k];
    if (dl[k] != 0) inside = false;
  }
  double dist;
  if (!inside) {
    double len = sqrt(dot3(dl, dl));
    dist = len - r;
    for (int k = 0; k < 3; k++) n[k] = -dl[k] / len;
    for (int k = 0; k < 3; k++) p[k] = cl[k] - n[k] * 0.5 * dist;
  } else {
    int best = 0;
    double bd = 1e300;
    for (int k = 0; k < 3; k++) {
      double dk = s[k] - fabs(c[k]);
      if (dk < bd) { bd = dk; best = k; }
    }
    double sg = c[best] >= 0 ? 1.0 : -1.0;
    n[0] = n[1] = n[2] = 0;
    for (int k = 0; k < 3; k++) { if (k == best) n[k] = -sg; p[k] = c[k]; }
    dist = -bd - r;
    for (int k = 0; k < 3; k++) if (k == best) p[k] = sg * s[k];
    for (int k = 0; k < 3; k++) p[k] -= n[k] * 0.5 * dist;
  }
  return dist;
}
// ---- hand geom : hand geom narrowphase (spheres and capsules): MuJoCo's analytic routines
// (engine_collision_primitive.c: mjraw_SphereSphere, mjc_SphereCapsule, mjraw_CapsuleCapsule), restated.
// A capsule is (centre, unit axis, radius, half length); outputs: signed distance, unit normal from
// geom 1 to geom 2, contact point mid-way between the surfaces.  Return the number of contacts.
RP_HD int sphere_sphere(double margin, const double* p1, const double* z1, double r1, const double* p2, const double* z2, double r2,
                        double* dist, double* normal, double* pos) {
  double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
  const double cd2 = dot3(dif, dif), mind = margin + r1 + r2;
  if (cd2 > mind * mind) return 0;
  double len = sqrt(cd2);
  if (len < RP_MINVAL) { dif[0] = 1; dif[1] = 0; dif[2] = 0; len = 0; }
  else for (int k = 0; k < 3; k++) dif[k] /= len;
  *dist = len - r1 - r2;
  for (int k = 0; k < 3; k++) normal[k] = dif[k];
  if (len < RP_MINVAL) {  // coincident centres: cross product of the z axes, or x
    cross3(normal, z1, z2);
    const double nn = sqrt(dot3(normal, normal));
    if (nn < RP_MINVAL) { normal[0] = 1; normal[1] = 0; normal[2] = 0; }
    else for (int k = 0; k < 3; k++) normal[k] /= nn;
  }
  for (int k = 0; k < 3; k++) pos[k] = p1[k] + normal[k] * (r1 + 0.5 * *dist);
  return 1;
}
RP_HD double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }
RP_HD int capsule_capsule(double margin, const double* p1, const double* ax1, double r1, double h1, const double* p2, const double* ax2,
                          double r2, double h2, double* dist, double* normal, double* pos) {  // outputs [2]
  const double dif[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
  const double ma = dot3(ax1, ax1), mb = -dot3(ax1, ax2), mc = dot3(ax2, ax2);
  const double u = -dot3(ax1, dif), v = dot3(ax2, dif);
  const double det = ma * mc - mb * mb;
  double v1[3], v2[3];
  if (fabs(det) >= RP_MINVAL) {
    double x1 = (mc * u - mb * v) / det, x2 = (ma * v - mb * u) / det;
    if (x1 > h1) { x1 = h1; x2 = (v - mb * h1) / mc; }
    else if (x1 < -h1) { x1 = -h1; x2 = (v + mb * h1) / mc; }
    if (x2 > h2) { x2 = h2; x1 = clipd((u - mb * h2) / ma, -h1, h1); }
    else if (x2 < -h2) { x2 = -h2; x1 = clipd((u + mb * h2) / ma, -h1, h1); }
    for (int k = 0; k < 3; k++) { v1[k] = p1[k] + ax1[k] * x1; v2[k] = p2[k] + ax2[k] * x2; }
    return sphere_sphere(margin, v1, ax1, r1, v2, ax2, r2, dist, normal, pos);
  }
  int n = 0;  // parallel axes: the ends of capsule 1 against segment 2, then the ends of capsule 2 against segment 1; at most two
  for (int pass = 0; pass < 4 && n < 2; pass++) {
    double x1, x2;
    if (pass == 0) { x1 = h1; x2 = clipd((v - mb * h1) / mc, -h2, h2); }
    else if (pass == 1) { x1 = -h1; x2 = clipd((v + mb * h1) / mc, -h2, h2); }
    else if (pass == 2) { x2 = h2; x1 = clipd((u - mb * h2) / ma, -h1, h1); }
    else { x2 = -h2; x1 = clipd((u + mb * h2) / ma, -h1, h1); }
    for (int k = 0; k < 3; k++) { v1[k] = p1[k] + ax1[k] * x1; v2[k] = p2[k] + ax2[k] * x2; }
    n += sphere_sphere(margin, v1, ax1, r1, v2, ax2, r2, dist + n, normal + 3 * n, pos + 3 * n);
  }
  return n;
}
// One pair; geom 1 must not be a capsule when geom 2 is a sphere (mj_collideGeoms orders by type).  h = 0: sphere.
RP_HD int sphere_capsule_pair(double margin, const double* p1, const double* z1, double r1, double h1, const double* p2, const double* z2,
                              double r2, double h2, double* dist, double* normal, double* pos) {
  if (h1 == 0 && h2 == 0) return sphere_sphere(margin, p1, z1, r1, p2, z2, r2, dist, normal, pos);
  if (h1 == 0) {  // mjc_SphereCapsule: nearest point of the segment, then sphere : sphere
    const double vec[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]};
    const double x = clipd(dot3(z2, vec), -h2, h2);
    const double q[3] = {p2[0] + z2[0] * x, p2[1] + z2[1] * x, p2[2] + z2[2] * x};
    return sphere_sphere(margin, p1, z1, r1, q, z2, r2, dist, normal, pos);
  }
  return capsule_capsule(margin, p1, z1, r1, h1, p2, z2, r2, h2, dist, normal, pos);
}
// separating-axis test of two oriented boxes (centre, rotation with the box axes as columns, row major; half sizes)
RP_HD bool obb_overlap(const double* ca, const double* Ra, const double* ha, const double* cb, const double* Rb, const double* hb) {
  double R[9], AR[9], t[3];
  const double d[3] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2]};
  for (int i = 0; i < 3; i++) {
    t[i] = d[0] * Ra[i] + d[1] * Ra[3 + i] + d[2] * Ra[6 + i];
    for (int j = 0; j < 3; j++) {
      R[3 * i + j] = Ra[i] * Rb[j] + Ra[3 + i] * Rb[3 + j] + Ra[6 + i] * Rb[6 + j];
      AR[3 * i + j] = fabs(R[3 * i + j]) + 1e-12;
    }
  }
  for (int i = 0; i < 3; i++) if (fabs(t[i]) > ha[i] + hb[0] * AR[3 * i] + hb[1] * AR[3 * i + 1] + hb[2] * AR[3 * i + 2]) return false;
  for (int j = 0; j < 3; j++)
    if (fabs(t[0] * R[j] + t[1] * R[3 + j] + t[2] * R[6 + j]) > hb[j] + ha[0] * AR[j] + ha[1] * AR[3 + j] + ha[2] * AR[6 + j]) return false;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double ra = ha[i1] * AR[3 * i2 + j] + ha[i2] * AR[3 * i1 + j], rb = hb[j1] * AR[3 * i + j2] + hb[j2] * AR[3 * i + j1];
      if (fabs(t[i2] * R[3 * i1 + j] - t[i1] * R[3 * i2 + j]) > ra + rb) return false;
    }
  return true;
}
RP_HD double point_box_d2(const double* c, const double* s) {
  double d2 = 0;
  for (int k = 0; k < 3; k++) {
    double e = fabs(c[k]) - s[k];
    if (e > 0) d2 += e * e;
  }
  return d2;
}
// t in [0,1] minimising dist(a + t (b-a), box): convex piecewise quadratic, breakpoints where a
// coordinate crosses a face plane; ties keep the smallest t.
RP_HD double segment_box_closest_t(const double* a, const double* b, const double* s) {
  double bp[8];
  int nbp = 0;
  bp[nbp++] = 0; bp[nbp++] = 1;
  double dir[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  for (int k = 0; k < 3; k++) {
    if (fabs(dir[k]) >= RP_MINVAL) {
      double t1 = (-s[k] - a[k]) / dir[k], t2 = (s[k] - a[k]) / dir[k];
      if (t1 > 0 && t1 < 1) bp[nbp++] = t1;
      if (t2 > 0 && t2 < 1) bp[nbp++] = t2;
    }
  }
  for (int i = 1; i < nbp; i++) {
    double v = bp[i];
    int j = i - 1;
    while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; j--; }
    bp[j + 1] = v;
  }
  double best_t = 0, best_d = 1e300;
  for (int i = 0; i + 1 < nbp; i++) {
    double t0 = bp[i], t1 = bp[i + 1];
    if (t1 - t0 >= 1e-14) {
    double tm = 0.5 * (t0 + t1), qa = 0, qb = 0;
    for (int k = 0; k < 3; k++) {
      double pm = a[k] + tm * dir[k];
      if (pm > s[k]) { qa += dir[k] * dir[k]; qb += 2 * dir[k] * (a[k] - s[k]); }
      else if (pm < -s[k]) { qa += dir[k] * dir[k]; qb += 2 * dir[k] * (a[k] + s[k]); }
    }
    double tc = t0;
    if (qa > RP_MINVAL) { tc = -qb / (2 * qa); tc = tc < t0 ? t0 : (tc > t1 ? t1 : tc); }
    double cand[3] = {t0, tc, t1};
    for (int c = 0; c < 3; c++) {
      double p[3] = {a[0] + cand[c] * dir[0], a[1] + cand[c] * dir[1], a[2] + cand[c] * dir[2]};
      double d2 = point_box_d2(p, s);
      if (d2 < best_d - 1e-18) { best_d = d2; best_t = cand[c]; }
    }
    }
  }
  return best_t;
}
// mju_makeFrame: frame[0:3] unit normal -> frame[3:6], frame[6:9] tangents
RP_HD void make_frame(double* frame) {
  double *x = frame, *y = frame + 3, *z = frame + 6;
  y[0] = y[1] = y[2] = 0;
  if (x[1] < 0.5 && x[1] > -0.5) y[1] = 1; else y[2] = 1;
  double t = dot3(x, y);
  for (int k = 0; k < 3; k++) y[k] -= t * x[k];
  double n = sqrt(dot3(y, y));
  for (int k = 0; k < 3; k++) y[k] /= n;
  cross3(z, x, y);
}
// mju_QCQP2
RP_HD int qcqp2(double* res, const double* Ain, const double* bin, const double* d, double r) {
  double b1 = bin[0] * d[0], b2 = bin[1] * d[1];
  double A11 = Ain[0] * d[0] * d[0], A22 = Ain[3] * d[1] * d[1], A12 = Ain[1] * d[0] * d[1];
  double la = 0, v1 = 0, v2 = 0;
  for (int iter = 0; iter < 20; iter++) {
    double det = (A11 + la) * (A22 + la) - A12 * A12;
    if (det < 1e-10) { res[0] = 0; res[1] = 0; return 0; }
    double detinv = rcp(det), P11 = (A22 + la) * detinv, P22 = (A11 + la) * detinv, P12 = -A12 * detinv;
    v1 = -P11 * b1 - P12 * b2;
    v2 = -P12 * b1 - P22 * b2;
    double val = v1 * v1 + v2 * v2 - r * r;
    if (val < 1e-10) break;
    double deriv = -2 * (P11 * v1 * v1 + 2 * P12 * v1 * v2 + P22 * v2 * v2);
    double delta = -val * rcp(deriv);
    if (delta < 1e-10) break;
    la += delta;
  }
  res[0] = v1 * d[0];
  res[1] = v2 * d[1];
  return la != 0;
}
// One PGS update of an elliptic condim-3 contact block.  A = 3x3 block of AR (row major, with R on
// the diagonal), res = residual b + AR*f at the current force, f = force (in/out), mu = friction.
// Returns the cost change (<= 0), reverting the update when it is positive.
RP_HD double pgs_elliptic_update(const double* A, const double* res, double* f, double mu) {
  double old[3] = {f[0], f[1], f[2]};
  double fr[2] = {mu, mu};
  if (f[0] < RP_MINVAL) {
    f[0] -= res[0] * rcp(A[0]);
    if (f[0] < 0) f[0] = 0;
    f[1] = f[2] = 0;
  } else {
    double v1[3];
    for (int r = 0; r < 3; r++) v1[r] = A[3 * r] * old[0] + A[3 * r + 1] * old[1] + A[3 * r + 2] * old[2];
    double denom = dot3(old, v1);
    if (denom >= RP_MINVAL) {
      double x = -dot3(old, res) * rcp(denom);
      if (f[0] + x * old[0] < 0) x = -1.0;
      for (int r = 0; r < 3; r++) f[r] += x * old[r];
    }
  }
  if (f[0] >= RP_MINVAL) {
    double Ac[4] = {A[4], A[5], A[7], A[8]}, bc[2], vv[2];
    for (int r = 1; r < 3; r++) {
      double t = res[r];
      for (int s = 1; s < 3; s++) t -= A[3 * r + s] * old[s];
      t += A[3 * r] * (f[0] - old[0]);
      bc[r - 1] = t;
    }
    int flg = qcqp2(vv, Ac, bc, fr, f[0]);
    if (flg) {
      double imu2 = rcp(fr[0] * fr[0]);
      double s = vv[0] * vv[0] * imu2 + vv[1] * vv[1] * imu2;
      s = sqrt(f[0] * f[0] * rcp(fmax(RP_MINVAL, s)));
      vv[0] *= s; vv[1] *= s;
    }
    f[1] = vv[0]; f[2] = vv[1];
  }
  double dl[3] = {f[0] - old[0], f[1] - old[1], f[2] - old[2]}, Ad[3];
  for (int r = 0; r < 3; r++) Ad[r] = A[3 * r] * dl[0] + A[3 * r + 1] * dl[1] + A[3 * r + 2] * dl[2];
  double change = 0.5 * dot3(dl, Ad) + dot3(dl, res);
  if (change > 1e-10) { f[0] = old[0]; f[1] = old[1]; f[2] = old[2]; change = 0; }
  return change;
}
// ---- The same update split in two, for the warp-wide Gauss-Seidel of the step kernel: everything that
// does not depend on the residual (the ray denominators, the lambda = 0 pass of the QCQP) is evaluated by
// pgs_elliptic_pre while the J*v reduction is still in flight, so that only the residual-dependent
// operations stay on the serial chain.  Same operations, same order, same results as pgs_elliptic_update.
struct QcqpPre { double A11, A22, A12, P11, P22, P12; int ok; };
struct EllPre { double old[3], rA0, rden, imu2; QcqpPre q; int lowf, den_ok; };
RP_HD void qcqp2_prepare(QcqpPre& q, const double* Ain, const double* d) {
  q.A11 = Ain[0] * d[0] * d[0]; q.A22 = Ain[3] * d[1] * d[1]; q.A12 = Ain[1] * d[0] * d[1];
  const double det = q.A11 * q.A22 - q.A12 * q.A12;  // (A11 + la) * (A22 + la) - A12^2 at la = 0
  q.ok = !(det < 1e-10);
  const double detinv = rcp(det);  // unused when !ok
  q.P11 = q.A22 * detinv; q.P22 = q.A11 * detinv; q.P12 = -q.A12 * detinv;
}
RP_HD int qcqp2_solve(double* res, const QcqpPre& q, const double* bin, const double* d, double r) {
  const double b1 = bin[0] * d[0], b2 = bin[1] * d[1];
  if (!q.ok) { res[0] = 0; res[1] = 0; return 0; }
  double la = 0, P11 = q.P11, P22 = q.P22, P12 = q.P12, v1, v2;
  for (int iter = 0;;) {
    v1 = -P11 * b1 - P12 * b2;
    v2 = -P12 * b1 - P22 * b2;
    const double val = v1 * v1 + v2 * v2 - r * r;
    if (val < 1e-10) break;
    const double deriv = -2 * (P11 * v1 * v1 + 2 * P12 * v1 * v2 + P22 * v2 * v2);
    const double delta = -val * rcp(deriv);
    if (delta < 1e-10) break;
    la += delta;
    if (++iter >= 20) break;
    const double det = (q.A11 + la) * (q.A22 + la) - q.A12 * q.A12;
    if (det < 1e-10) { res[0] = 0; res[1] = 0; return 0; }
    const double detinv = rcp(det);
    P11 = (q.A22 + la) * detinv; P22 = (q.A11 + la) * detinv; P12 = -q.A12 * detinv;
  }
  res[0] = v1 * d[0];
  res[1] = v2 * d[1];
  return la != 0;
}
RP_HD void pgs_elliptic_pre(EllPre& p, const double* A, const double* f, double mu) {
  p.old[0] = f[0]; p.old[1] = f[1]; p.old[2] = f[2];
  p.lowf = f[0] < RP_MINVAL;
  p.rA0 = rcp(A[0]);
  double v1[3];
  for (int r = 0; r < 3; r++) v1[r] = A[3 * r] * f[0] + A[3 * r + 1] * f[1] + A[3 * r + 2] * f[2];
  const double denom = dot3(f, v1);
  p.den_ok = denom >= RP_MINVAL;
  p.rden = rcp(denom);  // unused when !den_ok
  const double Ac[4] = {A[4], A[5], A[7], A[8]}, fr[2] = {mu, mu};
  qcqp2_prepare(p.q, Ac, fr);
  p.imu2 = rcp(fr[0] * fr[0]);
}
RP_HD double pgs_elliptic_post(const EllPre& p, const double* A, const double* res, double* f, double mu) {
  const double* old = p.old;
  const double fr[2] = {mu, mu};
  if (p.lowf) {
    f[0] -= res[0] * p.rA0;
    if (f[0] < 0) f[0] = 0;
    f[1] = f[2] = 0;
  } else if (p.den_ok) {
    double x = -dot3(old, res) * p.rden;
    if (f[0] + x * old[0] < 0) x = -1.0;
    for (int r = 0; r < 3; r++) f[r] += x * old[r];
  }
  if (f[0] >= RP_MINVAL) {
    double bc[2], vv[2];
    for (int r = 1; r < 3; r++) {
      double t = res[r];
      for (int s = 1; s < 3; s++) t -= A[3 * r + s] * old[s];
      t += A[3 * r] * (f[0] - old[0]);
      bc[r - 1] = t;
    }
    const int flg = qcqp2_solve(vv, p.q, bc, fr, f[0]);
    if (flg) {
      const double imu2 = p.imu2;
      double s = vv[0] * vv[0] * imu2 + vv[1] * vv[1] * imu2;
      s = sqrt(f[0] * f[0] * rcp(fmax(RP_MINVAL, s)));
      vv[0] *= s; vv[1] *= s;
    }
    f[1] = vv[0]; f[2] = vv[1];
  }
  const double dl[3] = {f[0] - old[0], f[1] - old[1], f[2] - old[2]};
  double Ad[3];
  for (int r = 0; r < 3; r++) Ad[r] = A[3 * r] * dl[0] + A[3 * r + 1] * dl[1] + A[3 * r + 2] * dl[2];
  double change = 0.5 * dot3(dl, Ad) + dot3(dl, res);
  if (change > 1e-10) { f[0] = old[0]; f[1] = old[1]; f[2] = old[2]; change = 0; }
  return change;
}
// ---- Three-way split used by the step kernel's solver (pgs_solve): (1) what depends only on the block A and the
// friction coefficient is evaluated once per substep and kept with the contact (EllInv: 1/A00, the scaled
// friction block and its inverse at lambda = 0, 1/mu^2); (2) what depends on the force of the previous sweep but not
// on the residual (the ray denominator) runs in the shadow of the J*v reduction; (3) the residual-dependent rest.
// Same operations in the same order as pgs_elliptic_update: the results are bit-identical (tests/test_hd_math.py).
struct EllInv { double rA0, A11, A22, A12, P11, P22, P12, imu2; };
RP_HD int pgs_elliptic_invariants(EllInv& q, const double* A, double mu) {  // returns the "friction block regular" flag
  q.rA0 = rcp(A[0]);
  const double Ac[4] = {A[4], A[5], A[7], A[8]}, fr[2] = {mu, mu};
  QcqpPre t;
  qcqp2_prepare(t, Ac, fr);
  q.A11 = t.A11; q.A22 = t.A22; q.A12 = t.A12; q.P11 = t.P11; q.P22 = t.P22; q.P12 = t.P12;
  q.imu2 = rcp(fr[0] * fr[0]);
  return t.ok;
}
RP_HD double pgs_elliptic_ray(const double* A, const double* f, int* den_ok) {  // 1 / (f' A f), valid when *den_ok
  double v1[3];
  for (int r = 0; r < 3; r++) v1[r] = A[3 * r] * f[0] + A[3 * r + 1] * f[1] + A[3 * r + 2] * f[2];
  const double denom = dot3(f, v1);
  *den_ok = denom >= RP_MINVAL;
  return rcp(denom);
}
RP_HD double pgs_elliptic_apply(const EllInv& q, int qok, double rden, int den_ok, const double* A, const double* res,
                                double* f, double mu) {
  const double old[3] = {f[0], f[1], f[2]};
  if (old[0] < RP_MINVAL) {
    f[0] -= res[0] * q.rA0;
    if (f[0] < 0) f[0] = 0;
    f[1] = f[2] = 0;
  } else if (den_ok) {
    double x = -dot3(old, res) * rden;
    if (f[0] + x * old[0] < 0) x = -1.0;
    for (int r = 0; r < 3; r++) f[r] += x * old[r];
  }
  if (f[0] >= RP_MINVAL) {
    double bc[2], vv[2] = {0, 0};
    for (int r = 1; r < 3; r++) {
      double t = res[r];
      for (int s = 1; s < 3; s++) t -= A[3 * r + s] * old[s];
      t += A[3 * r] * (f[0] - old[0]);
      bc[r - 1] = t;
    }
    // mju_QCQP2 from its second pass on (the lambda = 0 inverse comes with the contact)
    int flg = 0;
    if (qok) {
      const double b1 = bc[0] * mu, b2 = bc[1] * mu, rr = f[0] * f[0];
      double la = 0, P11 = q.P11, P22 = q.P22, P12 = q.P12, v1, v2;
      int okl = 1;
      for (int iter = 0;;) {
        v1 = -P11 * b1 - P12 * b2;
        v2 = -P12 * b1 - P22 * b2;
        const double val = v1 * v1 + v2 * v2 - rr;
        if (val < 1e-10) break;
        const double deriv = -2 * (P11 * v1 * v1 + 2 * P12 * v1 * v2 + P22 * v2 * v2);
        const double delta = -val * rcp(deriv);
        if (delta < 1e-10) break;
        la += delta;
        if (++iter >= 20) break;
        const double det = (q.A11 + la) * (q.A22 + la) - q.A12 * q.A12;
        if (det < 1e-10) { okl = 0; break; }
        const double detinv = rcp(det);
        P11 = (q.A22 + la) * detinv; P22 = (q.A11 + la) * detinv; P12 = -q.A12 * detinv;
      }
      if (okl) { vv[0] = v1 * mu; vv[1] = v2 * mu; flg = la != 0; }
    }
    if (flg) {
      double s = vv[0] * vv[0] * q.imu2 + vv[1] * vv[1] * q.imu2;
      s = sqrt(f[0] * f[0] * rcp(fmax(RP_MINVAL, s)));
      vv[0] *= s; vv[1] *= s;
    }
    f[1] = vv[0]; f[2] = vv[1];
  }
  const double dl[3] = {f[0] - old[0], f[1] - old[1], f[2] - old[2]};
  double Ad[3];
  for (int r = 0; r < 3; r++) Ad[r] = A[3 * r] * dl[0] + A[3 * r + 1] * dl[1] + A[3 * r + 2] * dl[2];
  double change = 0.5 * dot3(dl, Ad) + dot3(dl, res);
  if (change > 1e-10) { f[0] = old[0]; f[1] = old[1]; f[2] = old[2]; change = 0; }
  return change;
}
// primal map of an elliptic contact: jar (3) -> force (3)  (mj_constraintUpdate).  mu is the friction coefficient of
// the regularised cone, friction[0] * sqrt(R_t / R_n) (mj_makeImpedance), fr the contact's own friction coefficient.
RP_HD void elliptic_primal_force(const double* jar, const double* D, double mu, double fr, double* force) {
  double U0 = jar[0] * mu, U1 = jar[1] * fr, U2 = jar[2] * fr;
  double N = U0, T = sqrt(U1 * U1 + U2 * U2);
  if ((N >= mu * T) || (T <= 0 && N >= 0)) {
    force[0] = force[1] = force[2] = 0;
  } else if ((mu * N + T <= 0) || (T <= 0 && N < 0)) {
    for (int r = 0; r < 3; r++) force[r] = -D[r] * jar[r];
  } else {
    double Dm = D[0] / (mu * mu * (1 + mu * mu));
    double NmT = N - mu * T;
    force[0] = -Dm * NmT * mu;
    force[1] = -force[0] / T * U1 * fr;
    force[2] = -force[0] / T * U2 * fr;
  }
}
// geom-pair parameter mixing (mj_contactParam)
RP_HD void mix_params(const double* solref1, const double* solimp1, const double* fr1, double solmix1, int prio1,
                      const double* solref2, const double* solimp2, const double* fr2, double solmix2, int prio2,
                      double* solref, double* solimp, double* mu) {
  double mix;
  if (prio1 != prio2) mix = prio1 > prio2 ? 1.0 : 0.0;
  else if (solmix1 >= RP_MINVAL && solmix2 >= RP_MINVAL) mix = solmix1 / (solmix1 + solmix2);
  else if (solmix1 < RP_MINVAL && solmix2 < RP_MINVAL) mix = 0.5;
  else mix = solmix1 < RP_MINVAL ? 0.0 : 1.0;
  if (solref1[0] > 0 && solref2[0] > 0) for (int k = 0; k < 2; k++) solref[k] = mix * solref1[k] + (1 - mix) * solref2[k];
  else for (int k = 0; k < 2; k++) solref[k] = fmin(solref1[k], solref2[k]);
  for (int k = 0; k < 5; k++) solimp[k] = mix * solimp1[k] + (1 - mix) * solimp2[k];
  *mu = prio1 == prio2 ? fmax(fr1[0], fr2[0]) : (prio1 > prio2 ? fr1[0] : fr2[0]);
}


// dm_control.utils.rewards.tolerance(x, bounds=(lo, hi), margin, sigmoid="gaussian", value_at_margin=0.1)
RP_HD double tolerance_gaussian(double x, double lo, double hi, double margin) {
  if (x >= lo && x <= hi) return 1.0;
  const double dd = (x < lo ? lo - x : x - hi) / margin;
  const double sc = sqrt(-2 * log(0.1));
  return exp(-0.5 * (dd * sc) * (dd * sc));
}

// Rectangular linear sum assignment for nr <= nc: the shortest-augmenting-path algorithm scipy's
// linear_sum_assignment runs (Crouse 2016; rectangular_lsap: rows taken in order, the unvisited columns
// scanned from the last to the first, ties resolved towards a column that is still free), restated so that
// the optimal-transport fingering reward (piano_with_shadow_hands.py:333-369) picks the same pairs.
// cost: row-major [nr][nc]; col4row [nr] out.  wd: nr + 2 nc doubles, wi: 4 nc + nr ints of scratch.
RP_HD void lsa_rect(int nr, int nc, const double* cost, int* col4row, double* wd, int* wi) {
  double *u = wd, *v = wd + nr, *spc = wd + nr + nc;
  int *path = wi, *row4col = wi + nc, *remaining = wi + 2 * nc, *SC = wi + 3 * nc, *SR = wi + 4 * nc;
  for (int i = 0; i < nr; i++) { u[i] = 0; col4row[i] = -1; }
  for (int j = 0; j < nc; j++) { v[j] = 0; row4col[j] = -1; path[j] = -1; }
  for (int cur = 0; cur < nr; cur++) {
    double minVal = 0;
    int num_remaining = nc, i = cur, sink = -1;
    for (int it = 0; it < nc; it++) { remaining[it] = nc - it - 1; SC[it] = 0; spc[it] = HUGE_VAL; }
    for (int r = 0; r < nr; r++) SR[r] = 0;
    while (sink == -1) {
      int index = -1;
      double lowest = HUGE_VAL;
      SR[i] = 1;
      for (int it = 0; it < num_remaining; it++) {
        const int j = remaining[it];
        const double r = minVal + cost[i * nc + j] - u[i] - v[j];
        if (r < spc[j]) { path[j] = i; spc[j] = r; }
        if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
      }
      minVal = lowest;
      if (index < 0) return;  // infeasible (non-finite costs): leaves col4row of the remaining rows at -1
      const int j = remaining[index];
      if (row4col[j] == -1) sink = j; else i = row4col[j];
      SC[j] = 1;
      remaining[index] = remaining[--num_remaining];
    }
    u[cur] += minVal;
    for (int r = 0; r < nr; r++) if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
    for (int j = 0; j < nc; j++) if (SC[j]) v[j] -= minVal - spc[j];
    int j = sink;
    while (true) {
      const int r = path[j];
      row4col[j] = r;
      const int t = col4row[r]; col4row[r] = j; j = t;
      if (r == cur) break;
    }
  }
}

// _compute_ot_fingering_reward (piano_with_shadow_hands.py:333-369): fingertip[10][3] (left hand first, as the
// reference stacks them), keypos[K][3] in ascending key order; distance matrix [10][K], assignment, mean of
// tolerance(d, (0, 0.01), margin 0.1, gaussian) over the min(10, K) assigned pairs; 1 if nothing is to be
// pressed.  scipy transposes a tall matrix (K < 10) before solving: so does this.
// wd: 10 K + (10 + 2 max(10, K)) doubles, wi: 4 max(10, K) + 2 * 10 ints.
RP_HD double ot_fingering_reward(const double* fingertip, const double* keypos, int K, double* wd, int* wi) {
  if (K <= 0) return 1.0;
  const bool tall = K < 10;
  const int nr = tall ? K : 10, nc = tall ? 10 : K;
  double* cost = wd;
  for (int f = 0; f < 10; f++)
    for (int k = 0; k < K; k++) {
      const double dx = keypos[3 * k] - fingertip[3 * f], dy = keypos[3 * k + 1] - fingertip[3 * f + 1], dz = keypos[3 * k + 2] - fingertip[3 * f + 2];
      const double dist = sqrt(dx * dx + dy * dy + dz * dz);
      if (tall) cost[k * 10 + f] = dist; else cost[f * K + k] = dist;
    }
  int* col4row = wi;
  lsa_rect(nr, nc, cost, col4row, wd + 10 * K, wi + 10);
  double sum = 0;
  for (int r = 0; r < nr; r++) sum += tolerance_gaussian(cost[r * nc + col4row[r]], 0.0, 0.01, 0.1);
  return sum / nr;
}

}  // namespace rp
