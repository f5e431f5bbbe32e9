// Small fixed-size vector / quaternion / spatial helpers for the device engine.
// Templated on the arithmetic type (double for the parity path, float for the FP32 mode).
#pragma once
#include <cuda_runtime.h>

namespace b2 {

#define B2_DEV __device__ __forceinline__

template <typename T> struct Num;
template <> struct Num<double> {
  static B2_DEV double minval() { return 1e-15; }
  static B2_DEV double pi() { return 3.14159265358979323846; }
  static B2_DEV void sincos_(double x, double* s, double* c) { ::sincos(x, s, c); }
};
template <> struct Num<float> {
  static B2_DEV float minval() { return 1e-15f; }
  static B2_DEV float pi() { return 3.14159265358979323846f; }
  static B2_DEV void sincos_(float x, float* s, float* c) { ::sincosf(x, s, c); }
};

template <typename T> B2_DEV T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> B2_DEV T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> B2_DEV T tclip(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

template <typename T> B2_DEV T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> B2_DEV void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> B2_DEV T normalize3(T* a) {
  T n = sqrt(dot3(a, a));
  if (n < Num<T>::minval()) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { T inv = T(1) / n; a[0] *= inv; a[1] *= inv; a[2] *= inv; }
  return n;
}
template <typename T> B2_DEV void normalize4(T* q) {
  T n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < Num<T>::minval()) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (fabs(n - T(1)) > Num<T>::minval()) { T inv = T(1) / n; q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv; }
}
template <typename T> B2_DEV void quat_mul(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <typename T> B2_DEV void quat_axis_angle(T* r, const T* axis, T angle) {
  if (angle == T(0)) { r[0] = 1; r[1] = r[2] = r[3] = 0; return; }
  T s, c;
  Num<T>::sincos_(angle * T(0.5), &s, &c);
  r[0] = c; r[1] = axis[0] * s; r[2] = axis[1] * s; r[3] = axis[2] * s;
}
template <typename T> B2_DEV void quat_to_mat(T* m, const T* q) {
  T q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  T q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3], q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02); m[3] = 2 * (q12 + q03);
  m[5] = 2 * (q23 - q01); m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
template <typename T> B2_DEV void quat_rot(T* r, const T* v, const T* q) {
  T t0 = q[0] * v[0] + q[2] * v[2] - q[3] * v[1];
  T t1 = q[0] * v[1] + q[3] * v[0] - q[1] * v[2];
  T t2 = q[0] * v[2] + q[1] * v[1] - q[2] * v[0];
  T x = v[0] + 2 * (q[2] * t2 - q[3] * t1), y = v[1] + 2 * (q[3] * t0 - q[1] * t2), z = v[2] + 2 * (q[1] * t1 - q[2] * t0);
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> B2_DEV void mat_vec(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> B2_DEV void matT_vec(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2], z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
// quaternion exponential-map integration: q <- normalize(q) * exp(scale * w / 2)
template <typename T> B2_DEV void quat_integrate(T* q, const T* w, T scale) {
  T ax[3] = {w[0], w[1], w[2]}, rot[4];
  T angle = scale * normalize3(ax);
  quat_axis_angle(rot, ax, angle);
  normalize4(q);
  quat_mul(q, q, rot);
}
// rotation vector of unit quaternion q divided by dt
template <typename T> B2_DEV void quat_to_vel(T* r, const T* q, T dt) {
  T ax[3] = {q[1], q[2], q[3]};
  T s = normalize3(ax);
  T speed = 2 * atan2(s, q[0]);
  if (speed > Num<T>::pi()) speed -= 2 * Num<T>::pi();
  speed /= dt;
  r[0] = ax[0] * speed; r[1] = ax[1] * speed; r[2] = ax[2] * speed;
}
// spatial vectors are [angular(3); linear(3)]
template <typename T> B2_DEV void cross_motion(T* r, const T* v, const T* s) {
  r[0] = -v[2] * s[1] + v[1] * s[2];
  r[1] = v[2] * s[0] - v[0] * s[2];
  r[2] = -v[1] * s[0] + v[0] * s[1];
  r[3] = -v[2] * s[4] + v[1] * s[5];
  r[4] = v[2] * s[3] - v[0] * s[5];
  r[5] = -v[1] * s[3] + v[0] * s[4];
  r[3] += -v[5] * s[1] + v[4] * s[2];
  r[4] += v[5] * s[0] - v[3] * s[2];
  r[5] += -v[4] * s[0] + v[3] * s[1];
}
template <typename T> B2_DEV void cross_force(T* r, const T* v, const T* f) {
  r[0] = -v[2] * f[1] + v[1] * f[2];
  r[1] = v[2] * f[0] - v[0] * f[2];
  r[2] = -v[1] * f[0] + v[0] * f[1];
  r[3] = -v[2] * f[4] + v[1] * f[5];
  r[4] = v[2] * f[3] - v[0] * f[5];
  r[5] = -v[1] * f[3] + v[0] * f[4];
  r[0] += -v[5] * f[4] + v[4] * f[5];
  r[1] += v[5] * f[3] - v[3] * f[5];
  r[2] += -v[4] * f[3] + v[3] * f[4];
}
// 10-number spatial inertia (Ixx Iyy Izz Ixy Ixz Iyz, m*c, m) times a motion vector
template <typename T> B2_DEV void inert_mul(T* r, const T* i, const T* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
// rotate a diagonal inertia into world axes and shift it to an offset point
template <typename T> B2_DEV void inert_about(T* res, const T* diag, const T* R, const T* d, T mass) {
  T a0 = R[0] * diag[0], a1 = R[3] * diag[0], a2 = R[6] * diag[0];
  T b0 = R[1] * diag[1], b1 = R[4] * diag[1], b2 = R[7] * diag[1];
  T c0 = R[2] * diag[2], c1 = R[5] * diag[2], c2 = R[8] * diag[2];
  res[0] = R[0] * a0 + R[1] * b0 + R[2] * c0 + mass * (d[1] * d[1] + d[2] * d[2]);
  res[1] = R[3] * a1 + R[4] * b1 + R[5] * c1 + mass * (d[0] * d[0] + d[2] * d[2]);
  res[2] = R[6] * a2 + R[7] * b2 + R[8] * c2 + mass * (d[0] * d[0] + d[1] * d[1]);
  res[3] = R[0] * a1 + R[1] * b1 + R[2] * c1 - mass * d[0] * d[1];
  res[4] = R[0] * a2 + R[1] * b2 + R[2] * c2 - mass * d[0] * d[2];
  res[5] = R[3] * a2 + R[4] * b2 + R[5] * c2 - mass * d[1] * d[2];
  res[6] = mass * d[0]; res[7] = mass * d[1]; res[8] = mass * d[2]; res[9] = mass;
}
// complete a contact frame whose first row is the (unnormalised) normal
template <typename T> B2_DEV void make_frame(T* f) {
  normalize3(f);
  if (sqrt(dot3(f + 3, f + 3)) < T(0.5)) {
    f[3] = 0; f[4] = 0; f[5] = 0;
    if (f[1] < T(0.5) && f[1] > T(-0.5)) f[4] = 1; else f[5] = 1;
  }
  T t = dot3(f, f + 3);
  f[3] -= t * f[0]; f[4] -= t * f[1]; f[5] -= t * f[2];
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

}  // namespace b2
