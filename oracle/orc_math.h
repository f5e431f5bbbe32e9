/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the B200 batched physics path.
 *
 * Small vector / quaternion / spatial-algebra helpers restating the public algorithms of
 * MuJoCo's engine_util_blas.c, engine_util_spatial.c and engine_util_misc.c
 * (third-party dependency `mujoco>=3.1`, reference pyproject.toml:10-14; its C source is
 * not vendored under /root/reference).  Used by mjstep_oracle.c, which restates what the
 * reference reaches through mj.mj_step (reference mujoco_template/model.py:56-57).
 *
 * PARITY UNPINNED against a live MuJoCo: see DESIGN.md "Oracle".  Pinned instead by the
 * analytic and invariant tests in tests/test_oracle_*.py.
 */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <math.h>
#include <string.h>

#define ORC_MINVAL 1e-15
#define ORC_PI 3.14159265358979323846
#define ORC_MINIMP 0.0001
#define ORC_MAXIMP 0.9999

static inline double orc_max(double a, double b) { return a > b ? a : b; }
static inline double orc_min(double a, double b) { return a < b ? a : b; }
static inline double orc_clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

static inline void v3_copy(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void v3_zero(double* r) { r[0] = r[1] = r[2] = 0; }
static inline void v3_add(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void v3_sub(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v3_addto(double* r, const double* a) { r[0] += a[0]; r[1] += a[1]; r[2] += a[2]; }
static inline void v3_scl(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void v3_addscl(double* r, const double* a, const double* b, double s) { r[0] = a[0] + b[0] * s; r[1] = a[1] + b[1] * s; r[2] = a[2] + b[2] * s; }
static inline void v3_addtoscl(double* r, const double* a, double s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
static inline double v3_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v3_cross(double* r, const double* a, const double* b) {
  double t0 = a[1] * b[2] - a[2] * b[1], t1 = a[2] * b[0] - a[0] * b[2], t2 = a[0] * b[1] - a[1] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
static inline double v3_norm(const double* a) { return sqrt(v3_dot(a, a)); }
/* mju_normalize3: returns the norm; tiny vectors become (1,0,0) */
static inline double v3_normalize(double* a) {
  double n = v3_norm(a);
  if (n < ORC_MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { double inv = 1.0 / n; a[0] *= inv; a[1] *= inv; a[2] *= inv; }
  return n;
}
/* mju_normalize4 */
static inline double q_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < ORC_MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (fabs(n - 1) > ORC_MINVAL) { double inv = 1.0 / n; q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv; }
  return n;
}
static inline void q_mul(double* r, const double* a, const double* b) {
  double t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
static inline void q_neg(double* r, const double* a) { r[0] = a[0]; r[1] = -a[1]; r[2] = -a[2]; r[3] = -a[3]; }
/* mju_axisAngle2Quat */
static inline void q_axisangle(double* r, const double* axis, double angle) {
  if (angle == 0) { r[0] = 1; r[1] = r[2] = r[3] = 0; return; }
  double s = sin(angle * 0.5);
  r[0] = cos(angle * 0.5); r[1] = axis[0] * s; r[2] = axis[1] * s; r[3] = axis[2] * s;
}
/* mju_quat2Mat (row-major 3x3) */
static inline void q_tomat(double* m, const double* q) {
  if (q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0) {
    m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 0; m[4] = 1; m[5] = 0; m[6] = 0; m[7] = 0; m[8] = 1; return;
  }
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  double q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03); m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
/* mju_rotVecQuat */
static inline void q_rotvec(double* r, const double* v, const double* q) {
  if (v[0] == 0 && v[1] == 0 && v[2] == 0) { v3_zero(r); return; }
  if (q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0) { v3_copy(r, v); return; }
  double t0 = q[0] * v[0] + q[2] * v[2] - q[3] * v[1];
  double t1 = q[0] * v[1] + q[3] * v[0] - q[1] * v[2];
  double t2 = q[0] * v[2] + q[1] * v[1] - q[2] * v[0];
  double r0 = v[0] + 2 * (q[2] * t2 - q[3] * t1);
  double r1 = v[1] + 2 * (q[3] * t0 - q[1] * t2);
  double r2 = v[2] + 2 * (q[1] * t1 - q[2] * t0);
  r[0] = r0; r[1] = r1; r[2] = r2;
}
static inline void m3_mulvec(double* r, const double* m, const double* v) {
  double t0 = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double t1 = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double t2 = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
static inline void m3_multvec(double* r, const double* m, const double* v) {
  double t0 = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  double t1 = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  double t2 = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = t0; r[1] = t1; r[2] = t2;
}
/* mju_quatIntegrate */
static inline void q_integrate(double* q, const double* vel, double scale) {
  double tmp[3], qrot[4];
  v3_copy(tmp, vel);
  double angle = scale * v3_normalize(tmp);
  q_axisangle(qrot, tmp, angle);
  q_normalize(q);
  q_mul(q, q, qrot);
}
/* mju_quat2Vel */
static inline void q_tovel(double* r, const double* q, double dt) {
  double axis[3] = {q[1], q[2], q[3]};
  double sin_a_2 = v3_normalize(axis);
  double speed = 2 * atan2(sin_a_2, q[0]);
  if (speed > ORC_PI) speed -= 2 * ORC_PI;
  speed /= dt;
  v3_scl(r, axis, speed);
}
/* 6-D spatial vectors are [angular; linear] */
static inline void sp_cross_motion(double* r, const double* vel, const double* v) {
  r[0] = -vel[2] * v[1] + vel[1] * v[2];
  r[1] = vel[2] * v[0] - vel[0] * v[2];
  r[2] = -vel[1] * v[0] + vel[0] * v[1];
  r[3] = -vel[2] * v[4] + vel[1] * v[5];
  r[4] = vel[2] * v[3] - vel[0] * v[5];
  r[5] = -vel[1] * v[3] + vel[0] * v[4];
  r[3] += -vel[5] * v[1] + vel[4] * v[2];
  r[4] += vel[5] * v[0] - vel[3] * v[2];
  r[5] += -vel[4] * v[0] + vel[3] * v[1];
}
static inline void sp_cross_force(double* r, const double* vel, const double* f) {
  r[0] = -vel[2] * f[1] + vel[1] * f[2];
  r[1] = vel[2] * f[0] - vel[0] * f[2];
  r[2] = -vel[1] * f[0] + vel[0] * f[1];
  r[3] = -vel[2] * f[4] + vel[1] * f[5];
  r[4] = vel[2] * f[3] - vel[0] * f[5];
  r[5] = -vel[1] * f[3] + vel[0] * f[4];
  r[0] += -vel[5] * f[4] + vel[4] * f[5];
  r[1] += vel[5] * f[3] - vel[3] * f[5];
  r[2] += -vel[4] * f[3] + vel[3] * f[4];
}
/* mju_inertCom: 10-number inertia about an offset point, world axes */
static inline void sp_inert_com(double* res, const double* inert, const double* mat, const double* dif, double mass) {
  double tmp[9];
  tmp[0] = mat[0] * inert[0]; tmp[1] = mat[3] * inert[0]; tmp[2] = mat[6] * inert[0];
  tmp[3] = mat[1] * inert[1]; tmp[4] = mat[4] * inert[1]; tmp[5] = mat[7] * inert[1];
  tmp[6] = mat[2] * inert[2]; tmp[7] = mat[5] * inert[2]; tmp[8] = mat[8] * inert[2];
  res[0] = mat[0] * tmp[0] + mat[1] * tmp[3] + mat[2] * tmp[6];
  res[1] = mat[3] * tmp[1] + mat[4] * tmp[4] + mat[5] * tmp[7];
  res[2] = mat[6] * tmp[2] + mat[7] * tmp[5] + mat[8] * tmp[8];
  res[3] = mat[0] * tmp[1] + mat[1] * tmp[4] + mat[2] * tmp[7];
  res[4] = mat[0] * tmp[2] + mat[1] * tmp[5] + mat[2] * tmp[8];
  res[5] = mat[3] * tmp[2] + mat[4] * tmp[5] + mat[5] * tmp[8];
  res[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] -= mass * dif[0] * dif[1];
  res[4] -= mass * dif[0] * dif[2];
  res[5] -= mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
  res[9] = mass;
}
/* mju_mulInertVec */
static inline void sp_mul_inert(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
/* mju_dofCom */
static inline void sp_dof_com(double* r, const double* axis, const double* offset) {
  if (offset) { v3_copy(r, axis); v3_cross(r + 3, axis, offset); }
  else { v3_zero(r); v3_copy(r + 3, axis); }
}
/* mju_transformSpatial (motion or force vector to a new origin, optional rotation new->old) */
static inline void sp_transform(double* res, const double* vec, int flg_force, const double* newpos,
                                const double* oldpos, const double* rotnew2old) {
  double cros[3], dif[3], tran[6];
  memcpy(tran, vec, sizeof(tran));
  v3_sub(dif, newpos, oldpos);
  if (flg_force) { v3_cross(cros, dif, vec + 3); v3_sub(tran, vec, cros); }
  else { v3_cross(cros, dif, vec); v3_sub(tran + 3, vec + 3, cros); }
  if (rotnew2old) { m3_multvec(res, rotnew2old, tran); m3_multvec(res + 3, rotnew2old, tran + 3); }
  else memcpy(res, tran, sizeof(tran));
}
/* mju_makeFrame: complete a contact frame whose first row is the normal */
static inline void make_frame(double* frame) {
  v3_normalize(frame);
  if (v3_norm(frame + 3) < 0.5) {
    v3_zero(frame + 3);
    if (frame[1] < 0.5 && frame[1] > -0.5) frame[4] = 1; else frame[5] = 1;
  }
  double t = v3_dot(frame, frame + 3);
  v3_addtoscl(frame + 3, frame, -t);
  v3_normalize(frame + 3);
  v3_cross(frame + 6, frame, frame + 3);
}
#endif
