/* TEST INFRASTRUCTURE: operation-counting arithmetic type for the oracle.
 *
 * `make liborc_count.so` compiles mjstep_oracle.c as C++ with every `double` replaced by orc_real, a one-member class
 * whose operators do the IEEE arithmetic and bump per-thread counters.  Weights follow SURVEY.md section 8(d): add / sub /
 * mul / div / sqrt = 1 flop each, a sin / cos / atan2 / acos / exp / pow / tanh call = 20 (sincos of one angle counts as
 * two calls = 40); comparisons, negation, fabs, min / max and copies are free.  orc_count_reset() / orc_count_read() expose
 * the counters to oracle.py (op_count()).  The counted library computes bit-identical results (the arithmetic is the
 * same; it is only much slower), which the test suite checks. */
#ifndef ORC_COUNT_H
#define ORC_COUNT_H
#include <math.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef double orc_raw;
struct orc_counters { unsigned long long add, mul, div, sqrt_, trans; };
static thread_local orc_counters g_orc_ops;

struct orc_real {
  orc_raw v;
  orc_real() = default;
  orc_real(orc_raw x) : v(x) {}
  orc_real(int x) : v(x) {}
  orc_real(long x) : v((orc_raw)x) {}
  orc_real(unsigned x) : v(x) {}
  orc_real(unsigned long x) : v((orc_raw)x) {}
  explicit operator int() const { return (int)v; }
  explicit operator long() const { return (long)v; }
  explicit operator orc_raw() const { return v; }
  explicit operator bool() const { return v != 0; }
  orc_real& operator+=(orc_real o) { g_orc_ops.add++; v += o.v; return *this; }
  orc_real& operator-=(orc_real o) { g_orc_ops.add++; v -= o.v; return *this; }
  orc_real& operator*=(orc_real o) { g_orc_ops.mul++; v *= o.v; return *this; }
  orc_real& operator/=(orc_real o) { g_orc_ops.div++; v /= o.v; return *this; }
  orc_real operator-() const { return orc_real(-v); }
  orc_real operator+() const { return *this; }
};
#define ORC_BIN(op, ctr)                                                                              \
  static inline orc_real operator op(orc_real a, orc_real b) { g_orc_ops.ctr++; return orc_real(a.v op b.v); } \
  static inline orc_real operator op(orc_real a, orc_raw b) { g_orc_ops.ctr++; return orc_real(a.v op b); }   \
  static inline orc_real operator op(orc_raw a, orc_real b) { g_orc_ops.ctr++; return orc_real(a op b.v); }   \
  static inline orc_real operator op(orc_real a, int b) { g_orc_ops.ctr++; return orc_real(a.v op b); }       \
  static inline orc_real operator op(int a, orc_real b) { g_orc_ops.ctr++; return orc_real(a op b.v); }
ORC_BIN(+, add) ORC_BIN(-, add) ORC_BIN(*, mul) ORC_BIN(/, div)
#undef ORC_BIN
#define ORC_CMP(op)                                                                   \
  static inline bool operator op(orc_real a, orc_real b) { return a.v op b.v; }       \
  static inline bool operator op(orc_real a, orc_raw b) { return a.v op b; }          \
  static inline bool operator op(orc_raw a, orc_real b) { return a op b.v; }          \
  static inline bool operator op(orc_real a, int b) { return a.v op b; }              \
  static inline bool operator op(int a, orc_real b) { return a op b.v; }
ORC_CMP(<) ORC_CMP(>) ORC_CMP(<=) ORC_CMP(>=) ORC_CMP(==) ORC_CMP(!=)
#undef ORC_CMP
static inline bool operator!(orc_real a) { return a.v == 0; }
static inline orc_real sqrt(orc_real a) { g_orc_ops.sqrt_++; return orc_real(::sqrt(a.v)); }
static inline orc_real fabs(orc_real a) { return orc_real(::fabs(a.v)); }
static inline orc_real floor(orc_real a) { return orc_real(::floor(a.v)); }
static inline orc_real fmin(orc_real a, orc_real b) { return orc_real(::fmin(a.v, b.v)); }
static inline orc_real fmax(orc_real a, orc_real b) { return orc_real(::fmax(a.v, b.v)); }
static inline int isnan(orc_real a) { return std::isnan(a.v); }
static inline int isfinite(orc_real a) { return std::isfinite(a.v); }
#define ORC_TRANS1(fn) static inline orc_real fn(orc_real a) { g_orc_ops.trans++; return orc_real(::fn(a.v)); }
ORC_TRANS1(sin) ORC_TRANS1(cos) ORC_TRANS1(tan) ORC_TRANS1(acos) ORC_TRANS1(asin) ORC_TRANS1(atan) ORC_TRANS1(exp) ORC_TRANS1(log) ORC_TRANS1(tanh)
#undef ORC_TRANS1
static inline orc_real atan2(orc_real a, orc_real b) { g_orc_ops.trans++; return orc_real(::atan2(a.v, b.v)); }
static inline orc_real pow(orc_real a, orc_real b) { g_orc_ops.trans++; return orc_real(::pow(a.v, b.v)); }

extern "C" {
void orc_count_reset(void);
/* out[0..4] = add/sub, mul, div, sqrt, transcendental calls of the calling thread since the last reset */
void orc_count_read(unsigned long long* out);
}
#define double orc_real
#endif
