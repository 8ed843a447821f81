// hamilton_b200_trace.hpp — header-only C++ tracing number type.
//
// The reference's mkSystem takes `forall a. RealFloat a => Vector n a -> Vector m a`
// (src/Numeric/Hamilton.hs:212-215): a function polymorphic in its number type, which `ad`
// instantiates at its own dual/tower types.  The same trick crosses the C ABI here: instantiate
// the function at hb::Ex (C++ generic lambda / template) and every arithmetic operation is
// appended to a Wengert list (hb_op, include/hamilton_b200.h) that hb_system_from_tape compiles.
#pragma once
#include <cmath>
#include <vector>

#include "hamilton_b200.h"

namespace hb {

class Tape {
 public:
  std::vector<hb_op> ops;
  int push(int op, int a = 0, int b = 0, double c = 0.0) {
    hb_op o; o.op = op; o.a = a; o.b = b; o._pad = 0; o.c = c;
    ops.push_back(o);
    return (int)ops.size() - 1;
  }
};

struct Ex {
  Tape* t = nullptr;
  int id = -1;
  Ex() {}
  Ex(Tape* tp, int i) : t(tp), id(i) {}
  static Ex input(Tape& tp, int idx) { return Ex(&tp, tp.push(HB_OP_INPUT, idx)); }
  static Ex param(Tape& tp, int idx) { return Ex(&tp, tp.push(HB_OP_PARAM, idx)); }
  static Ex constant(Tape& tp, double c) { return Ex(&tp, tp.push(HB_OP_CONST, 0, 0, c)); }
};

inline Ex lift(const Ex& like, double c) { return Ex::constant(*like.t, c); }
inline Ex bin(int op, const Ex& a, const Ex& b) { return Ex(a.t, a.t->push(op, a.id, b.id)); }
inline Ex un(int op, const Ex& a) { return Ex(a.t, a.t->push(op, a.id)); }

inline Ex operator+(const Ex& a, const Ex& b) { return bin(HB_OP_ADD, a, b); }
inline Ex operator-(const Ex& a, const Ex& b) { return bin(HB_OP_SUB, a, b); }
inline Ex operator*(const Ex& a, const Ex& b) { return bin(HB_OP_MUL, a, b); }
inline Ex operator/(const Ex& a, const Ex& b) { return bin(HB_OP_DIV, a, b); }
inline Ex operator+(const Ex& a, double b) { return a + lift(a, b); }
inline Ex operator-(const Ex& a, double b) { return a - lift(a, b); }
inline Ex operator*(const Ex& a, double b) { return a * lift(a, b); }
inline Ex operator/(const Ex& a, double b) { return a / lift(a, b); }
inline Ex operator+(double a, const Ex& b) { return lift(b, a) + b; }
inline Ex operator-(double a, const Ex& b) { return lift(b, a) - b; }
inline Ex operator*(double a, const Ex& b) { return lift(b, a) * b; }
inline Ex operator/(double a, const Ex& b) { return lift(b, a) / b; }
inline Ex operator-(const Ex& a) { return un(HB_OP_NEG, a); }
inline Ex& operator+=(Ex& a, const Ex& b) { a = a + b; return a; }
inline Ex& operator-=(Ex& a, const Ex& b) { a = a - b; return a; }
inline Ex& operator*=(Ex& a, const Ex& b) { a = a * b; return a; }

inline Ex sin(const Ex& a) { return un(HB_OP_SIN, a); }
inline Ex cos(const Ex& a) { return un(HB_OP_COS, a); }
inline Ex tan(const Ex& a) { return un(HB_OP_TAN, a); }
inline Ex asin(const Ex& a) { return un(HB_OP_ASIN, a); }
inline Ex acos(const Ex& a) { return un(HB_OP_ACOS, a); }
inline Ex atan(const Ex& a) { return un(HB_OP_ATAN, a); }
inline Ex sinh(const Ex& a) { return un(HB_OP_SINH, a); }
inline Ex cosh(const Ex& a) { return un(HB_OP_COSH, a); }
inline Ex tanh(const Ex& a) { return un(HB_OP_TANH, a); }
inline Ex asinh(const Ex& a) { return un(HB_OP_ASINH, a); }
inline Ex acosh(const Ex& a) { return un(HB_OP_ACOSH, a); }
inline Ex atanh(const Ex& a) { return un(HB_OP_ATANH, a); }
inline Ex exp(const Ex& a) { return un(HB_OP_EXP, a); }
inline Ex log(const Ex& a) { return un(HB_OP_LOG, a); }
inline Ex sqrt(const Ex& a) { return un(HB_OP_SQRT, a); }
inline Ex abs(const Ex& a) { return un(HB_OP_ABS, a); }
inline Ex fabs(const Ex& a) { return un(HB_OP_ABS, a); }
inline Ex recip(const Ex& a) { return un(HB_OP_RECIP, a); }
inline Ex pow(const Ex& a, const Ex& b) { return bin(HB_OP_POW, a, b); }       // (**)
inline Ex pow(const Ex& a, double b) { return bin(HB_OP_POW, a, lift(a, b)); }
inline Ex powi(const Ex& a, int k) { return Ex(a.t, a.t->push(HB_OP_POWI, a.id, 0, (double)k)); }   // (^), (^^)
inline Ex atan2(const Ex& a, const Ex& b) { return bin(HB_OP_ATAN2, a, b); }

// The same names for plain doubles, so one generic function body serves both number types.
inline double powi(double a, int k) { return std::pow(a, k); }
inline double recip(double a) { return 1.0 / a; }

}  // namespace hb
