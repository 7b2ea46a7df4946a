// var.h - lazy script variables of the Karamelo input language.
//
// Host-side mirror of the reference's Var (reference src/var.h:18-58, src/var.cpp):
// a variable is {equation, value, constant}.  Constant operands fold to a value whose
// textual form is "%.15f"; an expression touching a non-constant operand stays symbolic
// and is RE-PARSED from its string on every result() (src/var.cpp:44-60), which sends its
// constants through the float-literal path of the parser again.  That behaviour is
// parity-critical and is reproduced here.
#pragma once
#include <string>

namespace kmlh {

class Input;

std::string fmt15(double v); // "%.15f", reference src/var.cpp:24-29

class Var {
public:
  Var() : value(0), constant(true) {}
  Var(double v) : equation(fmt15(v)), value(v), constant(true) {}
  Var(const std::string &eq, double v, bool c = false) : equation(eq), value(v), constant(c) {}

  double result(Input *in);            // re-evaluates when not constant
  double result() const { return value; }
  std::string str() const { return equation.empty() ? fmt15(value) : equation; }
  const std::string &eq() const { return equation; }
  bool is_constant() const { return constant; }
  void make_constant(Input *in);

  Var operator+(const Var &r) const { return bin(r, "+", value + r.value, constant && r.constant); }
  Var operator-(const Var &r) const { return bin(r, "-", value - r.value, constant && r.constant); }
  Var operator*(const Var &r) const { return bin(r, "*", value * r.value, false); }
  Var operator/(const Var &r) const { return bin(r, "/", value / r.value, false); }
  Var pow(const Var &r) const;
  Var operator>(const Var &r) const { return bin(r, ">", value > r.value, false); }
  Var operator>=(const Var &r) const { return bin(r, ">=", value >= r.value, false); }
  Var operator<(const Var &r) const { return bin(r, "<", value < r.value, false); }
  Var operator<=(const Var &r) const { return bin(r, "<=", value <= r.value, false); }
  Var operator==(const Var &r) const { return bin(r, "==", value == r.value, false); }
  Var operator!=(const Var &r) const { return bin(r, "!=", value != r.value, false); }
  Var operator-() const { return constant ? Var(-value) : Var("(-" + str() + ")", -value, false); }
  Var operator!() const { return constant ? Var((double)!value) : Var("(!" + str() + ")", (double)!value, false); }
  operator double() const { return value; }

private:
  // Both operands constant -> folded constant; otherwise "(a<op>b)" with the flag the reference
  // gives that operator (src/var.cpp:82-214: + and - propagate constness, the others never do).
  Var bin(const Var &r, const char *op, double v, bool c) const {
    if (constant && r.constant) return Var(v);
    return Var("(" + str() + op + r.str() + ")", v, c);
  }
  std::string equation;
  double value;
  bool constant;
};

// src/var.cpp:228-309
Var powv(int base, const Var &p);
Var fn1(const char *name, double (*f)(double), const Var &x);
Var atan2v(const Var &x, const Var &y);

} // namespace kmlh
