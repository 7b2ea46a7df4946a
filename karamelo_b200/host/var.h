// var.h - lazy script variables of the Karamelo input language.
//
// Host-side mirror of the reference's Var (reference src/var.h:18-58, src/var.cpp):
// a variable is {equation, value, constant}.  Constant operands fold to a value whose
// textual form is "%.15f"; an expression touching a non-constant operand stays symbolic
// and is RE-PARSED from its string on every result() (src/var.cpp:44-60), which sends its
// constants through the float-literal path of the parser again.  That behaviour is
// parity-critical and is reproduced here.
#pragma once
#include <memory>
#include <string>
#include <vector>

namespace kmlh {

class Input;

// Expression tracing (Input::compile): while a trace is active every Var produced by the parser also carries the operation tree that
// produced its value, with the particle variables x, y, z, x0, y0, z0 as leaves.  The tree is flattened into a postfix program
// (kml_expr, include/kml.h) that the engine evaluates per particle - the reference re-parses the expression string for every
// particle (e.g. src/fix_initial_velocity_particles.cpp:99-161).
struct XNode { int op; double val; std::shared_ptr<const XNode> a, b; };
typedef std::shared_ptr<const XNode> XRef;
bool trace_active();
void trace_set(bool on);

std::string fmt15(double v); // "%.15f", reference src/var.cpp:24-29

class Var {
public:
  Var() : value(0), constant(true) {}
  Var(double v) : equation(fmt15(v)), value(v), constant(true) {}
  Var(const std::string &eq, double v, bool c = false) : equation(eq), value(v), constant(c) {}
  Var(const std::string &eq, double v, bool c, XRef n) : equation(eq), value(v), constant(c), node(std::move(n)) {}
  XRef xnode() const; // the traced tree (a constant leaf when the value did not come from a particle variable)

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
  Var operator-() const { return constant ? Var(-value) : Var("(-" + str() + ")", -value, false, un(7)); }
  Var operator!() const { return constant ? Var((double)!value) : Var("(!" + str() + ")", (double)!value, false, un(8)); }
  operator double() const { return value; }

private:
  // Both operands constant -> folded constant; otherwise "(a<op>b)" with the flag the reference
  // gives that operator (src/var.cpp:82-214: + and - propagate constness, the others never do).
  Var bin(const Var &r, const char *op, double v, bool c) const {
    if (constant && r.constant) return Var(v);
    return Var("(" + str() + op + r.str() + ")", v, c, bin_node(r, op));
  }
  XRef bin_node(const Var &r, const char *op) const;
  XRef un(int opcode) const;
  std::string equation;
  double value;
  bool constant;
  XRef node;
};
XRef xnode_make(int op, double val, XRef a, XRef b);

// src/var.cpp:228-309
Var powv(int base, const Var &p);
Var fn1(const char *name, double (*f)(double), const Var &x);
Var atan2v(const Var &x, const Var &y);

} // namespace kmlh
