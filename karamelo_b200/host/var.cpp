#include "var.h"
#include "input.h"
#include <cmath>
#include <cstdio>

namespace kmlh {

std::string fmt15(double v) {
  char buf[400];
  int n = std::snprintf(buf, sizeof buf, "%.15f", v);
  return std::string(buf, n > 0 ? n : 0);
}

double Var::result(Input *in) {
  if (!constant && in) value = in->parsev(equation).value;
  return value;
}

void Var::make_constant(Input *in) {
  if (constant) return;
  if (in) value = in->parsev(equation).value;
  equation = fmt15(value);
  constant = true;
}

Var Var::pow(const Var &r) const { return bin(r, "^", std::pow(value, r.value), false); }

Var powv(int base, const Var &p) {
  double v = std::pow(base, p.result());
  if (p.is_constant()) return Var(v);
  return Var("pow(" + std::to_string(base) + "," + p.str() + ")", v, false);
}

Var fn1(const char *name, double (*f)(double), const Var &x) {
  double v = f(x.result());
  if (x.is_constant()) return Var(v);
  return Var(std::string(name) + "(" + x.str() + ")", v, false);
}

Var atan2v(const Var &x, const Var &y) {
  double v = std::atan2(x.result(), y.result());
  if (x.is_constant() && y.is_constant()) return Var(v);
  return Var("atan2(" + x.str() + ", " + y.str() + ")", v, false);
}

} // namespace kmlh
