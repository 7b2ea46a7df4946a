#include "var.h"
#include "input.h"
#include "../../include/kml.h"
#include <cmath>
#include <cstdio>
#include <cstring>

namespace kmlh {

static thread_local bool g_trace = false;
bool trace_active() { return g_trace; }
void trace_set(bool on) { g_trace = on; }
XRef xnode_make(int op, double val, XRef a, XRef b) { return std::make_shared<const XNode>(XNode{op, val, std::move(a), std::move(b)}); }
XRef Var::xnode() const { return node ? node : xnode_make(KML_X_CONST, value, nullptr, nullptr); }
XRef Var::un(int opcode) const { return g_trace ? xnode_make(opcode, 0, xnode(), nullptr) : nullptr; }
XRef Var::bin_node(const Var &r, const char *op) const {
  if (!g_trace) return nullptr;
  int code;
  if (!strcmp(op, "+")) code = KML_X_ADD; else if (!strcmp(op, "-")) code = KML_X_SUB; else if (!strcmp(op, "*")) code = KML_X_MUL;
  else if (!strcmp(op, "/")) code = KML_X_DIV; else if (!strcmp(op, "^")) code = KML_X_POW; else if (!strcmp(op, ">")) code = KML_X_GT;
  else if (!strcmp(op, ">=")) code = KML_X_GE; else if (!strcmp(op, "<")) code = KML_X_LT; else if (!strcmp(op, "<=")) code = KML_X_LE;
  else if (!strcmp(op, "==")) code = KML_X_EQ; else code = KML_X_NE;
  return xnode_make(code, 0, xnode(), r.xnode());
}

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
  return Var("pow(" + std::to_string(base) + "," + p.str() + ")", v, false,
             trace_active() ? xnode_make(KML_X_POW, 0, xnode_make(KML_X_CONST, (double)base, nullptr, nullptr), p.xnode()) : nullptr);
}

Var fn1(const char *name, double (*f)(double), const Var &x) {
  double v = f(x.result());
  if (x.is_constant()) return Var(v);
  XRef n;
  if (trace_active()) {
    const int code = !strcmp(name, "exp") ? KML_X_EXP : !strcmp(name, "sqrt") ? KML_X_SQRT : !strcmp(name, "cos") ? KML_X_COS : !strcmp(name, "sin") ? KML_X_SIN
                   : !strcmp(name, "tan") ? KML_X_TAN : KML_X_LOG;
    n = xnode_make(code, 0, x.xnode(), nullptr);
  }
  return Var(std::string(name) + "(" + x.str() + ")", v, false, n);
}

Var atan2v(const Var &x, const Var &y) {
  double v = std::atan2(x.result(), y.result());
  if (x.is_constant() && y.is_constant()) return Var(v);
  return Var("atan2(" + x.str() + ", " + y.str() + ")", v, false, trace_active() ? xnode_make(KML_X_ATAN2, 0, x.xnode(), y.xnode()) : nullptr);
}

} // namespace kmlh
