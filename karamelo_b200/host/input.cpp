// input.cpp - script interpreter; behaviour follows reference src/input.cpp (see input.h).
#include "input.h"
#include "sim.h"
#include <algorithm>
#include <cctype>
#include <cmath>
#include <fstream>
#include <iostream>
#include <stack>
#include <stdexcept>

namespace kmlh {

void fatal(const std::string &msg) { throw std::runtime_error(msg); }

Input::Input(Sim *s) : sim(s) {
  // built-in variables, reference src/input.cpp:52-62
  vars["time"] = Var("time", 0);
  vars["timestep"] = Var("timestep", 0);
  vars["dt"] = Var("dt", 0);
  for (const char *n : {"x", "y", "z", "x0", "y0", "z0"}) vars[n] = Var(n, 0);
  vars["PI"] = Var(M_PI);
}

bool Input::protected_variable(const std::string &n) const {
  return n == "x" || n == "y" || n == "z" || n == "time" || n == "dt"; // src/input.cpp:76-86
}

int Input::precedence(const std::string &op) {
  switch (op[0]) {
  case '>': case '<': case '=': case '!': return 1;
  case '+': case '-': return 2;
  case '/': return 3;
  case '^': return 4;
  case 'e': case 'E': return 5;
  }
  if (op == "*") return 3;
  if (op == "**") return 4;
  return 0;
}

Var Input::apply_op(const Var &a, const std::string &op, const Var &b) {
  if (op.size() == 1) {
    switch (op[0]) {
    case '+': return a + b;
    case '-': return a - b;
    case '*': return a * b;
    case '/': return a / b;
    case '^': return a.pow(b);
    case 'e': return a * powv(10, b); // src/input.cpp:191
    case '>': return a > b;
    case '<': return a < b;
    case '(': fatal("Error: unmatched parenthesis (\n");
    }
  } else {
    if (op == ">=") return a >= b;
    if (op == "<=") return a <= b;
    if (op == "==") return a == b;
    if (op == "!=") return a != b;
    if (op == "**") return a.pow(b);
  }
  fatal("Error: unknown operator " + op + "\n");
}

static bool is_operator(char c) { return c == '+' || c == '-' || c == '*' || c == '/' || c == '^' || c == '>' || c == '<' || c == '!'; }
static bool is_math_char(char c) { return is_operator(c) || c == '(' || c == ')' || c == '='; }

static std::string remove_whitespace(const std::string &s) { // src/input.cpp:352-365
  std::string o; bool quote = false;
  for (char c : s) {
    if (c == '"') quote = !quote;
    else if (c != ' ' || quote) o.push_back(c);
  }
  return o;
}

Var Input::parsev(std::string str) {
  std::stack<Var> values;
  std::stack<std::string> ops;
  std::string returnvar;
  str = remove_whitespace(str);
  const int n = (int)str.length();
  auto at = [&](int k) -> char { return k < n ? str[k] : '\0'; };
  bool negative = false;

  // collapse the operator on top of `ops` (unary when a single value is left)
  auto reduce_top = [&]() {
    if (values.empty()) fatal("Error: Ops is not empty with top element being " + ops.top() + ", while values is.\n");
    if (values.size() == 1) {
      const std::string op = ops.top();
      if (op == "-") { Var v = values.top(); values.pop(); ops.pop(); values.push(-v); }
      else if (op == "!") { Var v = values.top(); values.pop(); ops.pop(); values.push(!v); }
      else if (op == "+") ops.pop();
      else fatal("Error: do not know how to apply " + op + ", to " + values.top().eq() + ".\n");
    } else {
      Var b = values.top(); values.pop();
      Var a = values.top(); values.pop();
      std::string op = ops.top(); ops.pop();
      values.push(apply_op(a, op, b));
    }
  };
  auto announce = [&](const std::string &name) {
    double v = vars[name].result(this);
    if (echo) std::cout << name << " = " << v << std::endl;
  };

  for (int i = 0; i < n; i++) {
    const char c = str[i];
    if (isdigit((unsigned char)c)) {
      int j = 1;
      while (i + j < n && (isdigit((unsigned char)str[i + j]) || str[i + j] == '.')) j++;
      float f = std::stof(str.substr(i, j)); // literals are single precision in the reference (src/input.cpp:407,411)
      i += j - 1;
      if (negative) { values.push(Var((-1) * f)); negative = false; }
      else values.push(Var(f));
    } else if (c == '(') {
      ops.push("(");
      if (at(i + 1) == '-' && at(i + 2) != '(') { negative = true; i++; }
    } else if (c == ')') {
      if (ops.empty() && values.size() < 2) fatal("Error, unmatched parenthesis )\n");
      while (true) {
        if (ops.empty()) fatal("Error, unmatched parenthesis )\n");
        if (ops.top() == "(") break;
        reduce_top();
      }
      ops.pop();
    } else if (is_operator(c) || (c == '=' && at(i + 1) == '=')) {
      std::string op(1, c);
      if (values.empty() && at(i + 1) != '(' && op == "-") { negative = true; continue; }
      if (i + 1 >= n) fatal("Error: end-of-line character detected after operator " + op + ".\n");
      if (at(i + 1) == '*') { op.push_back('*'); i++; }
      else if (at(i + 1) == '=') { op.push_back('='); i++; }
      else if (is_operator(at(i + 1)) && at(i + 1) != '-') fatal("Error: unknown operator sequence " + op + at(i + 1) + ".\n");
      while (!ops.empty() && precedence(ops.top()) >= precedence(op)) {
        if (values.size() < 2) fatal("Error: malformed expression near operator " + op + ".\n");
        Var b = values.top(); values.pop();
        Var a = values.top(); values.pop();
        std::string top = ops.top(); ops.pop();
        values.push(apply_op(a, top, b));
      }
      ops.push(op);
      if (at(i + 1) == '-') { negative = true; i++; }
    } else {
      int j = 1;
      while (i + j < n && !is_math_char(str[i + j])) j++;
      std::string word = str.substr(i, j);
      i += j - 1;

      if (word == "E" || word == "e") { // power-of-ten operator: digit, e, sign (src/input.cpp:561-574)
        if (!values.empty() && i >= 1 && isdigit((unsigned char)str[i - 1]) && (at(i + 1) == '+' || at(i + 1) == '-')) {
          ops.push("e");
          if (at(i + 1) == '-') { negative = true; i++; }
          if (at(i + 1) == '+') i++;
          continue;
        }
      }

      auto it = vars.find(word);
      if (it != vars.end()) {
        if (at(i + 1) == '=' && at(i + 2) != '=' && i + 1 < n) {
          if (!values.empty() || !ops.empty()) fatal("Error: I do not understand when '=' is located in the middle of an expression\n");
          returnvar = word; i++;
        } else if (i + 1 >= n && values.empty() && ops.empty() && !trace_active()) {
          Var r = negative ? -vars[word] : vars[word];
          if (!returnvar.empty()) { vars[returnvar] = r; announce(returnvar); }
          return r;
        } else {
          Var leaf = vars[word];
          if (trace_active()) { // particle variables become the leaves of the traced tree
            static const char *pv[6] = {"x", "y", "z", "x0", "y0", "z0"};
            for (int k = 0; k < 6; k++) if (word == pv[k]) leaf = Var(word, leaf.result(), false, xnode_make(KML_X_VAR, (double)k, nullptr, nullptr));
          }
          if (negative) { values.push(-leaf); negative = false; }
          else values.push(leaf);
        }
      } else if (i + 1 < n && str[i + 1] == '=' && values.empty() && ops.empty()) {
        returnvar = word;
        if (protected_variable(returnvar)) fatal("Error: " + returnvar + " is a protected variable: it cannot be modified!\n");
        i++;
      } else if (i + 1 < n && str[i + 1] == '(') {
        i += 2;
        int k = 0, nl = 0, nr = 0;
        while (at(i + k) != ')' || nl != nr) {
          if (at(i + k) == '(') nl++;
          if (at(i + k) == ')') nr++;
          k++;
          if (i + k > n) fatal("Error: Unbalanced parenthesis '('.\n");
        }
        std::string arg = str.substr(i, k);
        i += k;
        // NB: the reference does not clear `negative` here (src/input.cpp:689-690)
        if (negative) values.push(-evaluate_function(word, arg));
        else values.push(evaluate_function(word, arg));
      } else {
        fatal("Error: " + word + " is unknown.\n");
      }
    }
  }

  while (!ops.empty()) reduce_top();

  if (values.empty()) {
    if (!returnvar.empty()) vars[returnvar] = Var(-1);
    return Var(-1);
  }
  if (!returnvar.empty()) { vars[returnvar] = values.top(); announce(returnvar); }
  return values.top();
}

Var Input::evaluate_function(const std::string &func, const std::string &arg) {
  // the reference splits on every comma, nested or not (src/input.cpp:246-263)
  std::vector<std::string> args;
  const int len = (int)arg.length();
  int start = 0;
  for (int i = 0; i < len; i++) {
    if (arg[i] == ',' || i == len - 1) {
      if (i == start && i != len - 1) fatal("Error: missing argument.\n");
      args.push_back(arg.substr(start, i - start + (i == len - 1)));
      start = i + 1;
    }
  }
  if (func == "exp") return fn1("exp", std::exp, parsev(arg));
  if (func == "sqrt") return fn1("sqrt", std::sqrt, parsev(arg));
  if (func == "cos") return fn1("cos", std::cos, parsev(arg));
  if (func == "sin") return fn1("sin", std::sin, parsev(arg));
  if (func == "tan") return fn1("tan", std::tan, parsev(arg));
  if (func == "log") return fn1("log", std::log, parsev(arg));
  if (func == "atan2") {
    if (args.size() != 2) fatal("Error: atan2 takes exactly two positional arguments.\n");
    return atan2v(parsev(args[0]), parsev(args[1]));
  }
  if (func == "value") { // src/input.cpp:856-865
    if (args.size() != 1) fatal("Error: value() takes exactly one argument.\n");
    Var v = parsev(args[0]); v.make_constant(this); return v;
  }
  if (func == "evaluate") return Var(parsev(arg).result(this));
  if (func == "print") {
    if (args.size() != 1) fatal("Error: print() takes exactly one argument.\n");
    Var v = parsev(args[0]);
    double r = v.result(this);
    if (echo) std::cout << args[0] << " = {equation=\"" << v.eq() << "\", value=" << r << ", constant=" << (v.is_constant() ? "true" : "false") << "}\n";
    return Var(0);
  }
  auto it = commands.find(func);
  if (it != commands.end()) return it->second(args);
  fatal("Error: Unknown function " + func + "\n");
}

// postorder emission; returns the evaluation stack depth the subtree needs (-1: the program does not fit)
static int flatten(const XRef &n, kml_expr *out) {
  if (!n) return -1;
  int need = 1;
  if (n->a) { need = flatten(n->a, out); if (need < 0) return -1; }
  if (n->b) { const int nb = flatten(n->b, out); if (nb < 0) return -1; need = std::max(need, 1 + nb); }
  if (out->n >= KML_EXPR_MAX) return -1;
  out->op[out->n] = n->op; out->val[out->n] = n->val; out->n++;
  return need;
}

bool Input::compile(const Var &v, kml_expr *out) {
  out->n = 0;
  if (v.is_constant()) { out->op[0] = KML_X_CONST; out->val[0] = v.result(); out->n = 1; return true; }
  const bool was_echo = echo; echo = false;
  trace_set(true);
  Var r;
  try { r = parsev(v.eq()); } catch (...) { trace_set(false); echo = was_echo; throw; }
  trace_set(false); echo = was_echo;
  const int need = flatten(r.xnode(), out);
  return need > 0 && need <= 16; // the device evaluates on a 16-entry stack
}

Var Input::line(const std::string &text) {
  std::string l = text.substr(0, text.find('#'));
  while (!l.empty() && (l.back() == '\r' || l.back() == '\n')) l.pop_back();
  Var v = parsev(l);
  v.result();
  return v;
}

void Input::file(const std::string &path) {
  std::ifstream f(path);
  if (!f) fatal("Cannot open input script " + path);
  std::string l;
  while (std::getline(f, l)) {
    std::string s = l.substr(0, l.find('#'));
    if (s == "quit") break;
    line(l);
  }
}

} // namespace kmlh
