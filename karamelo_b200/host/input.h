// input.h - the Karamelo input-script interpreter (host side).
//
// Mirrors the reference's Input (reference src/input.h, src/input.cpp:101-142 file(),
// :374-735 parsev(), :242-349 evaluate_function()): one line = one expression; numeric
// literals go through stof (float!), `e`/`E` after a digit is the power-of-ten operator,
// name(args...) dispatches to a command.  Commands are registered by the Sim.
#pragma once
#include "var.h"
#include "../../include/kml.h"
#include <functional>
#include <map>
#include <string>
#include <vector>

namespace kmlh {

class Sim;

class Input {
public:
  explicit Input(Sim *sim);
  void file(const std::string &path);  // src/input.cpp:101-142
  Var line(const std::string &text);   // one script line (comments stripped)
  Var parsev(std::string str);         // src/input.cpp:374-735
  Var evaluate_function(const std::string &func, const std::string &arg); // src/input.cpp:242-349
  // The expression of `v` as a postfix program over the particle variables (include/kml.h kml_expr): the equation string is parsed once
  // more with tracing on, so literals take the same float path and the operations the same order as the per-particle re-parse of the
  // reference.  Returns false when the program does not fit (the caller keeps the per-particle host evaluation).
  bool compile(const Var &v, struct ::kml_expr *out);

  std::map<std::string, Var> vars;
  typedef std::function<Var(std::vector<std::string> &)> Command;
  std::map<std::string, Command> commands;
  bool echo = true; // print "name = value" like the reference does on rank 0

private:
  Sim *sim;
  static int precedence(const std::string &op);                       // src/input.cpp:154-178
  Var apply_op(const Var &a, const std::string &op, const Var &b);    // src/input.cpp:186-218
  bool protected_variable(const std::string &name) const;             // src/input.cpp:926-932
};

[[noreturn]] void fatal(const std::string &msg); // error->all / error->one (src/error.cpp:33-76): throws

} // namespace kmlh
