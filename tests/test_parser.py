"""Host script interpreter: the parity-critical quirks of the reference's Input::parsev / Var
(reference src/input.cpp:374-735, src/var.cpp), checked against values printed by the
unmodified reference binary (oracle/_ref) for the same lines."""
import numpy as np
import pytest

from karamelo_b200.api import Engine, KmlError


def f32(x):
    return float(np.float32(x))


@pytest.fixture()
def eng(oracle_lib):
    e = Engine(oracle_lib)
    yield e
    e.close()


def test_literals_are_single_precision(eng):
    eng.line("nu = 0.3")
    assert eng.var("nu") == f32(0.3) != 0.3


def test_power_of_ten_operator(eng):
    eng.line("rho = 8.94e-06")
    assert eng.var("rho") == f32(8.94) * 10.0 ** -6
    eng.line("E = 1e+3")
    assert eng.var("E") == 1000.0
    with pytest.raises(KmlError):
        eng.line("bad = 1e3")  # no sign: "e3" is an unknown word in the reference too


def test_precedence_and_unary_minus(eng):
    eng.line("a = 2+3*4")
    assert eng.var("a") == 14
    eng.line("b = -2*3")
    assert eng.var("b") == -6
    eng.line("c = 2*-3")
    assert eng.var("c") == -6
    eng.line("d = 2^3**2")  # both are power, left-associative in the reference: (2^3)^2
    assert eng.var("d") == 64
    eng.line("e1 = (1-2*0.25)/(3*(1+0.5))")
    assert eng.var("e1") == (1 - 2 * 0.25) / (3 * (1 + 0.5))
    eng.line("K = 115/(3*(1-2*0.31))")
    assert eng.var("K") == 115 / (3 * (1 - 2 * f32(0.31)))


def test_lazy_variables_reparse_constants_through_float(eng):
    # a non-constant expression keeps its text; constants inside were serialised with %.15f
    # and are re-read by stof on every evaluation (reference src/var.cpp:24-60)
    eng.line("K = 115/(3*(1-2*0.31))")
    eng.line("g = K*time")
    eng.line("method(ulmpm, FLIP, linear, 0.99)")
    K = eng.var("K")
    assert K != f32(K)
    # time is 0 -> g is 0; make time visible through a second variable
    eng.line("h = K+time")
    assert eng.var("h") == f32(float("%.15f" % K))


def test_functions_and_value(eng):
    eng.line("s = sqrt(16)+exp(0)+cos(0)")
    assert eng.var("s") == 6
    eng.line("p = PI")
    assert eng.var("p") == np.pi
    eng.line("w = value(2*time+1)")
    assert eng.var("w") == 1


def test_comparison_operators(eng):
    eng.line("t1 = 3>2")
    eng.line("t2 = 3<=2")
    eng.line("t3 = 2==2")
    assert (eng.var("t1"), eng.var("t2"), eng.var("t3")) == (1, 0, 1)


def test_grid_and_particle_counts_match_reference_formulas(eng):
    # two-disks: 21 x 21 nodes, 2 x 208 particles
    from cases import two_disks
    eng.script(two_disks())
    assert eng.solid_info(0)["n"] == (21, 21, 1)
    assert [eng.solid_info(i)["np"] for i in range(2)] == [208, 208]


def test_unknown_command_is_an_error(eng):
    with pytest.raises(KmlError, match="Unknown function"):
        eng.line("frobnicate(1)")
