"""Parity cases: Karamelo scripts (setup only - the tests append their own run commands).

Each case is a reduced variant of one BASELINE.json config, written in the reference's own
input language so the unmodified reference (oracle/_ref), the CPU oracle and the CUDA engine
all run exactly the same text.  Discrepancies between BASELINE.json and the shipped example
files are resolved as SURVEY.md section 8 prescribes (scheme(usl) added to C1, cubic-spline
variant of C2, thermo-mechanical variant of C3, both contact fixes for C4).
"""

# C1: examples/two-disks.mpm - 2-D ULMPM, two elastic disks colliding, linear shape functions.
# The disks start closer together so that they collide within the first 100 steps.
def two_disks(scheme="usl", v=0.1, c=0.2, method="method(ulmpm, FLIP, linear, FLIP)"):
    return f"""
E   = 1e+3
nu  = 0.3
rho = 1000
L   = 1
hL  = 0.5*L
FLIP=1.0
{method}
scheme({scheme})
N        = 20
cellsize = L/N
dimension(2,-hL, hL, -hL, hL, cellsize)
R = 0.2
c = {c}
region(rBall1, cylinder, -c, -c, R)
region(rBall2, cylinder,  c,  c, R)
material(mat1, linear, rho, E, nu)
ppc1d = 2
solid(sBall1, region, rBall1, ppc1d, mat1, cellsize,0)
solid(sBall2, region, rBall2, ppc1d, mat1, cellsize,0)
group(gBall1, particles, region, rBall1, solid, sBall1)
group(gBall2, particles, region, rBall2, solid, sBall2)
v = {v}
fix(v0Ball1, initial_velocity_particles, gBall1,  v,  v, NULL)
fix(v0Ball2, initial_velocity_particles, gBall2, -v, -v, NULL)
set_dt(0.01)
"""


# C2: examples/Taylor-bar/cylindrical/ULMPM - 3-D ULMPM, Johnson-Cook + shock EOS, MUSL, a short
# bar one cell away from the rigid wall so that impact plasticity develops within 100 steps.
def taylor_bar(shape="cubic-spline", N=2, scheme="musl"):
    return f"""
E   = 115
nu  = 0.31
K   = E/(3*(1-2*nu))
G   = E/(2*(1+nu))
rho = 8.94e-06
sigmay  = 0.065
B       = 0.356
C       = 0.013
n       = 0.37
m       = 0
eps0dot = 1e-3
Tm      = 1600
hLx   = 3
R     = 1.6
Df    = 3
S     = 1.5
c0    = 3933
Gamma = 0
cv = 0
Tr = 25
FLIP = 0.99
method(ulmpm, FLIP, {shape}, FLIP)
scheme({scheme})
N        = {N}
cellsize = 1/N
x_wall = hLx + 2*cellsize
dimension(3, -hLx-1, x_wall, -Df, Df, -Df, Df, cellsize)
eos(eoss, shock, rho, K, c0, S, Gamma, cv, Tr, 0, 0)
strength(strengthJC, johnson_cook, G, sigmay, B, n, eps0dot, C, m, Tr, Tm)
material(mat, eos-strength, eoss, strengthJC)
region(cyl, cylinder, x, 0, 0, R, -hLx, hLx)
ppc1d = 2
solid(solid1, region, cyl, ppc1d, mat, cellsize, Tr)
region(region2, block, x_wall - cellsize, INF, INF, INF, INF, INF)
group(groupn2, nodes, region, region2, solid, solid1)
v = 190
group(gBall1, particles, region, cyl, solid, solid1)
fix(v0Ball1, initial_velocity_particles, gBall1, v, NULL, NULL)
fix(BC_Wall, velocity_nodes, groupn2, 0, NULL, NULL)
dt_factor(0.25)
"""


# C3: examples/Tensile_with_damage/Bernstein - TLMPM, Bernstein quadratic, JC strength + JC damage,
# optionally plastic-work heating.  Grip velocity raised and the JC failure strain lowered so that yield and
# damage (up to fully failed particles) occur within 100 steps.  The case is deliberately kept well conditioned:
# an earlier variant (artificial viscosity Q2 = 1.5, above the explicit stability limit) made the REFERENCE itself
# amplify a 1e-15 perturbation to 1e-5 in 100 steps, which no 1e-10 parity test can survive; this one amplifies
# it to ~1e-11 (measured with the oracle).  The oblique impact of C4 exists for the same reason: in a head-on
# collision the Coulomb friction direction v_t/|v_t| is rounding noise.
def tensile(thermal=False, shape="Bernstein-quadratic", vgrip=40, kind="tlmpm", Gamma=0, cv=0, m=0, C=0.01, d4=0, d5=0, alpha=0, cp="452e+6", depsdot0=1):
    """Gamma, cv: Mie-Gruneisen energy term e = cv rho0 (T - Tr) (src/eos_shock.cpp:99-133); m: JC thermal softening
    (src/strength_jc.cpp:127-133); d4, d5: strain-rate and temperature factors of the JC failure strain
    (src/damage_jc.cpp:120-129); alpha: thermal pressure alpha (T0 - T) (src/temperature_plastic_work.cpp:59-69, src/solid.cpp:1309-1312)"""
    method = f"method({kind}, FLIP, {shape}, FLIP" + (", thermo-mechanical)" if thermal else ")")
    tmat = f"temperature(tpw, plastic_work, 0.9, {cp}, 50, {alpha}, Tr, Tm)\n" if thermal else ""
    mat4 = "material(mat4, eos-strength, eoss, strengthjc, damagejc" + (", tpw)" if thermal else ")")
    return f"""
E = 211
nu = 0.33
K = E/(3*(1-2*nu))
G = E/(2*(1+nu))
rho = 7.75e-06
sigmay = 0.499
B = 0.382
n = 0.458
hLx = 10
hLy = 1.2
hLz = 1.2
S = 1.5
c0 = 5030
FLIP=0.99
cellsize = 0.8
{method}
dimension(3, -2*hLx, 2*hLx, -2*hLy, 2*hLy, -2*hLz, 2*hLz, cellsize)
region(box, block, -hLx, hLx, -hLy, hLy, -hLz, hLz)
Q1 = 0.06
Q2 = 0.1
Tr = 25
Tm = 1000
cv = {cv}
Gamma = {Gamma}
eos(eoss,   shock, rho, K, c0, S, Gamma, cv, Tr, Q1, Q2)
strength(strengthjc, johnson_cook, G, sigmay, B, n, 1, {C}, {m}, Tr, Tm)
d1 = 0.0382
d2 = 0.1162
d3 = -2.969
d4 = {d4}
d5 = {d5}
epsdot0 = {depsdot0}
damage(damagejc, damage_johnson_cook, d1, d2, d3, d4, d5, epsdot0, Tr, Tm)
{tmat}{mat4}
ppc = 2
solid(solid1, region, box, ppc, mat4, cellsize, Tr)
xBC = 9.7 - cellsize
region(region1, block, INF, -xBC, INF, INF, INF, INF)
group(groupn1, nodes, region, region1, solid, solid1)
region(region2, block, xBC, INF, INF, INF, INF, INF)
group(groupn2, nodes, region, region2, solid, solid1)
v = {vgrip}*(1.0-exp(-20000*time))
fix(BC_left, velocity_nodes, groupn1, -v, NULL, NULL)
fix(BC_right, velocity_nodes, groupn2, v, NULL, NULL)
"""


# C4: examples/Bouncing_balls/TLMPM/FLIP - TLMPM, two solids on private grids, contact fix.
def bouncing_balls(contact="minimize_penetration", v=0.5, method="method(tlmpm, FLIP, linear, alphaFLIP)"):
    fix = "fix(contact, contact/minimize_penetration, sBall1, sBall2, 0.3)" if contact == "minimize_penetration" else "fix(contact, contact/hertz, sBall1, sBall2)"
    return f"""
E   = 1e+3
nu  = 0.3
rho = 1000
L    = 1
hL   = 0.5*L
alphaFLIP=1
{method}
N        = 40
cellsize = L/N
dimension(2,-hL, hL, -hL, hL, cellsize)
R = 0.2
c = 0.16
region(rBall1, cylinder, -c, -c, R)
region(rBall2, cylinder, c, c, R)
material(mat1, linear, rho, E, nu)
ppc1d = 2
solid(sBall1, region, rBall1, ppc1d, mat1, cellsize,0)
solid(sBall2, region, rBall2, ppc1d, mat1, cellsize,0)
group(gBall1, particles, region, rBall1, solid, sBall1)
group(gBall2, particles, region, rBall2, solid, sBall2)
v = {v}
fix(v0Ball1, initial_velocity_particles, gBall1, v, 0.2*v, NULL)
fix(v0Ball2, initial_velocity_particles, gBall2, -v, -0.3*v, NULL)
{fix}
set_dt(0.001)
"""


# C5: synthetic 3-D ULMPM elastoplastic block, cubic B-splines, at test size.
def block(n=(8, 8, 8), scheme="musl", shape="cubic-spline", fixed_dt=False, a=2.5e-3, ppc=2, strength="plastic", drift=0.0, margin=4):
    """drift: uniform x velocity added to the squeeze (moves particles across slab cuts in the multi-GPU tests);
    margin: empty cells between the block and the box (4 keeps every stencil on interior nodes; 1 puts the outer
    particles on the boundary-modified splines, ntype -2/-1/1/2 of src/grid.cpp:236-240)"""
    nx, ny, nz = n
    a_m, a_e = ("%e" % a).split("e")
    a_txt = "%s%s%+d" % (a_m.rstrip("0").rstrip("."), "e", int(a_e))  # e.g. 2.5e-3: the parser needs a signed exponent
    strength_cmd = {"plastic": "strength(s, plastic, G, sigmay)", "linear": "strength(s, linear, G)",
                    "swift": "strength(s, swift, G, sigmay, 2, 0.001, 0.3)"}[strength]
    return f"""
E = 1000
nu = 0.3
rho = 1000
K = E/(3*(1-2*nu))
G = E/(2*(1+nu))
sigmay = 3
h = 1
method(ulmpm, FLIP, {shape}, 0.99)
scheme({scheme})
dimension(3, 0, {nx + 2 * margin}, 0, {ny + 2 * margin}, 0, {nz + 2 * margin}, h)
region(box, block, {margin}, {nx + margin}, {margin}, {ny + margin}, {margin}, {nz + margin})
eos(e, linear, rho, K)
{strength_cmd}
material(m, eos-strength, e, s)
solid(blk, region, box, {ppc}, m, h, 0)
group(gall, particles, region, box, solid, blk)
a = {a_txt}
cx = {margin + nx / 2}
cy = {margin + ny / 2}
cz = {margin + nz / 2}
fix(v0, initial_velocity_particles, gall, {drift}-a*(x-cx), 0.5*a*(y-cy), 0.5*a*(z-cz))
{"set_dt(0.2)" if fixed_dt else "dt_factor(0.5)"}
"""


# extra coverage of the functor table (not BASELINE configs): neo-Hookean bar under gravity (USF, quadratic
# splines) and a Tait-fluid column (EOS fluid + fluid strength, 2-D).
def neo_hookean_bar(scheme="usf", shape="quadratic-spline"):
    return f"""
E = 1e+6
nu = 0.3
rho = 1050
L = 1
hL = 0.5*L
method(ulmpm, FLIP, {shape}, 0.99)
scheme({scheme})
N = 4
cellsize = L/N
dimension(3,-hL-2*cellsize, hL+2*cellsize, -3*L, 2*cellsize, -hL-2*cellsize, hL+2*cellsize, cellsize)
region(box, block, -hL, hL, -L, 0, -hL, hL)
material(mat1, neo-hookean, rho, E, nu)
solid(solid1, region, box, 2, mat1, cellsize, 0)
region(rBCLX, block, INF, INF, -cellsize/4, INF, INF, INF)
group(gBCLX, nodes, region, rBCLX, solid, solid1)
fix(fBCLX, velocity_nodes, gBCLX, 0, 0, 0)
gravity = -300
fix(fbody, body_force, all, 0, gravity, 0)
dt_factor(0.2)
"""


def fluid_column():
    return """
gamma = 7
K     = 1.4e+6
G     = 0.001
rhoW  = 998
method(ulmpm, FLIP, linear, 1.0)
cellsize = 0.01
dimension(2, -0.2, 0, 0, 0.14, cellsize)
eos(eosf, fluid, rhoW, K, gamma)
strength(strengthf, fluid, G)
material(mat1, eos-strength, eosf, strengthf)
region(water, block, -0.1, 0, 0, 0.08)
solid(solidW, region, water, 2, mat1, cellsize,0)
region(rBottomW, block, INF, INF, INF, cellsize/4)
group(gBottomW, nodes, region, rBottomW, solid, solidW)
region(rRight, block, -cellsize/4, INF, -INF, INF)
group(gRight, nodes, region, rRight, solid, solidW)
fix(fBCLYW, velocity_nodes, gBottomW, NULL, 0)
fix(fBCRX, velocity_nodes, gRight,  0, NULL)
gravity = -9.81
fix(fbody, body_force, all, 0, gravity)
set_dt(1e-6)
"""


# convected particle domain interpolation (2-D), the reference's Vertical_bar CPDI example
# (examples/Vertical_bar/CPDI/CPDI-r4|Q4/vertical_bar.mpm) with the method() syntax the current parser accepts
# (method, sub-method, shape, ratio, "mechanical", style - src/update.cpp:108-196; the shipped files predate it).
def cpdi_bar(method="ulcpdi", style="R4", shape="linear"):
    # TLCPDI indexes its candidate nodes from domain->boxlo although the TL grid starts at the solid's own corner
    # (src/tlcpdi.cpp:200-222): the particles only find their nodes when the two coincide, so the TL variants use the
    # box the example keeps commented out (-L..0); UL needs the room below the bar (-3L..0).
    ylo = "-L" if method == "tlcpdi" else "-3*L"
    return f"""
E = 1e+6
nu = 0.3
rho = 1050
L = 1
hL = 0.5*L
FLIP=0.99
method({method}, FLIP, {shape}, FLIP, mechanical, {style})
N = 5
cellsize = L/N
dimension(2,-hL, hL, {ylo}, 0, cellsize)
region(box, block, -hL, hL, -L, 0)
material(mat1, neo-hookean, rho, E, nu)
solid(solid1, region, box, 2, mat1, cellsize, 0)
region(rBCLX, block, INF, INF, -cellsize/4, INF)
group(gBCLX, nodes, region, rBCLX, solid, solid1)
fix(fBCLX, velocity_nodes, gBCLX, NULL, 0, NULL)
gravity = -1e+3
fix(fbody, body_force, all, 0, gravity, 0)
dt_factor(0.2)
"""


# examples/Cantilever/2D/cantilever_force_tip.mpm at test size: TLMPM neo-Hookean beam, clamped nodes, tip load through
# fix force_nodes (shared by the massive nodes of the group), energies through the energy fixes.
def cantilever(shape="linear"):
    return f"""
E = 1e+6
nu = 0.3
rho = 1050
L = 1
alpha=0.99
N = 8
cellsize = L/N
method(tlmpm, FLIP, {shape}, alpha)
dimension(2, 0, 4*L, 0, L, cellsize)
region(box, block, 0, 4*L, 0, L)
material(mat1, neo-hookean, rho, E, nu)
solid(solid1, region, box, 2, mat1, cellsize, 0)
region(rBCLX, block, INF, cellsize/4, INF, INF)
group(gBCLX, nodes, region, rBCLX, solid, solid1)
fix(fBCLX, velocity_nodes, gBCLX, 0, 0)
region(rEND, block, 4*L-cellsize/4, INF, L-5*cellsize/4, INF)
group(gEND, nodes, region, rEND, solid, solid1)
F = -20000
fix(fEnd, force_nodes, gEND, 0, F, 0)
group(gAll, particles, region, box, solid, solid1)
fix(Ek, kinetic_energy, gAll)
fix(Es, strain_energy, gAll)
dt_factor(0.5)
"""


# Rigid bodies (material(..., rigid, rho), src/material.h:49): the second disk / ball is rigid; nodes inside its stencils carry
# Grid::rigid, take mass and momentum from the rigid solid only and keep v_update = v, so the deformable disk sees the rigid
# one as a moving velocity boundary condition (a6, a7, a11, a12).  No shipped example still parses its own
# rigid material line (the usage gained a density argument), so these variants of C1 / C4 stand in.
def rigid_disks(scheme="musl", method="method(ulmpm, FLIP, linear, 0.99)", tl=False):
    base = bouncing_balls("minimize_penetration", method=method) if tl else two_disks(scheme, method=method)
    marker, mat = "solid(sBall2, region, rBall2, ppc1d, mat1, cellsize,0)", "material(mat1, linear, rho, E, nu)\n"
    assert marker in base and mat in base
    # both materials are declared before the first solid: the reference keeps raw pointers into a growing vector<Mat>
    # (src/material.cpp:293-298, src/solid.cpp:72), so a material added after a solid leaves that solid's pointer dangling
    return base.replace(mat, mat + "material(matr, rigid, rho)\n").replace(marker, "solid(sBall2, region, rBall2, ppc1d, matr, cellsize,0)")


# Fixes that sit between the stages beyond those of the BASELINE configs: velocity_particles (a rigid tool
# driven at a prescribed, time-dependent velocity), initial_stress, initial_velocity_nodes, temperature_nodes / _particles.
def driven_tool():
    return rigid_disks("musl", method="method(ulmpm, FLIP, cubic-spline, 0.99)").replace(
        "fix(v0Ball2, initial_velocity_particles, gBall2, -v, -v, NULL)",
        "fix(vtool, velocity_particles, gBall2, -v*(0.5+time), -v)")


def driven_tool_nonuniform():  # the same fix with a particle-dependent value: evaluated per particle on the host, not by the kernel
    return driven_tool().replace("-v*(0.5+time), -v)", "-v*(0.5+time), -v*(1+0.5*x0))")


def prestressed_disks():
    return two_disks("musl", method="method(ulmpm, FLIP, linear, 0.99)") + """
fix(s0, initial_stress, gBall1, 0.5, -0.25+x0, NULL, NULL, NULL, 0.125)
region(rKick, block, -0.1, 0.1, -0.1, 0.1)
group(gKick, nodes, region, rKick, solid, sBall1)
fix(vn0, initial_velocity_nodes, gKick, 0.05*y0, -0.05, NULL)
"""


def heated_bar(**kw):
    return tensile(True, **kw) + """
region(rHot, block, INF, -hLx+cellsize, INF, INF, INF, INF)
group(gHotN, nodes, region, rHot, solid, solid1)
fix(fTn, temperature_nodes, gHotN, Tr+100*time/0.01)
region(rCold, block, hLx-cellsize, INF, INF, INF, INF, INF)
group(gColdP, particles, region, rCold, solid, solid1)
fix(fTp, temperature_particles, gColdP, Tr-10+x0)
"""


# delete_particles (src/delete_particles.cpp): particles removed before the run; the survivors keep the reference's swap-with-last order
def carved_disks(tl=False):
    if tl:
        return bouncing_balls("minimize_penetration") + "region(rCut2, cylinder, -0.16, -0.16, 0.08)\ndelete_particles(sBall1, region, rCut2)\n"
    return two_disks("musl") + ("region(rCut, block, 0.15, 0.3, INF, 0.25)\ndelete_particles(sBall2, region, rCut)\n"
                                "region(rCut2, cylinder, -0.2, -0.2, 0.08)\ndelete_particles(all, region, rCut2)\n")


# Two elastic spheres colliding in 3-D on one background grid: several solids through the generic 3-D scatter / gather kernels
# (the cell kernels serve one solid per grid), sphere regions, cubic B-splines
def two_spheres(scheme="musl", shape="cubic-spline"):
    return f"""
E   = 1e+3
nu  = 0.3
rho = 1000
L   = 1
hL  = 0.5*L
method(ulmpm, FLIP, {shape}, 0.99)
scheme({scheme})
N        = 12
cellsize = L/N
dimension(3,-hL, hL, -hL, hL, -hL, hL, cellsize)
R = 0.17
c = 0.14
region(rBall1, sphere, -c, -c, -c, R)
region(rBall2, sphere,  c,  c,  c, R)
material(mat1, linear, rho, E, nu)
ppc1d = 2
solid(sBall1, region, rBall1, ppc1d, mat1, cellsize,0)
solid(sBall2, region, rBall2, ppc1d, mat1, cellsize,0)
group(gBall1, particles, region, rBall1, solid, sBall1)
group(gBall2, particles, region, rBall2, solid, sBall2)
v = 0.2
fix(v0Ball1, initial_velocity_particles, gBall1,  v,  v, 0.8*v)
fix(v0Ball2, initial_velocity_particles, gBall2, -v, -v, -0.8*v)
dt_factor(0.1)
"""


# Axisymmetric 2-D (axisymmetric(true), src/domain.cpp:554-569): x is the radius.  Hoop terms: f_I[0] -= vol sigma_22 wf / x_p
# (src/solid.cpp:499-517, TL :472), L_22 += v_I[0] wf / x_p (src/solid.cpp:900-916, TL :818-828), mass = rho0 vol0 x0[0]
# (src/solid.cpp:2292-2298).  A short Taylor bar hitting a wall along its axis; UL cubic / linear, TL linear.
def axisym_bar(kind="ulmpm", shape="cubic-spline", scheme="musl"):
    return f"""
E   = 115
nu  = 0.31
K   = E/(3*(1-2*nu))
G   = E/(2*(1+nu))
rho = 8.94e-06
sigmay  = 0.065
B       = 0.356
C       = 0.013
n       = 0.37
eps0dot = 1e-3
Tm      = 1600
S     = 1.5
c0    = 3933
Tr = 25
FLIP = 0.99
method({kind}, FLIP, {shape}, FLIP)
scheme({scheme})
axisymmetric(true)
N        = 4
cellsize = 1/N
dimension(2, 0, 4, -cellsize, 6, cellsize)
eos(eoss, shock, rho, K, c0, S, 0, 0, Tr, 0, 0)
strength(strengthJC, johnson_cook, G, sigmay, B, n, eps0dot, C, 0, Tr, Tm)
material(mat, eos-strength, eoss, strengthJC)
region(cyl, block, 0, 1.6, {"cellsize" if kind == "ulmpm" else "0"}, 4)
solid(solid1, region, cyl, 2, mat, cellsize, Tr)
region(rWall, block, INF, INF, INF, cellsize/4)
group(gWall, nodes, region, rWall, solid, solid1)
region(rAxis, block, INF, cellsize/4, INF, INF)
group(gAxis, nodes, region, rAxis, solid, solid1)
group(gAll, particles, region, cyl, solid, solid1)
v = 190
fix(v0, initial_velocity_particles, gAll, NULL, -v, NULL)
fix(BC_Wall, velocity_nodes, gWall, NULL, 0, NULL)
fix(BC_Axis, velocity_nodes, gAxis, 0, NULL, NULL)
dt_factor(0.25)
"""


FULL_THERMAL = dict(Gamma=1.2, cv=500, m=1.1, d4=0.02, d5=0.6, alpha="2e-3", depsdot0=0.01)

# name -> (script, is_TL, thermal, steps)
CASES = {
    "c1_two_disks_usl": (two_disks("usl"), False, False, 100),
    "c1_two_disks_musl": (two_disks("musl"), False, False, 100),
    "c2_taylor_cubic": (taylor_bar("cubic-spline"), False, False, 100),
    "c2_taylor_linear": (taylor_bar("linear"), False, False, 100),
    "c3_tensile_bernstein": (tensile(False), True, False, 100),
    "c3_tensile_thermal": (tensile(True), True, True, 100),
    "c4_balls_minpen": (bouncing_balls("minimize_penetration"), True, False, 100),
    "c4_balls_hertz": (bouncing_balls("hertz"), True, False, 100),
    "c5_block_musl": (block((8, 8, 8), "musl"), False, False, 100),
    "c5_block_usl_fixed_dt": (block((6, 6, 6), "usl", fixed_dt=True), False, False, 100),
    "x_neo_hookean_usf": (neo_hookean_bar(), False, False, 100),
    "x_fluid_column": (fluid_column(), False, False, 100),
    # edge cases of the cell-centric kernels: boundary-modified splines on every face (margin 1, pure compression so that no
    # particle leaves the box), one and 27 particles per cell (staging rounds that start and end inside a cell), a segment
    # boundary inside the block (40 cells along z > 32)
    "e_block_boundary_splines": (block((6, 6, 6), "musl", a=2.5e-3, margin=1).replace("0.5*a*(y-cy), 0.5*a*(z-cz)", "-0.5*a*(y-cy), -0.5*a*(z-cz)"), False, False, 100),
    "e_block_ppc1": (block((8, 8, 8), "musl", ppc=1), False, False, 100),
    "e_block_ppc3": (block((5, 5, 5), "usl", ppc=3), False, False, 100),
    "e_block_two_segments": (block((3, 3, 40), "musl"), False, False, 60),
    "x_cantilever_force_nodes": (cantilever(), True, False, 100),
    # affine sub-methods (rows a7 _APIC, a8 _MLS, a12 ASFLIP, a15 _APIC of SURVEY section 8)
    "x_apic_ul_cubic": (two_disks("musl", method="method(ulmpm, APIC, cubic-spline)"), False, False, 100),
    "x_mls_ul_cubic": (two_disks("usl", method="method(ulmpm, MLS, cubic-spline)"), False, False, 100),
    "x_asflip_ul_quadratic": (two_disks("musl", method="method(ulmpm, ASFLIP, quadratic-spline, 0.99)"), False, False, 100),
    "x_aflip_ul_cubic_usf": (two_disks("usf", method="method(ulmpm, AFLIP, cubic-spline, 0.95)"), False, False, 100),
    "x_apic_tl_linear": (bouncing_balls("minimize_penetration", method="method(tlmpm, APIC, linear)"), True, False, 100),
    "x_cpdi_ul_r4": (cpdi_bar("ulcpdi", "R4"), False, False, 100),
    "x_cpdi_ul_q4": (cpdi_bar("ulcpdi", "Q4"), False, False, 100),
    "x_cpdi_tl_r4": (cpdi_bar("tlcpdi", "R4"), True, False, 100),
    "x_cpdi_tl_q4": (cpdi_bar("tlcpdi", "Q4"), True, False, 100),
    # rigid bodies: ULMPM linear MUSL, ULMPM cubic USL, TLMPM with contact
    "x_rigid_ul_linear_musl": (rigid_disks("musl"), False, False, 100),
    "x_rigid_ul_cubic_usl": (rigid_disks("usl", method="method(ulmpm, FLIP, cubic-spline, 0.99)"), False, False, 100),
    "x_rigid_tl_contact": (rigid_disks(tl=True, method="method(tlmpm, FLIP, linear, 0.99)"), True, False, 100),
    # gradient-enhanced momentum projection (method(..., mechanical, gradient-enhanced), src/solid.cpp:369-371)
    "x_ge_ul_cubic_musl": (two_disks("musl", method="method(ulmpm, FLIP, cubic-spline, 0.99, mechanical, gradient-enhanced)"), False, False, 100),
    "x_ge_ul_linear_usl": (two_disks("usl", method="method(ulmpm, FLIP, linear, 0.99, mechanical, gradient-enhanced)"), False, False, 100),
    # fixes between the stages beyond the BASELINE configs
    "x_fix_velocity_particles": (driven_tool(), False, False, 100),
    "x_fix_initial_stress_velocity_nodes": (prestressed_disks(), False, False, 100),
    "x_fix_temperature": (heated_bar(), True, True, 100),
    "x_fix_velocity_particles_x0": (driven_tool_nonuniform(), False, False, 100),
    "x_delete_particles_ul": (carved_disks(False), False, False, 100),
    "x_delete_particles_tl": (carved_disks(True), True, False, 100),
    "x_two_spheres_3d": (two_spheres(), False, False, 100),
    "x_tensile_tl_cubic": (tensile(False, shape="cubic-spline", vgrip=20), True, False, 100),
    # functor branches no BASELINE example reaches: Swift and linear
    # strength through the cell stress kernel; Mie-Gruneisen energy term (Gamma, cv != 0), JC thermal softening (m != 0), JC
    # failure strain with the rate (d4) and temperature (d5) factors, thermal pressure alpha (T0 - T) - total- and updated-Lagrangian,
    # thermo-mechanical, with temperatures imposed on both sides of Tr
    "p_block_swift": (block((8, 8, 8), "musl", strength="swift"), False, False, 100),
    "p_block_linear_strength": (block((8, 8, 8), "musl", strength="linear"), False, False, 100),
    "p_thermal_full_tl": (heated_bar(**FULL_THERMAL), True, True, 100),
    "p_thermal_full_ul": (heated_bar(kind="ulmpm", shape="cubic-spline", vgrip=20, **FULL_THERMAL), False, True, 100),
    # axisymmetric 2-D (hoop terms of the scatter, the gradient and the lattice masses)
    "p_axisym_ul_cubic_musl": (axisym_bar(), False, False, 100),
    "p_axisym_ul_linear_usl": (axisym_bar("ulmpm", "linear", "usl"), False, False, 100),
    "p_axisym_tl_linear": (axisym_bar("tlmpm", "linear"), True, False, 100),
}
