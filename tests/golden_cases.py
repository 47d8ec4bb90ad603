"""The reference's golden cases restated (values of their musubi.lua, not code), shared by
the oracle tests (CPU) and the device tests (GPU):

  gaussianPulse       mus/examples/fluid/benchmark/gaussianPulse/musubi.lua
  gaussianPulse (incompressible)  mus/examples/fluid_incompressible/benchmark/gaussianPulse/musubi.lua
  TGV_Simple_Re800    mus/examples/fluid_incompressible/benchmark/TaylorGreenVortex/TGV_Simple/
  TGV_Simple_Re1600   .../TGV_Simple_Re1600/musubi.lua
  tutorial Gausspulse mus/examples/tutorials/tutorial_cases/tutorial_gaussian_pulse/musubi.lua

Each setup returns an oracle Scheme holding the initial condition exactly as
mus_init_pdf (mus_flow_module.fpp:422-601) builds it, plus the unit conversion.
"""
import math
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD_PULSE_IC = {lv: [os.path.join(GOLDEN_DIR, "gaussianPulse-L%d_pressAlongLength_p0000%d_t0.000E+00.res" % (lv, r))
                      for r in range(3)] for lv in (5, 6)}
GOLD_PULSE = os.path.join(GOLDEN_DIR, "gaussianPulse_pressAlongLength_p00000_t10.001E+00.res")
GOLD_PULSE_INCOMP = {     # level -> (initial state or None, final state, steps)
    4: (None, os.path.join(GOLDEN_DIR, "incomp_gaussianPulse_pressAlongLength_p00000_t10.001E+00.res"), 9506),
    5: (os.path.join(GOLDEN_DIR, "incomp_gaussianPulse-L5_pressAlongLength_p00000_t0.000E+00.res"),
        os.path.join(GOLDEN_DIR, "incomp_gaussianPulse-L5_pressAlongLength_p00000_t10.000E+00.res"), 19011),
    6: (os.path.join(GOLDEN_DIR, "incomp_gaussianPulse-L6_pressAlongLength_p00000_t0.000E+00.res"),
        os.path.join(GOLDEN_DIR, "incomp_gaussianPulse-L6_pressAlongLength_p00000_t10.000E+00.res"), 38022),
}
GOLD_TGV800 = os.path.join(GOLDEN_DIR, "TGV_Simple_Re800_probeAtCenter_p00000.res")
GOLD_TGV1600 = os.path.join(GOLDEN_DIR, "TGV_Simple_Re1600_kE_all_p00000.res")
GOLD_TUTORIAL_PULSE = os.path.join(GOLDEN_DIR, "tutorial_Gausspulse_track_pressure_p00000.res")


def gaussian_pulse_setup(mo, nranks=1, rank=0, level=4, kind="fluid"):
    """kind = "fluid": mus/examples/fluid/benchmark/gaussianPulse/musubi.lua (IC from the Lua function
    ic_3Dgauss_pulse); "fluid_incompressible": mus/examples/fluid_incompressible/benchmark/
    gaussianPulse/musubi.lua (IC predefined 'gausspulse', tem_ic_predefs_module.f90:230-255 -- the
    same expression), otherwise the same values."""
    length = 10.0
    dx = length / 2.0 ** level
    nu_phy, cs_phy, rho0 = 0.01, 343.0, 1.0
    cs_lat = 1.0 / math.sqrt(3.0)
    dt = cs_lat / cs_phy * dx
    phys = mo.Physics(dx, dt, rho0)
    nu_lat = nu_phy / phys.fac_visc
    omega = 1.0 / (3.0 * nu_lat + 0.5)
    nsteps = int(math.ceil(10.0 / dt))
    ld = mo.build_level_desc(level, 19, "periodic", rank, nranks)
    sch = mo.Scheme(ld, "bgk", kind, omega=omega)
    sch.visc[:] = nu_lat
    bary = mo.barycenters(ld, (0.0, 0.0, 0.0), length)
    r = (bary[:, 0] - 5.0) ** 2 + (bary[:, 1] - 5.0) ** 2 + (bary[:, 2] - 5.0) ** 2
    p = rho0 * cs_phy ** 2 + 1.20 * np.exp((-math.log(2.0) / 1.0 ** 2) * r)
    rho = p * 3.0 * (1.0 / phys.fac_press)        # rho*cs2inv*inv_p, mus_flow_module.fpp:527
    sch.init_equilibrium(rho, np.zeros(3))
    return sch, phys, bary, nsteps


def tutorial_pulse_setup(mo):
    """tutorial_gaussian_pulse/musubi.lua: fluid / bgk / d3q19 on the predefined cube of edge 10 at
    refinement level 6 -- 64^3 periodic, the mesh and the kernel of BASELINE config 1 --, NO physics
    table, so every conversion factor is 1 (mus_load_physics, mus_physics_module.f90:231-236) and
    kinematic_viscosity = 0.03 is the lattice viscosity; IC: a plane pressure pulse
    p0 + 0.01 exp(-(x - 5)^2 / 2), fluid at rest; 50 steps; tracking: the element that holds the
    point (1, 1, 1), density / pressure / velocity after every step.  Returns (scheme, 0-based probe
    element, number of steps)."""
    level, length, nu = 6, 10.0, 0.03
    dx = length / 2.0 ** level
    ld = mo.build_level_desc(level, 19, "periodic")
    sch = mo.Scheme(ld, "bgk", "fluid", omega=1.0 / (3.0 * nu + 0.5))
    sch.visc[:] = nu
    bary = mo.barycenters(ld, (0.0, 0.0, 0.0), length)
    p = 1.0 * (1.0 / 3.0) + 0.01 * np.exp(-0.5 / 1.0 ** 2 * (bary[:, 0] - 5.0) ** 2)
    sch.init_equilibrium(p * 3.0, np.zeros(3))               # rho = p * cs2inv * inv_p, mus_flow_module.fpp:527
    c = (math.floor(1.0 / dx) + 0.5) * dx                    # barycentre of the element holding (1, 1, 1)
    probe = int(np.nonzero((np.abs(bary[:, 0] - c) < 1e-9) & (np.abs(bary[:, 1] - c) < 1e-9)
                           & (np.abs(bary[:, 2] - c) < 1e-9))[0][0])
    return sch, probe, 50


def tutorial_pulse_row(k, aux):
    """time (dt = 1), density, pressure = rho cs^2, velocity of the probe after k steps"""
    return [float(k), aux[0], aux[0] * (1.0 / 3.0), aux[1], aux[2], aux[3]]


def pulse_line_elements(sch, bary, level=4):
    """tracking shape canoND origin (0, 5, 5) vec (10,0,0): the 2^level cells whose lower face
    lies at y = z = 5 (barycentre 5 + dx/2: 5.3125 at level 4 as in the golden file), ascending x;
    0-based element indices."""
    c = 5.0 + 0.5 * 10.0 / 2.0 ** level
    sel = np.nonzero((np.abs(bary[:sch.ld.nFluid, 1] - c) < 1e-9)
                     & (np.abs(bary[:sch.ld.nFluid, 2] - c) < 1e-9))[0]
    return sel[np.argsort(bary[sel, 0])]


def pulse_track(aux4, sel, phys, bary):
    """density_phy, pressure_phy, velocity_phy of the tracked cells (mus_derQuan_module.fpp:659)."""
    aux = aux4[sel]
    dens = aux[:, 0] * phys.rho0
    press = aux[:, 0] * (1.0 / 3.0) * phys.fac_press
    vel = aux[:, 1:4] * phys.fac_vel
    return np.column_stack([bary[sel], dens, press, vel])


def tgv_setup(mo, Re, level, Ma, relaxation, origin):
    """fluid_incompressible Taylor-Green vortex in the periodic cube of edge 2 pi: IC from the
    analytic pressure / velocity / strain rate (acoustic f_neq)."""
    length = 2.0 * math.pi
    dx = length / 2.0 ** level
    rho0, cs_phy, u0 = 1.0, 343.0, 1.0
    nu_phy = 1.0 / Re
    cs_lat = math.sqrt(1.0 / 3.0)
    vel_lat = Ma * cs_lat
    dt = dx * vel_lat / 1.0
    phys = mo.Physics(dx, dt, rho0)
    nu_lat = nu_phy / phys.fac_visc
    omega = 1.0 / (3.0 * nu_lat + 0.5)
    if relaxation == "mrt":      # fluid = { bulk_viscosity = 2*nu_phy/3 }, mus_fluid_module.f90:468-485
        omega_bulk = mo.lib().ora_omega_bulk((2.0 * nu_phy / 3.0) / phys.fac_visc)
    else:
        omega_bulk = omega
    ld = mo.build_level_desc(level, 19, "periodic")
    sch = mo.Scheme(ld, relaxation, "fluid_incompressible", omega=omega, omega_bulk=omega_bulk)
    sch.visc[:] = nu_lat
    b = mo.barycenters(ld, origin, length)
    x, y, z = b[:, 0], b[:, 1], b[:, 2]
    p0 = rho0 * cs_phy ** 2
    vx = u0 * np.sin(x) * np.cos(y) * np.cos(z)
    vy = -u0 * np.cos(x) * np.sin(y) * np.cos(z)
    p1 = np.cos(2 * x) * (np.cos(2 * z) + 2.0)
    p2 = np.cos(2 * y) * (np.cos(2 * z) + 2.0)
    p = p0 + (p1 + p2) / 16.0
    sxx = np.cos(x) * np.cos(y) * np.cos(z)
    sxz = -0.5 * np.sin(x) * np.cos(y) * np.sin(z)
    syz = 0.5 * np.cos(x) * np.sin(y) * np.sin(z)
    inv_p, inv_v, inv_s = 1.0 / phys.fac_press, 1.0 / phys.fac_vel, 1.0 / phys.fac_strainRate
    zero = np.zeros_like(x)
    rho = p * 3.0 * inv_p
    vel = np.stack([vx * inv_v, vy * inv_v, zero], axis=1)
    S6 = np.stack([sxx * inv_s, -sxx * inv_s, zero, zero, syz * inv_s, sxz * inv_s], axis=1)
    sch.init_pdf(rho, vel, S6)
    return sch, phys, b, omega_bulk


def tgv800_setup(mo):
    """TGV_Simple_Re800: d3q19 mrt, level 6, Ma 0.09, probe at the cube centre every step."""
    sch, phys, b, ob = tgv_setup(mo, 800, 6, 0.09, "mrt", (0.0, 0.0, 0.0))
    nsteps = int(math.ceil(10.0 / phys.dt))
    c = math.pi + 0.5 * phys.dx                   # the cell whose lower corner is the centre
    probe = int(np.nonzero((np.abs(b[:, 0] - c) < 1e-9) & (np.abs(b[:, 1] - c) < 1e-9)
                           & (np.abs(b[:, 2] - c) < 1e-9))[0][0])
    return sch, phys, probe, nsteps, ob


def tgv800_row(k, aux, phys):
    """time, velocity_phy(3), pressure_phy of the probe after k steps"""
    return [k * phys.dt, aux[1] * phys.fac_vel, aux[2] * phys.fac_vel, aux[3] * phys.fac_vel,
            aux[0] * (1.0 / 3.0) * phys.fac_press]


def tgv1600_setup(mo):
    """TGV_Simple_Re1600 (shepherd): d3q19 bgk, level 7, Ma 0.15, sum of kinetic_energy_phy."""
    sch, phys, b, _ = tgv_setup(mo, 1600, 7, 0.15, "bgk", (-math.pi, -math.pi, -math.pi))
    nsteps = int(math.ceil(2.0 / phys.dt))
    return sch, phys, nsteps


def tgv1600_sample_steps(gold, phys):
    """the tracking interval (sim = 1/100) is coarser than dt: the golden rows are the steps
    0, 2, 4, ..., 470 and the last step 471; recovered from the file's time column."""
    k = np.round(gold[:, 0] / phys.dt).astype(int)
    assert np.max(np.abs(gold[:, 0] - k * phys.dt)) < 1e-10
    return k


def kinetic_energy_phy(aux4, nFluid, phys):
    """sum over the fluid elements of get_kineticEnergy_from_vel_dens_incompressible
    (sum(vel*vel)*0.5*rho0) times fac%energy = rho0 dx^5 / dt^2 (mus_physics_module.f90)."""
    v = aux4[:nFluid, 1:4]
    ke = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]) * 0.5 * 1.0
    fac_energy = phys.rho0 * phys.dx ** 5 / phys.dt ** 2
    return float(np.sum(ke * fac_energy))
