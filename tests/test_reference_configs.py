"""The restated golden cases (tests/golden_cases.py) against the REFERENCE'S OWN configuration
scripts, evaluated by the reference's own Lua interpreter (aotus/external/lua-5.4.8 compiled from
where it lies into oracle/_ref/liblua_ref.so, oracle/lua_ref.py) -- the way aotus reads them:
mesh, time step, viscosity -> omega, number of steps, scheme identification, tracking objects, and
the initial-condition FUNCTIONS called at every barycentre.

Runs in the build container only (reads /root/reference/mus/examples/.../musubi.lua); skipped where
the reference tree is absent."""
import math
import os

import numpy as np
import pytest

from golden_cases import gaussian_pulse_setup, tgv800_setup, tgv1600_setup, tutorial_pulse_setup

EX = "/root/reference/mus/examples"
PULSE = EX + "/fluid/benchmark/gaussianPulse/musubi.lua"
PULSE_INC = EX + "/fluid_incompressible/benchmark/gaussianPulse/musubi.lua"
TGV = EX + "/fluid_incompressible/benchmark/TaylorGreenVortex/TGV_Simple/TGV_Simple_Re%d/musubi.lua"
TUTORIAL = EX + "/tutorials/tutorial_cases/tutorial_gaussian_pulse/musubi.lua"

lua_ref = pytest.importorskip("oracle.lua_ref")
pytestmark = pytest.mark.skipif(not (os.path.exists(PULSE) and lua_ref.available()),
                                reason="needs /root/reference and oracle/_ref/liblua_ref.so (make -C oracle ref)")


@pytest.fixture
def script():
    opened = []

    def load(path):
        s = lua_ref.LuaScript(path)
        opened.append(s)
        return s
    yield load
    for s in opened:
        s.close()


def _check_common(cfg, sch, phys, nsteps, level, kind, relaxation):
    assert cfg.get("mesh.predefined") == "cube" and cfg.get("mesh.refinementLevel") == level == sch.ld.level
    assert cfg.get("identify.kind") == kind and cfg.get("identify.layout") == "d3q19"
    rel = cfg.get("identify.relaxation")
    assert (rel["name"] if isinstance(rel, dict) else rel) == relaxation
    assert cfg.get("physics.dt") == phys.dt and cfg.get("physics.rho0") == phys.rho0
    assert abs(cfg.get("mesh.length") / 2.0 ** level / phys.dx - 1.0) < 1e-15
    nu_lat = cfg.get("fluid.kinematic_viscosity") / (phys.dx ** 2 / phys.dt)
    assert abs(sch.visc[0] / nu_lat - 1.0) < 1e-15
    assert math.ceil(cfg.get("sim_control.time_control.max.sim") / cfg.get("physics.dt")) == nsteps == cfg.get("tmax_iter")


def test_gaussian_pulse_case_is_the_references_script(oracle, script):
    cfg = script(PULSE)
    sch, phys, bary, nsteps = gaussian_pulse_setup(oracle)
    _check_common(cfg, sch, phys, nsteps, 4, "fluid", "bgk")
    assert abs(cfg.get("omega") / (1.0 / (3.0 * sch.visc[0] + 0.5)) - 1.0) < 1e-15
    assert cfg.get("scaling") == "acoustic"
    # initial_condition.pressure is a Lua function: call it at every barycentre
    p = np.array([cfg.call("initial_condition.pressure", b[0], b[1], b[2], 0.0) for b in bary[:sch.ld.nFluid]])
    rho = p * 3.0 * (1.0 / phys.fac_press)
    got = sch.aux.reshape(-1, 4)[:sch.ld.nFluid]
    assert np.max(np.abs(got[:, 0] / rho - 1.0)) < 2e-15 and np.all(got[:, 1:] == 0.0)   # aux = moments of f_eq(rho, 0)
    for k in ("velocityX", "velocityY", "velocityZ"):
        assert cfg.get("initial_condition." + k) == 0.0
    # the tracking object the golden file comes from
    t = cfg.get("tracking.1")
    assert t["label"] == "pressAlongLength" and t["output"]["format"] == "asciiSpatial"
    assert t["variable"] == ["density_phy", "pressure_phy", "velocity_phy"]
    assert t["shape"]["kind"] == "canoND" and t["shape"]["object"] == {"origin": [0.0, 5.0, 5.0], "vec": [10.0, 0.0, 0.0]}
    assert cfg.get("simulation_name") == "gaussianPulse"


def test_incompressible_gaussian_pulse_case_is_the_references_script(oracle, script):
    cfg = script(PULSE_INC)
    sch, phys, bary, nsteps = gaussian_pulse_setup(oracle, kind="fluid_incompressible")
    _check_common(cfg, sch, phys, nsteps, 4, "fluid_incompressible", "bgk")
    ic = cfg.get("initial_condition.pressure")      # predefined = 'gausspulse' (tem_ic_predefs_module.f90:230-255)
    assert ic == {"predefined": "gausspulse", "center": [5.0, 5.0, 5.0], "halfwidth": 1.0, "amplitude": 1.2,
                  "background": 1.0 * 343.0 ** 2}
    b = bary[:sch.ld.nFluid]
    r = (b[:, 0] - ic["center"][0]) ** 2 + (b[:, 1] - ic["center"][1]) ** 2 + (b[:, 2] - ic["center"][2]) ** 2
    p = ic["background"] + ic["amplitude"] * np.exp((-math.log(2.0) / ic["halfwidth"] ** 2) * r)
    assert np.max(np.abs(sch.aux.reshape(-1, 4)[:sch.ld.nFluid, 0] / (p * 3.0 * (1.0 / phys.fac_press)) - 1.0)) < 2e-15


@pytest.mark.parametrize("Re", [800, 1600])
def test_taylor_green_cases_are_the_references_scripts(oracle, script, Re):
    cfg = script(TGV % Re)
    if Re == 800:
        sch, phys, probe, nsteps, omega_bulk = tgv800_setup(oracle)
        level, relaxation = 6, "mrt"
    else:
        sch, phys, nsteps = tgv1600_setup(oracle)
        level, relaxation = 7, "bgk"
    _check_common(cfg, sch, phys, nsteps, level, "fluid_incompressible", relaxation)
    assert cfg.get("Re") == Re
    o = cfg.get("mesh.origin")
    bary = oracle.barycenters(sch.ld, tuple(o), cfg.get("mesh.length"))[:sch.ld.nFluid]
    if Re == 800:
        nu_bulk_lat = cfg.get("fluid.bulk_viscosity") / (phys.dx ** 2 / phys.dt)
        assert abs(omega_bulk / oracle.lib().ora_omega_bulk(nu_bulk_lat) - 1.0) < 1e-15
        t = cfg.get("tracking.1")
        assert t["label"] == "probeAtCenter" and t["variable"] == ["velocity_phy", "pressure_phy"]
        assert np.allclose(bary[probe] - 0.5 * phys.dx, t["shape"]["object"]["origin"], rtol=0, atol=1e-12)
    # the initial-condition functions at a sample of the barycentres (every 37th element)
    aux = sch.aux.reshape(-1, 4)[:sch.ld.nFluid]
    sel = np.arange(0, sch.ld.nFluid, 37)
    ic = cfg.get("initial_condition")
    assert all(ic[k] == "<function>" for k in ("pressure", "velocityX", "velocityY"))
    for e in sel:
        x, y, z = bary[e]
        p = cfg.call("initial_condition.pressure", x, y, z, 0.0)
        vx = cfg.call("initial_condition.velocityX", x, y, z, 0.0)
        vy = cfg.call("initial_condition.velocityY", x, y, z, 0.0)
        vz = cfg.call("initial_condition.velocityZ", x, y, z, 0.0)
        assert abs(aux[e, 0] / (p * 3.0 / phys.fac_press) - 1.0) < 2e-15
        # aux = moments of the initial PDFs; the lattice density is p0 / fac_press ~ 950 here, so the
        # PDFs are O(50) and their first moment carries 1e-14 of rounding
        assert abs(aux[e, 1] - vx / phys.fac_vel) < 1e-13 and abs(aux[e, 2] - vy / phys.fac_vel) < 1e-13
        assert vz == 0 and abs(aux[e, 3]) < 1e-13


def test_tutorial_gaussian_pulse_case_is_the_references_script(oracle, script):
    cfg = script(TUTORIAL)
    sch, probe, nsteps = tutorial_pulse_setup(oracle)
    assert cfg.get("mesh.predefined") == "cube" and cfg.get("mesh.refinementLevel") == 6 == sch.ld.level
    assert cfg.get("mesh.origin") == [0.0, 0.0, 0.0] and cfg.get("mesh.length") == 10.0
    assert cfg.get("identify.kind") == "fluid" and cfg.get("identify.layout") == "d3q19"
    assert cfg.get("identify.relaxation") == "bgk"
    assert cfg.get("physics") is None                   # lattice units: every conversion factor is 1
    assert cfg.get("fluid.kinematic_viscosity") == sch.visc[0] == 0.03
    assert cfg.get("sim_control.time_control.max.iter") == nsteps == 50
    bary = oracle.barycenters(sch.ld, (0.0, 0.0, 0.0), 10.0)[:sch.ld.nFluid]
    aux = sch.aux.reshape(-1, 4)[:sch.ld.nFluid]
    for e in list(range(0, sch.ld.nFluid, 997)) + [probe]:
        p = cfg.call("initial_condition.pressure", bary[e, 0], bary[e, 1], bary[e, 2])
        assert abs(aux[e, 0] / (p * 3.0) - 1.0) < 2e-15 and np.all(aux[e, 1:] == 0.0)
    for k in ("velocityX", "velocityY", "velocityZ"):
        assert cfg.get("initial_condition." + k) == 0.0
    t = cfg.get("tracking")
    assert t["label"] == "track_pressure" and t["variable"] == ["density", "pressure", "velocity"]
    assert t["output"] == {"format": "ascii"} and t["shape"]["object"] == {"origin": [1.0, 1.0, 1.0]}
    assert t["time_control"]["interval"] == {"iter": 1} and t["time_control"]["min"] == {"iter": 1}
    dx = 10.0 / 64
    assert np.all(np.abs(bary[probe] - 1.0) <= 0.5 * dx)       # the probe element holds the point (1, 1, 1)
    assert cfg.get("simulation_name") == "Gausspulse"


def test_lua_bridge_basics():
    s = lua_ref.LuaScript(text="a = { 1, 2, { x = 'y' } }  function f(x, y) return { x + y, x * y } end  b = 2^0.5")
    assert s.get("a") == [1, 2, {"x": "y"}] and s.get("a.3.x") == "y" and s.get("nope.deeper") is None
    assert s.call("f", 2, 3) == [5.0, 6.0] and s.get("b") == math.sqrt(2.0)
    with pytest.raises(TypeError):
        s.call("a")
    s.close()
    with pytest.raises(RuntimeError):
        lua_ref.LuaScript(text="x = = 1")
