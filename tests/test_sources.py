"""Source terms and passive scalar ("next" rows n3 of SURVEY.md section 8f).

CPU part: analytic properties of the oracle restatement (the reference has no fixture for these
routines that is reproducible without Seeder: "parity unpinned by reference fixtures").
GPU part: the fused device kernels against the oracle, bit for bit."""
import math

import numpy as np
import pytest

KERNELS = [(r, q, k) for k in ("fluid", "fluid_incompressible") for q in (19, 27)
           for r in ("bgk", "trt", "mrt") if not (k == "fluid_incompressible" and r == "trt" and q == 27)]
F = np.array([3.0e-6, -2.0e-6, 1.0e-6])


def _periodic(mo, level, QQ, relax, kind, omega=1.6):
    ld = mo.build_level_desc(level, QQ, "periodic")
    s = mo.Scheme(ld, relax, kind, omega=omega, lambda_=0.25, omega_bulk=1.1)
    return ld, s


@pytest.mark.parametrize("relax,QQ,kind", KERNELS)
@pytest.mark.parametrize("order", [1, 2])
def test_uniform_force_accelerates_uniformly(oracle, relax, QQ, kind, order):
    """fluid at rest in a periodic box under a constant force: the momentum of every cell grows
    by exactly F per step (sum_i c_i S_i = F, sum_i S_i = 0); the auxField velocity carries the
    half-force shift of the second-order scheme."""
    ld, s = _periodic(oracle, 3, QQ, relax, kind)
    s.init_equilibrium(1.0, np.zeros(3))
    s.set_force(F, order=order)
    m0 = math.fsum(s.state[s.nNext][:ld.nFluid * QQ])
    n = 7
    s.run(n)
    c = oracle.cx_dir(QQ).astype(float)
    f = s.state[s.nNext][:ld.nFluid * QQ].reshape(ld.nFluid, QQ)
    mom = f @ c
    # momentum gained per step: F, except for trt with the second-order source, where the
    # reference combines the bgk prefactor (1 - omega/2) with an odd part relaxing at omega^-:
    # (1 - omega/2 + omega^-/2) F  (applySrc_force is selected for bgk AND trt,
    # mus_variable_module.f90:1047-1067)
    g = 1.0
    if relax == "trt" and order == 2:
        om = 1.6
        g = 1.0 - om / 2 + 0.5 / (0.5 + 0.25 / (1.0 / om - 0.5))
    assert np.max(np.abs(mom - n * g * F[None, :])) < 1e-15
    assert abs(math.fsum(f.ravel()) - m0) < 1e-13 * m0
    aux = s.aux[:ld.nFluid * 4].reshape(ld.nFluid, 4)
    rho = aux[:, :1] if kind == "fluid" else 1.0
    shift = 0.5 if order == 2 else 0.0
    assert np.max(np.abs(aux[:, 1:] * rho - ((n - 1) * g + shift) * F[None, :])) < 1e-15


@pytest.mark.parametrize("QQ", [19, 27])
def test_force_source_adds_no_mass_and_the_expected_momentum(oracle, QQ):
    """on a moving fluid applySrc_force adds (1 - omega/2) F of momentum (the collision with the
    shifted velocity supplies the rest), applySrc_force_MRT_* adds F (the momentum moments are
    not relaxed: s = 0); neither adds mass."""
    ld, a = _periodic(oracle, 3, QQ, "mrt", "fluid")
    _, b = _periodic(oracle, 3, QQ, "bgk", "fluid")
    rng = np.random.default_rng(5)
    vel = 0.05 * rng.standard_normal((ld.nElems, 3))
    rho = 1.0 + 0.01 * rng.standard_normal(ld.nElems)
    c = oracle.cx_dir(QQ).astype(float)
    out = []
    for s in (a, b):
        s.init_equilibrium(rho, vel)
        s.calc_aux(s.state[s.nNow])
        s.set_force(F)
        s.add_src_to_aux()
        before = s.state[s.nNext].copy()
        s.apply_source_terms()
        d = (s.state[s.nNext] - before)[:ld.nFluid * QQ].reshape(ld.nFluid, QQ)
        out.append((d.sum(axis=1), d @ c))
    for (dm, dp), g in zip(out, (1.0, 1.0 - 1.6 / 2)):
        assert np.max(np.abs(dm)) < 1e-16
        assert np.max(np.abs(dp - g * F[None, :])) < 1e-16


def test_force_on_an_element_subset_only_touches_it(oracle):
    ld, s = _periodic(oracle, 3, 19, "bgk", "fluid")
    s.init_equilibrium(1.0, np.zeros(3))
    pos = np.array([5, 17, 100], dtype=np.int32)
    s.set_force(np.tile(F, (3, 1)), posInTotal=pos)
    s.step()
    f = s.state[s.nNext][:ld.nFluid * 19].reshape(ld.nFluid, 19)
    mom = f @ oracle.cx_dir(19).astype(float)
    touched = np.zeros(ld.nFluid, bool)
    touched[pos - 1] = True
    assert np.max(np.abs(mom[~touched])) < 1e-16
    assert np.max(np.abs(mom[touched] - F[None, :])) < 1e-16


# ---- passive scalar ---------------------------------------------------------------------------
PS = [("bgk", "first"), ("bgk", "second"), ("trt", "standard")]


@pytest.mark.parametrize("relax,variant", PS)
@pytest.mark.parametrize("QQ", [19, 27])
def test_passive_scalar_conserves_mass_and_keeps_uniform_state(oracle, relax, variant, QQ):
    ld = oracle.build_level_desc(3, QQ, "periodic")
    s = oracle.PassiveScalarScheme(ld, relax, variant, diff_coeff=0.02)
    s.set_transport_velocity([0.03, -0.01, 0.02])
    rng = np.random.default_rng(3)
    s.init_equilibrium(1.0 + 0.1 * rng.random(ld.nElems))
    m0 = s.total_mass()
    s.run(10)
    assert abs(s.total_mass() - m0) < 1e-13 * m0
    u = oracle.PassiveScalarScheme(ld, relax, variant, diff_coeff=0.02)
    u.set_transport_velocity([0.0, 0.0, 0.0])     # at rest a uniform field is the equilibrium
    u.init_equilibrium(0.7)
    u.run(5)
    assert np.max(np.abs(u.aux[:ld.nFluid] - 0.7)) < 1e-14


@pytest.mark.parametrize("relax,variant", PS)
def test_passive_scalar_sine_wave_diffuses_and_advects(oracle, relax, variant):
    """C = 1 + a sin(k x) in a uniform flow u: amplitude decays with exp(-D k^2 t), D = diff_coeff
    (d_omega = 2/(1+6D): tau = 1/2 + 3D), and the phase moves with u t."""
    level, D, ux, n = 5, 0.05, 0.04, 100
    ld = oracle.build_level_desc(level, 19, "periodic")
    N = 1 << level
    x = oracle.barycenters(ld, (0.0, 0.0, 0.0), float(N))[:, 0]
    k = 2.0 * np.pi / N
    s = oracle.PassiveScalarScheme(ld, relax, variant, diff_coeff=D)
    s.set_transport_velocity([ux, 0.0, 0.0])
    s.init_equilibrium(1.0 + 0.1 * np.sin(k * x), np.tile([ux, 0.0, 0.0], (ld.nElems, 1)))
    s.run(n)
    s.step()                                       # aux = zeroth moment of the state after n steps
    C = s.aux[:ld.nFluid] - 1.0
    a_sin = 2.0 * np.mean(C * np.sin(k * x[:ld.nFluid]))
    a_cos = 2.0 * np.mean(C * np.cos(k * x[:ld.nFluid]))
    amp, phase = np.hypot(a_sin, a_cos), np.arctan2(-a_cos, a_sin)
    assert abs(amp / 0.1 - np.exp(-D * k * k * n)) < 0.01
    assert abs(phase - k * ux * n) < 0.02


# ---- device vs oracle -------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mbgpu():
    import musubi_b200 as mb
    mb.mus_init(0, 1, 0)
    yield mb
    mb.mus_finalize()


def _pair(mb, mo, level, QQ, relax, kind, omega=1.6):
    from musubi_b200 import cases
    ident = {"kind": kind, "relaxation": relax, "layout": "d3q%d" % QQ}
    ld = mb.LevelDesc(level, QQ, "periodic")
    old, ref = _periodic(mo, level, QQ, relax, kind, omega)
    rho, vel = cases.taylor_green(ld, mean=(0.01, -0.02, 0.015))
    ref.init_equilibrium(rho, vel)
    sch = mb.Scheme(ident, ld, float(1.0 / (3.0 * ref.visc[0] + 0.5)), lambda_=0.25, omega_bulk=1.1)
    sch.upload_state(level, ref.state[ref.nNow], ref.state[ref.nNext])
    return ld, ref, sch


@pytest.mark.gpu
@pytest.mark.parametrize("relax,QQ,kind", KERNELS)
@pytest.mark.parametrize("order,mode", [(2, "uniform"), (1, "uniform"), (2, "field"), (2, "subset")])
def test_force_source_device_matches_oracle(mbgpu, oracle, relax, QQ, kind, order, mode):
    level, n = 4, 20
    ld, ref, sch = _pair(mbgpu, oracle, level, QQ, relax, kind)
    rng = np.random.default_rng(17)
    if mode == "uniform":
        ref.set_force(F, order=order)
        sch.set_force(level, F, order=order)
    elif mode == "field":
        Ff = 1e-5 * rng.standard_normal((ld.nFluid, 3))
        ref.set_force(Ff, order=order)
        sch.set_force(level, Ff, order=order)
    else:
        pos = np.sort(rng.choice(ld.nFluid, ld.nFluid // 3, replace=False)).astype(np.int32) + 1
        Ff = 1e-5 * rng.standard_normal((pos.size, 3))
        ref.set_force(Ff, order=order, posInTotal=pos)
        sch.set_force(level, Ff, order=order, posInTotal=pos)
    ref.run(n)
    sch.do_computation(n)
    k = ld.nFluid * QQ
    got, exp = sch.download_state(level)[:k], ref.state[ref.nNext][:k]
    assert np.max(np.abs(got - exp) / np.abs(exp)) < 1e-10
    assert np.array_equal(got, exp)
    aux = sch.download_aux(level)[:ld.nFluid * 4]
    assert np.array_equal(aux, ref.aux[:ld.nFluid * 4])
    sch.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("relax,variant", PS)
@pytest.mark.parametrize("QQ", [19, 27])
@pytest.mark.parametrize("velocity", ["uniform", "field"])
def test_passive_scalar_device_matches_oracle(mbgpu, oracle, relax, variant, QQ, velocity):
    mb, level, n = mbgpu, 4, 25
    ident = {"kind": "passive_scalar", "relaxation": {"name": relax, "variant": variant},
             "layout": "d3q%d" % QQ}
    ld = mb.LevelDesc(level, QQ, "periodic")
    old = oracle.build_level_desc(level, QQ, "periodic")
    ref = oracle.PassiveScalarScheme(old, relax, variant, diff_coeff=0.03, lambda_=0.2)
    rng = np.random.default_rng(9)
    ref.init_equilibrium(1.0 + 0.2 * rng.random(old.nElems))
    sch = mb.Scheme(ident, ld, species={"diff_coeff": 0.03, "lambda": 0.2})
    vel = np.array([0.03, -0.02, 0.01]) if velocity == "uniform" else 0.05 * rng.standard_normal((ld.nFluid, 3))
    ref.set_transport_velocity(vel)
    sch.set_transport_velocity(level, vel)
    sch.upload_state(level, ref.state[ref.nNow], ref.state[ref.nNext])
    ref.run(n)
    sch.do_computation(n)
    k = ld.nFluid * QQ
    assert np.array_equal(sch.download_state(level)[:k], ref.state[ref.nNext][:k])
    assert np.array_equal(sch.download_aux(level)[:ld.nFluid], ref.aux[:ld.nFluid])
    sch.destroy()


@pytest.mark.gpu
def test_passive_scalar_coupled_to_flow_on_device(mbgpu, oracle):
    """BASELINE config 5's coupling on one level: the scalar's transport velocity is the flow's
    auxField velocity of the same step, read on the device (slot 1 <- slot 0)."""
    from musubi_b200._lib import check, lib
    mb, level, n, QQ = mbgpu, 4, 15, 19
    ld, fref, flow = _pair(mb, oracle, level, QQ, "bgk", "fluid")
    check(lib.musb200_set_aux_every_step(1))
    old = oracle.build_level_desc(level, QQ, "periodic")
    sref = oracle.PassiveScalarScheme(old, "bgk", "second", diff_coeff=0.02)
    rng = np.random.default_rng(2)
    sref.init_equilibrium(1.0 + 0.2 * rng.random(old.nElems))
    ps = mb.Scheme({"kind": "passive_scalar", "relaxation": {"name": "bgk", "variant": "second"},
                    "layout": "d3q19"}, ld, species={"diff_coeff": 0.02}, slot=1)
    ps.upload_state(level, sref.state[sref.nNow], sref.state[sref.nNext])
    ps.couple_transport_velocity(level, flow)
    for _ in range(n):
        fref.step()
        sref.set_transport_velocity(fref.aux[:old.nSolve * 4].reshape(-1, 4)[:, 1:])
        sref.step()
        flow.do_computation(1)
        ps.do_computation(1)
    k = ld.nFluid * QQ
    assert np.array_equal(ps.download_state(level)[:k], sref.state[sref.nNext][:k])
    assert np.array_equal(flow.download_state(level)[:k], fref.state[fref.nNext][:k])
    check(lib.musb200_set_aux_every_step(0))
    ps.destroy()
    flow.destroy()


@pytest.mark.gpu
def test_force_source_on_a_two_level_mesh_matches_oracle(mbgpu, oracle):
    """the body force acts on every level (fluid + ghostFromCoarser elements, in lattice units of
    that level: half the coarse value per finer level for a physically uniform force with acoustic
    scaling); device and oracle stay bit-identical through the interpolation schedule"""
    from test_multilevel import build
    from musubi_b200._lib import check, lib
    mb, QQ = mbgpu, 19
    lv, intp, tables, ms = build(oracle, 4, [(5, 11)], QQ, "linear")
    ident = {"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}
    omega = {l: float(1.0 / (3.0 * s.visc[0] + 0.5)) for l, s in ms.s.items()}
    visc = {l: float(s.visc[0]) for l, s in ms.s.items()}
    sch = mb.Scheme(ident, lv, omega, omega_bulk=1.2, intp=(tables, intp["order"]), viscosity=visc)
    for l, s in ms.s.items():
        sch.upload_state(l, s.state[s.nNow], s.state[s.nNext])
        check(lib.musb200_aux_upload(l, s.aux.ctypes.data))
        Fl = F * 0.5 ** (l - min(lv))            # body_force factor rho0 dx / dt^2 doubles per level
        s.set_force(Fl)
        sch.set_force(l, Fl)
    sch.do_computation(10)
    ms.run(10)
    for l, s in ms.s.items():
        n = lv[l].nElems * QQ
        assert np.array_equal(sch.download_state(l)[:n], s.state[s.nNext][:n])
    sch.destroy()


@pytest.mark.gpu
def test_passive_scalar_around_a_sphere_from_mesh_file(mbgpu, oracle, tmp_path):
    """passive scalar on a mesh that only exists as treelm files, bounce-back at the obstacle:
    device = oracle bit for bit and the scalar's mass is conserved"""
    from musubi_b200 import treelm_io as tio
    from test_treelm_io import _sphere_mesh
    mb, QQ = mbgpu, 19
    fd = tio.FileLevelDesc(_sphere_mesh(tmp_path), QQ)
    ref = oracle.PassiveScalarScheme(fd, "bgk", "second", diff_coeff=0.02)
    rng = np.random.default_rng(8)
    ref.init_equilibrium(1.0 + 0.3 * rng.random(fd.nElems))
    ref.set_transport_velocity([0.04, 0.01, -0.02])
    sch = mb.Scheme({"kind": "passive_scalar", "relaxation": {"name": "bgk", "variant": "second"},
                     "layout": "d3q19"}, fd, species={"diff_coeff": 0.02})
    sch.set_transport_velocity(fd.level, [0.04, 0.01, -0.02])
    sch.upload_state(fd.level, ref.state[ref.nNow], ref.state[ref.nNext])
    m0 = ref.total_mass()
    ref.run(30)
    sch.do_computation(30)
    k = fd.nFluid * QQ
    got = sch.download_state(fd.level)[:k]
    assert np.array_equal(got, ref.state[ref.nNext][:k])
    assert abs(math.fsum(got) / m0 - 1.0) < 1e-13
    sch.destroy()
