"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the
same seeded inputs.  Tolerances: north_star asks 1e-10 relative on PDFs, density
and velocity after N steps and 1e-13 on total mass; libmusb200 is built with
-fmad=false, so on periodic/cavity single-level cases the PDFs are expected to be
BIT-IDENTICAL to the oracle (asserted where it holds)."""
import ctypes

import numpy as np
import pytest

from helpers import make_pair, rel_diff

pytestmark = pytest.mark.gpu

KERNELS = [
    ({"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, 1.7),
    ({"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}, 1.7),
    ({"kind": "fluid", "relaxation": "mrt", "layout": "d3q19"}, 1.8),
    ({"kind": "fluid", "relaxation": "bgk", "layout": "d3q27"}, 1.6),
    ({"kind": "fluid", "relaxation": "trt", "layout": "d3q27"}, 1.7),
    ({"kind": "fluid", "relaxation": "mrt", "layout": "d3q27"}, 1.9),
    ({"kind": "fluid_incompressible", "relaxation": "bgk", "layout": "d3q19"}, 1.7),
    ({"kind": "fluid_incompressible", "relaxation": "trt", "layout": "d3q19"}, 1.7),
    ({"kind": "fluid_incompressible", "relaxation": "mrt", "layout": "d3q19"}, 1.8),
    ({"kind": "fluid_incompressible", "relaxation": "bgk", "layout": "d3q27"}, 1.6),
    ({"kind": "fluid_incompressible", "relaxation": "mrt", "layout": "d3q27"}, 1.9),
]


@pytest.fixture(scope="module")
def mb():
    import musubi_b200
    musubi_b200.mus_init(0, 1, 0)
    yield musubi_b200
    musubi_b200.mus_finalize()


@pytest.mark.parametrize("ident,omega", KERNELS, ids=lambda p: "-".join(p.values()) if isinstance(p, dict) else str(p))
def test_periodic_tgv_matches_oracle(mb, oracle, ident, omega):
    level, nsteps = 4, 100
    ld, old, ref, sch = make_pair(mb, oracle, level, ident, omega, omega_bulk=1.2)
    QQ = ld.QQ
    m0, _, nan0 = sch.reduce()
    assert nan0 == 0
    assert abs(m0 / ref.total_mass() - 1.0) < 1e-12   # different (tree) summation order
    sch.do_computation(nsteps)
    ref.run(nsteps)
    got = sch.download_state(level)
    exp = ref.state[ref.nNext]
    n = ld.nFluid * QQ
    assert rel_diff(got[:n], exp[:n]) < 1e-10          # the contract
    assert np.array_equal(got[:n], exp[:n])             # what -fmad=false delivers
    aux = sch.download_aux(level)
    assert rel_diff(aux[:ld.nFluid * 4].reshape(-1, 4)[:, 0], ref.aux[:ld.nFluid * 4].reshape(-1, 4)[:, 0]) < 1e-10
    assert np.max(np.abs(aux[:ld.nFluid * 4] - ref.aux[:ld.nFluid * 4])) < 1e-12
    m1, vmax, nan1 = sch.reduce()
    assert nan1 == 0 and 0.0 < vmax < 0.2
    assert abs(m1 / m0 - 1.0) < 1e-13                   # mass conservation over 100 steps
    sch.destroy()


def test_neighbour_list_bit_exact(mb, oracle):
    for QQ, kind in ((19, "periodic"), (27, "periodic"), (19, "cavity"), (27, "cavity")):
        ident = {"kind": "fluid", "relaxation": "bgk", "layout": "d3q%d" % QQ}
        ld, old, ref, sch = make_pair(mb, oracle, 4, ident, 1.5, kind=kind, ic="rest")
        got = sch.download_neigh(4)
        assert np.array_equal(got, old.neigh)
        sch.destroy()


@pytest.mark.parametrize("ident,omega", [KERNELS[1], KERNELS[5], KERNELS[6], KERNELS[8], KERNELS[10]],
                         ids=["trt-d3q19", "mrt-d3q27", "bgk-d3q19-incomp", "mrt-d3q19-incomp",
                              "mrt-d3q27-incomp"])
def test_lid_driven_cavity_matches_oracle(mb, oracle, ident, omega):
    level, nsteps = 4, 150
    ld, old, ref, sch = make_pair(mb, oracle, level, ident, omega, kind="cavity", ic="rest",
                                  lambda_=3.0 / 16.0, omega_bulk=1.1)
    sch.do_computation(nsteps)
    ref.run(nsteps)
    got = sch.download_state(level)
    exp = ref.state[ref.nNext]
    n = ld.nFluid * ld.QQ
    assert rel_diff(got[:n], exp[:n]) < 1e-10
    assert np.array_equal(got[:n], exp[:n])
    aux = sch.download_aux(level)[:ld.nFluid * 4].reshape(-1, 4)
    # the lid drives a flow: momentum entered the box
    assert np.abs(aux[:, 1]).max() > 1e-3
    sch.destroy()


@pytest.mark.parametrize("fused", [0, 1], ids=["two-phase-bcBuffer", "fused-per-element"])
def test_velocity_bounceback_paths_agree(mb, oracle, fused):
    """fill_bcBuffer + link loop (the reference's two phases) and the one-kernel-per-boundary form
    of velocity_bounceback both reproduce the oracle bit for bit"""
    from musubi_b200._lib import check, lib
    check(lib.musb200_set_fused_bc(fused))
    try:
        ident = {"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}
        level, nsteps = 4, 60
        ld, old, ref, sch = make_pair(mb, oracle, level, ident, 1.7, kind="cavity", ic="rest",
                                      lambda_=3.0 / 16.0)
        nl = ctypes.c_longlong()
        check(lib.musb200_timers_reset())
        sch.do_computation(nsteps)
        check(lib.musb200_launch_count(ctypes.byref(nl)))
        assert nl.value == nsteps * (2 if fused else 3)     # BC kernels + sweep per step
        ref.run(nsteps)
        n = ld.nFluid * ld.QQ
        assert np.array_equal(sch.download_state(level)[:n], ref.state[ref.nNext][:n])
        sch.destroy()
    finally:
        check(lib.musb200_set_fused_bc(1))


@pytest.mark.parametrize("graphs", [0, 1], ids=["direct-launches", "cuda-graph"])
def test_step_graph_replay_equals_direct_launches(mb, oracle, graphs):
    """a long musb200_step call replays a CUDA graph of two cycles; odd and even cycle counts,
    a changed boundary value in between (re-capture), all against the oracle"""
    from musubi_b200._lib import check, lib
    check(lib.musb200_set_graphs(graphs))
    try:
        ident = {"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}
        level = 4
        ld, old, ref, sch = make_pair(mb, oracle, level, ident, 1.7, kind="cavity", ic="rest",
                                      lambda_=3.0 / 16.0)
        nl = ctypes.c_longlong()
        check(lib.musb200_timers_reset())
        sch.do_computation(31)
        sch.do_computation(20)
        check(lib.musb200_launch_count(ctypes.byref(nl)))
        assert nl.value == 51 * 2
        ref.run(51)
        v = 2.0 * ref.bc_vel[2]
        ref.bc_vel[2] = v
        sch.set_bc_values(level, 2, v)
        sch.do_computation(9)
        ref.run(9)
        n = ld.nFluid * ld.QQ
        assert np.array_equal(sch.download_state(level)[:n], ref.state[ref.nNext][:n])
        assert np.array_equal(sch.download_aux(level)[:ld.nFluid * 4], ref.aux[:ld.nFluid * 4])
        sch.destroy()
    finally:
        check(lib.musb200_set_graphs(1))


@pytest.mark.parametrize("force", [False, True], ids=["plain", "with-force"])
def test_lazy_auxfield_on_demand_equals_the_sweeps(mb, oracle, force):
    """musb200_set_aux_every_step(2): nothing is materialised by the sweep; the probe of one
    element and the full download are computed on demand from state(:, now) and equal the
    oracle's auxField of the last step bit for bit"""
    from musubi_b200._lib import check, lib
    ident = {"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}
    level = 4
    ld, old, ref, sch = make_pair(mb, oracle, level, ident, 1.7, kind="cavity", ic="rest", lambda_=3.0 / 16.0)
    if force:
        ref.set_force([1e-5, -2e-5, 3e-6])
        sch.set_force(level, [1e-5, -2e-5, 3e-6])
    check(lib.musb200_set_aux_every_step(2))
    try:
        for k in (3, 1, 12):                      # 12 >= 8: the graph path
            sch.do_computation(k)
            ref.run(k)
            exp = ref.aux[:ld.nFluid * 4].reshape(-1, 4)
            for e in (1, 77, ld.nFluid):
                assert np.array_equal(sch.aux_probe(level, e), exp[e - 1])
        assert np.array_equal(sch.download_aux(level)[:ld.nFluid * 4], ref.aux[:ld.nFluid * 4])
        n = ld.nFluid * ld.QQ
        assert np.array_equal(sch.download_state(level)[:n], ref.state[ref.nNext][:n])
    finally:
        check(lib.musb200_set_aux_every_step(0))
        sch.destroy()


CHANNEL = [
    ({"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, "pressure_expol"),
    ({"kind": "fluid", "relaxation": "mrt", "layout": "d3q27"}, "pressure_antibounceback"),
    ({"kind": "fluid_incompressible", "relaxation": "trt", "layout": "d3q19"}, "pressure_antibounceback"),
    ({"kind": "fluid", "relaxation": "bgk", "layout": "d3q27"}, "pressure_expol"),
    ({"kind": "fluid_incompressible", "relaxation": "mrt", "layout": "d3q19"}, "pressure_expol"),
]


@pytest.mark.parametrize("ident,outlet", CHANNEL,
                         ids=["bgk19-expol", "mrt27-antibb", "trt19incomp-antibb", "bgk27-expol",
                              "mrt19incomp-expol"])
def test_channel_with_pressure_outlet_matches_oracle(mb, oracle, ident, outlet):
    """walls + velocity_bounceback inlet + pressure outlet (a12): 200 steps, bit-exact"""
    level, nsteps = 4, 200
    ld, old, ref, sch = make_pair(mb, oracle, level, ident, 1.6, kind="channel", ic="rest",
                                  omega_bulk=1.2, u_lid=(0.03, 0.0, 0.0), outlet=outlet, rho_out=1.0)
    sch.do_computation(nsteps)
    ref.run(nsteps)
    got = sch.download_state(level)
    exp = ref.state[ref.nNext]
    n = ld.nFluid * ld.QQ
    assert np.isfinite(exp[:n]).all()
    assert rel_diff(got[:n], exp[:n]) < 1e-10
    assert np.array_equal(got[:n], exp[:n])
    aux = sch.download_aux(level)[:ld.nFluid * 4].reshape(-1, 4)
    assert np.array_equal(aux, ref.aux[:ld.nFluid * 4].reshape(-1, 4))
    assert aux[:, 1].mean() > 0.02          # the inlet drives a through-flow
    sch.destroy()


def test_random_omega_per_element(mb, oracle):
    """per-element omega (viscosity spacetime function), fixed seed 12345, omega in [0.5, 1.95]."""
    ident = {"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}
    ld, old, ref, sch = make_pair(mb, oracle, 4, ident, 1.0)
    rng = np.random.default_rng(12345)
    om = rng.uniform(0.5, 1.95, ld.nSize)
    ref.visc[:] = (1.0 / om - 0.5) / 3.0
    oracle.lib().ora_update_omega(ref.omega.ctypes.data_as(oracle._dp), ref.visc.ctypes.data_as(oracle._dp), ld.nSolve)
    sch.set_relaxation(4, ref.omega[:ld.nSolve].copy(), 1.0)
    sch.do_computation(30)
    ref.run(30)
    got = sch.download_state(4)
    n = ld.nFluid * 19
    assert np.array_equal(got[:n], ref.state[ref.nNext][:n])
    sch.destroy()


def test_compute_host_single_element_like_reference_utest(mb, oracle):
    """mus/utests/mus_bgk_d3q19_compare_test.f90: one element, neigh = [1..QQ], random PDFs,
    optimised kernel vs the generic NoOpt kernel, tolerance 2500*eps."""
    import ctypes
    rng = np.random.default_rng(7)
    for relax, QQ in (("bgk", 19), ("mrt", 19), ("bgk", 27), ("mrt", 27)):
        nElems = 4
        neigh = np.zeros(QQ * nElems, dtype=np.int32)
        for d in range(QQ):
            for e in range(nElems):
                neigh[d * nElems + e] = e * QQ + d + 1
        w = oracle.weights(QQ)
        f = (w[None, :] * (1.0 + 0.05 * rng.standard_normal((nElems, QQ)))).ravel()
        omega = np.full(nElems, 1.7)
        out, aux = mb.compute_host({"kind": "fluid", "relaxation": relax, "layout": "d3q%d" % QQ},
                                   f, neigh, nElems, 1, omega, omega_bulk=1.3)
        ref = np.zeros_like(f)
        rp = oracle._Relax(0.25, 1.3)
        rc = oracle.lib().ora_compute_noopt(oracle.RELAX[relax], QQ, oracle._d(f), oracle._d(ref),
                                            oracle._d(aux), oracle._i(neigh), oracle._d(omega),
                                            nElems, 1, ctypes.byref(rp))
        assert rc == 0
        assert np.max(np.abs(out[:QQ] - ref[:QQ])) < 2500 * np.finfo(float).eps
        assert abs(out[:QQ].sum() - f[:QQ].sum()) < 1e-14
