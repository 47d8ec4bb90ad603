"""BASELINE config 5: a passive scalar transported by a flow on a multi-level mesh.  The reference
has no such run (it aborts for a passive scalar on a multi-level mesh and holds one scheme per
process), so parity is oracle-only: the extension is assembled from reference pieces -- the
recursive schedule with both schemes advancing inside every level step, the scalar's kernels, and
the reference's interpolation of arbitrary values (fillArbi*) applied to the scalar's PDFs.
CPU: properties of the oracle's coupled scheme.  GPU: musb200_step_schemes against it, bit for bit,
fluid and ghost elements of both schemes on every level."""
import numpy as np
import pytest

from test_multilevel import OMEGA_MIN, build


def _coupled(mo, boxes, method, relax="bgk", variant="first", QQ=19, scalar="blob"):
    lv, intp, tables, ms = build(mo, 4, boxes, QQ, method, omega_min=OMEGA_MIN[len(boxes)])
    cp = mo.CoupledMultiLevel(ms, relax, variant, diff_coeff_min=0.02, lambda_=0.2)
    for l, p in cp.ps.items():
        x = lv[l].bary_unit
        if scalar == "uniform":
            rho = np.full(lv[l].nElems, 1.25)
        else:
            r2 = ((x - np.array([0.45, 0.5, 0.55])) ** 2).sum(axis=1)
            rho = 1.0 + 0.5 * np.exp(-r2 / 0.02)
        p.init_equilibrium(rho)
    return lv, intp, tables, ms, cp


@pytest.mark.parametrize("boxes,method", [([(5, 11)], "linear"), ([(4, 12), (12, 20)], "quadratic")])
def test_uniform_scalar_in_a_fluid_at_rest_is_a_fixed_point(oracle, boxes, method):
    """a constant scalar in a fluid at rest stays constant on every level, ghosts included: the
    collision leaves f = w * rho, streaming and every interpolation (average, weighted average,
    least-square fit) reproduce constants"""
    lv, intp, tables, ms, cp = _coupled(oracle, boxes, method, scalar="uniform")
    for l, s in ms.s.items():
        s.init_equilibrium(np.ones(lv[l].nElems), np.zeros((lv[l].nElems, 3)))
    cp.run(6)
    w = oracle.weights(19)
    for l, p in cp.ps.items():
        f = p.state[p.nNext][:lv[l].nElems * 19].reshape(-1, 19)
        assert np.max(np.abs(f - 1.25 * w[None, :])) < 1e-13


def test_scalar_blob_is_transported_and_nearly_conserved(oracle):
    lv, intp, tables, ms, cp = _coupled(oracle, [(5, 11)], "linear")
    m0 = cp.scalar_mass()
    cp.run(20)
    m1 = cp.scalar_mass()
    assert abs(m1 / m0 - 1.0) < 2e-3                    # interpolated ghosts are not conservative
    for l, p in cp.ps.items():
        assert np.isfinite(p.state[p.nNext][:lv[l].nElems * 19]).all()
    # zero velocity and zero diffusion gradient across the interface: a second run with the
    # scalar's first moment shows the blob has moved with the mean flow (u = (0.02, -0.01, 0.015))
    x = lv[4].bary_unit[:lv[4].nFluid]
    rho = cp.ps[4].state[cp.ps[4].nNext][:lv[4].nFluid * 19].reshape(-1, 19).sum(axis=1)
    assert rho.max() > 1.0 + 1e-3


@pytest.fixture(scope="module")
def mbgpu():
    import musubi_b200
    musubi_b200.mus_init(0, 1, 0)
    yield musubi_b200
    musubi_b200.mus_finalize()


CASES = [([(5, 11)], "linear", "bgk", "first"), ([(5, 11)], "quadratic", "bgk", "second"),
         ([(5, 11)], "weighted_average", "trt", "standard"), ([(4, 12), (12, 20)], "linear", "bgk", "first")]


@pytest.mark.gpu
@pytest.mark.parametrize("boxes,method,relax,variant", CASES,
                         ids=["2lvl-linear-bgk1", "2lvl-quad-bgk2", "2lvl-wavg-trt", "3lvl-linear-bgk1"])
def test_coupled_flow_and_scalar_on_device_match_oracle(mbgpu, oracle, boxes, method, relax, variant):
    from musubi_b200._lib import check, lib
    mb, mo, QQ = mbgpu, oracle, 19
    lv, intp, tables, ms, cp = _coupled(mo, boxes, method, relax, variant)
    omega = {l: float(1.0 / (3.0 * s.visc[0] + 0.5)) for l, s in ms.s.items()}
    visc = {l: float(s.visc[0]) for l, s in ms.s.items()}
    flow = mb.Scheme({"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, lv, omega, omega_bulk=1.2,
                     intp=(tables, intp["order"]), viscosity=visc, slot=0)
    rel = {"name": relax, "variant": variant} if relax == "bgk" else relax
    ps = mb.Scheme({"kind": "passive_scalar", "relaxation": rel, "layout": "d3q19"}, lv,
                   species={"diff_coeff": cp.diff, "lambda": 0.2}, intp=(tables, intp["order"]), slot=1)
    for l, s in ms.s.items():
        flow.upload_state(l, s.state[s.nNow], s.state[s.nNext])
        flow._bind()
        check(lib.musb200_aux_upload(l, s.aux.ctypes.data))
        p = cp.ps[l]
        ps.upload_state(l, p.state[p.nNow], p.state[p.nNext])
        ps.couple_transport_velocity(l, flow)
    ncyc = 11            # >= 8: the coupled pair goes through CUDA-graph replay as well
    mb.step_schemes([flow, ps], ncyc)
    cp.run(ncyc)
    for l in sorted(lv):
        n = lv[l].nElems * QQ
        s, p = ms.s[l], cp.ps[l]
        assert np.isfinite(p.state[p.nNext][:n]).all()
        assert np.array_equal(flow.download_state(l)[:n], s.state[s.nNext][:n]), "flow, level %d" % l
        assert np.array_equal(ps.download_state(l)[:n], p.state[p.nNext][:n]), "scalar, level %d" % l
    ps.destroy()
    flow.destroy()
