"""Properties the reference's own unit tests check (mus/utests), restated
against the oracle, for every kernel on the path including the fluid_incompressible ones:

  mus_bgk_d3q19_compare_test.f90, mus_mrt_d3q19_test.f90   optimised vs NoOpt kernel, tol 2500*eps
  mus_bgk_d3q19_weights_test.f90, mus_trt_d3q27_weights_test.f90   rest state is a fixed point
  mus_mrt_matrix_D3Q19_test.f90 / _D3Q27_test.f90          M * M^-1 = I
  mus_fNeq_acoustic_test.f90                               f_neq from the strain rate
plus conservation of mass and momentum by the collision and the TRT -> BGK limit."""
import ctypes

import numpy as np
import pytest

EPS = np.finfo(float).eps
ALL = [(r, q, k) for k in (0, 1) for q in (19, 27) for r in ("bgk", "trt", "mrt")
       if not (k == 1 and r == "trt" and q == 27)]     # the reference has no such kernel


def ident_neigh(QQ, n):
    ng = np.zeros(QQ * n, dtype=np.int32)
    for d in range(QQ):
        ng[d * n:(d + 1) * n] = np.arange(n) * QQ + d + 1
    return ng


def run_kernel(mo, relax, QQ, incomp, f, omega, lam=0.25, omega_bulk=1.3, noopt=False):
    n = f.size // QQ
    ng = ident_neigh(QQ, n)
    aux = np.zeros(n * 4)
    (mo.lib().ora_calc_aux_incomp if incomp else mo.lib().ora_calc_aux)(QQ, mo._d(aux), mo._d(f), mo._i(ng), n, n)
    out = np.zeros_like(f)
    om = np.full(n, float(omega))
    rp = mo._Relax(lam, omega_bulk)
    if noopt:
        rc = mo.lib().ora_compute_noopt_kind(mo.RELAX[relax], QQ, incomp, mo._d(f), mo._d(out), mo._d(aux),
                                             mo._i(ng), mo._d(om), n, n, ctypes.byref(rp))
    else:
        rc = mo.lib().ora_compute(mo.RELAX[relax], QQ, incomp, mo._d(f), mo._d(out), mo._d(aux),
                                  mo._i(ng), mo._d(om), n, n, ctypes.byref(rp))
    assert rc == 0
    return out.reshape(n, QQ), aux.reshape(n, 4)


def random_pdfs(mo, QQ, n, seed):
    rng = np.random.default_rng(seed)
    w = mo.weights(QQ)
    return (w[None, :] * (1.0 + 0.05 * rng.standard_normal((n, QQ)))).ravel()


@pytest.mark.parametrize("relax,QQ,incomp", [c for c in ALL if c[0] != "trt"])
def test_optimised_kernel_equals_noopt_kernel(oracle, relax, QQ, incomp):
    f = random_pdfs(oracle, QQ, 64, 11 + QQ + incomp)
    a, _ = run_kernel(oracle, relax, QQ, incomp, f, 1.7)
    b, _ = run_kernel(oracle, relax, QQ, incomp, f, 1.7, noopt=True)
    assert np.max(np.abs(a - b)) < 2500 * EPS


@pytest.mark.parametrize("relax,QQ,incomp", ALL)
def test_rest_state_is_a_fixed_point(oracle, relax, QQ, incomp):
    w = oracle.weights(QQ)
    f = np.tile(w, 8)
    out, aux = run_kernel(oracle, relax, QQ, incomp, f, 1.8)
    assert np.max(np.abs(out - w[None, :])) < 4 * EPS
    assert np.max(np.abs(aux[:, 0] - 1.0)) < 4 * EPS and np.max(np.abs(aux[:, 1:])) < 4 * EPS


@pytest.mark.parametrize("relax,QQ,incomp", ALL)
def test_collision_conserves_mass_and_momentum(oracle, relax, QQ, incomp):
    f = random_pdfs(oracle, QQ, 256, 5)
    out, _ = run_kernel(oracle, relax, QQ, incomp, f, 1.6)
    cx = oracle.cx_dir(QQ).astype(np.float64)
    fin = f.reshape(-1, QQ)
    assert np.max(np.abs(out.sum(axis=1) - fin.sum(axis=1))) < 20 * EPS
    assert np.max(np.abs(out @ cx - fin @ cx)) < 20 * EPS


@pytest.mark.parametrize("incomp", [0, 1])
def test_trt_d3q19_with_equal_rates_is_bgk(oracle, incomp):
    """lambda = (1/omega - 1/2)^2 makes omega^- = omega^+ = omega"""
    omega = 1.45
    lam = (1.0 / omega - 0.5) ** 2
    f = random_pdfs(oracle, 19, 128, 3)
    a, _ = run_kernel(oracle, "trt", 19, incomp, f, omega, lam=lam)
    b, _ = run_kernel(oracle, "bgk", 19, incomp, f, omega)
    assert np.max(np.abs(a - b)) < 50 * EPS


@pytest.mark.parametrize("QQ", [19, 27])
def test_mrt_matrix_times_inverse_is_identity(oracle, QQ):
    L = oracle.lib()
    M = np.ctypeslib.as_array(L.ora_mrt_matrix(QQ, 0), shape=(QQ, QQ))
    Mi = np.ctypeslib.as_array(L.ora_mrt_matrix(QQ, 1), shape=(QQ, QQ))
    assert np.max(np.abs(M @ Mi - np.eye(QQ))) < 1e-14
    assert np.max(np.abs(Mi @ M - np.eye(QQ))) < 1e-14


@pytest.mark.parametrize("QQ", [19, 27])
def test_fneq_acoustic_moments(oracle, QQ):
    """getNEq_acoustic: zero mass and momentum, and its stress moment returns the strain rate:
    sum_i c_ia c_ib f_neq_i (pre-collision) = -2 rho0 nu/(cs2 omega ...) -> checked as symmetry and
    proportionality  Pi_ab = -(2/(3 omega)) S_ab * (pre-collision), traceless S."""
    rng = np.random.default_rng(2)
    S = rng.standard_normal((3, 3))
    S = 0.5 * (S + S.T)
    S -= np.eye(3) * np.trace(S) / 3.0
    omega = 1.3
    nEq = np.zeros(QQ)
    Sf = np.ascontiguousarray(S.T.ravel())
    oracle.lib().ora_nEq_acoustic(QQ, omega, oracle._d(Sf), oracle._d(nEq))
    cx = oracle.cx_dir(QQ).astype(np.float64)
    assert abs(nEq.sum()) < 1e-15 and np.max(np.abs(nEq @ cx)) < 1e-15
    pre = nEq / (1.0 - omega)                    # back to pre-collision (convPrePost, PULL build)
    Pi = np.einsum("i,ia,ib->ab", pre, cx, cx)
    nu = (1.0 / omega - 0.5) / 3.0
    expect = -(2.0 * nu * S) * 2.0 / (2.0 - omega)      # -tau * cs4inv/(2-omega) * sum_i w_i Q_iab Q_icd
    assert np.max(np.abs(Pi - expect)) < 1e-14
