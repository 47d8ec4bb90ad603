"""The ghost interpolation ARITHMETIC held by a second implementation (a14 - a16): from the
barycentres of the elements alone -- no child numbers, no 0.25 * childPosition table, no
precomputed least-square matrices, no weights -- a numpy re-computation of what the reference's
routines define:
  coarse -> fine   least-square fit (numpy lstsq) of a linear / quadratic polynomial through the
                   sources' f_eq and f_neq at their geometric offsets from the parent, evaluated at
                   the ghost's offset; weighted average with w ~ prod_k (1 - |dx_k|); f_neq scaled
                   by omega_c (1 - omega_f) / (2 (1 - omega_c) omega_f)
  fine -> coarse   mean of the children's f_eq, mean of their f_neq times
                   2 omega_f (1 - omega_c) / ((1 - omega_f) omega_c)
compared with the ghosts the oracle's restatement (oracle/intp.c, driven by the dependency lists
and matrices) holds after a few cycles of a two-level run.  The source SETS are taken from the
lists (they have their own independent restatement, tests/test_oracle_dependencies.py)."""
import numpy as np
import pytest

from test_multilevel import build
from test_oracle_dense_lbm import _feq, _stencil


def _poly(order, d):
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    cols = [np.ones_like(x), x, y, z]
    if order == 2:
        cols += [x * x, y * y, z * z, x * y, y * z, z * x]
    return np.stack(cols, axis=-1)


def _wrap(d):
    return d - np.round(d)            # periodic unit cube: shortest offset


def _eq_neq(mo, QQ, state, aux, pos):
    cx, w, _ = _stencil(mo, QQ)
    a = aux.reshape(-1, 4)[pos - 1]
    f = state.reshape(-1, QQ)[pos - 1]
    fe = _feq(0, a[:, 0], a[:, 1:].T, cx, w).T
    return fe, f - fe


@pytest.mark.parametrize("QQ,method", [(19, "linear"), (19, "quadratic"), (19, "weighted_average"),
                                       (27, "linear"), (27, "quadratic")])
def test_ghost_values_equal_a_geometric_recomputation(oracle, QQ, method):
    mo = oracle
    lv, intp, tables, ms = build(mo, 4, [(5, 11)], QQ, method)
    ms.run(3)
    C, F = lv[4], lv[5]
    sc, sf = ms.s[4], ms.s[5]
    om_c = 1.0 / (3.0 * sc.visc[0] + 0.5)
    om_f = 1.0 / (3.0 * sf.visc[0] + 0.5)
    dxc = 1.0 / (1 << 4)
    # ---- coarse -> fine ---------------------------------------------------------------------
    fac_c2f = 0.5 * om_c * (1.0 - om_f) / ((1.0 - om_c) * om_f)
    st_f = sf.state[sf.nNext].reshape(-1, QQ)
    worst, nchk = 0.0, 0
    for order in range(intp["order"] + 1):
        t = tables.get((5, ("fromCoarser", order)))
        if t is None or len(t["targets"]) == 0:
            continue
        for i, tgt in enumerate(t["targets"]):
            src = t["srcPos"][t["srcOffset"][i]:t["srcOffset"][i + 1]]
            xt = F.bary_unit[tgt - 1]
            # the parent: the coarse cell that contains the ghost's barycentre
            pc = (np.floor(xt / dxc) + 0.5) * dxc
            ds = _wrap(C.bary_unit[src - 1] - pc) / dxc          # source offsets in coarse cells
            dt = _wrap(xt - pc) / dxc                            # +-0.25 per axis
            assert np.allclose(np.abs(dt), 0.25)
            fe, fn = _eq_neq(mo, QQ, sc.state[sc.nNext], sc.aux, src)
            if order == 0:
                wgt = np.prod(1.0 - np.abs(ds - dt[None, :]), axis=1)
                wgt = wgt / wgt.sum()
                te, tn = wgt @ fe, wgt @ fn
            else:
                A = _poly(order, ds)
                ce = np.linalg.lstsq(A, fe, rcond=None)[0]
                cn = np.linalg.lstsq(A, fn, rcond=None)[0]
                pt = _poly(order, dt[None, :])[0]
                te, tn = pt @ ce, pt @ cn
            exp = te + tn * fac_c2f
            got = st_f[tgt - 1]
            worst = max(worst, float(np.max(np.abs(got - exp) / np.abs(exp))))
            nchk += 1
    assert nchk == F.nGhostFromCoarser and worst < 5e-11, worst
    # ---- fine -> coarse ---------------------------------------------------------------------
    fac_f2c = 2.0 * om_f * (1.0 - om_c) / ((1.0 - om_f) * om_c)
    t = tables[(4, "fromFiner")]
    st_c = sc.state[sc.nNext].reshape(-1, QQ)
    worst = 0.0
    for i, tgt in enumerate(t["targets"]):
        src = t["srcPos"][t["srcOffset"][i]:t["srcOffset"][i + 1]]
        # geometric children: the fine elements whose barycentres lie in the coarse cell
        d = _wrap(F.bary_unit[src - 1] - C.bary_unit[tgt - 1]) / dxc
        assert np.all(np.abs(d) < 0.5) and len(src) == 8
        fe, fn = _eq_neq(mo, QQ, sf.state[sf.nNext], sf.aux, src)
        exp = fe.mean(axis=0) + fn.mean(axis=0) * fac_f2c
        worst = max(worst, float(np.max(np.abs(st_c[tgt - 1] - exp) / np.abs(exp))))
        a_exp = sf.aux.reshape(-1, 4)[src - 1].mean(axis=0)
        assert np.allclose(sc.aux.reshape(-1, 4)[tgt - 1], a_exp, rtol=1e-13, atol=1e-15)
    assert worst < 5e-12, worst
