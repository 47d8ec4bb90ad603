"""Where ncu's "sectors per request" of the sweep comes from: the element order is treelm's Morton
order, a warp of 32 consecutive elements is a 4 x 4 x 2 brick, and the gather of direction q reads
the brick shifted by -c_q.  Counting the distinct 32-byte sectors of every such request -- plus the
coalesced 4-byte neighbour words, 4 sectors per warp -- reproduces the averages ncu measured on the
B200 (profiles/r01_sweep_trt_d3q19_256.md: 8.42, profiles/r02_sweep_mrt_d3q27_256.md: 9.21), so
those figures are a property of the layout, not of the kernel's code."""
import numpy as np
import pytest


def _morton(x, y, z):
    r = np.zeros_like(x)
    for b in range(8):
        r |= ((x >> b) & 1) << (3 * b) | ((y >> b) & 1) << (3 * b + 1) | ((z >> b) & 1) << (3 * b + 2)
    return r


@pytest.mark.parametrize("QQ,ncu_value,by_len", [(19, 8.42, {0: 8.0, 1: 10.67, 2: 14.0}),
                                                 (27, 9.21, {0: 8.0, 1: 10.67, 2: 14.0, 3: 18.0})])
def test_sectors_per_request_of_morton_ordered_gathers(oracle, QQ, ncu_value, by_len):
    mo, n = oracle, 64
    ids = np.arange(n ** 3, dtype=np.int64)
    x, y, z = mo.coord_of_morton(ids)
    assert (_morton(x, y, z) == ids).all()
    cx = mo.cx_dir(QQ)
    per = np.empty(QQ)
    for q in range(QQ):
        src = _morton((x - cx[q, 0]) % n, (y - cx[q, 1]) % n, (z - cx[q, 2]) % n)
        sectors = np.sort((src // 4).reshape(-1, 32)[:2048], axis=1)          # 4 doubles per 32-byte sector
        per[q] = (1 + (np.diff(sectors, axis=1) != 0).sum(axis=1)).mean()
    for length, want in by_len.items():
        got = per[(cx ** 2).sum(axis=1) == length].mean()
        assert abs(got - want) < 0.01, (length, got)
    # all global loads of the sweep: QQ 8-byte gathers + (QQ - 1) coalesced 4-byte index loads (4 sectors)
    overall = (per.sum() + 4.0 * (QQ - 1)) / (2 * QQ - 1)
    assert abs(overall - ncu_value) < 0.02, overall
