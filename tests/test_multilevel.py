"""Multi-level (config 4): ghost interpolation + the level-sync schedule.

CPU: the synthetic multi-level generator against the oracle's restatement of
mus_construct_connectivity, and analytic properties of the oracle's interpolation
(the reference's own interpolation utests are deactivated -> "parity unpinned"):
uniform flow is a fixed point, linear/quadratic interpolation reproduce linear fields.
GPU: libmusb200 against the oracle on the same two- and three-level meshes, bit-exact."""
import numpy as np
import pytest


def tgv_like(bary_unit, u0=0.03):
    x = 2.0 * np.pi * bary_unit
    vel = np.stack([u0 * np.sin(x[:, 0]) * np.cos(x[:, 1]) * np.cos(x[:, 2]) + 0.02,
                    -u0 * np.cos(x[:, 0]) * np.sin(x[:, 1]) * np.cos(x[:, 2]) - 0.01,
                    0.015 + 0.0 * x[:, 0]], axis=1)
    rho = 1.0 + 3.0 * (u0 * u0 / 16.0) * (np.cos(2 * x[:, 0]) + np.cos(2 * x[:, 1])) * (np.cos(2 * x[:, 2]) + 2.0)
    return rho, vel


def build(mo, min_level, boxes, QQ, method, relaxation="bgk", cylinder=None, omega_min=1.6):
    import musubi_b200 as mb
    from musubi_b200 import treelm_multilevel as tm
    lv, intp = tm.build_multilevel(min_level, boxes, QQ=QQ, cylinder=cylinder, intp_method=method)
    tables = mb.multilevel_tables(lv, intp)
    ms = mo.MultiLevelScheme(lv, tables, relaxation, "fluid", omega_min=omega_min, omega_bulk=1.2,
                             order=intp["order"])
    for l, s in ms.s.items():
        rho, vel = tgv_like(lv[l].bary_unit)
        s.init_equilibrium(rho, vel)
    return lv, intp, tables, ms


@pytest.mark.parametrize("QQ", [19, 27])
def test_generator_connectivity_matches_oracle(oracle, QQ):
    from musubi_b200 import treelm_multilevel as tm
    lv, _ = tm.build_multilevel(5, [(10, 22)], QQ=QQ, cylinder=(32.0, 32.0, 3.0, 29, 35))
    for l, L in lv.items():
        ng = np.zeros(QQ * L.nSize, dtype=np.int32)
        oracle.lib().ora_construct_connectivity(oracle._i(ng), L.nSize, L.nElems, QQ,
                                                oracle._i(np.ascontiguousarray(L.nghElems)),
                                                oracle._l(L.property), L.nFluid, L.nElems)
        assert np.array_equal(ng, L.neigh)
        assert np.all(np.diff(L.total[:L.nFluid]) > 0)
        a, b = L.nFluid, L.nFluid + L.nGhostFromCoarser
        assert np.all(np.diff(L.total[a:b]) > 0) and np.all(np.diff(L.total[b:]) > 0)
    assert lv[5].nGhostFromFiner > 0 and lv[6].nGhostFromCoarser > 0
    assert (lv[6].nghElems[:lv[6].nFluid] == -1).any()      # the cylinder is seen as a wall


@pytest.mark.parametrize("method", ["weighted_average", "linear", "quadratic"])
@pytest.mark.parametrize("QQ", [19, 27])
def test_uniform_flow_is_a_fixed_point(oracle, method, QQ):
    import musubi_b200 as mb
    from musubi_b200 import treelm_multilevel as tm
    lv, intp = tm.build_multilevel(4, [(5, 11)], QQ=QQ, intp_method=method)
    ms = oracle.MultiLevelScheme(lv, mb.multilevel_tables(lv, intp), "bgk", "fluid", omega_min=1.6,
                                 order=intp["order"])
    u = np.array([0.02, 0.01, -0.015])
    for l, s in ms.s.items():
        s.init_equilibrium(np.ones(lv[l].nElems), np.tile(u, (lv[l].nElems, 1)))
    ref = ms.s[4].state[ms.s[4].nNext][:QQ].copy()
    m0 = ms.total_mass()
    ms.run(6)
    for l, s in ms.s.items():
        got = s.state[s.nNext][:lv[l].nFluid * QQ].reshape(-1, QQ)
        assert np.abs(got - ref).max() < 5e-15
    assert abs(ms.total_mass() / m0 - 1.0) < 1e-14


@pytest.mark.parametrize("method,order", [("linear", 1), ("quadratic", 2)])
def test_interpolation_reproduces_linear_fields(oracle, method, order):
    """a density field linear in x,y,z at rest (f = f_eq) must be interpolated exactly to the
    fine ghosts by the least-square linear and quadratic interpolation."""
    import musubi_b200 as mb
    from musubi_b200 import treelm_multilevel as tm
    QQ = 19
    lv, intp = tm.build_multilevel(4, [(5, 11)], QQ=QQ, intp_method=method)
    tables = mb.multilevel_tables(lv, intp)
    ms = oracle.MultiLevelScheme(lv, tables, "bgk", "fluid", omega_min=1.6, order=order)
    g = np.array([0.3, -0.2, 0.1])
    for l, s in ms.s.items():
        b = lv[l].bary_unit
        # wrap-free coordinate: the refined box is in the interior of the cube
        rho = 1.0 + (b - 0.5) @ g
        s.init_equilibrium(rho, np.zeros((lv[l].nElems, 3)))
    ms._from_coarser(4)
    f = ms.s[5]
    L5 = lv[5]
    gh = slice(L5.nFluid, L5.nFluid + L5.nGhostFromCoarser)
    got_rho = f.state[f.nNext].reshape(-1, QQ)[gh].sum(axis=1)
    exp_rho = 1.0 + (L5.bary_unit[gh] - 0.5) @ g
    assert np.abs(got_rho - exp_rho).max() < 1e-13


# ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mbgpu():
    import musubi_b200
    musubi_b200.mus_init(0, 1, 0)
    yield musubi_b200
    musubi_b200.mus_finalize()


CASES = [
    (4, [(5, 11)], 19, "linear", "bgk", None),
    (5, [(10, 22)], 19, "linear", "bgk", (32.0, 32.0, 3.0, 29, 35)),
    (4, [(5, 11)], 27, "quadratic", "mrt", None),
    (4, [(5, 11)], 19, "weighted_average", "trt", None),
    (4, [(4, 12), (12, 20)], 19, "linear", "bgk", None),
]
# omega of the coarsest level.  Acoustic scaling doubles nu per finer level; the three-level case
# must avoid omega_fine == 1 exactly (omega_min = 1.6 gives it on level min+2), where the reference's
# f_neq rescale factor 2 w_f (1 - w_c) / ((1 - w_f) w_c) divides by zero.
OMEGA_MIN = {1: 1.6, 2: 1.75}


@pytest.mark.gpu
@pytest.mark.parametrize("tiled", [1, 0], ids=["tiled", "target-major"])
@pytest.mark.parametrize("min_level,boxes,QQ,method,relax,cyl", CASES,
                         ids=["2lvl-linear-bgk19", "2lvl-cylinder", "2lvl-quad-mrt27", "2lvl-wavg-trt19",
                              "3lvl-linear-bgk19"])
def test_multilevel_gpu_matches_oracle(mbgpu, oracle, min_level, boxes, QQ, method, relax, cyl, tiled):
    """both forms of the coarse -> fine interpolation kernel (shared-memory tiles of sources /
    one thread per (target, direction)) against the oracle, bit for bit"""
    mb = mbgpu
    from musubi_b200._lib import check, lib
    check(lib.musb200_set_intp_tiled(tiled))
    lv, intp, tables, ms = build(oracle, min_level, boxes, QQ, method, relax, cyl,
                                 omega_min=OMEGA_MIN[len(boxes)])
    ident = {"kind": "fluid", "relaxation": relax, "layout": "d3q%d" % QQ}
    omega = {l: float(1.0 / (3.0 * s.visc[0] + 0.5)) for l, s in ms.s.items()}
    visc = {l: float(s.visc[0]) for l, s in ms.s.items()}
    sch = mb.Scheme(ident, lv, omega, lambda_=0.25, omega_bulk=1.2, intp=(tables, intp["order"]),
                    viscosity=visc)
    from musubi_b200._lib import check, lib
    for l, s in ms.s.items():
        sch.upload_state(l, s.state[s.nNow], s.state[s.nNext])
        check(lib.musb200_aux_upload(l, s.aux.ctypes.data))
    ncyc = 12
    sch.do_computation(ncyc)
    ms.run(ncyc)
    for l, s in ms.s.items():
        L = lv[l]
        got = sch.download_state(l)[:L.nElems * QQ].reshape(-1, QQ)
        exp = s.state[s.nNext][:L.nElems * QQ].reshape(-1, QQ)
        nf = L.nFluid
        assert np.isfinite(exp).all(), "oracle run is not finite"
        rel = np.max(np.abs(got[:nf] - exp[:nf]) / np.abs(exp[:nf]))
        assert rel < 1e-10, (l, rel)
        assert np.array_equal(got[:nf], exp[:nf]), "fluid PDFs of level %d not bit-identical" % l
        # ghosts filled by interpolation
        assert np.array_equal(got[nf:], exp[nf:]), "ghost PDFs of level %d differ" % l
        aux = sch.download_aux(l)[:L.nElems * 4]
        assert np.array_equal(aux[:nf * 4], s.aux[:nf * 4])
    sch.destroy()
    check(lib.musb200_set_intp_tiled(1))


# ---- several ranks -----------------------------------------------------------------------------
def _partitioned(mo, min_level, boxes, QQ, method, nranks, relax="bgk", cyl=None, omega_min=1.6,
                 balanced=False):
    import musubi_b200 as mb
    from musubi_b200 import treelm_multilevel as tm
    lv, intp, tables, ms = build(mo, min_level, boxes, QQ, method, relax, cyl, omega_min)
    ranks = tm.partition_multilevel(lv, nranks, weights=tm.level_weights(lv) if balanced else None)
    rtables = [mb.multilevel_tables(rl, intp) for rl in ranks]
    mr = mo.MultiRankMultiLevel(ranks, rtables, relaxation=relax, kind="fluid", omega_min=omega_min,
                                omega_bulk=1.2, order=intp["order"])
    for m in mr.r:
        for l, s in m.s.items():
            rho, vel = tgv_like(s.ld.bary_unit)
            s.init_equilibrium(rho, vel)
    return lv, intp, ms, ranks, rtables, mr


@pytest.mark.parametrize("min_level,boxes,QQ,method,relax,cyl,nranks", [
    (4, [(5, 11)], 19, "linear", "bgk", None, 2),
    (4, [(5, 11)], 19, "linear", "bgk", None, 3),
    (5, [(10, 22)], 19, "linear", "bgk", (32.0, 32.0, 3.0, 29, 35), 4),
    (4, [(5, 11)], 27, "quadratic", "mrt", None, 2),
    (4, [(4, 12), (12, 20)], 19, "linear", "bgk", None, 4),
], ids=["2lvl-2ranks", "2lvl-3ranks", "2lvl-cylinder-4ranks", "2lvl-quad-mrt27-2ranks", "3lvl-4ranks"])
def test_partitioned_multilevel_equals_single_domain(oracle, min_level, boxes, QQ, method, relax, cyl, nranks):
    """the multi-level mesh cut along the global space-filling curve: every rank's fluid elements
    evolve bit-identically to the single-domain run (ghosts are recomputed locally, only fluid
    elements travel through the halo buffers)"""
    lv, intp, ms, ranks, rtables, mr = _partitioned(oracle, min_level, boxes, QQ, method, nranks, relax, cyl,
                                                    OMEGA_MIN[len(boxes)])
    # partition sanity: every fluid element owned exactly once, equal shares
    for l in lv:
        owned = np.concatenate([rl[l].globalPos[:rl[l].nFluid] for rl in ranks])
        assert np.array_equal(np.sort(owned), np.arange(1, lv[l].nFluid + 1))
    tot = [sum(rl[l].nFluid for l in lv) for rl in ranks]
    assert max(tot) - min(tot) <= 1
    ncyc = 6
    ms.run(ncyc)
    mr.run(ncyc)
    for r, m in enumerate(mr.r):
        for l, s in m.s.items():
            M = ranks[r][l]
            g = M.globalPos[:M.nFluid] - 1
            got = s.state[s.nNext][:M.nFluid * QQ].reshape(-1, QQ)
            exp = ms.s[l].state[ms.s[l].nNext].reshape(-1, QQ)[g]
            assert np.array_equal(got, exp), "rank %d level %d fluid PDFs differ" % (r, l)


def test_sparta_split_rule():
    """tem_balance_sparta restated for one rank: every splitter sits after the element whose
    weight prefix sum is closest to k * W / nParts"""
    from musubi_b200 import treelm_multilevel as tm
    assert list(tm.sparta_split(np.ones(12), 4)) == [3, 3, 3, 3]
    assert list(tm.sparta_split(np.ones(10), 1)) == [10]
    # the reference's own known-answer test (tem/utests/tem_sparta_test.f90:46-86): five ranks with five
    # elements each and these weights end up with 4 / 3 / 5 / 5 / 8 elements at offsets 0 / 4 / 7 / 12 / 17
    w = np.array([5, 3, 1, 2, 1, 4, 6, 1, 3, 2, 1, 3, 1, 1, 1, 1, 9, 1, 1, 1, 1, 1, 1, 1, 1], dtype=np.float64)
    cnt = tm.sparta_split(w, 5)
    assert list(cnt) == [4, 3, 5, 5, 8]
    assert list(np.concatenate([[0], np.cumsum(cnt)[:-1]])) == [0, 4, 7, 12, 17]
    rng = np.random.default_rng(3)
    for n, parts in ((1000, 7), (513, 8), (64, 3)):
        w = rng.choice([1.0, 2.0, 4.0], size=n)
        cnt = tm.sparta_split(w, parts)
        assert cnt.sum() == n and np.all(cnt > 0)
        pre = np.cumsum(w)
        ends = np.cumsum(cnt)[:-1]
        for k, e in enumerate(ends):
            target = (k + 1) * pre[-1] / parts
            assert abs(pre[e - 1] - target) <= np.min(np.abs(pre - target)) + 1e-9
        load = np.add.reduceat(w, np.concatenate([[0], ends]))
        assert load.max() - pre[-1] / parts <= 4.0          # never more than one element off per cut


@pytest.mark.parametrize("min_level,boxes,nranks", [(4, [(5, 11)], 2), (4, [(5, 11)], 3),
                                                    (4, [(4, 12), (12, 20)], 4)],
                         ids=["2lvl-2ranks", "2lvl-3ranks", "3lvl-4ranks"])
def test_weighted_partition_balances_the_work_and_stays_bit_identical(oracle, min_level, boxes, nranks):
    """the SPartA cut with level weights 2^(l - minLevel): level steps per coarse cycle are spread
    evenly (the equal-count cut of treelm's first distribution is not), results unchanged"""
    from musubi_b200 import treelm_multilevel as tm
    QQ = 19
    lv, intp, ms, ranks, rtables, mr = _partitioned(oracle, min_level, boxes, QQ, "linear", nranks,
                                                    omega_min=OMEGA_MIN[len(boxes)], balanced=True)
    minL = min(lv)
    work = [sum(rl[l].nFluid * 2 ** (l - minL) for l in lv) for rl in ranks]
    plain = [sum(rl[l].nFluid * 2 ** (l - minL) for l in lv) for rl in tm.partition_multilevel(lv, nranks)]
    finest = 2 ** (max(lv) - minL)
    assert max(work) - min(work) <= 2 * finest          # within one finest element per cut
    # never worse than the equal-count cut; equal where the octant symmetry of the centred boxes
    # already balances it (2 and 4 ranks), better otherwise
    assert max(work) <= max(plain) and (nranks in (2, 4) or max(work) < max(plain))
    for l in lv:
        owned = np.concatenate([rl[l].globalPos[:rl[l].nFluid] for rl in ranks])
        assert np.array_equal(np.sort(owned), np.arange(1, lv[l].nFluid + 1))
    ncyc = 4
    ms.run(ncyc)
    mr.run(ncyc)
    for r, m in enumerate(mr.r):
        for l, s in m.s.items():
            M = ranks[r][l]
            g = M.globalPos[:M.nFluid] - 1
            got = s.state[s.nNext][:M.nFluid * QQ].reshape(-1, QQ)
            exp = ms.s[l].state[ms.s[l].nNext].reshape(-1, QQ)[g]
            assert np.array_equal(got, exp), "rank %d level %d fluid PDFs differ" % (r, l)


@pytest.mark.parametrize("min_level,boxes,QQ,method,relax,nranks", [
    (4, [(5, 11)], 19, "linear", "bgk", 2), (4, [(5, 11)], 19, "linear", "bgk", 3),
    (4, [(5, 11)], 27, "quadratic", "mrt", 2), (4, [(4, 12), (12, 20)], 19, "linear", "bgk", 4)],
    ids=["2lvl-2ranks", "2lvl-3ranks", "2lvl-quad-mrt27-2ranks", "3lvl-4ranks"])
def test_ghosts_delegated_through_the_from_coarser_and_from_finer_buffers(oracle, min_level, boxes, QQ, method,
                                                                         relax, nranks):
    """the reference's form of the same run: shared ghosts are interpolated by one rank and reach
    the others through sendBufferFromCoarser / sendBufferFromFiner (+ the auxField of
    ghostFromFiner elements) right after the interpolation that fills them; the fluid PDFs of
    every rank stay bit-identical to the single-domain run"""
    import musubi_b200 as mb
    from musubi_b200 import treelm_multilevel as tm
    lv, intp, tables, ms = build(oracle, min_level, boxes, QQ, method, relax, None, OMEGA_MIN[len(boxes)])
    ranks = tm.partition_multilevel(lv, nranks)
    rtables = [mb.multilevel_tables(rl, intp) for rl in ranks]
    dtables, comm = tm.delegate_shared_ghosts(ranks, rtables, lv)
    n_del = {k: sum(len(c["elemPos"]) for r in range(nranks) for l in lv for c in comm[r][l][k]["recv"])
             for k in ("fromCoarser", "fromFiner")}
    assert n_del["fromCoarser"] > 0 and n_del["fromFiner"] > 0
    # every message has its counterpart, element for element (same treeIDs in the same order)
    for r in range(nranks):
        for l in lv:
            for k in ("fromCoarser", "fromFiner"):
                for snd in comm[r][l][k]["send"]:
                    rcv = [c for c in comm[snd["proc"]][l][k]["recv"] if c["proc"] == r]
                    assert len(rcv) == 1 and rcv[0]["pos"].size == snd["pos"].size
                    assert np.array_equal(ranks[r][l].total[snd["elemPos"] - 1],
                                          ranks[snd["proc"]][l].total[rcv[0]["elemPos"] - 1])
    # the removed targets are exactly the received elements
    for r in range(nranks):
        for l in lv:
            got = set()
            for k in ("fromCoarser", "fromFiner"):
                for c in comm[r][l][k]["recv"]:
                    got |= set(int(e) for e in c["elemPos"])
            before = set()
            after = set()
            for key in rtables[r]:
                if key[0] == l:
                    before |= set(int(t) for t in rtables[r][key]["targets"])
                    after |= set(int(t) for t in dtables[r][key]["targets"])
            assert before - after == got and after <= before
    mr = oracle.MultiRankMultiLevel(ranks, dtables, ghost_comm=comm, relaxation=relax, kind="fluid",
                                    omega_min=OMEGA_MIN[len(boxes)], omega_bulk=1.2, order=intp["order"])
    for m in mr.r:
        for l, s in m.s.items():
            rho, vel = tgv_like(s.ld.bary_unit)
            s.init_equilibrium(rho, vel)
    ncyc = 5
    ms.run(ncyc)
    mr.run(ncyc)
    for r, m in enumerate(mr.r):
        for l, s in m.s.items():
            M = ranks[r][l]
            g = M.globalPos[:M.nFluid] - 1
            got = s.state[s.nNext][:M.nFluid * QQ].reshape(-1, QQ)
            exp = ms.s[l].state[ms.s[l].nNext].reshape(-1, QQ)[g]
            assert np.array_equal(got, exp), "rank %d level %d fluid PDFs differ" % (r, l)


def test_shear_wave_crosses_the_refinement_interface(oracle):
    """physics pin for the ghost interpolation (no reference fixture reaches it): u_x = U sin(2 pi y)
    on a 32^3 periodic cube with a 12^3-cell box refined once.  Acoustic scaling keeps the physical
    viscosity across levels (f_neq rescaled at the interface), so the wave must decay at the coarse
    level's nu k^2 and stay a sine on both levels; the quadratic interpolation leaves the smallest
    distortion."""
    import math
    N, k, U, n = 32, 2.0 * math.pi, 1.0e-3, 100
    resid = {}
    for method in ("weighted_average", "linear", "quadratic"):
        lv, intp, tables, ms = build(oracle, 5, [(10, 22)], 19, method, "bgk", None, 1.6)
        for s in ms.s.values():
            vel = np.zeros((s.ld.nElems, 3))
            vel[:, 0] = U * np.sin(k * s.ld.bary_unit[:, 1])
            s.init_equilibrium(np.ones(s.ld.nElems), vel)
        nu0 = float(ms.s[5].visc[0])
        assert abs(float(ms.s[6].visc[0]) / nu0 - 2.0) < 1e-14          # lattice viscosity doubles per level
        ms.run(n)
        worst = 0.0
        for l, s in ms.s.items():
            nf = s.ld.nFluid
            aux, sn = s.aux.reshape(-1, 4)[:nf], np.sin(k * s.ld.bary_unit[:nf, 1])
            amp = float((aux[:, 1] * sn).sum() / (sn * sn).sum())
            if l == 5:      # the coarse level covers (almost) whole periods: its projection is the decay
                rate = -math.log(amp / U) / n
                assert abs(rate / (nu0 * (k / N) ** 2) - 1.0) < 0.05, (method, rate)
            worst = max(worst, float(np.max(np.abs(aux[:, 1] - amp * sn))) / abs(amp),
                        float(np.max(np.abs(aux[:, 2]))) / U, float(np.max(np.abs(aux[:, 3]))) / U)
            assert np.max(np.abs(aux[:, 0] - 1.0)) < 1e-5
        resid[method] = worst
    assert resid["weighted_average"] < 0.03 and resid["linear"] < 0.03 and resid["quadratic"] < 0.008
    assert resid["quadratic"] < 0.5 * resid["linear"]
