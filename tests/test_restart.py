"""Restart bridge: the device state serialised in the order of
mus_pdf_serialize (mus/source/mus_buffer_module.fpp:80-137) -- the payload of the reference's
restart *.lsb -- and read back with mus_pdf_unserialize."""
import numpy as np
import pytest


def test_level_of_treeid_and_global_tree_order(oracle):
    mo = oracle
    assert [mo.level_of(t) for t in (0, 1, 8, 9, 72, 73, 584, 585)] == [0, 1, 1, 2, 2, 3, 3, 4]
    from musubi_b200 import treelm_multilevel as tm
    lv, _ = tm.build_multilevel(4, [(5, 11)], QQ=19)
    tid, lp = mo.global_tree(lv)
    assert tid.size == sum(L.nFluid for L in lv.values())
    # space-filling-curve order: every element's first finest-level descendant id is increasing
    maxL = max(lv)
    lvl = np.array([mo.level_of(int(t)) for t in tid])
    key = np.array([(int(t) - mo.first_id_at_level(int(l))) << (3 * (maxL - int(l))) for t, l in zip(tid, lvl)])
    assert np.all(np.diff(key) > 0)
    for l, L in lv.items():
        sel = lvl == l
        assert np.array_equal(np.asarray(L.total)[lp[sel] - 1], tid[sel])


def test_oracle_serialize_roundtrip(oracle):
    mo = oracle
    ld = mo.build_level_desc(3, 19, "periodic")
    s = mo.Scheme(ld, "bgk", "fluid", omega=1.7)
    rng = np.random.default_rng(4)
    s.state[s.nNext][:] = rng.random(s.state[s.nNext].size)
    tid, lp = mo.global_tree({3: ld})
    buf = mo.pdf_serialize({3: s}, tid, lp)
    assert np.array_equal(buf, s.state[s.nNext][:ld.nFluid * 19])   # single level: list order = SFC
    t = mo.Scheme(ld, "bgk", "fluid", omega=1.7)
    mo.pdf_unserialize({3: t}, tid, lp, buf)
    assert np.array_equal(t.state[t.nNext][:ld.nFluid * 19], buf)


@pytest.fixture(scope="module")
def mbgpu():
    import musubi_b200 as mb
    mb.mus_init(0, 1, 0)
    yield mb
    mb.mus_finalize()


@pytest.mark.gpu
def test_device_restart_dump_equals_reference_order_multilevel(mbgpu, oracle):
    """two-level mesh, a few cycles, dump in two chunks: bytes equal the oracle's serialisation;
    then restore into a fresh scheme and continue: identical to the uninterrupted run."""
    from test_multilevel import build
    mb, mo, QQ = mbgpu, oracle, 19
    lv, intp, tables, ms = build(mo, 4, [(5, 11)], QQ, "linear")
    ident = {"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}
    omega = {l: float(1.0 / (3.0 * s.visc[0] + 0.5)) for l, s in ms.s.items()}
    visc = {l: float(s.visc[0]) for l, s in ms.s.items()}
    from musubi_b200._lib import check, lib

    def fresh():
        sch = mb.Scheme(ident, lv, omega, omega_bulk=1.2, intp=(tables, intp["order"]), viscosity=visc)
        for l, s in ms.s.items():
            sch.upload_state(l, s.state[s.nNow], s.state[s.nNext])
            check(lib.musb200_aux_upload(l, s.aux.ctypes.data))
        return sch

    sch = fresh()
    sch.do_computation(5)
    ms.run(5)
    tid, lp = mo.global_tree(lv)
    exp = mo.pdf_serialize(ms.s, tid, lp)
    h = tid.size // 2 + 3                      # chunked like tem_restart_writeData
    got = np.concatenate([sch.pdf_serialize(tid[:h], lp[:h]), sch.pdf_serialize(tid[h:], lp[h:])])
    assert got.tobytes() == exp.tobytes()
    # restore: ghosts and nNow come from the live scheme here (the reference re-fills the ghosts by
    # interpolation after a restart read); overwrite the fluid PDFs with a perturbed dump and back
    pert = got + 1.0
    sch.pdf_unserialize(tid, lp, pert)
    assert np.array_equal(sch.pdf_serialize(tid, lp), pert)
    sch.pdf_unserialize(tid[:h], lp[:h], got[:h * QQ])
    sch.pdf_unserialize(tid[h:], lp[h:], got[h * QQ:])
    sch.do_computation(3)
    ms.run(3)
    assert sch.pdf_serialize(tid, lp).tobytes() == mo.pdf_serialize(ms.s, tid, lp).tobytes()
    sch.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("boxes,method", [([(5, 11)], "linear"), ([(4, 12), (12, 20)], "quadratic")],
                         ids=["2lvl-linear", "3lvl-quadratic"])
def test_restart_into_a_freshly_created_multilevel_scheme(mbgpu, oracle, tmp_path, boxes, method):
    """mus_readRestart + mus_init_flow (mus_flow_module.fpp:206-240): the restart file holds fluid
    elements only, so a FRESH scheme (ghost, auxField rows all zero after level_create) has to
    rebuild auxField and ghosts before its first step -- musb200_fill_helper_elements, called by
    Scheme.read_restart.  The restarted device run equals the oracle doing the same (bit for bit)
    and continues the uninterrupted run to rounding (ghosts re-interpolated from the moments of
    the stored PDFs instead of the pre-collision ones; the reference has the same property)."""
    from test_multilevel import build
    from musubi_b200._lib import check, lib
    mb, mo, QQ = mbgpu, oracle, 19
    from test_multilevel import OMEGA_MIN
    om_min = OMEGA_MIN[len(boxes)]       # keeps every level's omega away from 1 (the f_neq factor's pole)
    lv, intp, tables, ms = build(mo, 4, boxes, QQ, method, omega_min=om_min)
    ident = {"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}
    omega = {l: float(1.0 / (3.0 * s.visc[0] + 0.5)) for l, s in ms.s.items()}
    visc = {l: float(s.visc[0]) for l, s in ms.s.items()}
    kw = dict(omega_bulk=1.2, intp=(tables, intp["order"]), viscosity=visc)
    sch = mb.Scheme(ident, lv, omega, **kw)
    for l, s in ms.s.items():
        sch.upload_state(l, s.state[s.nNow], s.state[s.nNext])
        check(lib.musb200_aux_upload(l, s.aux.ctypes.data))
    sch.do_computation(4)
    ms.run(4)
    tid, lp = mo.global_tree(lv)
    dump = mo.pdf_serialize(ms.s, tid, lp)
    _, hdr = sch.write_restart(str(tmp_path) + "/", "fresh", dict(sim=4.0, iter=4))
    sch.do_computation(3)                                           # the uninterrupted run
    cont = sch.pdf_serialize(tid, lp)
    sch.destroy()

    # fresh device scheme: nothing but the file
    sch2 = mb.Scheme(ident, lv, omega, **kw)
    sch2.read_restart(hdr)
    assert np.array_equal(sch2.pdf_serialize(tid, lp), dump)
    # fresh oracle scheme fed with the same fluid PDFs, same fill
    ms2 = mo.MultiLevelScheme(lv, tables, "bgk", "fluid", omega_min=om_min, omega_bulk=1.2, order=intp["order"])
    mo.pdf_unserialize(ms2.s, tid, lp, dump)
    ms2.fill_helper_elements()
    for l, s in ms2.s.items():                                      # ghosts and auxField after the fill
        n = lv[l].nElems
        assert np.array_equal(sch2.download_state(l)[:n * QQ], s.state[s.nNext][:n * QQ]), "level %d" % l
        nf = lv[l].nFluid
        assert np.array_equal(sch2.download_aux(l)[:nf * 4], s.aux[:nf * 4])
    sch2.do_computation(3)
    ms2.run(3)
    got = sch2.pdf_serialize(tid, lp)
    assert got.tobytes() == mo.pdf_serialize(ms2.s, tid, lp).tobytes()
    assert not np.isnan(got).any()
    assert np.max(np.abs(got - cont) / np.abs(cont)) < 1e-11       # continues the uninterrupted run
    sch2.destroy()


@pytest.mark.parametrize("boxes,method", [([(5, 11)], "linear"), ([(4, 12), (12, 20)], "quadratic")])
def test_oracle_restart_with_helper_fill_continues_the_run(oracle, boxes, method):
    """the oracle's own restart: fluid PDFs into a fresh MultiLevelScheme, fill_helper_elements,
    continue -- equal to the uninterrupted run to rounding, nothing NaN (a fresh scheme without the
    fill pulls zeros from its ghosts)"""
    from test_multilevel import OMEGA_MIN, build
    mo = oracle
    om_min = OMEGA_MIN[len(boxes)]
    lv, intp, tables, ms = build(mo, 4, boxes, 19, method, omega_min=om_min)
    ms.run(3)
    tid, lp = mo.global_tree(lv)
    dump = mo.pdf_serialize(ms.s, tid, lp)
    ms.run(2)
    cont = mo.pdf_serialize(ms.s, tid, lp)
    ms2 = mo.MultiLevelScheme(lv, tables, "bgk", "fluid", omega_min=om_min, omega_bulk=1.2, order=intp["order"])
    mo.pdf_unserialize(ms2.s, tid, lp, dump)
    ms2.fill_helper_elements()
    ms2.run(2)
    got = mo.pdf_serialize(ms2.s, tid, lp)
    assert not np.isnan(got).any()
    assert np.max(np.abs(got - cont) / np.abs(cont)) < 1e-11
