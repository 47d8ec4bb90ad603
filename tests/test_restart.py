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
