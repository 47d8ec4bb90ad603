"""CPU-side tests: the C-ABI library loads and exports every declared symbol,
the host mirror reproduces the reference's dispatch/error behaviour, and the
product's mesh generator agrees bit-for-bit with the oracle's restatement of
the treelm / mus_construct index lists.  No compute call is made (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import musubi_b200._lib as L
    hdr = open(os.path.join(ROOT, "include", "musb200.h")).read()
    declared = set(re.findall(r"\bint\s+(musb200_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L.lib, name), name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)


def test_scheme_select_mirrors_reference_dispatch():
    import musubi_b200 as mb
    assert mb.select_kernel({"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}) == (0, 0, 19)
    assert mb.select_kernel({"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}) == (1, 0, 19)
    assert mb.select_kernel({"kind": "fluid", "relaxation": {"name": "mrt", "variant": "standard"},
                             "layout": "d3q27"}) == (2, 0, 27)
    assert mb.select_kernel({"kind": "fluid_incompressible", "relaxation": "bgk", "layout": "d3q19"}) == (0, 1, 19)
    for bad in ({"kind": "multispecies_gas", "relaxation": "bgk", "layout": "d3q19"},
                {"kind": "fluid", "relaxation": "cumulant", "layout": "d3q27"},
                {"kind": "fluid", "relaxation": "bgk", "layout": "d2q9"},
                {"kind": "fluid", "relaxation": {"name": "bgk", "variant": "improved"}, "layout": "d3q19"}):
        with pytest.raises(mb.Musb200Error) as ei:
            mb.select_kernel(bad)
        assert ei.value.code == 4


def test_no_cpu_fallback_without_gpu():
    import musubi_b200 as mb
    n = ctypes.c_int()
    rc = mb._lib.lib.musb200_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(mb.Musb200Error):
        mb.mus_init(0, 1, 0)
    with pytest.raises(mb.Musb200Error) as ei:
        mb._lib.check(mb._lib.lib.musb200_step(1, 1, 1))
    assert ei.value.code == 6


@pytest.mark.parametrize("level,QQ,kind,nranks", [
    (3, 19, "periodic", 1), (4, 27, "periodic", 1), (4, 19, "cavity", 1), (4, 27, "cavity", 1),
    (4, 19, "periodic", 2), (4, 27, "periodic", 4), (4, 19, "cavity", 3), (5, 27, "periodic", 8),
    (4, 19, "channel", 1), (4, 27, "channel", 1), (4, 19, "channel", 2), (4, 27, "channel", 8)])
def test_index_lists_bit_exact_vs_oracle(oracle, level, QQ, kind, nranks):
    import musubi_b200 as mb
    for r in range(nranks):
        a = mb.LevelDesc(level, QQ, kind, r, nranks)
        b = oracle.build_level_desc(level, QQ, kind, r, nranks)
        assert (a.nFluid, a.nHalo, a.nElems, a.nSize) == (b.nFluid, b.nHalo, b.nElems, b.nSize)
        for k in ("total", "property", "nghElems", "neigh", "bc_elemBuffer"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        assert [c["proc"] for c in a.recv] == [c["proc"] for c in b.recv]
        assert [c["proc"] for c in a.send] == [c["proc"] for c in b.send]
        for x, y in zip(a.recv + a.send, b.recv + b.send):
            assert np.array_equal(x["pos"], y["pos"]) and np.array_equal(x["elemPos"], y["elemPos"])
        for x, y in zip(a.bc, b.bc):
            assert x["id"] == y["id"] and x["kind"] == y["kind"]
            for k in ("elems", "links", "outPos", "posInBuffer", "iDir", "normalInd", "posInBcElemBuf",
                      "neighPos", "iElemOfLink", "statePos"):
                assert np.array_equal(x[k], y[k]), k


@pytest.mark.parametrize("QQ,kind,octants,nranks", [(19, "cavity", 1, 1), (19, "cavity", 2, 2),
                                                      (27, "cavity", 4, 4), (19, "channel", 2, 2),
                                                      (27, "channel", 4, 3)])
def test_partial_cube_lists_bit_exact_vs_oracle(oracle, QQ, kind, octants, nranks):
    """weak-scaling meshes: the first 1/2/4 octants of the cube (what bench.py --gpus N uses)"""
    import musubi_b200 as mb
    for r in range(nranks):
        a = mb.LevelDesc(4, QQ, kind, r, nranks, octants=octants)
        b = oracle.build_level_desc(4, QQ, kind, r, nranks, octants=octants)
        assert a.nFluid + sum(mb.LevelDesc(4, QQ, kind, q, nranks, octants=octants).nFluid
                              for q in range(nranks) if q != r) == octants * 8 ** 3
        for k in ("total", "property", "nghElems", "neigh", "bc_elemBuffer"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), k
        for x, y in zip(a.recv + a.send, b.recv + b.send):
            assert x["proc"] == y["proc"] and np.array_equal(x["pos"], y["pos"])
        for x, y in zip(a.bc, b.bc):
            for k in ("elems", "links", "outPos", "posInBuffer", "iDir", "normalInd", "neighPos"):
                assert np.array_equal(x[k], y[k]), k


def test_send_and_recv_lists_pair_up(oracle):
    """what rank p receives from q is exactly what q sends to p (same length, same links)."""
    import musubi_b200 as mb
    nr = 4
    lds = [mb.LevelDesc(4, 27, "periodic", r, nr) for r in range(nr)]
    for p in range(nr):
        for rcv in lds[p].recv:
            q = rcv["proc"]
            snd = [s for s in lds[q].send if s["proc"] == p]
            assert len(snd) == 1
            assert len(snd[0]["pos"]) == len(rcv["pos"])
            # same treeIDs and directions on both sides
            te = lds[p].total[(rcv["pos"] - 1) // 27]
            se = lds[q].total[(snd[0]["pos"] - 1) // 27]
            assert np.array_equal(te, se)
            assert np.array_equal((rcv["pos"] - 1) % 27, (snd[0]["pos"] - 1) % 27)
    # reduced link set of a z-slab partition: 9 of 27 links per face halo
    two = mb.LevelDesc(4, 27, "periodic", 0, 2)
    assert len(two.recv[0]["pos"]) == 2 * 16 * 16 * 9


def test_default_omega_bulk_follows_the_reference():
    """fluid table without bulk_viscosity (mus_fluid_module.f90:205-262): d3q27 -> 1.54,
    incompressible d3q19 -> 1.19, compressible d3q19 aborts; finer levels scale acoustically"""
    from musubi_b200.scheme import default_omega_bulk
    f27 = {"kind": "fluid", "relaxation": "mrt", "layout": "d3q27"}
    i19 = {"kind": "fluid_incompressible", "relaxation": "mrt", "layout": "d3q19"}
    assert abs(default_omega_bulk(f27, 5, 5) - 1.54) < 1e-15
    assert abs(default_omega_bulk(i19, 7, 7) - 1.19) < 1e-15
    nu = (2.0 / 9.0) * (1.0 / 1.54 - 0.5)
    assert abs(default_omega_bulk(f27, 6, 5) - 1.0 / (9.0 * (2.0 * nu) / 2.0 + 0.5)) < 1e-15
    with pytest.raises(ValueError, match="bulk_viscosity"):
        default_omega_bulk({"kind": "fluid", "relaxation": "mrt", "layout": "d3q19"}, 4, 4)


def test_periodic_level1_cube_neighbours_are_the_references_known_answers(oracle):
    """treelm's own unit tests hold the neighbour relations of the predefined periodic cube at
    refinement level 1 (tem/utests/tem_serial_singlelevel_test.f90:95-260, cube from
    tem_utestEnv_module.f90:43-49): the right neighbour of element 2 in x is 1, of 6 is 5; of 3 in
    y is 1; of 1 in z is 5, of 4 is 8 (treeIDs).  Here: both connectivity generators, through the
    neigh list -- direction d with c_d = -e_axis pulls from the right neighbour"""
    import musubi_b200 as mb
    for QQ in (19, 27):
        for ld in (mb.LevelDesc(1, QQ, "periodic"), oracle.build_level_desc(1, QQ, "periodic")):
            assert ld.nFluid == 8 and list(ld.total[:8]) == list(range(1, 9))     # level-1 treeIDs 1..8
            cx = oracle.cx_dir(QQ)
            for left, right, axis in ((2, 1, 0), (6, 5, 0), (3, 1, 1), (1, 5, 2), (4, 8, 2)):
                want = -np.eye(3, dtype=cx.dtype)[axis]
                d = int(np.nonzero((cx == want).all(axis=1))[0][0])
                pos = int(ld.neigh[d * ld.nSize + (left - 1)])           # 1-based state position
                assert (pos - 1) // QQ == right - 1 and (pos - 1) % QQ == d, (QQ, left, right, axis)
