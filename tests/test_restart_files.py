"""Restart FILES of the reference (tem_restart_module.f90, mus_restart_module.f90): header
script, binary layout, time stamps, partitioned reads -- musubi_b200/restart_io.py.

Known answers come from the reference itself: the header printed in
mus/examples/tutorials/tut_05_restart.md:39-68 (values restated below), and the time stamps in
the names of its golden result files ('..._t10.001E+00.res' after 9506 steps of the gaussianPulse
case, '..._t0.000E+00.res', C2D '..._t2.362E+00.res')."""
import math
import os

import numpy as np
import pytest

from musubi_b200 import restart_io as rio

# the header of the reference's restart tutorial (tut_05_restart.md:39-68), same keys and values
TUTORIAL_HEADER = """
 binary_name = {
    'restart/channel_9.021E-03.lsb'
}
 solver_configFile = 'musubi.lua'
 mesh = './mesh/'
 weights = ''
 time_point = {
    sim =    9.021097956087902E-03,
    iter = 5,
    clock =  103.734125999999996E-03
}
 nElems = 2048
 nDofs = 1
 solver = 'Musubi_v2.0'
 varsys = {
    systemname = 'fluid_incompressible',
    variable = {
        {
            name = 'pdf',
            ncomponents = 19,
            state_varpos = { 1, 2, 3, 4, 5, 6, 7, 8,
                9, 10, 11, 12, 13, 14, 15, 16,
                17, 18, 19 }
        }
    },
    nScalars = 19,
    nStateVars = 1,
    nAuxScalars = 4,
    nAuxVars = 2
}
"""


def test_time_stamp_is_fortran_en12_3():
    dx = 10.0 / 16
    dt = (1.0 / math.sqrt(3.0)) / 343.0 * dx
    assert rio.time_stamp(9506 * dt) == "10.001E+00"        # gaussianPulse_..._t10.001E+00.res
    assert rio.time_stamp(0.0) == "0.000E+00"
    assert rio.time_stamp(2.362) == "2.362E+00"
    assert rio.time_stamp(9.021097956087902e-03) == "9.021E-03"     # the tutorial's file names
    assert rio.time_stamp(999.9996) == "1.000E+03" and rio.time_stamp(0.9999996) == "1.000E+00"
    assert rio.time_stamp(123456.789) == "123.457E+03" and rio.time_stamp(-5.5e7) == "-55.000E+06"
    assert rio.fortran_en(103.734125999999996e-03, 15, 24) == " 103.734125999999996E-03"   # EN24.15
    assert rio.fortran_en(9.021097956087902e-03, 15, 24) == "   9.021097956087902E-03"
    with pytest.raises(ValueError):
        rio.time_stamp(float("nan"))


def test_reader_takes_the_tutorial_header():
    h = rio.parse_lua_assignments(TUTORIAL_HEADER)
    assert h["binary_name"] == ["restart/channel_9.021E-03.lsb"]
    assert h["mesh"] == "./mesh/" and h["weights"] == "" and h["solver"] == "Musubi_v2.0"
    assert h["time_point"] == {"sim": 9.021097956087902e-03, "iter": 5, "clock": 103.734125999999996e-03}
    assert h["nElems"] == 2048 and h["nDofs"] == 1
    v = h["varsys"]
    assert v["systemname"] == "fluid_incompressible" and v["nScalars"] == 19 and v["nAuxVars"] == 2
    assert v["variable"] == [{"name": "pdf", "ncomponents": 19, "state_varpos": list(range(1, 20))}]


def test_writer_reproduces_the_tutorial_header():
    """same keys, order, number formats and line structure as tem_restart_writeHeader printed"""
    txt = rio.header_text("restart/channel_9.021E-03.lsb",
                          dict(sim=9.021097956087902e-03, iter=5, clock=103.734125999999996e-03), 2048,
                          rio.fluid_varsys("fluid_incompressible", 19))
    assert txt.split() == TUTORIAL_HEADER.split()
    assert rio.parse_lua_assignments(txt) == rio.parse_lua_assignments(TUTORIAL_HEADER)


def test_lua_subset_parser_edge_cases():
    h = rio.parse_lua_assignments("""-- comment
      a = -1.5e3; b = "q'uote" c = { 1, 2; 3 }  --[[ block
      comment ]] d = { x = true, y = { }, 'pos' } e = 'it\\'s'
      mesh = { predefined = 'cube', origin = { 0.0, 0.0, 0.0 }, length = 10.0, refinementLevel = 4 }""")
    assert h["a"] == -1500.0 and h["b"] == "q'uote" and h["c"] == [1, 2, 3] and h["e"] == "it's"
    assert h["d"] == {"x": True, "y": [], 1: "pos"}
    assert h["mesh"]["predefined"] == "cube" and h["mesh"]["refinementLevel"] == 4
    with pytest.raises(ValueError):
        rio.parse_lua_assignments("a = math.sqrt(2)")        # expressions are not evaluated
    with pytest.raises(ValueError):
        rio.parse_lua_assignments("a = 'open")


def test_dump_layout_names_and_partitioned_read(tmp_path):
    QQ, n = 19, 37
    rng = np.random.default_rng(7)
    data = rng.random((n, QQ))
    vs = rio.fluid_varsys("fluid", QQ)
    time = dict(sim=0.25, iter=40)
    prefix = str(tmp_path / "restart") + os.sep
    # three "ranks" write their shares into the one file (rank 0 the headers), out of order
    parts = []
    for r in (2, 0, 1):
        share, rem = divmod(n, 3)
        off, cnt = r * share + min(r, rem), share + (1 if r < rem else 0)
        parts.append((off, cnt))
        b, hname = rio.write_restart(prefix, "pulse", data[off:off + cnt], time, vs, elem_offset=off,
                                     nElems_global=n, write_header=(r == 0),
                                     mesh=dict(predefined="cube", origin=(0.0, 0.0, 0.0), length=10.0,
                                               refinementLevel=4))
    assert sorted(os.listdir(tmp_path / "restart")) == [
        "pulse_250.000E-03.lsb", "pulse_header_250.000E-03.lua", "pulse_lastHeader.lua"]
    assert np.fromfile(b, dtype="<f8").tobytes() == data.tobytes()      # element-major, nothing else
    assert open(hname).read() == open(prefix + "pulse_lastHeader.lua").read()
    rf = rio.RestartFile(prefix + "pulse_lastHeader.lua", base_dir="/")   # binary found beside the header
    assert rf.nElems == n and rf.nScalars == QQ and rf.systemname == "fluid"
    assert rf.time == dict(sim=0.25, iter=40, clock=None)
    assert rf.mesh == dict(predefined="cube", origin=[0.0, 0.0, 0.0], length=10.0, refinementLevel=4)
    assert np.array_equal(rf.read(), data)
    # treelm's distribution: the first `remainder` ranks hold one element more
    assert [rf.part(r, 3) for r in range(3)] == [(0, 13), (13, 12), (25, 12)]
    for nranks in (1, 2, 3, 5):
        got = np.vstack([rio.read_restart(hname, r, nranks, base_dir="/")[2] for r in range(nranks)])
        assert np.array_equal(got, data)
    with pytest.raises(ValueError):
        rf.read(30, 10)
    with pytest.raises(ValueError):
        rio.write_restart(prefix, "pulse", data[:, :18], time, vs)       # not a multiple of nScalars
    with pytest.raises(ValueError):
        rio.write_restart(prefix, "pulse", data, time, vs, elem_offset=5, nElems_global=n)


def test_reader_rejects_a_truncated_binary_and_missing_keys(tmp_path):
    vs = rio.fluid_varsys("fluid", 19)
    prefix = str(tmp_path) + os.sep
    b, h = rio.write_restart(prefix, "x", np.zeros((4, 19)), dict(sim=1.0, iter=1), vs)
    with open(b, "r+b") as fh:
        fh.truncate(4 * 19 * 8 - 8)
    with pytest.raises(ValueError):
        rio.RestartFile(h)
    os.remove(b)
    with pytest.raises(FileNotFoundError):
        rio.RestartFile(h)
    bad = tmp_path / "bad.lua"
    bad.write_text("nElems = 4\n")
    with pytest.raises(ValueError):
        rio.RestartFile(str(bad))


def test_tree_order_equals_the_oracles_global_tree(oracle):
    from musubi_b200 import treelm_multilevel as tm
    for minL, boxes in ((4, [(5, 11)]), (4, [(4, 12), (12, 20)])):
        lv, _ = tm.build_multilevel(minL, boxes, QQ=19)
        tid, lp = rio.tree_order(lv)
        etid, elp = oracle.global_tree(lv)
        assert np.array_equal(tid, etid) and np.array_equal(lp, elp)
        # a leaf list in treelm order: parents' ranges never overlap, finest-descendant index increases
        assert tid.size == sum(L.nFluid for L in lv.values())


def test_oracle_state_through_a_restart_file_continues_identically(tmp_path, oracle):
    """two-level run: dump after 3 cycles through the file format, read it back on 1 and on 2
    "ranks" into a scheme that kept only its ghosts, continue: identical to the uninterrupted run"""
    from test_multilevel import build
    mo = oracle
    lv, intp, tables, ms = build(mo, 4, [(5, 11)], 19, "linear")
    ms.run(3)
    tid, lp = rio.tree_order(lv)
    buf = mo.pdf_serialize(ms.s, tid, lp)
    prefix = str(tmp_path / "restart") + os.sep
    _, hname = rio.write_restart(prefix, "twolevel", buf, dict(sim=3 * 0.5, iter=3),
                                 rio.fluid_varsys("fluid", 19))
    keep = {l: s.state[s.nNext].copy() for l, s in ms.s.items()}
    for nranks in (1, 2):
        for l, s in ms.s.items():                       # wipe the fluid PDFs, keep the ghosts
            s.state[s.nNext][:lv[l].nFluid * 19] = -1.0
        for r in range(nranks):
            rf, off, data = rio.read_restart(hname, r, nranks, base_dir="/")
            mo.pdf_unserialize(ms.s, tid[off:off + len(data)], lp[off:off + len(data)], data.ravel())
        assert rf.time["iter"] == 3 and rf.nElems == tid.size
        for l, s in ms.s.items():
            assert np.array_equal(s.state[s.nNext], keep[l])


@pytest.fixture(scope="module")
def mbgpu():
    import musubi_b200 as mb
    mb.mus_init(0, 1, 0)
    yield mb
    mb.mus_finalize()


@pytest.mark.gpu
def test_device_restart_file_round_trip_continues_bit_identically(mbgpu, oracle, tmp_path):
    """Scheme.write_restart after 4 cycles of a two-level run: the file's bytes equal the oracle's
    serialisation; a perturbed device state restored with Scheme.read_restart -- which, as
    mus_init_flow does after mus_readRestart, rebuilds auxField and ghosts from the stored fluid
    PDFs (musb200_fill_helper_elements) -- continues exactly like the oracle that does the same
    fill, and to rounding like the uninterrupted run."""
    from test_multilevel import build
    from musubi_b200._lib import check, lib
    mb, mo, QQ = mbgpu, oracle, 19
    lv, intp, tables, ms = build(mo, 4, [(5, 11)], QQ, "linear")
    ident = {"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}
    omega = {l: float(1.0 / (3.0 * s.visc[0] + 0.5)) for l, s in ms.s.items()}
    visc = {l: float(s.visc[0]) for l, s in ms.s.items()}
    sch = mb.Scheme(ident, lv, omega, omega_bulk=1.2, intp=(tables, intp["order"]), viscosity=visc)
    for l, s in ms.s.items():
        sch.upload_state(l, s.state[s.nNow], s.state[s.nNext])
        check(lib.musb200_aux_upload(l, s.aux.ctypes.data))
    sch.do_computation(4)
    ms.run(4)
    prefix = str(tmp_path / "restart") + os.sep
    bname, hname = sch.write_restart(prefix, "cyl", dict(sim=4.0, iter=4))
    tid, lp = rio.tree_order(lv)
    exp = mo.pdf_serialize(ms.s, tid, lp)
    assert open(bname, "rb").read() == exp.tobytes()
    rf = rio.RestartFile(hname, base_dir="/")
    assert rf.nElems == tid.size and rf.nScalars == QQ and rf.systemname == "fluid"
    sch.pdf_unserialize(tid, lp, exp + 1.0)                   # spoil the fluid PDFs
    t = sch.read_restart(hname, base_dir="/")
    assert t["iter"] == 4 and t["sim"] == 4.0
    assert sch.pdf_serialize(tid, lp).tobytes() == exp.tobytes()
    import copy
    cont = copy.deepcopy(ms)                                  # the uninterrupted run
    cont.run(3)
    ms.fill_helper_elements()                                 # the oracle's restart: same fill
    sch.do_computation(3)
    ms.run(3)
    got = sch.pdf_serialize(tid, lp)
    assert got.tobytes() == mo.pdf_serialize(ms.s, tid, lp).tobytes()
    ref = mo.pdf_serialize(cont.s, tid, lp)
    assert np.max(np.abs(got - ref) / np.abs(ref)) < 1e-11
    sch.destroy()


def _write_share(args):
    prefix, rank, nranks, nGlob, nScal = args
    from musubi_b200 import restart_io as rio
    lo, hi = rank * nGlob // nranks, (rank + 1) * nGlob // nranks
    data = (np.arange(lo, hi, dtype=np.float64)[:, None] * 100.0 + np.arange(nScal)[None, :]).ravel()
    vs = rio.fluid_varsys("fluid", nScal)
    rio.write_restart(prefix, "race", data, dict(sim=1.5, iter=3), vs, elem_offset=lo, nElems_global=nGlob,
                      write_header=(rank == 0))
    from musubi_b200 import treelm_io as tio
    tio.dump_weights(prefix + "weights.lsb", np.arange(lo, hi, dtype=np.float64), elem_offset=lo, nElems_global=nGlob)
    return hi - lo


def test_all_ranks_write_one_dump_concurrently(tmp_path):
    """every rank of a multi-rank dump opens the file at the same time: no rank may truncate
    what another has already written (the file is opened without O_TRUNC and written with
    pwrite at the rank's offset); repeated many times to give the race a chance"""
    import multiprocessing as mp
    nranks, nGlob, nScal = 8, 40000, 19
    ctx = mp.get_context("fork")
    for trial in range(6):
        prefix = str(tmp_path / ("t%d_" % trial))
        with ctx.Pool(nranks) as pool:
            pool.map(_write_share, [(prefix, r, nranks, nGlob, nScal) for r in range(nranks)])
        got = np.fromfile(prefix + "race_" + rio.time_stamp(1.5) + rio.ENDIAN_SUFFIX).reshape(nGlob, nScal)
        exp = np.arange(nGlob, dtype=np.float64)[:, None] * 100.0 + np.arange(nScal)[None, :]
        assert np.array_equal(got, exp)
        assert np.array_equal(np.fromfile(prefix + "weights.lsb"), np.arange(nGlob, dtype=np.float64))
