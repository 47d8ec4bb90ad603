"""Tracking files of the reference (libharvesting/hvs_ascii_module.f90) written by
musubi_b200/tracking.py: number format, headers, canoND element selection and the derived
variables, pinned by the reference's own golden .res files -- header lines and coordinate columns
byte for byte, values by the reference's criterion numpy.allclose(rtol=1e-10, atol=1e-5)
(pysys-extensions/apes/apeshelper.py:90-123) between the file WRITTEN HERE and the golden file."""
import math
import os

import numpy as np
import pytest

from golden_cases import (GOLD_PULSE, GOLD_PULSE_IC, GOLD_PULSE_INCOMP, GOLD_TGV800, GOLD_TGV1600,
                          gaussian_pulse_setup, tgv800_setup, tgv1600_sample_steps, tgv1600_setup)
from musubi_b200 import tracking as tr


def _tokens(path, limit=4000):
    out = []
    for line in open(path):
        if not line.startswith("#"):
            out += line.split()
            if len(out) > limit:
                break
    return out


@pytest.mark.parametrize("path", [GOLD_PULSE, GOLD_TGV800, GOLD_TGV1600, GOLD_PULSE_INCOMP[6][1]])
def test_e24_16e3_reproduces_every_number_the_reference_printed(path):
    toks = _tokens(path)
    assert len(toks) > 100
    for t in toks:
        assert tr.fortran_e(float(t)).strip() == t
    assert tr.fortran_e(0.0) == " 0.0000000000000000E+000" and len(tr.fortran_e(-1.5e-300)) == 24
    assert tr.fortran_e(0.99999999999999999) == " 0.1000000000000000E+001"


def test_line_selection_and_barycentres_equal_the_golden_coordinates(oracle):
    """canoND line origin (0, 5, 5) vec (10, 0, 0) lies ON cell faces at every level: the half-open
    cube test keeps the upper row only; coordinates printed identically at levels 4, 5, 6"""
    for level, files in ((4, [GOLD_PULSE]), (5, GOLD_PULSE_IC[5]), (6, GOLD_PULSE_IC[6])):
        ld = oracle.build_level_desc(level, 19, "periodic")
        tid = np.asarray(ld.total[:ld.nFluid])
        sel = tr.select_line(tid, (0.0, 5.0, 5.0), (10.0, 0.0, 0.0), (0.0, 0.0, 0.0), 10.0)
        assert sel.size == 1 << level
        bary = tr.barycenters_of(tid[sel], (0.0, 0.0, 0.0), 10.0)
        got = sorted(" ".join(tr.fortran_e(x) for x in b) for b in bary)
        gold = sorted(line[1:75] for f in files for line in open(f) if not line.startswith("#"))
        assert got == gold
    # a line that starts and ends on cell faces inside the mesh: the cell whose FAR face touches
    # the start is kept (t_far = 0 is not < 0, proj = 0), the one beginning at the end is not (proj < 1)
    ld = oracle.build_level_desc(3, 19, "periodic")
    tid = np.asarray(ld.total[:ld.nFluid])
    sel = tr.select_line(tid, (2.5, 0.0, 0.0), (5.0, 0.0, 0.0), (0.0, 0.0, 0.0), 10.0)
    assert sorted(tr.barycenters_of(tid[sel], (0, 0, 0), 10.0)[:, 0]) == [1.875, 3.125, 4.375, 5.625, 6.875]


def test_point_selection_is_the_cell_whose_lower_corner_is_the_point(oracle):
    sch, phys, probe, nsteps, _ = tgv800_setup(oracle)
    L = 2.0 * math.pi
    tid = np.asarray(sch.ld.total[:sch.ld.nFluid])
    assert tr.select_point(tid, (0.5 * L, 0.5 * L, 0.5 * L), (0.0, 0.0, 0.0), L) == probe
    assert tr.select_point(tid, (-1.0, 0.0, 0.0), (0.0, 0.0, 0.0), L) == 0            # clamped
    # leaves of two levels: a point in the coarse part finds the coarse leaf through its ancestors
    from musubi_b200 import treelm_multilevel as tm
    from musubi_b200.restart_io import tree_order
    lv, _ = tm.build_multilevel(4, [(5, 11)], QQ=19)
    leaves, _ = tree_order(lv)
    k = tr.select_point(leaves, (0.01, 0.01, 0.01), (0.0, 0.0, 0.0), 1.0)
    assert leaves[k] == tm.first_id(4)
    k = tr.select_point(leaves, (0.5, 0.5, 0.5), (0.0, 0.0, 0.0), 1.0)
    assert tr._level_of(int(leaves[k])) == 5


def test_pulse_line_file_written_here_passes_the_references_check(oracle, tmp_path):
    """fluid/benchmark/gaussianPulse end to end: 9506 steps, tracking object 'pressAlongLength'
    written as the reference writes it; file name, header and coordinates identical, values close"""
    sch, phys, _, nsteps = gaussian_pulse_setup(oracle)
    sch.run(nsteps)
    tid = np.asarray(sch.ld.total[:sch.ld.nFluid])
    sel = tr.select_line(tid, (0.0, 5.0, 5.0), (10.0, 0.0, 0.0), (0.0, 0.0, 0.0), 10.0)
    p = tr.Physics(phys.dx, phys.dt, phys.rho0)
    variables = ["density_phy", "pressure_phy", "velocity_phy"]
    vals = tr.track(variables, sch.aux.reshape(-1, 4)[sel], p)
    name = tr.write_ascii_spatial(str(tmp_path) + os.sep, "gaussianPulse", "pressAlongLength", nsteps * phys.dt,
                                  tr.barycenters_of(tid[sel], (0.0, 0.0, 0.0), 10.0), vals, variables)
    assert os.path.basename(name) == os.path.basename(GOLD_PULSE)
    mine, gold = open(name).read().splitlines(), open(GOLD_PULSE).read().splitlines()
    assert mine[:2] == gold[:2] and len(mine) == len(gold)
    assert [l[:76] for l in mine[2:]] == [l[:76] for l in gold[2:]]
    a, b = np.loadtxt(name, comments="#"), np.loadtxt(GOLD_PULSE, comments="#")
    assert np.allclose(a, b, rtol=1e-10, atol=1e-5)


def test_probe_series_file_has_the_references_header_and_rows(oracle, tmp_path):
    """TGV_Simple_Re800 'probeAtCenter' (ascii format): header identical, first rows close"""
    sch, phys, probe, nsteps, _ = tgv800_setup(oracle)
    p = tr.Physics(phys.dx, phys.dt, phys.rho0)
    variables = ["velocity_phy", "pressure_phy"]
    t = tr.AsciiTracker(str(tmp_path) + os.sep, "TGV_Simple_Re800", "probeAtCenter", variables)
    for k in range(6):
        t.dump(k * phys.dt, tr.track(variables, sch.aux.reshape(-1, 4)[probe], p, incompressible=True))
        sch.run(1)
    t.close()
    assert os.path.basename(t.name) == os.path.basename(GOLD_TGV800)
    mine, gold = open(t.name).read().splitlines(), open(GOLD_TGV800).read().splitlines()
    assert mine[:2] == gold[:2]
    a, b = np.loadtxt(t.name, comments="#"), np.loadtxt(GOLD_TGV800, comments="#")[:6]
    assert np.allclose(a, b, rtol=1e-10, atol=1e-5)
    # appended to after a restart: no second header
    t2 = tr.AsciiTracker(str(tmp_path) + os.sep, "TGV_Simple_Re800", "probeAtCenter", variables)
    t2.dump(6 * phys.dt, np.zeros(4))
    t2.close()
    assert sum(l.startswith("#") for l in open(t.name)) == 2 and len(open(t.name).readlines()) == 9
    # reduced variables carry the '_red' suffix (TGV_Simple_Re1600 'kE_all')
    r = tr.AsciiTracker(str(tmp_path) + os.sep, "TGV_Simple_Re1600", "kE_all", ["kinetic_energy_phy"], reduced=True)
    r.close()
    assert open(r.name).read().splitlines() == open(GOLD_TGV1600).read().splitlines()[:2]


def test_derive_rejects_unknown_variables_and_shapes(tmp_path):
    with pytest.raises(ValueError):
        tr.derive("wss_phy", np.zeros((1, 4)), tr.Physics(1.0, 1.0))
    with pytest.raises(ValueError):
        tr.write_ascii_spatial(str(tmp_path) + os.sep, "a", "b", 0.0, np.zeros((2, 3)), np.zeros((2, 2)),
                               ["velocity_phy"])


def test_reduced_series_file_equals_the_references_kinetic_energy_golden(oracle, tmp_path):
    """TGV_Simple_Re1600 'kE_all': shape = all, reduction = sum of kinetic_energy_phy, ascii
    format -- the first samples (steps 0, 2, 4) written here against the golden rows"""
    sch, phys, nsteps = tgv1600_setup(oracle)
    gold = np.loadtxt(GOLD_TGV1600, comments="#")
    steps = tgv1600_sample_steps(gold, phys)[:3]
    p = tr.Physics(phys.dx, phys.dt, phys.rho0)
    t = tr.AsciiTracker(str(tmp_path) + os.sep, "TGV_Simple_Re1600", "kE_all", ["kinetic_energy_phy"], reduced=True)
    k = 0
    for target in steps:
        sch.run(int(target) - k)
        k = int(target)
        ke = tr.derive("kinetic_energy_phy", sch.aux.reshape(-1, 4)[:sch.ld.nFluid], p, incompressible=True)
        t.dump(k * phys.dt, tr.reduce_spatial(ke, "sum"))
    t.close()
    assert os.path.basename(t.name) == os.path.basename(GOLD_TGV1600)
    assert open(t.name).read().splitlines()[:2] == open(GOLD_TGV1600).read().splitlines()[:2]
    assert np.allclose(np.loadtxt(t.name, comments="#"), gold[:3], rtol=1e-10, atol=1e-5)


def test_spatial_reductions():
    v = np.array([[1.0, -2.0], [3.0, 4.0], [-5.0, 0.5]])
    vol = np.array([1.0, 0.125, 0.125])
    assert list(tr.reduce_spatial(v, "sum")) == [-1.0, 2.5]
    assert np.allclose(tr.reduce_spatial(v, "average"), [-1.0 / 3, 2.5 / 3], rtol=1e-15)
    assert np.allclose(tr.reduce_spatial(v, "l2norm", vol), np.sqrt([1 + 9 / 8 + 25 / 8, 4 + 2 + 0.25 / 8]))
    assert np.allclose(tr.reduce_spatial(v, "l2normalized", vol),
                       np.sqrt(np.array([1 + 9 / 8 + 25 / 8, 4 + 2 + 0.25 / 8]) / 1.25))
    assert list(tr.reduce_spatial(v, "linfnorm")) == [5.0, 4.0]
    assert list(tr.reduce_spatial(v, "max")) == [3.0, 4.0] and list(tr.reduce_spatial(v, "min")) == [-5.0, -2.0]
    assert list(tr.reduce_spatial(np.array([1.0, 2.0]), "sum")) == [3.0]
    with pytest.raises(ValueError):
        tr.reduce_spatial(v, "median")


def test_number_formats_round_trip_random_doubles():
    """e24.16e3 carries 16 significant digits: reading a printed value back gives the double to
    within one unit of the 16th digit, and printing that again is a fixed point; EN24.15 (the
    restart header's reals) likewise with 16 to 18 digits"""
    from musubi_b200.restart_io import fortran_en
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.standard_normal(300) * 10.0 ** rng.integers(-300, 300, 300),
                        rng.random(100), [1.0, -1.0, 0.1, 999.9999999999999, 1e-320, 1.7e308]])
    for v in x:
        s = tr.fortran_e(v)
        assert len(s) == 24 and s[-5] == "E" and s.strip()[:3] in ("0.0", "0.1", "0.2", "0.3", "0.4", "0.5", "0.6",
                                                               "0.7", "0.8", "0.9", "-0.")
        back = float(s)
        assert back == v or abs(back / v - 1.0) < 1.0e-15
        assert tr.fortran_e(back) == s
        e = fortran_en(v, 15, 24)
        mant = abs(float(e.split("E")[0]))
        assert int(e.split("E")[1]) % 3 == 0 and (1.0 <= mant < 1000.0)
        backe = float(e)
        assert backe == v or abs(backe / v - 1.0) < 1.0e-15
        assert fortran_en(backe, 15, 24) == e
