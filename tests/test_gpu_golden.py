"""GPU vs the REFERENCE'S OWN golden vectors, through the C ABI: the device path is run on the
golden cases (initial state built by the host as mus_init_pdf does) and compared with the
committed .res files by the reference's criterion numpy.allclose(rtol=1e-10, atol=1e-5)
(pysys-extensions/apes/apeshelper.py:90-123), plus tighter bounds that hold for this build."""
import numpy as np
import pytest

from golden_cases import (GOLD_PULSE, GOLD_TGV800, GOLD_TGV1600, GOLD_TUTORIAL_PULSE, gaussian_pulse_setup,
                          kinetic_energy_phy, pulse_line_elements, pulse_track, tgv800_row,
                          tgv800_setup, tgv1600_sample_steps, tgv1600_setup, tutorial_pulse_row,
                          tutorial_pulse_setup)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import musubi_b200
    musubi_b200.mus_init(0, 1, 0)
    yield musubi_b200
    musubi_b200.mus_finalize()


def device_scheme(mb, ref, ident, level, omega_bulk=None):
    ld = mb.LevelDesc(level, 19, "periodic")
    assert np.array_equal(ld.neigh, ref.ld.neigh)
    omega = float(1.0 / (3.0 * ref.visc[0] + 0.5))
    sch = mb.Scheme(ident, ld, omega, omega_bulk=omega if omega_bulk is None else omega_bulk)
    sch.upload_state(level, ref.state[ref.nNow], ref.state[ref.nNext])
    return ld, sch


def test_gaussian_pulse_device_matches_reference_golden(mb, oracle):
    gold = np.loadtxt(GOLD_PULSE, comments="#")
    ref, phys, bary, nsteps = gaussian_pulse_setup(oracle)
    ld, sch = device_scheme(mb, ref, {"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, 4)
    sch.do_computation(nsteps)                       # 9506 steps in one call
    aux = sch.download_aux(4).reshape(-1, 4)
    got = pulse_track(aux, pulse_line_elements(ref, bary), phys, bary)
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got[:, 3] / gold[:, 3] - 1.0)) < 1e-13     # density_phy
    assert np.max(np.abs(got[:, 4] / gold[:, 4] - 1.0)) < 1e-13     # pressure_phy
    assert np.max(np.abs(got[:, 5:] - gold[:, 5:])) < 2e-11         # velocity_phy
    sch.destroy()


def test_tgv_re800_device_probe_series_matches_reference_golden(mb, oracle):
    """fluid_incompressible / mrt / d3q19 at 64^3: all 1962 samples of the centre probe."""
    gold = np.loadtxt(GOLD_TGV800, comments="#")
    ref, phys, probe, nsteps, omega_bulk = tgv800_setup(oracle)
    ident = {"kind": "fluid_incompressible", "relaxation": "mrt", "layout": "d3q19"}
    ld, sch = device_scheme(mb, ref, ident, 6, omega_bulk)
    rows = [tgv800_row(0, ref.aux.reshape(-1, 4)[probe], phys)]     # initAuxField of the host
    for k in range(1, nsteps + 1):
        sch.do_computation(1)
        rows.append(tgv800_row(k, sch.aux_probe(6, probe + 1), phys))
    got = np.array(rows)
    assert got.shape == gold.shape
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got[:, 1:4] - gold[:, 1:4])) < 5e-11
    assert np.max(np.abs(got[:, 4] / gold[:, 4] - 1.0)) < 1e-14
    # and the device equals the oracle bit for bit on the first 60 steps
    ref.run(60)
    ld2, sch2 = device_scheme(mb, tgv800_setup(oracle)[0], ident, 6, omega_bulk)
    sch2.do_computation(60)
    n = ld2.nFluid * 19
    assert np.array_equal(sch2.download_state(6)[:n], ref.state[ref.nNext][:n])
    sch.destroy()


def test_tutorial_gaussian_pulse_device_probe_series_matches_reference_golden(mb, oracle):
    """fluid / bgk / d3q19 at 64^3 in lattice units -- BASELINE config 1's mesh and kernel: the
    reference's point probe after every one of its 50 steps"""
    gold = np.loadtxt(GOLD_TUTORIAL_PULSE, comments="#")
    ref, probe, nsteps = tutorial_pulse_setup(oracle)
    ld, sch = device_scheme(mb, ref, {"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, 6)
    rows = []
    for k in range(1, nsteps + 1):
        sch.do_computation(1)
        rows.append(tutorial_pulse_row(k, sch.aux_probe(6, probe + 1)))
    got = np.array(rows)
    assert got.shape == gold.shape
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got[:, 1:3] / gold[:, 1:3] - 1.0)) < 1e-13      # density, pressure
    assert np.max(np.abs(got[:, 3:] - gold[:, 3:])) < 1e-13              # velocity, absolute
    # and the device equals the oracle bit for bit after the 50 steps
    ref.run(nsteps)
    n = ld.nFluid * 19
    assert np.array_equal(sch.download_state(6)[:n], ref.state[ref.nNext][:n])
    sch.destroy()


def test_tgv_re1600_device_kinetic_energy_matches_reference_golden(mb, oracle):
    """fluid_incompressible / bgk / d3q19 at 128^3: all 237 samples of the summed kinetic energy."""
    gold = np.loadtxt(GOLD_TGV1600, comments="#")
    ref, phys, nsteps = tgv1600_setup(oracle)
    steps = set(tgv1600_sample_steps(gold, phys).tolist())
    ident = {"kind": "fluid_incompressible", "relaxation": "bgk", "layout": "d3q19"}
    ld, sch = device_scheme(mb, ref, ident, 7)
    got = [kinetic_energy_phy(ref.aux.reshape(-1, 4), ld.nFluid, phys)]
    for k in range(1, nsteps + 1):
        sch.do_computation(1)
        if k in steps:
            got.append(kinetic_energy_phy(sch.download_aux(7).reshape(-1, 4), ld.nFluid, phys))
    got = np.array(got)
    assert got.shape == (gold.shape[0],)
    assert np.allclose(got, gold[:, 1], rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got / gold[:, 1] - 1.0)) < 1e-11
    sch.destroy()


@pytest.mark.parametrize("level", [4, 5, 6])
def test_incompressible_gaussian_pulse_device_matches_reference_golden(mb, oracle, level):
    """mus/examples/fluid_incompressible/benchmark/gaussianPulse at the reference's three
    resolutions (16^3 / 32^3 / 64^3, 9506 / 19011 / 38022 steps in one musb200_step call)."""
    from golden_cases import GOLD_PULSE_INCOMP
    _, fin, steps = GOLD_PULSE_INCOMP[level]
    gold = np.loadtxt(fin, comments="#")
    ref, phys, bary, nsteps = gaussian_pulse_setup(oracle, level=level, kind="fluid_incompressible")
    assert nsteps == steps
    ident = {"kind": "fluid_incompressible", "relaxation": "bgk", "layout": "d3q19"}
    ld, sch = device_scheme(mb, ref, ident, level)
    sch.do_computation(nsteps)
    aux = sch.download_aux(level).reshape(-1, 4)
    got = pulse_track(aux, pulse_line_elements(ref, bary, level), phys, bary)
    assert got.shape == gold.shape
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got[:, 3] / gold[:, 3] - 1.0)) < 1e-13     # density_phy
    assert np.max(np.abs(got[:, 4] / gold[:, 4] - 1.0)) < 1e-13     # pressure_phy
    assert np.max(np.abs(got[:, 5:] - gold[:, 5:])) < 2e-11         # velocity_phy
    sch.destroy()


def test_device_tracking_file_passes_the_references_own_check(mb, oracle, tmp_path):
    """the reference's regression test end to end on the device: run gaussianPulse, write the
    'pressAlongLength' tracking object in the reference's asciiSpatial format
    (musubi_b200/tracking.py), compare the FILE with the golden file as apeshelper.assertIsClose
    does (numpy.loadtxt + allclose(rtol=1e-10, atol=1e-5)); name, header, coordinates identical."""
    import os
    from musubi_b200 import tracking as tr
    ref, phys, bary, nsteps = gaussian_pulse_setup(oracle)
    ld, sch = device_scheme(mb, ref, {"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, 4)
    sch.do_computation(nsteps)
    aux = sch.download_aux(4).reshape(-1, 4)
    tid = np.asarray(ld.total[:ld.nFluid])
    sel = tr.select_line(tid, (0.0, 5.0, 5.0), (10.0, 0.0, 0.0), (0.0, 0.0, 0.0), 10.0)
    variables = ["density_phy", "pressure_phy", "velocity_phy"]
    name = tr.write_ascii_spatial(str(tmp_path) + os.sep, "gaussianPulse", "pressAlongLength", nsteps * phys.dt,
                                  tr.barycenters_of(tid[sel], (0.0, 0.0, 0.0), 10.0),
                                  tr.track(variables, aux[sel], tr.Physics(phys.dx, phys.dt, phys.rho0)), variables)
    assert os.path.basename(name) == os.path.basename(GOLD_PULSE)
    mine, gold = open(name).read().splitlines(), open(GOLD_PULSE).read().splitlines()
    assert mine[:2] == gold[:2] and [l[:76] for l in mine[2:]] == [l[:76] for l in gold[2:]]
    assert np.allclose(np.loadtxt(name, comments="#"), np.loadtxt(GOLD_PULSE, comments="#"), rtol=1e-10, atol=1e-5)
    sch.destroy()
