"""Pins the CPU oracle against the reference's own golden vectors for this path, compared
exactly as the reference's pysys tests do: numpy.allclose(rtol=1e-10, atol=1e-5)
(pysys-extensions/apes/apeshelper.py:90-123):
  gaussianPulse      fluid / bgk / d3q19, level 4, periodic, np=2, 9506 steps, line sample
  TGV_Simple_Re800   fluid_incompressible / mrt / d3q19, 64^3, np=12, centre probe EVERY step
                     (1962 samples) -- also pins mus_init_pdf with the acoustic f_neq
  TGV_Simple_Re1600  fluid_incompressible / bgk / d3q19, 128^3, np=8, sum of kinetic energy
  gaussianPulse (fluid_incompressible/benchmark)  fluid_incompressible / bgk / d3q19 at levels 4
                     (9506 steps) and 5 (19011 steps), initial states at levels 5 and 6; level 6
                     (38022 steps of 64^3) runs on the device only (tests/test_gpu_golden.py)"""
import numpy as np
import pytest

from golden_cases import (GOLD_PULSE as GOLD, GOLD_TGV800, GOLD_TGV1600, gaussian_pulse_setup,
                          kinetic_energy_phy, pulse_line_elements, pulse_track, tgv800_row,
                          tgv800_setup, tgv1600_sample_steps, tgv1600_setup)


def track_line(sch, phys, bary):
    return pulse_track(sch.aux.reshape(-1, 4), pulse_line_elements(sch, bary), phys, bary)


def test_gaussian_pulse_matches_reference_golden(oracle):
    gold = np.loadtxt(GOLD, comments="#")
    sch, phys, bary, nsteps = gaussian_pulse_setup(oracle)
    assert nsteps == 9506
    m0 = sch.total_mass()
    sch.run(nsteps)
    got = track_line(sch, phys, bary)
    assert got.shape == gold.shape
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)
    # far tighter than the reference's own criterion on the well-conditioned columns
    assert np.max(np.abs(got[:, 3] / gold[:, 3] - 1.0)) < 1e-13     # density_phy
    assert np.max(np.abs(got[:, 4] / gold[:, 4] - 1.0)) < 1e-13     # pressure_phy
    assert np.max(np.abs(got[:, 5:] - gold[:, 5:])) < 2e-11         # velocity_phy (abs; fac_vel=594 => 3e-14 lattice, rounding noise)
    assert abs(sch.total_mass() / m0 - 1.0) < 1e-11   # 9506 steps of rounding; 1e-13 is checked per 100 steps elsewhere


def test_gaussian_pulse_two_ranks_identical(oracle):
    """the reference ran this case with np=2; the partitioned oracle must give the
    same line (halo exchange moves bytes only)."""
    gold = np.loadtxt(GOLD, comments="#")
    runs = [gaussian_pulse_setup(oracle, nranks=2, rank=r) for r in range(2)]
    schemes = [r[0] for r in runs]
    oracle.run_multi(schemes, 400)
    single, phys, bary, _ = gaussian_pulse_setup(oracle)
    single.run(400)
    ref = single.state[single.nNext].reshape(-1, 19)[:single.ld.nFluid]
    off = 0
    for s in schemes:
        got = s.state[s.nNext].reshape(-1, 19)[:s.ld.nFluid]
        assert np.array_equal(got, ref[off:off + s.ld.nFluid])
        off += s.ld.nFluid
    assert gold.shape == (16, 8)


def test_tgv_re800_probe_series_matches_reference_golden(oracle):
    """every one of the 1962 samples of the reference's centre probe (u, v, w, p)"""
    gold = np.loadtxt(GOLD_TGV800, comments="#")
    sch, phys, probe, nsteps, _ = tgv800_setup(oracle)
    assert nsteps == 1961 and gold.shape == (nsteps + 1, 5)
    rows = []
    for k in range(nsteps + 1):
        rows.append(tgv800_row(k, sch.aux.reshape(-1, 4)[probe], phys))
        if k < nsteps:
            sch.step()
    got = np.array(rows)
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)            # the reference's criterion
    assert np.max(np.abs(got[:, 1:4] - gold[:, 1:4])) < 5e-11       # velocity_phy, absolute
    assert np.max(np.abs(got[:, 4] / gold[:, 4] - 1.0)) < 1e-14     # pressure_phy
    assert np.max(np.abs(got[:, 0] - gold[:, 0])) < 1e-12           # time axis = k*dt


def test_tgv_re1600_kinetic_energy_matches_reference_golden(oracle):
    """the first 13 of the 237 samples (the GPU test covers all of them)"""
    gold = np.loadtxt(GOLD_TGV1600, comments="#")
    sch, phys, nsteps = tgv1600_setup(oracle)
    steps = tgv1600_sample_steps(gold, phys)
    assert nsteps == 471 and steps[-1] == nsteps and gold.shape == (237, 2)
    n = 24
    got = []
    for k in range(n + 1):
        if k in steps:
            got.append(kinetic_energy_phy(sch.aux.reshape(-1, 4), sch.ld.nFluid, phys))
        if k < n:
            sch.step()
    got = np.array(got)
    assert len(got) == 13
    assert np.allclose(got, gold[:len(got), 1], rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got / gold[:len(got), 1] - 1.0)) < 1e-12


@pytest.mark.parametrize("level", [5, 6])
def test_gaussian_pulse_initial_state_matches_reference_goldens_at_levels_5_and_6(oracle, level):
    """gaussianPulse-L5 / -L6 ..._t0.000E+00.res (three ranks' shares of the line): the initial
    condition (mus_init_pdf + initial auxField) and the unit conversion at two more levels"""
    from golden_cases import GOLD_PULSE_IC
    gold = np.vstack([np.loadtxt(f, comments="#", ndmin=2) for f in GOLD_PULSE_IC[level]])
    gold = gold[np.argsort(gold[:, 0])]
    sch, phys, bary, _ = gaussian_pulse_setup(oracle, level=level)
    got = pulse_track(sch.aux.reshape(-1, 4), pulse_line_elements(sch, bary, level), phys, bary)
    assert got.shape == gold.shape == (1 << level, 8)
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got[:, :3] - gold[:, :3])) < 1e-14       # barycentres
    assert np.max(np.abs(got[:, 3] / gold[:, 3] - 1.0)) < 1e-15    # density_phy
    assert np.max(np.abs(got[:, 4] / gold[:, 4] - 1.0)) < 1e-15    # pressure_phy
    assert np.all(got[:, 5:] == 0.0) and np.all(gold[:, 5:] == 0.0)


@pytest.mark.parametrize("level", [4, 5])
def test_incompressible_gaussian_pulse_matches_reference_golden(oracle, level):
    """mus/examples/fluid_incompressible/benchmark/gaussianPulse: the final line sample after
    ceil(10 / dt) steps (and the initial one where the reference ships it)"""
    from golden_cases import GOLD_PULSE_INCOMP
    ic, fin, steps = GOLD_PULSE_INCOMP[level]
    sch, phys, bary, nsteps = gaussian_pulse_setup(oracle, level=level, kind="fluid_incompressible")
    assert nsteps == steps
    sel = pulse_line_elements(sch, bary, level)
    if ic is not None:
        gold0 = np.loadtxt(ic, comments="#")
        got0 = pulse_track(sch.aux.reshape(-1, 4), sel, phys, bary)
        assert got0.shape == gold0.shape == (1 << level, 8)
        assert np.max(np.abs(got0[:, 3] / gold0[:, 3] - 1.0)) < 1e-15
        assert np.max(np.abs(got0[:, 4] / gold0[:, 4] - 1.0)) < 1e-15
    gold = np.loadtxt(fin, comments="#")
    sch.run(nsteps)
    got = pulse_track(sch.aux.reshape(-1, 4), sel, phys, bary)
    assert got.shape == gold.shape
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)
    assert np.max(np.abs(got[:, 3] / gold[:, 3] - 1.0)) < 1e-13     # density_phy
    assert np.max(np.abs(got[:, 4] / gold[:, 4] - 1.0)) < 1e-13     # pressure_phy
    assert np.max(np.abs(got[:, 5:] - gold[:, 5:])) < 2e-11         # velocity_phy


def test_incompressible_gaussian_pulse_initial_state_at_level_6(oracle):
    from golden_cases import GOLD_PULSE_INCOMP
    gold0 = np.loadtxt(GOLD_PULSE_INCOMP[6][0], comments="#")
    sch, phys, bary, nsteps = gaussian_pulse_setup(oracle, level=6, kind="fluid_incompressible")
    assert nsteps == GOLD_PULSE_INCOMP[6][2]
    got0 = pulse_track(sch.aux.reshape(-1, 4), pulse_line_elements(sch, bary, 6), phys, bary)
    assert got0.shape == gold0.shape == (64, 8)
    assert np.max(np.abs(got0[:, :3] - gold0[:, :3])) < 1e-14
    assert np.max(np.abs(got0[:, 3] / gold0[:, 3] - 1.0)) < 1e-15
    assert np.max(np.abs(got0[:, 4] / gold0[:, 4] - 1.0)) < 1e-15


def test_tutorial_gaussian_pulse_probe_series_matches_reference_golden(oracle):
    """the tutorial's reference run (mus/examples/tutorials/tutorial_cases/tutorial_gaussian_pulse):
    fluid / bgk / d3q19 on 64^3 -- BASELINE config 1's mesh and kernel -- in lattice units; density,
    pressure and velocity of one element after EVERY one of its 50 steps"""
    from golden_cases import GOLD_TUTORIAL_PULSE, tutorial_pulse_row, tutorial_pulse_setup
    gold = np.loadtxt(GOLD_TUTORIAL_PULSE, comments="#")
    sch, probe, nsteps = tutorial_pulse_setup(oracle)
    assert gold.shape == (nsteps, 6) and sch.ld.nFluid == 64 ** 3
    rows = []
    for k in range(1, nsteps + 1):
        sch.step()
        rows.append(tutorial_pulse_row(k, sch.aux.reshape(-1, 4)[probe]))
    got = np.array(rows)
    assert np.allclose(got, gold, rtol=1e-10, atol=1e-5)             # the reference's criterion
    assert np.array_equal(got[:, 0], gold[:, 0])                     # time axis: iterations, dt = 1
    assert np.max(np.abs(got[:, 1] / gold[:, 1] - 1.0)) < 1e-14      # density   (measured 2.4e-15)
    assert np.max(np.abs(got[:, 2] / gold[:, 2] - 1.0)) < 1e-14      # pressure  (2.1e-15)
    assert np.max(np.abs(got[:, 3:] - gold[:, 3:])) < 5e-15          # velocity, absolute (8e-16)
    assert np.max(np.abs(got[:, 3] / gold[:, 3] - 1.0)) < 1e-10      # u_x grows from 2e-6 to 2e-3: relative too
