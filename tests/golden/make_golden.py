#!/usr/bin/env python3
"""Copies the reference's own golden result files that pin the hot path into
tests/golden/ (they are DATA of the reference's test-suite) so that the oracle and
the device path can be checked on machines without /root/reference.

  mus/examples/fluid/benchmark/gaussianPulse/reference/
     gaussianPulse_pressAlongLength_p00000_t10.001E+00.res
        fluid / bgk / d3q19, level 4, np=2, line sample at t=10.001 (9506 steps)
  mus/examples/fluid_incompressible/benchmark/TaylorGreenVortex/TGV_Simple/
     TGV_Simple_Re800/reference/TGV_Simple_Re800_probeAtCenter_p00000.res
        fluid_incompressible / mrt / d3q19, level 6 (64^3), np=12, centre probe every step
        (1962 samples of velocity_phy, pressure_phy)
     TGV_Simple_Re1600/reference/TGV_Simple_Re1600_kE_all_p00000.res
        fluid_incompressible / bgk / d3q19, level 7 (128^3), np=8, sum of kinetic_energy_phy
        over all elements every step (237 samples)

     gaussianPulse-L5_... / -L6_..._p0000{0,1,2}_t0.000E+00.res
        the same case at refinement levels 5 and 6, np=3, line sample of the INITIAL state (the
        three files are the three ranks' shares of the line)

  mus/examples/fluid_incompressible/benchmark/gaussianPulse/reference/  (stored with the prefix
  "incomp_": the file names are those of the fluid case)
     gaussianPulse_pressAlongLength_p00000_t10.001E+00.res              level 4, 9506 steps
     gaussianPulse-L5_..._t0.000E+00.res / _t10.000E+00.res             level 5, 19011 steps
     gaussianPulse-L6_..._t0.000E+00.res / _t10.000E+00.res             level 6, 38022 steps
        fluid_incompressible / bgk / d3q19, IC pressure = predefined 'gausspulse'
        (tem_ic_predefs_module.f90:230-255), line sample of the initial and the final state

  mus/examples/tutorials/tutorial_cases/tutorial_gaussian_pulse/ref/
     Gausspulse_track_pressure_p00000.res   (stored as tutorial_Gausspulse_track_pressure_p00000.res)
        fluid / bgk / d3q19, level 6 (64^3 -- the mesh and the kernel of BASELINE config 1), no physics
        table (lattice units), plane pressure pulse; point probe of density, pressure, velocity after
        every one of its 50 steps

Run in the build container only:  python tests/golden/make_golden.py
"""
import os
import shutil

EX = "/root/reference/mus/examples"
TGV = EX + "/fluid_incompressible/benchmark/TaylorGreenVortex/TGV_Simple"
HERE = os.path.dirname(os.path.abspath(__file__))
for d, f in ((EX + "/fluid/benchmark/gaussianPulse/reference",
              "gaussianPulse_pressAlongLength_p00000_t10.001E+00.res"),
             *[(EX + "/fluid/benchmark/gaussianPulse/reference",
                "gaussianPulse-L%d_pressAlongLength_p0000%d_t0.000E+00.res" % (lv, r))
               for lv in (5, 6) for r in range(3)],
             (TGV + "/TGV_Simple_Re800/reference", "TGV_Simple_Re800_probeAtCenter_p00000.res"),
             (TGV + "/TGV_Simple_Re1600/reference", "TGV_Simple_Re1600_kE_all_p00000.res")):
    shutil.copy(os.path.join(d, f), os.path.join(HERE, f))
    print("copied", f)
INC = EX + "/fluid_incompressible/benchmark/gaussianPulse/reference"
for f in sorted(os.listdir(INC)):
    shutil.copy(os.path.join(INC, f), os.path.join(HERE, "incomp_" + f))
    print("copied", f, "-> incomp_" + f)
TUT = EX + "/tutorials/tutorial_cases/tutorial_gaussian_pulse/ref"
shutil.copy(os.path.join(TUT, "Gausspulse_track_pressure_p00000.res"),
            os.path.join(HERE, "tutorial_Gausspulse_track_pressure_p00000.res"))
print("copied Gausspulse_track_pressure_p00000.res -> tutorial_Gausspulse_track_pressure_p00000.res")
