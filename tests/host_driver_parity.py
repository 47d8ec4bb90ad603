"""GPU pass of the compiled host (mus_b200_host): initial state from a file, N steps, restart dump
bit-compared with the oracle.  Run by tests/test_host_driver.py (gpu test) and
scripts/gpu_verify_1gpu.sh; green on a B200 in round 2 (it found the missing auxField
initialisation of the host, since closed by musb200_fill_helper_elements)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import musoracle as mo  # noqa: E402
from musubi_b200 import cases  # noqa: E402
import musubi_b200 as mb  # noqa: E402

EXE = os.path.join(ROOT, "musubi_b200", "mus_b200_host")
bad = 0
for mesh, layout, relax, kind, extra in (
        ("periodic", "d3q19", "bgk", "fluid", []),
        ("periodic", "d3q27", "mrt", "fluid", ["--omega-bulk", "1.3"]),
        ("cavity", "d3q19", "trt", "fluid", ["--lid", "0.05", "0.02", "0.0", "--lambda", "0.1875"]),
        ("channel", "d3q19", "bgk", "fluid_incompressible", ["--lid", "0.03", "0.0", "0.0"])):
    QQ, level, steps = (19 if layout == "d3q19" else 27), 5, 40
    ld = mo.build_level_desc(level, QQ, mesh)
    ref = mo.Scheme(ld, relax, kind, omega=1.7, lambda_=0.1875 if relax == "trt" else 0.25,
                    omega_bulk=1.3 if relax == "mrt" else 1.7)
    gld = mb.LevelDesc(level, QQ, mesh)
    if mesh == "periodic":
        rho, vel = cases.taylor_green(gld, mean=(0.01, -0.02, 0.015))
    else:
        rho, vel = cases.cavity_rest(gld)
        ref.bc_vel[2] = cases.lid_values(gld, (0.05, 0.02, 0.0) if mesh == "cavity" else (0.03, 0.0, 0.0))
        if mesh == "channel":
            ref.bc_kind[3] = "pressure_expol"
            ref.bc_rho[3] = 1.0
    ref.init_equilibrium(rho, vel)
    with tempfile.TemporaryDirectory() as d:
        fin, fout = os.path.join(d, "in.lsb"), os.path.join(d, "out.lsb")
        ref.state[ref.nNext][:ld.nElems * QQ].tofile(fin)
        r = subprocess.run([EXE, "--level", str(level), "--mesh", mesh, "--layout", layout, "--relaxation", relax,
                            "--kind", kind, "--omega", "1.7", "--steps", str(steps), "--check-interval", "20",
                            "--state-in", fin, "--state-out", fout] + extra, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
        if r.returncode != 0:
            print("FAILED rc=%d: %s" % (r.returncode, r.stderr[-500:]))
            bad += 1
            continue
        got = np.fromfile(fout)
    ref.run(steps)
    exp = ref.state[ref.nNext][:ld.nFluid * QQ]
    nd = int((got != exp).sum())
    print("host driver %s %s %s %s: ndiff=%d of %d" % (mesh, layout, relax, kind, nd, exp.size))
    bad += nd != 0
print("host driver parity:", "OK" if bad == 0 else "%d case(s) FAILED" % bad)
sys.exit(1 if bad else 0)
