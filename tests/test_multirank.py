"""N > 1 paths.  CPU: world_size-2/4 gloo runs of the host-side halo logic (send/recv
position lists of the product's mesh generator moved over a real process boundary).
GPU: the same partitions through libmusb200 + NCCL, bit-compared with the single-domain oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(nproc, extra, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "parity_multi.py")] + extra
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("nproc,layout,kind", [(2, "d3q19", "periodic"), (4, "d3q27", "periodic"),
                                               (3, "d3q19", "cavity"), (4, "d3q19", "channel")])
def test_halo_lists_over_gloo(nproc, layout, kind):
    r = _launch(nproc, ["--mode", "lists", "--layout", layout, "--kind", kind, "--level", "4"], 29611 + nproc)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("halo links verified") == nproc


def _ngpu():
    try:
        import ctypes
        import musubi_b200._lib as L
        n = ctypes.c_int()
        return n.value if L.lib.musb200_device_count(ctypes.byref(n)) == 0 else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("layout,relax,kind", [("d3q27", "mrt", "periodic"), ("d3q19", "trt", "cavity"),
                                               ("d3q19", "bgk", "channel")])
def test_multi_gpu_matches_single_domain_oracle(layout, relax, kind):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 2 if n < 4 else 4
    r = _launch(nproc, ["--mode", "gpu", "--layout", layout, "--relaxation", relax, "--kind", kind,
                        "--level", "5", "--steps", "40"], 29651)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ndiff=0") == nproc
