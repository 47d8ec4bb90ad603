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
           os.path.join(ROOT, "tests", "parity_multi.py")] + extra
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("nproc,layout,kind", [(2, "d3q19", "periodic"), (3, "d3q19", "cavity"),
                                               (4, "d3q19", "channel"), (8, "d3q27", "periodic")])
def test_halo_lists_over_gloo(nproc, layout, kind):
    """incl. the 8-rank octant partition of bench.py --gpus 8 (cavity: face and edge peers;
    periodic D3Q27: face, edge and corner peers) and the host part of the peer-memory set-up"""
    r = _launch(nproc, ["--mode", "lists", "--layout", layout, "--kind", kind, "--level", "4"],
                29611 + nproc + (40 if layout == "d3q27" else 0))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("halo links verified") == nproc


@pytest.mark.parametrize("nproc,layout,levels", [(3, "d3q27", 2), (4, "d3q19", 3)])
def test_multilevel_halo_lists_over_gloo(nproc, layout, levels):
    r = _launch(nproc, ["--mode", "lists-ml", "--layout", layout, "--levels", str(levels)], 29631 + nproc)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("multi-level halo links verified") == nproc


def _ngpu():
    try:
        import ctypes
        import musubi_b200._lib as L
        n = ctypes.c_int()
        return n.value if L.lib.musb200_device_count(ctypes.byref(n)) == 0 else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("layout,relax,kind,extra", [
    ("d3q27", "mrt", "periodic", []), ("d3q19", "trt", "cavity", []), ("d3q19", "bgk", "channel", []),
    ("d3q27", "mrt", "periodic", ["--overlap"]),             # without peer memory: plain exchange
    ("d3q19", "trt", "cavity", ["--octants", "2"]),          # the weak-scaling mesh of bench.py
    ("d3q27", "mrt", "periodic", ["--p2p"]),                 # peer-memory halo exchange
    ("d3q19", "trt", "cavity", ["--octants", "2", "--p2p"]),
    ("d3q19", "bgk", "channel", ["--p2p", "--overlap"]),     # push overlapped with the halo-free CTAs
    ("d3q19", "trt", "cavity", ["--octants", "2", "--p2p", "--overlap"]),
    ("d3q27", "mrt", "periodic", ["--p2p", "--overlap"]),
    ("d3q27", "mrt", "periodic", ["--p2p", "--fused-push"]),      # links stored by the sweep itself
    ("d3q19", "trt", "cavity", ["--octants", "2", "--p2p", "--fused-push"]),
    ("d3q27", "mrt", "periodic", ["--p2p", "--sweep-wait"]),      # wait inside the next sweep (halo CTAs)
    ("d3q19", "trt", "cavity", ["--octants", "2", "--p2p", "--no-graphs"]),   # direct launches
    ("d3q19", "bgk", "channel", ["--p2p"])],                      # pressure boundary reads halo neighbours
    ids=["mrt27-periodic", "trt19-cavity", "bgk19-channel", "mrt27-periodic-overlap", "trt19-cavity-2oct",
         "mrt27-periodic-p2p", "trt19-cavity-2oct-p2p", "bgk19-channel-p2p-overlap", "trt19-cavity-2oct-p2p-overlap",
         "mrt27-periodic-p2p-overlap",
         "mrt27-periodic-p2p-fusedpush", "trt19-cavity-2oct-p2p-fusedpush", "mrt27-periodic-p2p-sweepwait",
         "trt19-cavity-2oct-p2p-nographs", "bgk19-channel-p2p"])
def test_multi_gpu_matches_single_domain_oracle(layout, relax, kind, extra):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 2 if n < 4 else 4
    if "--octants" in extra:
        nproc = 2
    r = _launch(nproc, ["--mode", "gpu", "--layout", layout, "--relaxation", relax, "--kind", kind,
                        "--level", "5", "--steps", "40"] + extra, 29651)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ndiff=0") == nproc


@pytest.mark.gpu
@pytest.mark.parametrize("layout,relax,levels,method,extra", [
    ("d3q19", "bgk", 2, "linear", []), ("d3q27", "mrt", 2, "quadratic", []), ("d3q19", "bgk", 3, "linear", []),
    ("d3q19", "bgk", 2, "linear", ["--p2p"]), ("d3q27", "mrt", 2, "quadratic", ["--p2p", "--no-graphs"]),
    ("d3q19", "bgk", 3, "linear", ["--p2p"]), ("d3q19", "bgk", 2, "linear", ["--restart"]),
    ("d3q19", "bgk", 3, "linear", ["--restart", "--p2p"]), ("d3q19", "bgk", 2, "linear", ["--ghost-exchange"])],
    ids=["2lvl-linear-bgk19", "2lvl-quad-mrt27", "3lvl-linear-bgk19", "2lvl-linear-bgk19-p2p",
         "2lvl-quad-mrt27-p2p-nographs", "3lvl-linear-bgk19-p2p", "2lvl-restart", "3lvl-restart-p2p",
         "2lvl-ghost-exchange"])
def test_multi_gpu_multilevel_matches_single_domain_oracle(layout, relax, levels, method, extra):
    """multi-level mesh partitioned along the global space-filling curve, one GPU per rank:
    state + auxField halo exchange per level through NCCL or peer memory (--p2p: one push kernel
    per level step, CUDA-graph replay of the cycle), ghosts interpolated locally; --restart: fluid
    PDFs into a fresh scheme + musb200_fill_helper_elements; --ghost-exchange: the reference's
    FromCoarser / FromFiner buffers"""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    nproc = 2 if n < 4 else 4
    r = _launch(nproc, ["--mode", "gpu-ml", "--layout", layout, "--relaxation", relax, "--levels", str(levels),
                        "--method", method, "--steps", "10"] + extra, 29653)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ndiff=0") >= nproc * (levels + 1)


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--sweep-wait"], ["--overlap"]],
                         ids=["wait-after-push", "wait-in-sweep", "overlapped"])
def test_exchange_timeout_surfaces_as_an_error_instead_of_a_hang(extra):
    """a rank that stops stepping: the others' waits give up after the configured timeout and the
    next synchronising call returns MUSB200_ERR_NCCL (the reference aborts all ranks through
    tem_abort, tem/source/tem_aux_module.f90:457-478)"""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _launch(2, ["--mode", "gpu-timeout", "--layout", "d3q19", "--relaxation", "bgk", "--kind", "periodic",
                    "--level", "5"] + extra, 29655)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("timeout reported") == 1 and "NO ERROR" not in r.stdout
