#!/usr/bin/env python3
"""bench.py -- MLUPS of the per-level LBM time step on N B200s (one process per
GPU) with the HBM roofline and the reference's CPU algorithm timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl musb200|reference]
                  [--workload cfg1|cfg2|cfg3|cfg3-256|cfg4|cfg5] [--level L]

Workloads (BASELINE.json configs):
  N = 1 : cfg2  D3Q19 TRT lid-driven cavity 256^3 (level 8), bounce-back walls +
                velocity_bounceback lid  -- the configuration the metric is quoted on
  N = 2, 4, 8 : the same cavity WEAK-scaled, 256^3 cells per GPU: the first N octants of the
                level-9 cube (512x256x256, 512x512x256, 512^3 -- the "512^3-equivalent mesh" of
                north_star at 8 GPUs), SFC-partitioned into N equal Morton ranges, halo exchange
                through peer memory over NVLink (NCCL send/recv with --no-p2p)
  Without --workload the line also carries
    check.multirank   (N > 1) the same workload kinds at level 6 run on N ranks and on rank 0 alone,
                      fluid PDFs bit-compared (no oracle involved), for every exchange path
    cfg3              BASELINE config 3, D3Q27 MRT periodic 512^3 STRONG-scaled over the N ranks
                      (N = 1: connectivity generated on the device, 76 GB on one B200)
  --workload cfg4 : BASELINE config 4, two-level octree with linear ghost interpolation; a step is
                one coarse cycle (1 coarse + 2 fine level steps); with --gpus N the mesh is cut along
                the global space-filling curve (strong scaling)
  --workload cfg5 : BASELINE config 5, three-level octree, D3Q19 BGK flow + passive scalar coupled
                on the device (two schemes stepped together); no reference behaviour exists, reported
                separately
  --workload cfg3 : the cfg3 block alone as the main line (also cfg3-256 on one GPU)
A "step" is one level time step (set_boundary, swap, fused aux+stream+collide,
halo exchange) over the whole mesh.  `value` is device-timed with inputs resident
in HBM; `e2e` goes through the public C ABI with HOST buffers: state upload from
pinned host memory, every step the host->device copy of the boundary values and
a device->host read of the tracked probe element, and the final state download.
"""
import argparse
import ctypes
import hashlib
import json
import math
import os
import re
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BYTES_PER_LUP = {19: 2 * 19 * 8 + 19 * 4, 27: 2 * 27 * 8 + 27 * 4}  # 380 / 540

WORKLOADS = {
    "cfg1": dict(ident={"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, level=6,
                 kind="periodic", omega=1.8, name="D3Q19 BGK Taylor-Green vortex 64^3 periodic"),
    "cfg2": dict(ident={"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}, level=8,
                 kind="cavity", omega=1.7, name="D3Q19 TRT lid-driven cavity 256^3, bounce-back walls"),
    "cfg3": dict(ident={"kind": "fluid", "relaxation": "mrt", "layout": "d3q27"}, level=9,
                 kind="periodic", omega=1.9, name="D3Q27 MRT periodic channel 512^3"),
    "cfg4": dict(ident={"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, level=7,
                 kind="multilevel", omega=1.7, boxes=[(32, 96)], cylinder=(128.0, 128.0, 16.0, 72, 184),
                 name="two-level octree: level-7 periodic cube, 64^3 coarse cells refined to level 8 "
                      "around a solid cylinder, D3Q19 BGK, linear ghost interpolation"),
    "cfg5": dict(ident={"kind": "fluid", "relaxation": "bgk", "layout": "d3q19"}, level=7,
                 kind="multilevel", omega=1.8, boxes=[(32, 96), (96, 160)], cylinder=None, scalar=True,
                 name="three-level octree (levels 7/8/9, nested 128^3-cell boxes), D3Q19 BGK flow + passive scalar "
                      "(bgk, first order) transported by it, linear ghost interpolation"),
    "cfg3-256": dict(ident={"kind": "fluid", "relaxation": "mrt", "layout": "d3q27"}, level=8,
                     kind="periodic", omega=1.9, name="D3Q27 MRT periodic channel 256^3"),
}
# the CPU arm runs the SAME configuration up to this level (256^3: 7 GB of host arrays, 0.2 s per
# step); beyond it (N > 1: 512x256x256 ... 512^3) it runs one GPU's share, 256^3, and says so
CPU_MAX_LEVEL = 8


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def kernel_source_hash():
    """identifies the sweep kernel a profile was taken of: sha256 over the sources it is built from"""
    h = hashlib.sha256()
    for f in ("sweep_kernel.cuh", "collide.cuh", "equilibrium.cuh", "kernels.cuh"):
        with open(os.path.join(ROOT, "musubi_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def measured_traffic(kernel, cells, path=None):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/traffic.json: entries keyed by kernel, cell count AND the hash of the kernel's
    sources at capture time); None when no capture matches -- a capture of an older kernel never
    passes for the current one."""
    try:
        entries = json.load(open(path or os.path.join(ROOT, "profiles", "traffic.json")))
        sha = kernel_source_hash()
        for t in entries:
            if t["kernel"] == kernel and int(t["cells"]) == int(cells) and t["source_sha"] == sha:
                return {"bytes": float(t["traffic_gb"]) * 1e9, "algorithmic_bytes": None,
                        "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)",
                        "source": t["source"], "source_sha": sha}
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """SM clock and clock-event reasons through NVML, in process, every ~2 ms from the warm-up
    on; `mark()` brackets the timed regions so that the median is taken under load."""

    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
               ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
               ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"),
               ("hw_power_brake", "nvmlClocksEventReasonHwPowerBrakeSlowdown"))

    def __init__(self, index=0, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.windows, self.stop_flag, self.err = [], [], False, None
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may remap local ranks
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                ids = [v for v in vis.split(",") if v.strip() != ""]
                if self.index < len(ids) and ids[self.index].strip().isdigit():
                    idx = int(ids[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                t = time.perf_counter()
                self.samples.append((t, float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), int(get_reasons(h))))
                time.sleep(self.period)
            nv.nvmlShutdown()
        except Exception as ex:          # no NVML on this host: the line says so
            self.err = repr(ex)

    def mark(self, t0, t1):
        self.windows.append((t0, t1))

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "error": self.err}
        import pynvml as nv
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in self.windows)]
        use = inside or self.samples
        bits = 0
        for s in use:
            bits |= s[2]
        reasons = [name for name, attr in self.REASONS if bits & int(getattr(nv, attr, 0))]
        return {"sm_mhz": float(np.median([s[1] for s in use])), "sm_max_mhz": self.max_mhz,
                "sm_min_mhz": float(min(s[1] for s in use)), "reasons": reasons,
                "samples": len(use), "samples_total": len(self.samples),
                "window": "timed regions" if inside else "warm-up + timed regions",
                "how": "NVML in-process, %.0f ms period" % (self.period * 1e3)}


# ---------------------------------------------------------------------------
def cpu_baseline(wl, level, steps, warmup=2, target_s=None):
    """the reference's CPU algorithm (oracle port: AOS, two-pass, per-element omega),
    OpenMP over the host cores, on a bounded sample of the same workload: `steps` level steps
    after `warmup` untimed ones; with target_s the step count is chosen from a short probe so
    that the timed part is about target_s seconds of CPU work."""
    from oracle import musoracle as mo
    QQ = 19 if wl["ident"]["layout"] == "d3q19" else 27
    t_setup = time.perf_counter()
    ld = mo.build_level_desc(level, QQ, wl["kind"])
    sch = mo.Scheme(ld, wl["ident"]["relaxation"], wl["ident"]["kind"], omega=wl["omega"],
                    lambda_=3.0 / 16.0, omega_bulk=wl["omega"])
    if wl["kind"] == "cavity":
        rho, vel = np.ones(ld.nElems), np.zeros((ld.nElems, 3))
        for bc in ld.bc:
            if bc["id"] == 2:
                sch.bc_vel[2] = np.tile(np.array([0.05, 0.0, 0.0]), (len(bc["links"]), 1))
    else:
        x = mo.barycenters(ld, (0.0, 0.0, 0.0), 2.0 * math.pi)
        u0 = 0.09 / math.sqrt(3.0)
        vel = np.stack([u0 * np.sin(x[:, 0]) * np.cos(x[:, 1]) * np.cos(x[:, 2]) + 0.05,
                        -u0 * np.cos(x[:, 0]) * np.sin(x[:, 1]) * np.cos(x[:, 2]),
                        np.zeros(ld.nElems)], axis=1)
        rho = np.ones(ld.nElems)
    sch.init_equilibrium(rho, vel)
    del rho, vel
    setup_s = time.perf_counter() - t_setup
    sch.run(max(1, warmup))
    if target_s is not None:
        t0 = time.perf_counter()
        sch.run(2)
        steps = int(min(1000, max(steps, target_s / ((time.perf_counter() - t0) / 2.0))))
    t0 = time.perf_counter()
    sch.run(steps)
    dt = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return dict(value=ld.nFluid * steps / dt / 1e6, unit="MLUPS", cores=cores, kind="port",
                sample="%s at level %d (%d^3 = %d cells), %d steps, oracle C port of the reference "
                       "algorithm with OpenMP (Fortran toolchain unavailable)" % (
                           re.sub(r",? \d+\^3.*", "", wl["name"]), level, 1 << level, ld.nFluid, steps),
                ms_per_step=dt / steps * 1e3, level=level, setup_s=round(setup_s, 1))


def run_reference(args, wl_name, wl, gpu_level):
    """the reference arm: nothing of the product is imported or loaded on this path"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        # torchrun pins every rank to one OpenMP thread; the CPU arm runs on rank 0 alone and
        # takes all host cores (read by libgomp when the oracle library is loaded below)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    level = min(gpu_level, CPU_MAX_LEVEL)
    t0 = time.perf_counter()
    cb = cpu_baseline(wl, level, max(1, args.steps), warmup=min(args.warmup, 3))
    same = (level == gpu_level and args.gpus == 1)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": cb["value"], "unit": "MLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak" if wl_name == "cfg2" else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name + ": " + wl["name"], "sample": cb["sample"], "level": level,
                   "cells": int((1 << level) ** 3), "relaxation": wl["ident"]["relaxation"],
                   "layout": wl["ident"]["layout"], "omega": wl["omega"],
                   "same_config_as_gpu_arm": same,
                   "note": ("the GPU arm's configuration, step for step" if same else
                            "N > 1: the GPU arm runs %d x 256^3 cells weak-scaled; the CPU arm (one host, rank 0) "
                            "runs one GPU's share, the 256^3 cavity -- MLUPS is size-normalised, the host does "
                            "not get faster with more cells" % args.gpus)},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
class Comm:
    """the host-side rendezvous of the ranks (gloo): what MPI is for the Fortran host"""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
            self.dist = dist

    def unique_id(self, mb):
        if self.dist is None:
            return None
        import torch
        t = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            t = torch.frombuffer(bytearray(mb.get_unique_id()), dtype=torch.uint8).clone()
        self.dist.broadcast(t, 0)
        return bytes(t.numpy().tobytes())

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def allred(self, x, op="MAX"):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t[0])

    def gather_arrays(self, a):
        """rank 0 receives every rank's float64 array (list by rank), others get None"""
        if self.dist is None:
            return [a]
        import torch
        n = torch.tensor([a.size], dtype=torch.int64)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
        self.dist.all_gather(sizes, n)
        out = None
        if self.rank == 0:
            out = [a]
            for r in range(1, self.world):
                buf = torch.zeros(int(sizes[r][0]), dtype=torch.float64)
                self.dist.recv(buf, r)
                out.append(buf.numpy())
        else:
            self.dist.send(torch.from_numpy(np.ascontiguousarray(a)), 0)
        return out

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def connect_p2p(comm, sch, level, want):
    """peer-memory halo exchange when every rank can do it (the choice is collective)"""
    from musubi_b200._lib import check, lib
    if comm.world == 1 or not want:
        return False
    ok = 1.0
    try:
        sch.p2p_connect(comm.dist, level)
    except Exception as ex:       # no peer access on this box: stay on the NCCL path
        sys.stderr.write("rank %d: peer-memory halo exchange unavailable (%s)\n" % (comm.rank, ex))
        ok = 0.0
    if comm.allred(ok, "MIN") > 0.5:
        return True
    check(lib.musb200_p2p_enable(level, 0))
    return False


def timed_steps(comm, sch, K, sampler=None):
    """exactly K steps bracketed by barrier + synchronize on both sides, device-timed with CUDA
    events on the library's stream, max over ranks"""
    from musubi_b200._lib import check, lib
    comm.barrier()
    sch.synchronize()
    t0 = time.perf_counter()
    check(lib.musb200_event_mark(0))
    sch.do_computation(K)
    check(lib.musb200_event_mark(1))
    sch.synchronize()
    t1 = time.perf_counter()
    comm.barrier()
    if sampler is not None:
        sampler.mark(t0, t1)
    ms = ctypes.c_double()
    check(lib.musb200_event_elapsed(ctypes.byref(ms)))
    return comm.allred(ms.value, "MAX")


# ---------------------------------------------------------------------------
def multirank_check(comm, mb, want_p2p):
    """N ranks against rank 0 alone, no oracle: the workload kinds of this bench at level 6
    (cavity: the first N octants, TRT D3Q19 + lid; periodic: D3Q27 MRT), 24 steps, every halo
    exchange path; the fluid PDFs of all ranks are gathered on rank 0 and compared bit for bit
    with the single-domain run of the same library (which the -m gpu tests pin to the oracle).
    Returns {"multirank_ndiff": total, "multirank": {case: ndiff}, ...}."""
    from musubi_b200 import cases
    from musubi_b200._lib import check, lib
    steps, level, out, t0 = 24, 6, {}, time.perf_counter()
    paths = [("p2p", dict(p2p=True, sweep_wait=0, graphs=1, overlap=0), "peer-memory push + wait kernel, CUDA graph"),
             ("p2p-overlap", dict(p2p=True, sweep_wait=0, graphs=0, overlap=1),
              "push on a second stream overlapping the next sweep, whose CTAs that pull from halo rows run last and wait"),
             ("p2p-sweepwait", dict(p2p=True, sweep_wait=1, graphs=0, overlap=0),
              "peer-memory push, wait inside the next sweep (halo CTAs), direct launches"),
             ("nccl", dict(p2p=False, sweep_wait=0, graphs=1, overlap=0), "pack / ncclSend / ncclRecv / unpack")]
    if not want_p2p:
        paths = paths[3:]
    for name, ident, kind, octants, omega in (
            ("cfg2", WORKLOADS["cfg2"]["ident"], "cavity", comm.world, 1.7),
            ("cfg3", WORKLOADS["cfg3"]["ident"], "periodic", 8, 1.9)):
        QQ = 19 if ident["layout"] == "d3q19" else 27
        gld = mb.LevelDesc(level, QQ, kind, 0, 1, octants=octants)          # the single domain
        rho, vel = cases.taylor_green(gld, mean=(0.01, -0.02, 0.015))
        init_g = cases.equilibrium_state(QQ, rho, vel, gld.nSize).reshape(-1, QQ)
        lid_g = cases.lid_values(gld, (0.05, 0.02, 0.0)) if kind == "cavity" else None
        ref = None
        if comm.rank == 0:                                                  # rank 0 alone, scheme slot 1
            one = mb.Scheme(ident, gld, omega, lambda_=3.0 / 16.0, omega_bulk=omega, slot=1)
            one.upload_state(level, init_g.ravel(), init_g.ravel())
            if lid_g is not None:
                one.set_bc_values(level, 2, lid_g)
            one.do_computation(steps)
            ref = one.download_state(level)[:gld.nFluid * QQ].reshape(-1, QQ)
            one.destroy()
        ld = mb.LevelDesc(level, QQ, kind, comm.rank, comm.world, octants=octants)
        gpos = (ld.total - gld.total[0]).astype(np.int64)
        init = np.zeros(ld.nSize * QQ)
        init[:ld.nElems * QQ] = init_g[gpos].ravel()
        for pname, opt, _ in paths:
            check(lib.musb200_set_sweep_wait(opt["sweep_wait"]))
            check(lib.musb200_set_graphs(opt["graphs"]))
            check(lib.musb200_set_overlap(opt["overlap"]))
            sch = mb.Scheme(ident, ld, omega, lambda_=3.0 / 16.0, omega_bulk=omega)
            sch.upload_state(level, init, init)
            on = connect_p2p(comm, sch, level, opt["p2p"])
            if lid_g is not None:
                sch.set_bc_values(level, 2, cases.lid_values(ld, (0.05, 0.02, 0.0)))
            sch.do_computation(steps)
            got = sch.download_state(level)[:ld.nFluid * QQ]
            sch.synchronize()
            comm.barrier()
            sch.destroy()
            parts = comm.gather_arrays(got)
            if comm.rank == 0:
                allv = np.concatenate(parts).reshape(-1, QQ)
                nd = int(np.count_nonzero(allv != ref)) if allv.shape == ref.shape else -1
                out["%s/%s%s" % (name, pname, "" if (on or not opt["p2p"]) else " (fell back to nccl)")] = nd
    check(lib.musb200_set_sweep_wait(0))
    check(lib.musb200_set_graphs(1))
    check(lib.musb200_set_overlap(0))
    if comm.rank != 0:
        return None
    return {"multirank_ndiff": int(sum(max(v, 0) for v in out.values()) + sum(1 for v in out.values() if v < 0)),
            "multirank": out,
            "multirank_how": "level 6 (cavity: first %d octants, D3Q19 TRT + lid; periodic: D3Q27 MRT), %d steps on "
                             "%d ranks vs the same library on rank 0 alone, fluid PDFs compared bit for bit; paths: %s"
                             % (comm.world, steps, comm.world, "; ".join("%s = %s" % (p[0], p[2]) for p in paths)),
            "multirank_s": round(time.perf_counter() - t0, 1)}


# ---------------------------------------------------------------------------
def run_single_level(args, comm, mb, wl_name, wl, octants, K, W, sampler, with_e2e, with_cpu, weak):
    """one single-level workload on the ranks of `comm`: returns the JSON line (rank 0) or None"""
    from musubi_b200 import cases
    from musubi_b200._lib import check, lib
    rank, world = comm.rank, comm.world
    ident, level, QQ = wl["ident"], wl["level"], (19 if wl["ident"]["layout"] == "d3q19" else 27)
    t_setup = time.perf_counter()
    device_mesh = (world == 1 and wl["kind"] == "periodic" and (8 ** level) * QQ > 2 ** 31 - 1)
    if device_mesh:
        # the 32-bit host list ends at nSize*QQ < 2^31: connectivity generated on the device
        ld = mb.DeviceCube(level, QQ, "periodic")
    else:
        ld = mb.LevelDesc(level, QQ, wl["kind"], rank, world, octants=octants)
    overlap = args.overlap and world > 1
    check(lib.musb200_set_overlap(1 if overlap else 0))
    check(lib.musb200_set_fused_push(1 if args.fused_push else 0))
    check(lib.musb200_set_sweep_wait(1 if args.sweep_wait else 0))
    nbytes = ld.nSize * QQ * 8
    sch = mb.Scheme(ident, ld, wl["omega"], lambda_=3.0 / 16.0, omega_bulk=wl["omega"])
    host_state, hp = None, None
    if wl["kind"] == "cavity":
        rho, vel = cases.cavity_rest(ld)
    else:
        rho, vel = cases.taylor_green(ld, mean=(0.05, 0.0, 0.0))
    if with_e2e:
        hp = ctypes.c_void_p()
        check(lib.musb200_host_alloc(nbytes, ctypes.byref(hp)))            # pinned host mirror of state
        host_state = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(ld.nSize * QQ,))
        host_state[:] = cases.equilibrium_state(QQ, rho, vel, ld.nSize)
        check(lib.musb200_state_upload(level, 2, host_state.ctypes.data))
        check(lib.musb200_set_now_next(level, 1, 2))
        check(lib.musb200_state_copy_next_to_now(level))
    else:
        sch.init_equilibrium(level, rho, vel)        # f_eq(rho, u) evaluated on the device
    del rho, vel
    p2p_on = connect_p2p(comm, sch, level, not args.no_p2p)
    halo_path = "none (1 rank)"
    if world > 1:
        if p2p_on and args.fused_push:
            halo_path = "peer-memory stores over NVLink fused into the sweep kernel + arrival flags"
        elif p2p_on:
            halo_path = "peer-memory stores over NVLink (one push kernel per step), " + (
                "on a second stream overlapping the next sweep, whose CTAs that pull from halo rows run last and "
                "wait for the arrival" if overlap else
                "arrival waited for inside the next sweep by the CTAs that pull from halo rows" if args.sweep_wait
                else "arrival waited for right after the push")
        else:
            halo_path = "NCCL send/recv (pack, group, unpack) after compute"
    lid = cases.lid_values(ld) if wl["kind"] == "cavity" else None
    lid_pinned = None
    if lid is not None and lid.size:
        lp = ctypes.c_void_p()
        check(lib.musb200_host_alloc(lid.nbytes, ctypes.byref(lp)))
        lid_pinned = np.ctypeslib.as_array(ctypes.cast(lp, ctypes.POINTER(ctypes.c_double)), shape=(lid.size,))
        lid_pinned[:] = lid.ravel()
        sch.set_bc_values(level, 2, lid_pinned)
    sch.synchronize()
    setup_s = time.perf_counter() - t_setup
    nFluid_total = comm.allred(float(ld.nFluid), "SUM")

    # ---------------- device-resident throughput ---------------------------
    sch.do_computation(W)
    sch.synchronize()
    # region 1: exactly K steps, nothing but the step's own launches on the stream -> `value`
    check(lib.musb200_timers_reset())
    t_ms = timed_steps(comm, sch, K, sampler)
    nl = ctypes.c_longlong()
    check(lib.musb200_launch_count(ctypes.byref(nl)))
    launches = int(nl.value)
    # region 2: the same K steps with CUDA events around every stage (per-kernel durations
    # for the roofline; the extra event records cost a fraction of a percent, so they are
    # kept out of `value`)
    check(lib.musb200_set_profiling(1))
    check(lib.musb200_timers_reset())
    t2_ms = timed_steps(comm, sch, K, sampler)
    cm, bm, com, im = (ctypes.c_double() for _ in range(4))
    check(lib.musb200_timers(ctypes.byref(cm), ctypes.byref(bm), ctypes.byref(com), ctypes.byref(im)))
    check(lib.musb200_set_profiling(0))
    sweep_ms = comm.allred(cm.value, "MAX") / K
    comm_ms = comm.allred(com.value, "MAX") / K
    mass, vmax, nan = sch.reduce()
    value = nFluid_total * K / (t_ms * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_LUP[QQ] * float(ld.nFluid) / (sweep_ms * 1e-3) / 1e9   # per GPU, dominant kernel
    kernel_name = "sweepKernel<%d,%s>" % (QQ, ident["relaxation"])
    traffic = measured_traffic(kernel_name, ld.nFluid)
    if traffic is not None:
        traffic["algorithmic_bytes"] = BYTES_PER_LUP[QQ] * float(ld.nFluid)

    # ---------------- end to end through the C ABI with host buffers --------
    e2e = None
    if with_e2e:
        # the probe is tracked every iteration: lazy auxField, the tracked element's moments are
        # computed on demand (musb200_aux_probe) instead of materialising auxField every step
        check(lib.musb200_set_aux_every_step(2))
        probe = np.zeros(4)
        Ke = K
        comm.barrier()
        sch.synchronize()
        t0 = time.perf_counter()
        check(lib.musb200_state_upload(level, 2, host_state.ctypes.data))
        check(lib.musb200_set_now_next(level, 1, 2))
        check(lib.musb200_state_copy_next_to_now(level))
        if lid_pinned is not None:
            check(lib.musb200_bc_set_values(level, 2, lid_pinned.size, lid_pinned.ctypes.data))
        sch.synchronize()
        t_up = time.perf_counter()
        for it in range(Ke):
            check(lib.musb200_step(level, level, 1))
            if lid_pinned is not None and it + 1 < Ke:
                # the next step's boundary values go up while this step runs (copy stream)
                check(lib.musb200_bc_set_values(level, 2, lid_pinned.size, lid_pinned.ctypes.data))
            check(lib.musb200_aux_probe(level, 1, probe.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        t_st = time.perf_counter()
        check(lib.musb200_state_download(level, sch.now_next(level)[1], host_state.ctypes.data))
        sch.synchronize()
        t_dn = time.perf_counter()
        comm.barrier()
        dt = comm.allred(time.perf_counter() - t0, "MAX")
        up_s, st_s, dn_s = (comm.allred(x, "MAX") for x in (t_up - t0, t_st - t_up, t_dn - t_st))
        bc_bytes = int(lid_pinned.nbytes) if lid_pinned is not None else 0
        e2e = {"value": nFluid_total * Ke / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(nbytes / Ke + bc_bytes), "d2h_bytes_per_step": int(nbytes / Ke + 32),
               "steps": Ke, "wall_s": dt, "pcie_s": up_s + dn_s, "compute_s": st_s,
               "pcie_share": (up_s + dn_s) / max(dt, 1e-12),
               "pcie_gbs_per_gpu": 2.0 * nbytes / max(up_s + dn_s, 1e-12) / 1e9,
               "protocol": "pinned-host state upload (pcie_s) + K x (BC values H2D on the copy stream, level "
                           "step, probe element computed on demand + D2H) (compute_s) + state download (pcie_s); "
                           "with few steps this is a PCIe figure: %d ranks share the host's links" % world}
        check(lib.musb200_set_aux_every_step(0))

    cb = None
    if rank == 0 and with_cpu:
        try:    # about 10 s of CPU work on the host cores, the SAME configuration (N = 1 only)
            cb = cpu_baseline(wl, min(level, CPU_MAX_LEVEL), 10, target_s=10.0)
            cb["same_config"] = (cb["level"] == level)
        except Exception as ex:  # the oracle is optional test infrastructure
            cb = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
    line = None
    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_ms / K, "higher_is_better": True,
            "scaling": "weak" if weak else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name + ": " + wl["name"], "level": level, "cells": int(nFluid_total),
                       "partition": "treelm SFC, %d equal Morton ranges" % world,
                       "cells_per_gpu": int(nFluid_total / world),
                       "mesh": ("predefined cube, connectivity generated on the device" if device_mesh else
                                "host lists (C++ generator) through the C ABI"),
                       "halo_exchange": halo_path,
                       "relaxation": ident["relaxation"], "layout": ident["layout"], "omega": wl["omega"],
                       "l2": "state of %.2f GB per buffer per GPU >> 126 MB L2, no flush needed" % (nbytes / 1e9),
                       "aux_every_step": False, "setup_s": round(setup_s, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (traffic["bytes"] if traffic else None),
                         "traffic_detail": traffic, "peak_source": peak_src,
                         "kernel": kernel_name, "kernel_source_sha": kernel_source_hash(),
                         "bytes_per_lup": BYTES_PER_LUP[QQ], "kernel_ms": sweep_ms,
                         "share_of_step": sweep_ms / (t2_ms / K),
                         "timed": "CUDA events around every sweep launch of a second K-step region "
                                  "(%.4f ms per step with the stage events in)" % (t2_ms / K)},
            "timers_ms_per_step": {"compute": sweep_ms, "bc": bm.value / K, "comm": comm_ms,
                                   "intp": im.value / K,
                                   "note": "max over ranks; with the peer-memory exchange `comm` is the push "
                                           "kernel, the wait for the peers' links sits inside `compute`"},
            "cpu_baseline": cb, "e2e": e2e, "gpu_launches": launches,
            "check": {"total_mass": mass, "max_vel": vmax, "nan": nan},
        }
    sch.synchronize()
    comm.barrier()                    # peers may still store into this rank's halo rows
    sch.destroy()
    if hp is not None:
        check(lib.musb200_host_free(hp))
    return line


# ---------------------------------------------------------------------------
def run_multilevel(args, comm, mb, wl_name, wl):
    """cfg4: K coarse cycles of do_recursive_multiLevel, device timed; on N > 1 ranks the mesh is
    cut along the global space-filling curve (strong scaling), halo exchange per level through
    peer memory (state + auxField in one push kernel) or NCCL."""
    from musubi_b200 import cases
    from musubi_b200 import treelm_multilevel as tm
    from musubi_b200._lib import check, lib
    world, rank = comm.world, comm.rank
    W, K = max(3, args.warmup), max(1, args.steps)
    t_setup = time.perf_counter()
    minL = args.level or WORKLOADS[wl_name]["level"]
    scale = 2.0 ** (minL - WORKLOADS[wl_name]["level"])      # boxes / cylinder are given at the base level
    boxes = [(int(lo * scale), int(hi * scale)) for lo, hi in wl["boxes"]]
    cyl = None
    if wl.get("cylinder"):
        cyl = tuple(c * scale for c in wl["cylinder"][:3]) + tuple(int(c * scale) for c in wl["cylinder"][3:])
    glob, intp = tm.build_multilevel(minL, boxes, QQ=19, cylinder=cyl, intp_method="linear")
    weights = tm.level_weights(glob) if args.balance else None      # SPartA cut by level steps per cycle
    lv = glob if world == 1 else tm.partition_multilevel(glob, world, weights=weights)[rank]
    tables = mb.multilevel_tables(lv, intp)
    levels = sorted(lv)
    nu0 = (1.0 / wl["omega"] - 0.5) / 3.0
    visc = {l: nu0 * 2.0 ** (l - levels[0]) for l in levels}       # acoustic scaling
    omega = {l: 1.0 / (3.0 * visc[l] + 0.5) for l in levels}
    sch = mb.Scheme(wl["ident"], lv, omega, intp=(tables, intp["order"]), viscosity=visc)
    scalar = None
    if wl.get("scalar"):
        # BASELINE config 5: a passive scalar in scheme slot 1, transported by the flow of slot 0
        # (velocity read from the flow's auxField on the device), diffusivity scaled acoustically
        diff = {l: 0.02 * 2.0 ** (l - levels[0]) for l in levels}
        scalar = mb.Scheme({"kind": "passive_scalar", "relaxation": {"name": "bgk", "variant": "first"},
                            "layout": "d3q19"}, lv, species={"diff_coeff": diff, "lambda": 0.25},
                           intp=(tables, intp["order"]), slot=1)
    host, nbytes = {}, 0
    for l in levels:
        L = lv[l]
        x = 2.0 * math.pi * L.bary_unit
        u0 = 0.03
        vel = np.stack([u0 * np.sin(x[:, 0]) * np.cos(x[:, 1]) * np.cos(x[:, 2]) + 0.02,
                        -u0 * np.cos(x[:, 0]) * np.sin(x[:, 1]) * np.cos(x[:, 2]),
                        np.zeros(L.nElems)], axis=1)
        host[l] = cases.equilibrium_state(19, np.ones(L.nElems), vel, L.nSize)
        nbytes += host[l].nbytes
        sch.upload_state(l, host[l])
        aux = np.zeros(L.nSize * 4)
        aux[:L.nElems * 4] = np.concatenate([np.ones((L.nElems, 1)), vel], axis=1).ravel()
        sch._bind()
        check(lib.musb200_aux_upload(l, aux.ctypes.data))
        if scalar is not None:
            r2 = ((L.bary_unit - 0.5) ** 2).sum(axis=1)
            conc = 1.0 + 0.5 * np.exp(-r2 / 0.02)
            scalar.upload_state(l, cases.equilibrium_state(19, conc, np.zeros((L.nElems, 3)), L.nSize))
            scalar.couple_transport_velocity(l, sch)
    p2p_on = False
    if world > 1 and not args.no_p2p:
        p2p_on = all([connect_p2p(comm, sc, l, True) for sc in ([sch] + ([scalar] if scalar else [])) for l in levels])
    nsch = 2 if scalar is not None else 1

    def do_cycles(n):
        if scalar is None:
            sch.do_computation(n)
        else:
            mb.step_schemes([sch, scalar], n)

    class _Stepper:               # what timed_steps drives
        do_computation = staticmethod(do_cycles)
        synchronize = staticmethod(sch.synchronize)
    sch.synchronize()
    setup_s = time.perf_counter() - t_setup
    upd = {l: 2 ** (l - levels[0]) for l in levels}                # level steps per coarse cycle
    lups_cycle = nsch * comm.allred(float(sum(lv[l].nFluid * upd[l] for l in levels)), "SUM")
    solve_cycle = nsch * sum((lv[l].nFluid + lv[l].nGhostFromCoarser) * upd[l] for l in levels)   # this rank

    do_cycles(W)
    sch.synchronize()
    sampler = ClockSampler(comm.local_rank)
    if rank == 0:
        sampler.start()
    check(lib.musb200_timers_reset())
    t_ms = timed_steps(comm, _Stepper, K, sampler)      # region 1: nothing but the cycle's own launches
    nl = ctypes.c_longlong()
    check(lib.musb200_launch_count(ctypes.byref(nl)))
    check(lib.musb200_set_profiling(1))         # region 2: CUDA events around every stage
    check(lib.musb200_timers_reset())
    t2_ms = timed_steps(comm, _Stepper, K, sampler)
    clocks = sampler.finish() if rank == 0 else None
    cm, bm, com, im = (ctypes.c_double() for _ in range(4))
    check(lib.musb200_timers(ctypes.byref(cm), ctypes.byref(bm), ctypes.byref(com), ctypes.byref(im)))
    check(lib.musb200_set_profiling(0))
    value = lups_cycle * K / (t_ms * 1e-3) / 1e6
    # the reference's own figure for the same run (mus_perf_measure / calc_MLUPS,
    # mus_tools_module.f90:498-531, 658-691): coarse levels scaled by 1 / sf^(maxLevel - l), the
    # main loop's iteration count (coarse cycles) divided by sf^(maxLevel - minLevel) once more
    from musubi_b200 import timing
    value_ref_formula = timing.perf_measure({l: int(glob[l].nFluid) for l in levels}, K, t_ms * 1e-3,
                                            max(cm.value, 1e-9) * 1e-3)[0]
    sweep_ms = cm.value / K
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_LUP[19] * float(solve_cycle) / (sweep_ms * 1e-3) / 1e9
    mass = sum(sch.reduce(l)[0] / 8.0 ** (l - levels[0]) for l in levels) if world == 1 else None
    smass = (sum(scalar.reduce(l)[0] / 8.0 ** (l - levels[0]) for l in levels)
             if (world == 1 and scalar is not None) else None)

    e2e = None
    if not args.no_e2e and scalar is None:
        comm.barrier()
        t0 = time.perf_counter()
        for l in levels:
            sch.upload_state(l, host[l])
        sch.do_computation(K)
        for l in levels:
            host[l] = sch.download_state(l)
        sch.synchronize()
        comm.barrier()
        dt = comm.allred(time.perf_counter() - t0, "MAX")
        e2e = {"value": lups_cycle * K / dt / 1e6, "unit": "MLUPS", "h2d_bytes_per_step": int(2 * nbytes / K),
               "d2h_bytes_per_step": int(nbytes / K), "steps": K, "wall_s": dt,
               "protocol": "state upload of every level (pageable host arrays) + K coarse cycles + state download"}
    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name + ": " + wl["name"], "levels": levels,
                       "cells": {str(l): int(glob[l].nFluid) for l in levels},
                       "rank0": {str(l): {"fluid": int(lv[l].nFluid), "ghostFromCoarser": int(lv[l].nGhostFromCoarser),
                                          "ghostFromFiner": int(lv[l].nGhostFromFiner), "halo": int(lv[l].nHalo)}
                                 for l in levels},
                       "partition": "global space-filling curve over all levels, %d %s ranges; ghosts "
                                    "interpolated locally, fluid-only halos (state + auxField) through %s" % (
                                        world, "SPartA-weighted" if args.balance else "equal",
                                        "peer memory, one push kernel per level step" if p2p_on else "NCCL"),
                       "step": "one coarse cycle = %s level steps" % "+".join(str(upd[l]) for l in levels),
                       "mlups_definition": "sum_l nFluid(l) * 2^(l-minLevel) per coarse cycle / time (SURVEY 8d)" + (
                           ", flow and scalar updates both counted (two schemes)" if scalar is not None else ""),
                       "schemes": (["fluid bgk d3q19 (slot 0)", "passive_scalar bgk/first d3q19 (slot 1), transport "
                                    "velocity = the flow's auxField of the same level step, ghosts by the reference's "
                                    "arbitrary-value interpolation of its PDFs"] if scalar is not None else
                                   ["fluid bgk d3q19"]),
                       "reference_behaviour": ("none: the reference aborts for a passive scalar on a multi-level mesh "
                                               "(mus_scheme_module.f90:166-190); parity is against the oracle's "
                                               "coupled scheme only" if scalar is not None else "do_recursive_multiLevel"),
                       "mlups_by_reference_formula": value_ref_formula,
                       "interpolation": "linear", "omega": {str(l): omega[l] for l in levels},
                       "l2": "state %.2f GB per rank > 126 MB L2" % (2 * nbytes / 1e9), "setup_s": round(setup_s, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "kernel": "sweepKernel<19,bgk>%s (all level steps of a cycle, rank 0)" % (
                             " + passiveScalarKernel<19>" if scalar is not None else ""),
                         "bytes_per_lup": BYTES_PER_LUP[19], "kernel_ms": sweep_ms,
                         "share_of_step": sweep_ms / (t2_ms / K)},
            "timers_ms_per_step": {"compute": cm.value / K, "bc": bm.value / K, "comm": com.value / K,
                                   "intp": im.value / K},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(nl.value), "clocks": clocks,
            "check": {"total_mass": mass, "scalar_mass": smass},
        }
        print(json.dumps(line), flush=True)
    sch.synchronize()
    comm.barrier()
    if scalar is not None:
        scalar.destroy()
    sch.destroy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="musb200", choices=["musb200", "reference"])
    ap.add_argument("--workload", default=None, choices=list(WORKLOADS))
    ap.add_argument("--level", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg3 block of the default line")
    ap.add_argument("--no-check", action="store_true", help="skip the multi-rank bit-compare of the default line")
    ap.add_argument("--overlap", action="store_true",
                    help="peer-memory halo exchange overlapped with the next sweep: push on a second stream, the "
                         "CTAs that pull from halo rows moved to the end of the launch (measured 2 %% slower than "
                         "push + wait kernel after compute at 256^3 per GPU, which stays the default)")
    ap.add_argument("--fused-push", action="store_true",
                    help="peer-memory halo exchange with the link stores fused into the sweep kernel")
    ap.add_argument("--sweep-wait", action="store_true",
                    help="peer-memory halo exchange: wait inside the next sweep (halo CTAs) instead of a wait kernel")
    ap.add_argument("--balance", action="store_true",
                    help="cfg4 on N > 1 ranks: cut the space-filling curve by tem_balance_sparta with level "
                         "weights 2^(l - minLevel) instead of equal element counts")
    ap.add_argument("--no-p2p", action="store_true",
                    help="halo exchange through pack/ncclSend/ncclRecv/unpack instead of peer memory")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl_name = args.workload or "cfg2"
    wl = dict(WORKLOADS[wl_name])
    octants = 8
    if wl_name == "cfg2" and world > 1:
        if world not in (2, 4, 8):
            raise SystemExit("the weak-scaled cavity needs 1, 2, 4 or 8 ranks (whole octants per rank)")
        wl["level"], octants = wl["level"] + 1, world       # 256^3 cells per GPU
        wl["name"] = "D3Q19 TRT lid-driven cavity, 256^3 cells per GPU (first %d octants of 512^3)" % world
    if args.level:
        wl["level"] = args.level
    weak = (wl_name == "cfg2")
    if args.impl == "reference":
        if wl["kind"] == "multilevel":
            raise SystemExit("--impl reference times the single-level workloads")
        # the CPU arm runs on one host: the GPU arm's own configuration up to 256^3, beyond that
        # (weak-scaled N > 1, cfg3) one GPU's share of it
        run_reference(args, wl_name, dict(WORKLOADS[wl_name]), wl["level"])
        return
    W = max(3, args.warmup)
    K = max(1, args.steps)

    comm = Comm()
    import musubi_b200 as mb
    mb.mus_init(comm.rank, world, comm.local_rank, comm.unique_id(mb))
    if wl["kind"] == "multilevel":
        run_multilevel(args, comm, mb, wl_name, wl)
        mb.mus_finalize()
        comm.close()
        return

    full = args.workload is None and args.level is None      # the driver's invocation: the whole line
    check_block = None
    if full and world > 1 and not args.no_check:
        check_block = multirank_check(comm, mb, not args.no_p2p)
    sampler = ClockSampler(comm.local_rank)
    if comm.rank == 0:
        sampler.start()
    line = run_single_level(args, comm, mb, wl_name, wl, octants, K, W, sampler,
                            with_e2e=not args.no_e2e, with_cpu=(world == 1 and not args.no_cpu_baseline), weak=weak)
    clocks = sampler.finish() if comm.rank == 0 else None
    cfg3 = None
    if full and not args.no_cfg3:
        # BASELINE config 3 strong-scaled over the same ranks: fewer steps (11 ms each at N = 1)
        s3 = ClockSampler(comm.local_rank)
        if comm.rank == 0:
            s3.start()
        K3 = max(5, min(K, 60))
        l3 = run_single_level(args, comm, mb, "cfg3", dict(WORKLOADS["cfg3"]), 8, K3, W, s3,
                              with_e2e=False, with_cpu=False, weak=False)
        c3 = s3.finish() if comm.rank == 0 else None
        if l3 is not None:
            cfg3 = {"workload": l3["config"]["workload"], "value": l3["value"], "unit": "MLUPS",
                    "scaling": "strong", "n_gpus": world, "steps": K3, "ms_per_step": l3["ms_per_step"],
                    "sweep_ms": l3["roofline"]["kernel_ms"], "comm_ms": l3["timers_ms_per_step"]["comm"],
                    "frac": l3["roofline"]["frac"], "achieved_gbs": l3["roofline"]["achieved"],
                    "cells": l3["config"]["cells"], "cells_per_gpu": l3["config"]["cells_per_gpu"],
                    "mesh": l3["config"]["mesh"], "halo_exchange": l3["config"]["halo_exchange"],
                    "setup_s": l3["config"]["setup_s"], "gpu_launches": l3["gpu_launches"],
                    "clocks": c3, "check": l3["check"],
                    "note": "strong-scaling efficiency = value(N) / (N * value(1)) across the driver's per-N lines"}
    if comm.rank == 0:
        line["clocks"] = clocks
        if check_block:
            line["check"].update(check_block)
        if cfg3 is not None:
            line["cfg3"] = cfg3
        print(json.dumps(line), flush=True)
    mb.mus_finalize()
    comm.close()


if __name__ == "__main__":
    main()
