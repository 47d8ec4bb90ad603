#!/usr/bin/env python3
"""bench.py -- MLUPS of the per-level LBM time step on N B200s (one process per
GPU) with the HBM roofline and the reference's CPU algorithm timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl musb200|reference]
                  [--workload cfg1|cfg2|cfg3|cfg3-256|cfg4] [--level L]

Workloads (BASELINE.json configs):
  N = 1 : cfg2  D3Q19 TRT lid-driven cavity 256^3 (level 8), bounce-back walls +
                velocity_bounceback lid  -- the configuration the metric is quoted on
  N = 2, 4, 8 : the same cavity WEAK-scaled, 256^3 cells per GPU: the first N octants of the
                level-9 cube (512x256x256, 512x512x256, 512^3 -- the "512^3-equivalent mesh" of
                north_star at 8 GPUs), SFC-partitioned into N equal Morton ranges, halo exchange
                through peer memory over NVLink (NCCL send/recv with --no-p2p)
  --workload cfg4 : BASELINE config 4, two-level octree with linear ghost interpolation; a step is
                one coarse cycle (1 coarse + 2 fine level steps); with --gpus N the mesh is cut along
                the global space-filling curve (strong scaling)
  --workload cfg3 : D3Q27 MRT periodic 512^3 (level 9) strong-scaled over N ranks (BASELINE
                config 3; also cfg3-256 on one GPU)
A "step" is one level time step (set_boundary, swap, fused aux+stream+collide,
halo exchange) over the whole mesh.  `value` is device-timed with inputs resident
in HBM; `e2e` goes through the public C ABI with HOST buffers: state upload from
pinned host memory, every step the host->device copy of the boundary values and
a device->host read of the tracked probe element, and the final state download.
"""
import argparse
import ctypes
import json
import math
import os
import re
import subprocess
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
    "cfg3-256": dict(ident={"kind": "fluid", "relaxation": "mrt", "layout": "d3q27"}, level=8,
                     kind="periodic", omega=1.9, name="D3Q27 MRT periodic channel 256^3"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def measured_traffic(kernel, cells):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r01_traffic.json); None when no capture matches this kernel and cell count."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(kernel)
        if t and int(t["cells"]) == int(cells):
            return {"bytes": float(t["traffic_gb"]) * 1e9, "algorithmic_bytes": None,
                    "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)",
                    "source": t["source"]}
    except Exception:
        pass
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------
def cpu_baseline(wl, level, steps, warmup=2, target_s=None):
    """the reference's CPU algorithm (oracle port: AOS, two-pass, per-element omega),
    OpenMP over the host cores, on a bounded sample of the same workload: `steps` level steps
    after `warmup` untimed ones; with target_s the step count is chosen from a short probe so
    that the timed part is about target_s seconds of CPU work."""
    from oracle import musoracle as mo
    QQ = 19 if wl["ident"]["layout"] == "d3q19" else 27
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
    sch.run(max(1, warmup))
    if target_s is not None:
        t0 = time.perf_counter()
        sch.run(3)
        steps = int(min(1000, max(steps, target_s / ((time.perf_counter() - t0) / 3.0))))
    t0 = time.perf_counter()
    sch.run(steps)
    dt = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    return dict(value=ld.nFluid * steps / dt / 1e6, unit="MLUPS", cores=cores, kind="port",
                sample="%s at level %d (%d^3 = %d cells), %d steps, oracle C port of the reference "
                       "algorithm with OpenMP (Fortran toolchain unavailable)" % (
                           re.sub(r",? \d+\^3.*", "", wl["name"]), level, 1 << level, ld.nFluid, steps),
                ms_per_step=dt / steps * 1e3)


def run_reference(args, wl_name, wl):
    """the reference arm: nothing of the product is imported or loaded on this path"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        # torchrun pins every rank to one OpenMP thread; the CPU arm runs on rank 0 alone and
        # takes all host cores (read by libgomp when the oracle library is loaded below)
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    level = min(wl["level"], 7)
    t0 = time.perf_counter()
    cb = cpu_baseline(wl, level, max(1, args.steps), warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "MLUPS", "value": cb["value"], "unit": "MLUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak" if wl_name == "cfg2" else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name + ": " + wl["name"], "sample": cb["sample"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def run_multilevel(args, wl_name, wl):
    """cfg4: K coarse cycles of do_recursive_multiLevel, device timed; on N > 1 ranks the mesh is
    cut along the global space-filling curve (strong scaling), halo exchange per level over NCCL."""
    import musubi_b200 as mb
    from musubi_b200 import cases
    from musubi_b200 import treelm_multilevel as tm
    from musubi_b200._lib import check, lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W, K = max(3, args.warmup), max(1, args.steps)
    dist, uid = None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(mb.get_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(t, 0)
        uid = bytes(t.numpy().tobytes())
    mb.mus_init(rank, world, local_rank, uid)

    def allred(x, op):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t[0])

    def barrier():
        if dist is not None:
            dist.barrier()

    t_setup = time.perf_counter()
    minL = args.level or WORKLOADS[wl_name]["level"]
    scale = 2.0 ** (minL - WORKLOADS[wl_name]["level"])      # boxes / cylinder are given at the base level
    boxes = [(int(lo * scale), int(hi * scale)) for lo, hi in wl["boxes"]]
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
        check(lib.musb200_aux_upload(l, aux.ctypes.data))
    sch.synchronize()
    setup_s = time.perf_counter() - t_setup
    upd = {l: 2 ** (l - levels[0]) for l in levels}                # level steps per coarse cycle
    lups_cycle = allred(float(sum(lv[l].nFluid * upd[l] for l in levels)), "SUM")
    solve_cycle = sum((lv[l].nFluid + lv[l].nGhostFromCoarser) * upd[l] for l in levels)   # this rank

    def timed_region():
        barrier()
        sch.synchronize()
        check(lib.musb200_event_mark(0))
        sch.do_computation(K)
        check(lib.musb200_event_mark(1))
        sch.synchronize()
        barrier()
        ms = ctypes.c_double()
        check(lib.musb200_event_elapsed(ctypes.byref(ms)))
        return allred(ms.value, "MAX")

    sch.do_computation(W)
    sch.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    check(lib.musb200_timers_reset())
    t_ms = timed_region()                       # region 1: nothing but the cycle's own launches
    nl = ctypes.c_longlong()
    check(lib.musb200_launch_count(ctypes.byref(nl)))
    check(lib.musb200_set_profiling(1))         # region 2: CUDA events around every stage
    check(lib.musb200_timers_reset())
    t2_ms = timed_region()
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
    mass = allred(sum(sch.reduce(l)[0] / 8.0 ** (l - levels[0]) for l in levels), "SUM") if world == 1 else None

    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        for l in levels:
            sch.upload_state(l, host[l])
        sch.do_computation(K)
        for l in levels:
            host[l] = sch.download_state(l)
        sch.synchronize()
        barrier()
        dt = allred(time.perf_counter() - t0, "MAX")
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
                                    "interpolated locally, fluid-only halos over NCCL" % (
                                        world, "SPartA-weighted" if args.balance else "equal"),
                       "step": "one coarse cycle = %s level steps" % "+".join(str(upd[l]) for l in levels),
                       "mlups_definition": "sum_l nFluid(l) * 2^(l-minLevel) per coarse cycle / time (SURVEY 8d)",
                       "mlups_by_reference_formula": value_ref_formula,
                       "interpolation": "linear", "omega": {str(l): omega[l] for l in levels},
                       "l2": "state %.2f GB per rank > 126 MB L2" % (2 * nbytes / 1e9), "setup_s": round(setup_s, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "kernel": "sweepKernel<19,bgk> (all level steps of a cycle, rank 0)",
                         "bytes_per_lup": BYTES_PER_LUP[19], "kernel_ms": sweep_ms,
                         "share_of_step": sweep_ms / (t2_ms / K)},
            "timers_ms_per_step": {"compute": cm.value / K, "bc": bm.value / K, "comm": com.value / K,
                                   "intp": im.value / K},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(nl.value), "clocks": clocks,
            "check": {"total_mass": mass},
        }
        print(json.dumps(line), flush=True)
    sch.synchronize()
    barrier()
    sch.destroy()
    mb.mus_finalize()
    if dist is not None:
        dist.destroy_process_group()


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
    ap.add_argument("--overlap", action="store_true",
                    help="sweep the send-halo elements first and overlap their exchange with the rest")
    ap.add_argument("--fused-push", action="store_true",
                    help="peer-memory halo exchange with the link stores fused into the sweep kernel")
    ap.add_argument("--balance", action="store_true",
                    help="cfg4 on N > 1 ranks: cut the space-filling curve by tem_balance_sparta with level "
                         "weights 2^(l - minLevel) instead of equal element counts")
    ap.add_argument("--no-p2p", action="store_true",
                    help="halo exchange through pack/ncclSend/ncclRecv/unpack instead of peer memory")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
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
        run_reference(args, wl_name, wl)
        return
    if wl["kind"] == "multilevel":
        run_multilevel(args, wl_name, wl)
        return
    W = max(3, args.warmup)
    K = max(1, args.steps)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)

    import musubi_b200 as mb
    from musubi_b200 import cases
    from musubi_b200._lib import check, lib

    uid = None
    if world > 1:
        import torch
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(mb.get_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(t, 0)
        uid = bytes(t.numpy().tobytes())
    mb.mus_init(rank, world, local_rank, uid)

    def barrier():
        if dist is not None:
            dist.barrier()

    def allmax(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    ident, level, QQ = wl["ident"], wl["level"], (19 if wl["ident"]["layout"] == "d3q19" else 27)
    t_setup = time.perf_counter()
    ld = mb.LevelDesc(level, QQ, wl["kind"], rank, world, octants=octants)
    check(lib.musb200_set_overlap(1 if args.overlap else 0))
    check(lib.musb200_set_fused_push(1 if args.fused_push else 0))
    if wl["kind"] == "cavity":
        rho, vel = cases.cavity_rest(ld)
    else:
        rho, vel = cases.taylor_green(ld, mean=(0.05, 0.0, 0.0))
    nbytes = ld.nSize * QQ * 8
    hp = ctypes.c_void_p()
    check(lib.musb200_host_alloc(nbytes, ctypes.byref(hp)))            # pinned host mirror of state
    host_state = np.ctypeslib.as_array(ctypes.cast(hp, ctypes.POINTER(ctypes.c_double)), shape=(ld.nSize * QQ,))
    host_state[:] = cases.equilibrium_state(QQ, rho, vel, ld.nSize)
    del rho, vel
    sch = mb.Scheme(ident, ld, wl["omega"], lambda_=3.0 / 16.0, omega_bulk=wl["omega"])
    halo_path = "NCCL send/recv"
    if world > 1 and not args.no_p2p:
        ok = 1.0
        try:
            sch.p2p_connect(dist, level)
        except Exception as ex:       # no peer access on this box: stay on the NCCL path
            sys.stderr.write("rank %d: peer-memory halo exchange unavailable (%s)\n" % (rank, ex))
            ok = 0.0
        # the choice is collective: one rank without peer access puts every rank on NCCL
        import torch
        t = torch.tensor([ok], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if float(t[0]) > 0.5:
            halo_path = ("peer-memory stores over NVLink fused into the sweep kernel + arrival handshake"
                         if args.fused_push else "peer-memory stores over NVLink (one kernel)")
        else:
            check(lib.musb200_p2p_enable(level, 0))
    lid = cases.lid_values(ld) if wl["kind"] == "cavity" else None
    lid_pinned = None
    if lid is not None and lid.size:
        lp = ctypes.c_void_p()
        check(lib.musb200_host_alloc(lid.nbytes, ctypes.byref(lp)))
        lid_pinned = np.ctypeslib.as_array(ctypes.cast(lp, ctypes.POINTER(ctypes.c_double)), shape=(lid.size,))
        lid_pinned[:] = lid.ravel()
        sch.set_bc_values(level, 2, lid_pinned)
    check(lib.musb200_state_upload(level, 2, host_state.ctypes.data))
    check(lib.musb200_set_now_next(level, 1, 2))
    check(lib.musb200_state_copy_next_to_now(level))
    sch.synchronize()
    setup_s = time.perf_counter() - t_setup
    nFluid_total = allsum(float(ld.nFluid))

    # ---------------- device-resident throughput ---------------------------
    sch.do_computation(W)
    sch.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # region 1: exactly K steps, nothing but the step's own launches on the stream -> `value`
    check(lib.musb200_timers_reset())
    barrier()
    sch.synchronize()
    check(lib.musb200_event_mark(0))
    sch.do_computation(K)
    check(lib.musb200_event_mark(1))
    sch.synchronize()
    barrier()
    ms = ctypes.c_double()
    check(lib.musb200_event_elapsed(ctypes.byref(ms)))
    t_ms = allmax(ms.value)
    nl = ctypes.c_longlong()
    check(lib.musb200_launch_count(ctypes.byref(nl)))
    launches = int(nl.value)
    # region 2: the same K steps with CUDA events around every stage (per-kernel durations
    # for the roofline; the extra event records cost a fraction of a percent, so they are
    # kept out of `value`)
    check(lib.musb200_set_profiling(1))
    check(lib.musb200_timers_reset())
    barrier()
    sch.synchronize()
    check(lib.musb200_event_mark(0))
    sch.do_computation(K)
    check(lib.musb200_event_mark(1))
    sch.synchronize()
    barrier()
    ms2 = ctypes.c_double()
    check(lib.musb200_event_elapsed(ctypes.byref(ms2)))
    t2_ms = allmax(ms2.value)
    clocks = sampler.finish() if rank == 0 else None
    cm, bm, com, im = (ctypes.c_double() for _ in range(4))
    check(lib.musb200_timers(ctypes.byref(cm), ctypes.byref(bm), ctypes.byref(com), ctypes.byref(im)))
    check(lib.musb200_set_profiling(0))
    sweep_ms = allmax(cm.value) / K
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
    if not args.no_e2e:
        # the probe is tracked every iteration: lazy auxField, the tracked element's moments are
        # computed on demand (musb200_aux_probe) instead of materialising auxField every step
        check(lib.musb200_set_aux_every_step(2))
        probe = np.zeros(4)
        Ke = K
        barrier()
        sch.synchronize()
        t0 = time.perf_counter()
        check(lib.musb200_state_upload(level, 2, host_state.ctypes.data))
        check(lib.musb200_set_now_next(level, 1, 2))
        check(lib.musb200_state_copy_next_to_now(level))
        if lid_pinned is not None:
            check(lib.musb200_bc_set_values(level, 2, lid_pinned.size, lid_pinned.ctypes.data))
        for it in range(Ke):
            check(lib.musb200_step(level, level, 1))
            if lid_pinned is not None and it + 1 < Ke:
                # the next step's boundary values go up while this step runs (copy stream)
                check(lib.musb200_bc_set_values(level, 2, lid_pinned.size, lid_pinned.ctypes.data))
            check(lib.musb200_aux_probe(level, 1, probe.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
        check(lib.musb200_state_download(level, sch.now_next(level)[1], host_state.ctypes.data))
        sch.synchronize()
        barrier()
        dt = allmax(time.perf_counter() - t0)
        bc_bytes = int(lid_pinned.nbytes) if lid_pinned is not None else 0
        e2e = {"value": nFluid_total * Ke / dt / 1e6, "unit": "MLUPS",
               "h2d_bytes_per_step": int(nbytes / Ke + bc_bytes), "d2h_bytes_per_step": int(nbytes / Ke + 32),
               "steps": Ke, "wall_s": dt,
               "protocol": "pinned-host state upload + K x (BC values H2D on the copy stream, level step, "
                           "probe element computed on demand + D2H) + state download"}
        check(lib.musb200_set_aux_every_step(0))

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:    # about 10 s of CPU work on the host cores (N = 1 only)
            cb = cpu_baseline(wl, min(level, 7), 10, target_s=10.0)
        except Exception as ex:  # the oracle is optional test infrastructure
            cb = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "port", "sample": "failed: %r" % (ex,)}
    if rank == 0:
        line = {
            "metric": "MLUPS", "value": value, "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_ms / K, "higher_is_better": True,
            "scaling": "weak" if weak else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name + ": " + wl["name"], "level": level, "cells": int(nFluid_total),
                       "partition": "treelm SFC, %d equal Morton ranges" % world,
                       "cells_per_gpu": int(nFluid_total / world),
                       "halo_exchange": ("none (1 rank)" if world == 1 else
                                         "%s, %s" % (halo_path, "overlapped with the interior sweep"
                                                     if args.overlap else "after compute")),
                       "relaxation": ident["relaxation"], "layout": ident["layout"], "omega": wl["omega"],
                       "l2": "state of %.2f GB per buffer per GPU >> 126 MB L2, no flush needed" % (nbytes / 1e9),
                       "aux_every_step": False, "setup_s": round(setup_s, 2)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (traffic["bytes"] if traffic else None),
                         "traffic_detail": traffic, "peak_source": peak_src,
                         "kernel": kernel_name,
                         "bytes_per_lup": BYTES_PER_LUP[QQ], "kernel_ms": sweep_ms,
                         "share_of_step": sweep_ms / (t2_ms / K),
                         "timed": "CUDA events around every sweep launch of a second K-step region "
                                  "(%.4f ms per step with the stage events in)" % (t2_ms / K)},
            "timers_ms_per_step": {"compute": cm.value / K, "bc": bm.value / K, "comm": com.value / K,
                                   "intp": im.value / K},
            "cpu_baseline": cb, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "check": {"total_mass": mass, "max_vel": vmax, "nan": nan},
        }
        print(json.dumps(line), flush=True)
    sch.synchronize()
    barrier()                    # peers may still store into this rank's halo rows
    sch.destroy()
    mb.mus_finalize()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
