"""Multi-rank parity driver, launched by torch.distributed.run (one process per rank).

  --mode lists : CPU only (gloo).  Every rank builds its SFC partition with the product's
                 mesh generator, fills a state array with a value that encodes (treeID, dir),
                 packs its send lists, exchanges them over gloo and unpacks; every received
                 halo link must carry the value of the element that owns it.
  --mode lists-ml: the same for a partitioned multi-level mesh: per level, every halo element
                 must receive all QQ links and its four auxField entries from the rank that owns
                 it, and every element a local solved element pulls from must be present.
  --mode gpu   : one GPU per rank through libmusb200 + NCCL; after K steps each rank's fluid
                 PDFs must be bit-identical to the single-domain oracle run.
  --mode gpu-ml: a multi-level mesh (nested refined boxes) cut along the global space-filling
                 curve, one GPU per rank, halo exchange of state and auxField per level through
                 NCCL, ghosts interpolated locally; fluid PDFs of every level bit-identical to the
                 single-domain oracle after K coarse cycles.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def lists_multilevel(a, dist, torch, rank, world, QQ):
    """CPU only: the halo lists of partition_multilevel moved over gloo"""
    from musubi_b200 import treelm_multilevel as tm
    boxes = [(5, 11)] if a.levels == 2 else [(4, 12), (12, 20)]
    lv, _ = tm.build_multilevel(4, boxes, QQ=QQ, intp_method=a.method)
    mine = tm.partition_multilevel(lv, world)[rank]
    nchk = 0
    for l, M in sorted(mine.items()):
        code = lambda tid, d: tid.astype(np.float64) * 100.0 + d  # noqa: E731
        state = np.full(M.nSize * QQ, -1.0)
        aux = np.full(M.nSize * 4, -1.0)
        own = np.arange(M.nFluid)
        for d in range(QQ):
            state[own * QQ + d] = code(M.total[:M.nFluid], d + 1)
        for k in range(4):
            aux[own * 4 + k] = code(M.total[:M.nFluid], 50 + k)
        reqs, bufs = [], {}
        for s in M.send:
            e = s["elemPos"].astype(np.int64) - 1
            payload = np.concatenate([state[s["pos"] - 1], aux.reshape(-1, 4)[e].ravel()])
            reqs.append(dist.isend(torch.from_numpy(payload), s["proc"], tag=l))
        for r in M.recv:
            bufs[r["proc"]] = torch.zeros(len(r["pos"]) + 4 * len(r["elemPos"]), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[r["proc"]], r["proc"], tag=l))
        for q in reqs:
            q.wait()
        for r in M.recv:
            buf = bufs[r["proc"]].numpy()
            n = len(r["pos"])
            state[r["pos"] - 1] = buf[:n]
            e = r["elemPos"].astype(np.int64) - 1
            aux.reshape(-1, 4)[e] = buf[n:].reshape(-1, 4)
            el, d = (r["pos"] - 1) // QQ, (r["pos"] - 1) % QQ + 1
            assert np.array_equal(state[r["pos"] - 1], code(M.total[el], d)), "halo carries foreign PDFs"
            assert np.array_equal(aux.reshape(-1, 4)[e], np.stack([code(M.total[e], 50 + k) for k in range(4)], 1))
            nchk += n
        # every halo element complete; everything a solved element pulls from is present locally
        h0 = M.nFluid + M.nGhostFromCoarser + M.nGhostFromFiner
        assert np.all(state[h0 * QQ:M.nElems * QQ] >= 0.0) and np.all(aux[h0 * 4:M.nElems * 4] >= 0.0)
        pulled = M.neigh.reshape(QQ, M.nSize)[:, :M.nFluid]
        src = (pulled - 1) // QQ
        assert np.all((src >= 0) & (src < M.nElems))
        # a fluid element never bounces back where the single-domain mesh has a neighbour
        g = M.globalPos[:M.nFluid] - 1
        ref = lv[l].neigh.reshape(QQ, lv[l].nSize)[:, g]
        ref_bounce = (ref - 1) // QQ == g[None, :]
        my_bounce = src == np.arange(M.nFluid)[None, :]
        assert np.array_equal(my_bounce, ref_bounce)
    print("rank %d: %d multi-level halo links verified" % (rank, nchk))


def run_multilevel(a, mb, dist, torch, rank, world, QQ):
    from oracle import musoracle as mo
    from musubi_b200 import treelm_multilevel as tm
    from musubi_b200._lib import check, lib
    from test_multilevel import OMEGA_MIN, build
    boxes = [(5, 11)] if a.levels == 2 else [(4, 12), (12, 20)]
    cyl = None
    minL = 4
    # every rank builds the global mesh and its partition (deterministic), and the single-domain
    # oracle run that is the truth for all of them
    lv, intp, tables, ms = build(mo, minL, boxes, QQ, a.method, a.relaxation, cyl, OMEGA_MIN[len(boxes)])
    ranks = tm.partition_multilevel(lv, world)
    mine = ranks[rank]
    my_tables, ghost_comm = mb.multilevel_tables(mine, intp), None
    if a.ghost_exchange:
        # the reference's form: shared ghosts interpolated by one rank, shipped through the
        # FromCoarser / FromFiner buffers of the level (state links + auxField of ghostFromFiner)
        dt, comm = tm.delegate_shared_ghosts(ranks, [mb.multilevel_tables(rl, intp) for rl in ranks], lv)
        my_tables, ghost_comm = dt[rank], comm[rank]
        print("rank %d: %d ghosts received through ghost buffers" % (
            rank, sum(len(c["elemPos"]) for l in ghost_comm for k in ghost_comm[l] for c in ghost_comm[l][k]["recv"])))
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(mb.get_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(t, 0)
    mb.mus_init(rank, world, int(os.environ.get("LOCAL_RANK", rank)), bytes(t.numpy().tobytes()))
    ident = {"kind": "fluid", "relaxation": a.relaxation, "layout": a.layout}
    omega = {l: float(1.0 / (3.0 * s.visc[0] + 0.5)) for l, s in ms.s.items()}
    visc = {l: float(s.visc[0]) for l, s in ms.s.items()}
    kw = dict(lambda_=0.25, omega_bulk=1.2, intp=(my_tables, intp["order"]), viscosity=visc, ghost_comm=ghost_comm)
    check(lib.musb200_set_graphs(0 if a.no_graphs else 1))

    def connect(sc):
        if a.p2p:      # peer-memory exchange of state + auxField halos on every level (collective)
            for l in sorted(mine):
                sc.p2p_connect(dist, l)

    if a.restart:
        # mus_readRestart into a FRESH scheme on every rank: only the fluid PDFs are known; halos,
        # ghosts and auxField come from musb200_fill_helper_elements (collective).  Truth: the
        # single-domain oracle restarted the same way.
        from musubi_b200 import restart_io
        ms.run(3)
        gtid, glp = mo.global_tree(lv)
        dump = mo.pdf_serialize(ms.s, gtid, glp)
        ms2 = mo.MultiLevelScheme(lv, tables, a.relaxation, "fluid", omega_min=OMEGA_MIN[len(boxes)],
                                  omega_bulk=1.2, order=intp["order"])
        mo.pdf_unserialize(ms2.s, gtid, glp, dump)
        ms2.fill_helper_elements()
        sch = mb.Scheme(ident, mine, omega, **kw)
        connect(sch)
        tid, lp = restart_io.tree_order(mine)
        where = {int(t): i for i, t in enumerate(gtid)}
        sel = np.array([where[int(t)] for t in tid], dtype=np.int64)
        sch.pdf_unserialize(tid, lp, dump.reshape(-1, QQ)[sel].ravel())
        sch.fill_helper_elements()
        ms = ms2
    else:
        sch = mb.Scheme(ident, mine, omega, **kw)
        connect(sch)
    for l, s in ms.s.items():
        if a.restart:
            break
        M = mine[l]
        g = M.globalPos - 1
        st = np.zeros(M.nSize * QQ)
        st[:M.nElems * QQ] = s.state[s.nNext].reshape(-1, QQ)[g].ravel()
        sch.upload_state(l, st, st)
        aux = np.zeros(M.nSize * 4)
        aux[:M.nElems * 4] = s.aux.reshape(-1, 4)[g].ravel()
        check(lib.musb200_aux_upload(l, aux.ctypes.data))
    sch.do_computation(a.steps)
    ms.run(a.steps)
    nd_total = 0
    for l, s in ms.s.items():
        M = mine[l]
        g = M.globalPos[:M.nFluid] - 1
        got = sch.download_state(l)[:M.nFluid * QQ].reshape(-1, QQ)
        exp = s.state[s.nNext].reshape(-1, QQ)[g]
        nd = int((got != exp).sum())
        nd_total += nd
        print("rank %d/%d level %d: %d fluid + %d/%d ghosts + %d halos, ndiff=%d" % (
            rank, world, l, M.nFluid, M.nGhostFromCoarser, M.nGhostFromFiner, M.nHalo, nd))
    print("rank %d/%d multilevel %s %s: ndiff=%d" % (rank, world, a.relaxation, a.layout, nd_total))
    assert nd_total == 0
    sch.synchronize()
    dist.barrier()
    sch.destroy()
    mb.mus_finalize()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="lists")
    ap.add_argument("--level", type=int, default=4)
    ap.add_argument("--layout", default="d3q27")
    ap.add_argument("--relaxation", default="mrt")
    ap.add_argument("--kind", default="periodic")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--outlet", default="pressure_expol")
    ap.add_argument("--octants", type=int, default=8)
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--p2p", action="store_true")
    ap.add_argument("--levels", type=int, default=2, help="gpu-ml: 2 or 3 levels")
    ap.add_argument("--method", default="linear", help="gpu-ml: interpolation method")
    ap.add_argument("--ghost-exchange", action="store_true",
                    help="gpu-ml: shared ghosts delegated to one rank and exchanged through the "
                         "FromCoarser / FromFiner buffers (the reference's form of the run)")
    ap.add_argument("--fused-push", action="store_true",
                    help="peer-memory exchange with the push fused into the sweep kernel")
    ap.add_argument("--sweep-wait", action="store_true",
                    help="peer-memory exchange: wait inside the next sweep (halo CTAs) instead of a wait kernel")
    ap.add_argument("--no-graphs", action="store_true", help="direct launches instead of CUDA-graph replay")
    ap.add_argument("--restart", action="store_true",
                    help="gpu-ml: restart into a fresh scheme (fluid PDFs only + musb200_fill_helper_elements)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import musubi_b200 as mb
    QQ = 19 if a.layout == "d3q19" else 27
    if a.mode == "lists-ml":
        lists_multilevel(a, dist, torch, rank, world, QQ)
        dist.barrier()
        dist.destroy_process_group()
        return
    if a.mode == "gpu-ml":
        run_multilevel(a, mb, dist, torch, rank, world, QQ)
        dist.barrier()
        dist.destroy_process_group()
        return
    ld = mb.LevelDesc(a.level, QQ, a.kind, rank, world, octants=a.octants)

    if a.mode == "gpu-timeout":
        # failure handling of the peer-memory exchange: the last rank "dies" (stops stepping) while
        # the others go on; their waits give up after the timeout, the stream drains and the next
        # synchronising call returns MUSB200_ERR_NCCL naming the silent rank -- no hang
        from musubi_b200 import cases
        from musubi_b200._lib import Musb200Error, check, lib
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(mb.get_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(t, 0)
        mb.mus_init(rank, world, int(os.environ.get("LOCAL_RANK", rank)), bytes(t.numpy().tobytes()))
        check(lib.musb200_set_exchange_timeout(1.5))
        check(lib.musb200_set_sweep_wait(1 if a.sweep_wait else 0))
        check(lib.musb200_set_overlap(1 if a.overlap else 0))
        ident = {"kind": "fluid", "relaxation": a.relaxation, "layout": a.layout}
        sch = mb.Scheme(ident, ld, 1.7, lambda_=0.25, omega_bulk=1.3)
        rho, vel = cases.taylor_green(ld)
        st = cases.equilibrium_state(QQ, rho, vel, ld.nSize)
        sch.upload_state(a.level, st, st)
        sch.p2p_connect(dist, a.level)
        sch.do_computation(3)
        sch.synchronize()
        dist.barrier()
        import time
        t0 = time.time()
        if rank == world - 1:
            print("rank %d: going silent" % rank)
        else:
            try:
                sch.do_computation(4)
                sch.synchronize()
                print("rank %d: NO ERROR after %.1f s" % (rank, time.time() - t0))
            except Musb200Error as ex:
                ok = ex.code == 3 and "timed out" in str(ex)
                print("rank %d: %s after %.1f s -> %s" % (rank, "timeout reported" if ok else "unexpected", time.time() - t0, ex))
        dist.barrier()
        mb.mus_finalize()
        dist.barrier()
        dist.destroy_process_group()
        return

    if a.mode == "lists":
        code = lambda tid, d: tid.astype(np.float64) * 100.0 + d  # noqa: E731
        state = np.full(ld.nSize * QQ, -1.0)
        e = np.arange(ld.nFluid)
        for d in range(QQ):
            state[e * QQ + d] = code(ld.total[:ld.nFluid], d + 1)
        reqs, bufs = [], {}
        for s in ld.send:
            t = torch.from_numpy(state[s["pos"] - 1].copy())
            reqs.append(dist.isend(t, s["proc"]))
        for r in ld.recv:
            bufs[r["proc"]] = torch.zeros(len(r["pos"]), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[r["proc"]], r["proc"]))
        for q in reqs:
            q.wait()
        nchk = 0
        for r in ld.recv:
            state[r["pos"] - 1] = bufs[r["proc"]].numpy()
            el, d = (r["pos"] - 1) // QQ, (r["pos"] - 1) % QQ + 1
            assert np.array_equal(state[r["pos"] - 1], code(ld.total[el], d)), "halo carries foreign data"
            nchk += len(r["pos"])
        # every link a local element pulls from a halo must have been received
        pulled = ld.neigh[:].reshape(QQ, ld.nSize)[:, :ld.nFluid].ravel()
        from_halo = pulled[(pulled - 1) // QQ >= ld.nFluid]
        assert np.all(state[from_halo - 1] >= 0.0), "a pulled halo link was never received"
        # the peer-memory path: every receiver ships its recv position list to its sender
        # (Scheme.p2p_connect's host part), then the sender "pushes" (value, remote position) and the
        # receiver stores each value where the sender says -- what pushHaloKernel does over NVLink
        from musubi_b200.scheme import exchange_recv_lists
        proc, nVals, rpos = exchange_recv_lists(dist, ld)
        assert list(proc) == [s_["proc"] for s_ in ld.send] and rpos.size == int(nVals.sum())
        state2 = np.full(ld.nSize * QQ, -1.0)
        state2[:ld.nFluid * QQ] = state[:ld.nFluid * QQ]
        reqs, bufs, off = [], {}, 0
        for s_ in ld.send:
            n = len(s_["pos"])
            pay = np.concatenate([state[s_["pos"] - 1], rpos[off:off + n].astype(np.float64)])
            reqs.append(dist.isend(torch.from_numpy(pay), s_["proc"], tag=7))
            off += n
        for r in ld.recv:
            bufs[r["proc"]] = torch.zeros(2 * len(r["pos"]), dtype=torch.float64)
            reqs.append(dist.irecv(bufs[r["proc"]], r["proc"], tag=7))
        for q in reqs:
            q.wait()
        for r in ld.recv:
            b = bufs[r["proc"]].numpy()
            n = len(r["pos"])
            where = b[n:].astype(np.int64)
            assert np.array_equal(np.sort(where), np.sort(r["pos"])), "pushed positions are not my recv list"
            state2[where - 1] = b[:n]
        assert np.array_equal(state2, state), "peer push lands elsewhere than the recv unpack"
        print("rank %d: %d halo links verified, %d pulled-from-halo links covered, %d peers" % (
            rank, nchk, from_halo.size, len(ld.send)))
    else:
        from oracle import musoracle as mo
        from musubi_b200 import cases
        ident = {"kind": "fluid", "relaxation": a.relaxation, "layout": a.layout}
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(mb.get_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(t, 0)
        mb.mus_init(rank, world, int(os.environ.get("LOCAL_RANK", rank)), bytes(t.numpy().tobytes()))
        # single-domain oracle = the truth for every rank
        mb._lib.check(mb._lib.lib.musb200_set_overlap(1 if a.overlap else 0))
        mb._lib.check(mb._lib.lib.musb200_set_fused_push(1 if a.fused_push else 0))
        mb._lib.check(mb._lib.lib.musb200_set_sweep_wait(1 if a.sweep_wait else 0))
        mb._lib.check(mb._lib.lib.musb200_set_graphs(0 if a.no_graphs else 1))
        gl = mo.build_level_desc(a.level, QQ, a.kind, octants=a.octants)
        ref = mo.Scheme(gl, a.relaxation, "fluid", omega=1.7, lambda_=0.25, omega_bulk=1.3)
        gld = mb.LevelDesc(a.level, QQ, a.kind, 0, 1, octants=a.octants)
        if a.kind == "cavity":
            rho, vel = cases.cavity_rest(gld)
            ref.bc_vel[2] = cases.lid_values(gld, (0.05, 0.02, 0.0))
        elif a.kind == "channel":
            rho, vel = cases.cavity_rest(gld)
            ref.bc_vel[2] = cases.lid_values(gld, (0.03, 0.0, 0.0))
            ref.bc_kind[3] = a.outlet
            ref.bc_rho[3] = 1.0
        else:
            rho, vel = cases.taylor_green(gld, mean=(0.01, -0.02, 0.015))
        ref.init_equilibrium(rho, vel)
        first = int(ld.total[0] - gld.total[0])
        # local initial state: own fluid elements + halos, looked up by treeID
        gpos = (ld.total - gld.total[0]).astype(np.int64)
        init = np.zeros(ld.nSize * QQ)
        init[:ld.nElems * QQ] = ref.state[ref.nNext].reshape(-1, QQ)[gpos].ravel()
        sch = mb.Scheme(ident, ld, float(1.0 / (3.0 * ref.visc[0] + 0.5)), lambda_=0.25, omega_bulk=1.3,
                        bc_kind={3: a.outlet} if a.kind == "channel" else None)
        sch.upload_state(a.level, init, init)
        if a.p2p:
            sch.p2p_connect(dist, a.level)
        if a.kind == "cavity":
            sch.set_bc_values(a.level, 2, cases.lid_values(ld, (0.05, 0.02, 0.0)))
        if a.kind == "channel":
            sch.set_bc_values(a.level, 2, cases.lid_values(ld, (0.03, 0.0, 0.0)))
            nOut = len([b for b in ld.bc if b["id"] == 3][0]["elems"])
            if nOut:
                sch.set_bc_values(a.level, 3, np.full(nOut, 1.0))
            aux0 = np.zeros(ld.nSize * 4)
            aux0[:ld.nElems * 4] = ref.aux.reshape(-1, 4)[gpos].ravel()
            from musubi_b200._lib import check, lib
            check(lib.musb200_aux_upload(a.level, aux0.ctypes.data))
        m0 = sch.reduce()[0]
        sch.do_computation(a.steps)
        ref.run(a.steps)
        got = sch.download_state(a.level)[:ld.nFluid * QQ].reshape(-1, QQ)
        exp = ref.state[ref.nNext].reshape(-1, QQ)[first:first + ld.nFluid]
        nd = int((got != exp).sum())
        rel = float(np.max(np.abs(got - exp) / np.abs(exp)))
        m1 = sch.reduce()[0]
        print("rank %d/%d %s %s %s: ndiff=%d maxrel=%.2e mass drift=%.2e" % (
            rank, world, a.kind, a.relaxation, a.layout, nd, rel, abs(m1 / m0 - 1.0)))
        assert rel < 1e-10 and nd == 0
        assert abs(m1 / ref.total_mass() - 1.0) < 1e-12
        if a.kind == "periodic":
            assert abs(m1 / m0 - 1.0) < 1e-13
        sch.synchronize()
        dist.barrier()           # peers may still store into this rank's halo rows
        sch.destroy()
        mb.mus_finalize()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
