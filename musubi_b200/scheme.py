"""Host-side mirror of the reference's plugin surface for the per-level time
step, above the C ABI:

  identify{kind, relaxation, layout}  -> scheme%compute pointee
        (mus_init_advRel_fluid, mus/source/init/mus_initFluid_module.f90:102-409)
  mus_scheme_type%{state, pdf, auxField, levelDesc}
        (mus/source/scheme/mus_scheme_type_module.f90:88-180)
  control%do_computation(minLevel)
        (mus/source/mus_control_module.f90:149-221)

All arithmetic runs in libmusb200.so on the GPU; this module only moves the
host arrays through the C ABI, as the Fortran shim does.
"""
import ctypes

import numpy as np

from . import _lib  # noqa: F401  (musubi_b200._lib is part of the package surface: bench.py, tests)
from ._lib import P_DBL, P_I32, P_I64, check, lib, ptr

P2P_BLOB = 512          # MUSB200_P2P_BLOB of include/musb200.h
BC_KIND = {"wall": 0, "velocity_bounceback": 1, "pressure_antibounceback": 2, "pressure_expol": 3}
_initialized = False


def mus_init(rank=0, nranks=1, device=None, unique_id=None):
    """binds this process to one GPU (tem_start analogue for the device side)."""
    global _initialized
    if _initialized:
        return
    if device is None:
        device = rank
    uid = None
    if nranks > 1:
        if unique_id is None:
            raise ValueError("nranks > 1 needs the NCCL unique id")
        uid = ctypes.c_char_p(bytes(unique_id))
    check(lib.musb200_init(rank, nranks, device, uid))
    _initialized = True


def mus_finalize():
    global _initialized
    check(lib.musb200_finalize())
    _initialized = False


def get_unique_id():
    buf = ctypes.create_string_buffer(128)
    check(lib.musb200_get_unique_id(buf))
    return buf.raw


PS_VARIANT = {"first": 1, "second": 2}


def _relaxation_of(identify):
    rel = identify.get("relaxation", "bgk")
    variant = "standard"
    if isinstance(rel, dict):
        variant = rel.get("variant", "standard")
        rel = rel.get("name", "bgk")
    return rel, variant


def default_omega_bulk(identify, level, minLevel):
    """fluid%omegaBulkLvl(level) when the fluid table has no bulk_viscosity
    (mus_fluid_module.f90:205-262, 468-485): kind 'fluid' aborts unless the layout is d3q27
    (omegaBulk 1.54 on the coarsest level); 'fluid_incompressible' takes 1.54 (d3q27) / 1.19 (d3q19).
    That value defines viscBulk_phy on minLevel; acoustic scaling doubles the lattice bulk
    viscosity per finer level, omegaBulk = 1 / (9 nu_bulk / (5 - 9 cs2) + 1/2)."""
    kind, layout = identify.get("kind", "fluid"), identify.get("layout", "d3q19")
    if layout == "d3q27":
        ob = 1.54
    elif kind == "fluid":
        raise ValueError('"bulk_viscosity" is not provided in fluid table (kind fluid, layout %s): '
                         "pass omega_bulk" % layout)
    else:
        ob = 1.19
    cs2 = 1.0 / 3.0
    visc_bulk = ((5.0 - 9.0 * cs2) / 9.0) * (1.0 / ob - 0.5) * 2.0 ** (level - minLevel)
    return 1.0 / (9.0 * visc_bulk / (5.0 - 9.0 * cs2) + 0.5)


def select_kernel(identify):
    """mus_init_advRel_fluid / _fluid_incompressible / _lbm_ps: the (kind, relaxation, variant,
    layout) dispatch."""
    rel = identify.get("relaxation", "bgk")
    variant = "standard"
    if isinstance(rel, dict):
        variant = rel.get("variant", "standard")
        rel = rel.get("name", "bgk")
    relax, kind, QQ = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    check(lib.musb200_scheme_select(identify.get("kind", "fluid").encode(), rel.encode(),
                                    variant.encode(), identify.get("layout", "d3q19").encode(),
                                    ctypes.byref(relax), ctypes.byref(kind), ctypes.byref(QQ)))
    return relax.value, kind.value, QQ.value


class Scheme:
    """mus_scheme_type for one level range on one rank, device resident."""

    def __init__(self, identify, levelDescs, omega=None, lambda_=0.25, omega_bulk=None, intp=None,
                 viscosity=None, bc_kind=None, slot=0, species=None, ghost_comm=None):
        """intp: None or (tables, order) with tables as built by multilevel_tables();
        viscosity: {level: lattice viscosity} (fluid%viscKine%dataOnLvl) for the interpolation;
        bc_kind: {boundary id: kind} binds the mesh's 'pressure' boundaries to pressure_expol or
        pressure_antibounceback (the boundary_condition table of the Lua configuration);
        slot: scheme slot of the library (several schemes on one mesh);
        species: {"diff_coeff": D, "lambda": 0.25} of a passive_scalar scheme;
        ghost_comm: {level: {"fromCoarser" | "fromFiner": {"send" | "recv": [dict(proc, pos)]}}},
        the level's sendBufferFromCoarser / FromFiner lists of a host that ships interpolated ghosts
        between ranks as treelm's construction does (treelm_multilevel.delegate_shared_ghosts)."""
        self.slot = int(slot)
        self._bind()
        self.relax, self.kind, self.QQ = select_kernel(identify)
        self.scheme_kind = identify.get("kind", "fluid")       # varSys%SystemName of the restart header
        self.passive_scalar = identify.get("kind", "fluid") == "passive_scalar"
        self.nAux = 1 if self.passive_scalar else 4
        if self.passive_scalar:
            if species is None:
                raise ValueError("passive_scalar needs species = {diff_coeff, lambda}")
            self.ps_variant = PS_VARIANT.get(_relaxation_of(identify)[1], 0)
        elif omega is None:
            raise ValueError("a fluid scheme needs omega")
        if not isinstance(levelDescs, dict):
            levelDescs = {levelDescs.level: levelDescs}
        self.levelDesc = levelDescs
        self.minLevel, self.maxLevel = min(levelDescs), max(levelDescs)
        self.lambda_ = lambda_
        for lvl, ld in levelDescs.items():
            if ld.QQ != self.QQ:
                raise ValueError("levelDesc built for another stencil")
            if getattr(ld, "device_generated", False):
                check(lib.musb200_level_create_cube(lvl, ld.level, self.QQ, 1 if ld.kind == "walls" else 0))
            else:
                check(lib.musb200_level_create(lvl, self.QQ, self.QQ, self.nAux, ld.nSize, ld.nFluid,
                                               ld.nGhostFromCoarser, ld.nGhostFromFiner, ld.nHalo,
                                               ptr(ld.neigh, P_I32), ptr(ld.property, P_I64),
                                               ptr(ld.total, P_I64)))
            if self.passive_scalar:
                dc = species["diff_coeff"]          # lattice units of the level: {level: D} or one value
                check(lib.musb200_set_species(lvl, self.relax, self.ps_variant,
                                              float(dc[lvl] if isinstance(dc, dict) else dc),
                                              float(species.get("lambda", 0.25))))
            else:
                om = omega[lvl] if isinstance(omega, dict) else omega
                ob = omega_bulk[lvl] if isinstance(omega_bulk, dict) else omega_bulk
                if ob is None:
                    # bgk and trt never read it; mrt gets the reference's default (or its abort)
                    ob = (default_omega_bulk(identify, lvl, self.minLevel)
                          if _relaxation_of(identify)[0] == "mrt" else 1.0)
                self.set_relaxation(lvl, om, ob)
            if len(ld.bc_elemBuffer):
                check(lib.musb200_bc_elembuffer(lvl, len(ld.bc_elemBuffer), ptr(ld.bc_elemBuffer, P_I32)))
            for bc in ld.bc:
                kind = (bc_kind or {}).get(bc["id"], bc["kind"])
                if kind not in BC_KIND:
                    raise ValueError("boundary %d: kind %r needs a bc_kind binding" % (bc["id"], kind))
                check(lib.musb200_bc_register(lvl, bc["id"], BC_KIND[kind], len(bc["links"]),
                                              ptr(bc["links"], P_I32), ptr(bc["outPos"], P_I32),
                                              ptr(bc["posInBuffer"], P_I32), ptr(bc["iDir"], P_I32)))
                if kind.startswith("pressure") and len(bc["elems"]):
                    npos = np.ascontiguousarray(bc["neighPos"], dtype=np.int32)
                    check(lib.musb200_bc_register_elems(
                        lvl, bc["id"], len(bc["elems"]), ptr(bc["elems"], P_I32),
                        ptr(bc["posInBcElemBuf"], P_I32), ptr(bc["normalInd"], P_I32),
                        npos.shape[1], ptr(npos, P_I32), ptr(bc["iElemOfLink"], P_I32)))
            if viscosity is not None:
                v = viscosity[lvl]
                if np.isscalar(v):
                    check(lib.musb200_set_viscosity(lvl, None, float(v)))
                else:
                    v = np.ascontiguousarray(v, dtype=np.float64)
                    check(lib.musb200_set_viscosity(lvl, ptr(v, P_DBL), 0.0))
            bufs = [(0, 0, ld.send), (0, 1, ld.recv)]
            for kind_id, name in ((1, "fromCoarser"), (2, "fromFiner")):
                gc = (ghost_comm or {}).get(lvl, {}).get(name, {})
                bufs += [(kind_id, 0, gc.get("send", [])), (kind_id, 1, gc.get("recv", []))]
            for kind_id, d, lists in bufs:
                if not lists:
                    continue
                proc = np.array([c["proc"] for c in lists], dtype=np.int32)
                nVals = np.array([len(c["pos"]) for c in lists], dtype=np.int32)
                pos = np.concatenate([c["pos"] for c in lists]).astype(np.int32)
                check(lib.musb200_comm_register(lvl, kind_id, d, len(lists), ptr(proc, P_I32),
                                                ptr(nVals, P_I32), ptr(pos, P_I32)))

        if intp is not None:
            tables, order = intp
            for (lvl, what), t in tables.items():
                direction, o = (0, 0) if what == "fromFiner" else (1, what[1])
                if len(t["targets"]) == 0:
                    continue
                coord = np.ascontiguousarray(t["coord"], dtype=np.float64)
                wts = np.ascontiguousarray(t["weights"], dtype=np.float64)
                mats = np.ascontiguousarray(t["matrices"], dtype=np.float64)
                check(lib.musb200_intp_register(
                    lvl, direction, o, len(t["targets"]), ptr(t["targets"], P_I32),
                    ptr(t["srcOffset"], P_I32), ptr(t["srcPos"], P_I32),
                    ptr(wts, P_DBL) if wts.size else None,
                    ptr(t["posInMat"], P_I32) if len(t["posInMat"]) else None, int(t["nMat"]),
                    ptr(t["matOffset"], P_I32), ptr(mats, P_DBL) if mats.size else None,
                    ptr(coord, P_DBL) if coord.size else None))

    def _bind(self):
        check(lib.musb200_scheme_bind(self.slot))

    # -- source = { force = ... } (lattice units) -------------------------------
    def set_force(self, level, force, order=2, posInTotal=None):
        """force: 3 values (uniform) or [n][3] for the elements posInTotal (default 1..nSolve)"""
        self._bind()
        F = np.ascontiguousarray(force, dtype=np.float64)
        pos = None if posInTotal is None else np.ascontiguousarray(posInTotal, dtype=np.int32)
        ld = self.levelDesc[level]
        n = (ld.nFluid + ld.nGhostFromCoarser) if pos is None else pos.size
        uniform = 1 if (F.ndim == 1 and F.size == 3) else 0
        check(lib.musb200_source_force(level, int(order), int(n), ptr(pos, P_I32), ptr(F, P_DBL), uniform))

    # -- scheme%transVar (passive scalar) ----------------------------------------
    def set_transport_velocity(self, level, vel):
        self._bind()
        v = np.ascontiguousarray(vel, dtype=np.float64)
        check(lib.musb200_set_transport_velocity(level, v.size // 3, ptr(v, P_DBL), 1 if v.size == 3 else 0))

    def couple_transport_velocity(self, level, flow):
        """device-side coupling: read the velocity from the auxField of the Scheme `flow`"""
        self._bind()
        check(lib.musb200_couple_transport_velocity(level, flow.slot, level))

    # -- fluid%viscKine%omLvl / lambda / omegaBulkLvl ------------------------
    def set_relaxation(self, level, omega, omega_bulk):
        self._bind()
        if np.isscalar(omega):
            check(lib.musb200_set_relaxation(level, self.relax, self.kind, None, float(omega),
                                             float(self.lambda_), float(omega_bulk)))
        else:
            om = np.ascontiguousarray(omega, dtype=np.float64)
            check(lib.musb200_set_relaxation(level, self.relax, self.kind, ptr(om, P_DBL), 0.0,
                                             float(self.lambda_), float(omega_bulk)))

    # -- state(level)%val(:, 1:2), AOS, host layout ---------------------------
    def upload_state(self, level, aos_now, aos_next=None, nNow=1, nNext=2):
        self._bind()
        a = np.ascontiguousarray(aos_now, dtype=np.float64)
        check(lib.musb200_state_upload(level, nNow, a.ctypes.data))
        b = a if aos_next is None else np.ascontiguousarray(aos_next, dtype=np.float64)
        check(lib.musb200_state_upload(level, nNext, b.ctypes.data))
        check(lib.musb200_set_now_next(level, nNow, nNext))

    def init_equilibrium(self, level, rho, vel):
        """mus_init_pdf with zero strain rate, evaluated on the device: auxField <- (rho, vel),
        both state buffers <- f_eq(rho, vel)"""
        self._bind()
        ld = self.levelDesc[level]
        aux = np.zeros((ld.nSize, 4))
        aux[:ld.nElems, 0] = rho
        aux[:ld.nElems, 1:] = vel
        check(lib.musb200_aux_upload(level, aux.ctypes.data))
        check(lib.musb200_state_init_equilibrium(level))

    def now_next(self, level):
        self._bind()
        a, b = ctypes.c_int(), ctypes.c_int()
        check(lib.musb200_get_now_next(level, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def download_state(self, level, which=None):
        self._bind()
        ld = self.levelDesc[level]
        out = np.empty(ld.nSize * self.QQ)
        if which is None:
            which = self.now_next(level)[1]
        check(lib.musb200_state_download(level, which, out.ctypes.data))
        return out

    def download_aux(self, level):
        self._bind()
        out = np.empty(self.levelDesc[level].nSize * self.nAux)
        check(lib.musb200_aux_download(level, out.ctypes.data))
        return out

    def aux_probe(self, level, elemPos):
        self._bind()
        """tracking of one element (1-based position in the level's total list): rho, ux, uy, uz"""
        out = np.empty(self.nAux)
        check(lib.musb200_aux_probe(level, int(elemPos), ptr(out, P_DBL)))
        return out

    def download_neigh(self, level):
        self._bind()
        ld = self.levelDesc[level]
        out = np.zeros(ld.nSize * self.QQ, dtype=np.int32)
        check(lib.musb200_neigh_download(level, ptr(out, P_I32)))
        return out

    def set_bc_values(self, level, bc_id, vals):
        self._bind()
        v = np.ascontiguousarray(vals, dtype=np.float64)
        check(lib.musb200_bc_set_values(level, bc_id, v.size, v.ctypes.data))

    # -- restart: mus_pdf_serialize / mus_pdf_unserialize ---------------------------
    def pdf_serialize(self, treeID, levelPointer):
        """state(:, nNext) of the chunk's elements in treeID order, QQ values per element"""
        self._bind()
        t = np.ascontiguousarray(treeID, dtype=np.int64)
        lp = np.ascontiguousarray(levelPointer, dtype=np.int32)
        buf = np.empty(t.size * self.QQ)
        check(lib.musb200_pdf_serialize(t.size, ptr(t, P_I64), ptr(lp, P_I32), ptr(buf, P_DBL)))
        return buf

    def pdf_unserialize(self, treeID, levelPointer, buffer):
        self._bind()
        t = np.ascontiguousarray(treeID, dtype=np.int64)
        lp = np.ascontiguousarray(levelPointer, dtype=np.int32)
        b = np.ascontiguousarray(buffer, dtype=np.float64)
        if b.size != t.size * self.QQ:
            raise ValueError("restart buffer: QQ values per element expected")
        check(lib.musb200_pdf_unserialize(t.size, ptr(t, P_I64), ptr(lp, P_I32), ptr(b, P_DBL)))

    def write_restart(self, prefix, sim_name, time, mesh="./mesh/", elem_offset=0, nElems_global=None,
                      write_header=True, **header_kw):
        """mus_writeRestart (mus_restart_module.f90:57-166): the fluid PDFs of every level in
        tree order -> <prefix><sim_name>_<stamp>.lsb + header scripts (restart_io.write_restart).
        time = dict(sim=..., iter=...).  Returns (binary path, header path)."""
        from . import restart_io
        tid, lp = restart_io.tree_order(self.levelDesc)
        vs = restart_io.fluid_varsys(self.scheme_kind, self.QQ)
        if self.passive_scalar:
            vs.update(nAuxScalars=1, nAuxVars=1)
        return restart_io.write_restart(prefix, sim_name, self.pdf_serialize(tid, lp), time, vs, mesh=mesh,
                                        elem_offset=elem_offset, nElems_global=nElems_global,
                                        write_header=write_header, **header_kw)

    def read_restart(self, header_path, rank=0, nranks=1, base_dir=None, elem_offset=None):
        """mus_readRestart (mus_restart_module.f90:172-246): this rank's share of the binary file
        -> state(:, nNext) of the fluid elements; returns the header's time_point.  Ghosts and
        halos are not in the file (the reference re-fills them by interpolation / exchange).
        elem_offset: tree%elemOffset of this rank when the mesh is not distributed in treelm's
        equal shares (a weighted partition); the element count is the scheme's own."""
        from . import restart_io
        tid, lp = restart_io.tree_order(self.levelDesc)
        if elem_offset is None:
            rf, off, data = restart_io.read_restart(header_path, rank, nranks, base_dir)
        else:
            rf = restart_io.RestartFile(header_path, base_dir)
            data = rf.read(int(elem_offset), tid.size)
        if rf.nScalars * rf.nDofs != self.QQ or data.shape[0] != tid.size:
            raise ValueError("restart file %s: %d elements x %d scalars, this scheme holds %d x %d"
                             % (header_path, data.shape[0], rf.nScalars * rf.nDofs, tid.size, self.QQ))
        self.pdf_unserialize(tid, lp, data.ravel())
        # the file holds fluid elements only: halos, ghosts and auxField are rebuilt as
        # mus_init_flow does after mus_readRestart (mus_flow_module.fpp:206-240)
        self.fill_helper_elements()
        return rf.time

    def fill_helper_elements(self):
        """mus_initAuxField + fillHelperElementsFineToCoarse / CoarseToFine on the device: auxField
        of the fluid elements from state(:, nNext), halo exchange, ghost interpolation
        (collective over the ranks)"""
        self._bind()
        check(lib.musb200_fill_helper_elements(self.minLevel, self.maxLevel))

    # -- peer-memory halo exchange ------------------------------------------------
    def p2p_connect(self, dist, level=None):
        """set-up of the peer-memory halo exchange over a torch.distributed (gloo) group: the
        blobs are all-gathered and every receiver ships its recv position list to its sender
        -- what the Fortran shim does with MPI_Allgather / MPI_Sendrecv."""
        import torch
        level = self.minLevel if level is None else level
        ld = self.levelDesc[level]
        blob = ctypes.create_string_buffer(P2P_BLOB)
        check(lib.musb200_p2p_export(level, blob))
        mine = torch.frombuffer(bytearray(blob.raw), dtype=torch.uint8).clone()
        allb = [torch.zeros(P2P_BLOB, dtype=torch.uint8) for _ in range(dist.get_world_size())]
        dist.all_gather(allb, mine)
        proc, nVals, rpos = exchange_recv_lists(dist, ld)
        blobs = b"".join(bytes(allb[int(p)].numpy().tobytes()) for p in proc)
        check(lib.musb200_p2p_connect(level, len(proc), ptr(proc, P_I32), ctypes.c_char_p(blobs),
                                      ptr(nVals, P_I32), ptr(rpos, P_I32)))
        dist.barrier()

    # -- control%do_computation ----------------------------------------------
    def do_computation(self, nCycles=1):
        self._bind()
        check(lib.musb200_step(self.minLevel, self.maxLevel, int(nCycles)))

    def synchronize(self):
        check(lib.musb200_synchronize())

    def reduce(self, level=None):
        self._bind()
        level = self.minLevel if level is None else level
        m, v, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        check(lib.musb200_reduce(level, ctypes.byref(m), ctypes.byref(v), ctypes.byref(n)))
        return m.value, v.value, n.value

    def destroy(self):
        self._bind()
        for lvl in list(self.levelDesc):
            lib.musb200_level_destroy(lvl)


def step_schemes(schemes, nCycles=1):
    """several schemes on one mesh stepped together (musb200_step_schemes): within every level step
    they advance in the order given -- the flow first, then the passive scalar that reads its
    velocity -- and then fill their ghosts"""
    slots = (ctypes.c_int * len(schemes))(*[sc.slot for sc in schemes])
    lo, hi = schemes[0].minLevel, schemes[0].maxLevel
    if any((sc.minLevel, sc.maxLevel) != (lo, hi) for sc in schemes):
        raise ValueError("coupled schemes must live on the same levels")
    check(lib.musb200_step_schemes(len(schemes), slots, lo, hi, int(nCycles)))


def exchange_recv_lists(dist, ld):
    """host part of the peer-memory set-up over a torch.distributed group: every receiver ships
    its recv position list to the rank it receives from (MPI_Sendrecv in the Fortran shim).
    Returns (proc, nVals, remotePos) in the order of ld.send: remotePos = for every link I send,
    the state position it has on the receiving rank."""
    import torch
    reqs, got = [], {}
    for r in ld.recv:                                   # my recv list goes to its sender
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(r["pos"], dtype=np.int32)), r["proc"]))
    for s_ in ld.send:                                  # the receiver's list for my message
        got[s_["proc"]] = torch.zeros(len(s_["pos"]), dtype=torch.int32)
        reqs.append(dist.irecv(got[s_["proc"]], s_["proc"]))
    for q in reqs:
        q.wait()
    proc = np.array([s_["proc"] for s_ in ld.send], dtype=np.int32)
    nVals = np.array([len(s_["pos"]) for s_ in ld.send], dtype=np.int32)
    rpos = (np.concatenate([got[int(p)].numpy() for p in proc]).astype(np.int32)
            if len(proc) else np.zeros(0, dtype=np.int32))
    return proc, nVals, rpos


def compute_host(identify, inState, neigh, nElems, nSolve, omega, lambda_=0.25, omega_bulk=1.0):
    """strict drop-in of the `kernel` interface with host arrays
    (mus_scheme_type_module.f90:204-235); returns (outState, auxField)."""
    relax, kind, QQ = select_kernel(identify)
    a = np.ascontiguousarray(inState, dtype=np.float64)
    out = np.zeros_like(a)
    aux = np.zeros(nElems * 4)
    om = np.ascontiguousarray(omega, dtype=np.float64)
    ng = np.ascontiguousarray(neigh, dtype=np.int32)
    check(lib.musb200_compute_host(relax, kind, QQ, ptr(a, P_DBL), ptr(out, P_DBL), ptr(aux, P_DBL),
                                   ptr(ng, P_I32), int(nElems), int(nSolve), ptr(om, P_DBL),
                                   float(lambda_), float(omega_bulk)))
    return out, aux


def multilevel_tables(levels, intp):
    """flat dependency tables {(level, 'fromFiner' | ('fromCoarser', order)): arrays} of a
    multi-level mesh (levelDesc%depFromFiner / depFromCoarser / intpFromCoarser(order))."""
    from . import treelm_multilevel as tm
    t = {}
    for lvl, L in levels.items():
        if L.nGhostFromFiner:
            t[(lvl, "fromFiner")] = tm.intp_tables(L, intp, "fromFiner")
        if L.nGhostFromCoarser:
            for o in range(intp["order"] + 1):
                t[(lvl, ("fromCoarser", o))] = tm.intp_tables(L, intp, "fromCoarser", o)
    return t
