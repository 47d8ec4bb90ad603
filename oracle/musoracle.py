"""musoracle.py -- CPU ORACLE driver (TEST INFRASTRUCTURE ONLY).

numpy restatement of the host-side pieces of the reference that surround the
per-level LBM time step, plus ctypes access to the C restatement of the
kernels (oracle/*.c -> oracle/_build/libmusoracle.so).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (musubi_b200/) never does.

What is restated here (reference file:line):
  * Morton treeIDs, periodic wrap      tem/source/tem_topology_module.f90:88-108, 528-638
  * predefined cube + SFC partition    tem/source/treelmesh_module.f90:1224-1318 (:1276-1296)
  * total list [fluid|gFC|gFF|halo]    tem/source/tem_construction_module.f90:2358-2460
  * state connectivity                 mus/source/mus_connectivity_module.fpp:73-179 (C side)
  * reduced halo link lists            mus/source/mus_construction_module.fpp:1162-1356
  * boundary element / link lists      mus/source/mus_construction_module.fpp:2203-2376,
                                       mus/source/bc/mus_bc_header_module.fpp:1702-1739, 1876-1967
  * initial condition                  mus/source/mus_flow_module.fpp:484-589
  * unit conversion                    mus/source/mus_physics_module.f90:511-580
  * single-level schedule              mus/source/mus_control_module.f90:507-701
All index lists are kept 1-based exactly as the Fortran arrays hold them.

Pinning (tests/test_oracle_golden.py, tests/golden/).  PINNED by the
reference's own golden result files: connectivity, auxField, omega update, BGK D3Q19 (fluid and
fluid_incompressible), MRT D3Q19 fluid_incompressible, mus_init_pdf incl. the acoustic f_neq, unit
conversion, halo lists (two partitions) -- gaussianPulse (fluid: level 4; fluid_incompressible:
levels 4, 5, 6 + initial states), TGV_Simple_Re800, TGV_Simple_Re1600; the restated cases are
checked against the reference's own musubi.lua scripts (oracle/lua_ref.py).  Restated utest
properties pin BGK / MRT optimised vs NoOpt kernels, rest-state fixed points, M M^-1 = I, f_neq.
PARITY UNPINNED by any reference fixture reproducible here: TRT D3Q19, BGK / TRT / MRT D3Q27 as
whole runs, velocity_bounceback and the pressure boundaries in 3-D, all ghost interpolation
routines, the force source terms, the passive-scalar kernels (their reference cases need Seeder
meshes or have deactivated utests); these rest on analytic checks and oracle <-> device equality.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BGK, TRT, MRT = 0, 1, 2
RELAX = {"bgk": BGK, "trt": TRT, "mrt": MRT}
PRP_FLUID, PRP_SOLID, PRP_HASBND, PRP_SENDHALO = 1, 2, 3, 12

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_int64)


class _Relax(ctypes.Structure):
    _fields_ = [("lambda_", ctypes.c_double), ("omegaBulk", ctypes.c_double)]


def build(force=False):
    """compile oracle/*.c with the committed Makefile (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "_build", "libmusoracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ora_omega_bulk.restype = ctypes.c_double
        _LIB.ora_omega_bulk.argtypes = [ctypes.c_double]
        _LIB.ora_total_mass.restype = ctypes.c_double
        _LIB.ora_cxDir.restype = ctypes.POINTER(ctypes.c_int)
        _LIB.ora_cxDirInv.restype = ctypes.POINTER(ctypes.c_int)
        _LIB.ora_weights.restype = _dp
        _LIB.ora_mrt_matrix.restype = _dp
        _LIB.ora_nEq_acoustic.argtypes = [ctypes.c_int, ctypes.c_double, _dp, _dp]
        _LIB.ora_apply_src_force.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp,
                                             ctypes.c_double, ctypes.c_int, _ip, _dp]
        _LIB.ora_compute_passive_scalar.argtypes = [ctypes.c_int, ctypes.c_int, _dp, _dp, _ip,
                                                    ctypes.c_int, ctypes.c_int, _dp, ctypes.c_double,
                                                    ctypes.c_double]
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _l(a):
    return a.ctypes.data_as(_lp)


# --------------------------------------------------------------------------
# stencil tables
# --------------------------------------------------------------------------
def cx_dir(QQ):
    p = lib().ora_cxDir(QQ)
    return np.array([[p[3 * i + k] for k in range(3)] for i in range(QQ)], dtype=np.int64)


def cx_dir_inv(QQ):
    p = lib().ora_cxDirInv(QQ)
    return np.array([p[i] for i in range(QQ)], dtype=np.int64)  # 1-based values


def weights(QQ):
    p = lib().ora_weights(QQ)
    return np.array([p[i] for i in range(QQ)])


# --------------------------------------------------------------------------
# treelm topology
# --------------------------------------------------------------------------
def first_id_at_level(level):
    return (8 ** level - 1) // 7


def _spread3(v):
    """insert two zero bits between the bits of v (21-bit input)."""
    v = v.astype(np.uint64)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def _compact3(v):
    v = v.astype(np.uint64) & np.uint64(0x1249249249249249)
    v = (v | (v >> np.uint64(2))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v >> np.uint64(4))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v >> np.uint64(8))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v >> np.uint64(16))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v >> np.uint64(32))) & np.uint64(0x1FFFFF)
    return v


def morton_of_coord(x, y, z):
    """x -> bit 0, y -> bit 1, z -> bit 2 of every octal digit (tem_IdOfCoord)."""
    return (_spread3(np.asarray(x)) | (_spread3(np.asarray(y)) << np.uint64(1))
            | (_spread3(np.asarray(z)) << np.uint64(2))).astype(np.int64)


def coord_of_morton(m):
    m = np.asarray(m).astype(np.uint64)
    return (_compact3(m).astype(np.int64), _compact3(m >> np.uint64(1)).astype(np.int64),
            _compact3(m >> np.uint64(2)).astype(np.int64))


def id_of_coord(x, y, z, level):
    n = 1 << level
    return first_id_at_level(level) + morton_of_coord(np.mod(x + n, n), np.mod(y + n, n), np.mod(z + n, n))


def coord_of_id(tid, level):
    return coord_of_morton(np.asarray(tid) - first_id_at_level(level))


def partition_ranges(nElems, nParts):
    """contiguous equal shares; the first `remainder` parts get one more."""
    share, rem = divmod(nElems, nParts)
    first, out = 0, []
    for p in range(nParts):
        n = share + (1 if p < rem else 0)
        out.append((first, first + n))
        first += n
    return out


# --------------------------------------------------------------------------
# level descriptor of a single-level box (periodic cube or walled cavity)
# --------------------------------------------------------------------------
class LevelDesc:
    pass


def box_extents(level, octants):
    """the domain is the first `octants` (1, 2, 4, 8) octants of the level-L universe cube in
    Morton order: x doubles first, then y, then z (treelm meshes are sparse octrees; absent
    octants are simply not in the element list)."""
    assert octants in (1, 2, 4, 8)
    n = 1 << level
    h = n // 2
    return (n if octants >= 2 else h, n if octants >= 4 else h, n if octants >= 8 else h)


def box_boundary_id(xn, yn, zn, n, kind):
    """BC id seen when looking from a fluid cell to position (xn,yn,zn).
    kind 'periodic': no boundaries (treelm wraps at the universe cube).
    kind 'cavity'  : id 2 ('lid', velocity_bounceback) where only the top plane
                     z = n is crossed; id 1 ('wall') for every other exit.
    kind 'channel' : id 1 ('wall') where a y or z plane is crossed; else id 2 ('inlet',
                     velocity_bounceback) at x < 0 and id 3 ('outlet', a pressure boundary) at x >= n.
    The synthetic-mesh convention (Seeder would store these ids in bnd.lsb)."""
    if kind == "periodic":
        return np.zeros(xn.shape, dtype=np.int64)
    nx, ny, nz = n if isinstance(n, tuple) else (n, n, n)
    if kind == "channel":
        out_yz = (yn < 0) | (yn >= ny) | (zn < 0) | (zn >= nz)
        bid = np.zeros(xn.shape, dtype=np.int64)
        bid[out_yz] = 1
        bid[(~out_yz) & (xn < 0)] = 2
        bid[(~out_yz) & (xn >= nx)] = 3
        return bid
    out_xy = (xn < 0) | (xn >= nx) | (yn < 0) | (yn >= ny)
    bid = np.zeros(xn.shape, dtype=np.int64)
    bid[out_xy | (zn < 0)] = 1
    bid[(~out_xy) & (zn >= nz)] = 2
    return bid


def build_level_desc(level, QQ, kind="periodic", rank=0, nranks=1, comm_reduced=True,
                     _with_send=True, octants=8):
    """total list, neighbour lists, connectivity, halo and BC lists of one rank."""
    ld = LevelDesc()
    n = 1 << level
    dims = box_extents(level, octants)
    if octants != 8 and kind == "periodic":
        raise ValueError("treelm wraps at the universe cube: a periodic mesh must fill it")
    nGlob = octants * (n // 2) ** 3
    cx = cx_dir(QQ)
    inv = cx_dir_inv(QQ)
    QQN = QQ - 1
    ranges = partition_ranges(nGlob, nranks)
    lo, hi = ranges[rank]
    nFluid = hi - lo
    first = first_id_at_level(level)
    mloc = np.arange(lo, hi, dtype=np.int64)
    x, y, z = coord_of_morton(mloc)

    # neighbour morton index per direction (or -bcid)
    ngh_m = np.empty((nFluid, QQN), dtype=np.int64)
    for d in range(QQN):
        xn, yn, zn = x + cx[d, 0], y + cx[d, 1], z + cx[d, 2]
        bid = box_boundary_id(xn, yn, zn, dims, kind)
        m = morton_of_coord(np.mod(xn, n), np.mod(yn, n), np.mod(zn, n))
        ngh_m[:, d] = np.where(bid > 0, -bid, m)
    remote = (ngh_m >= 0) & ((ngh_m < lo) | (ngh_m >= hi))
    halo_m = np.unique(ngh_m[remote])
    nHalo = halo_m.size
    nElems = nFluid + nHalo
    ld.level, ld.QQ, ld.kind, ld.octants = level, QQ, kind, octants
    ld.nFluid, ld.nGhostFromCoarser, ld.nGhostFromFiner, ld.nHalo = nFluid, 0, 0, nHalo
    ld.nElems = nElems
    ld.nSize = ((nElems + 3) // 4) * 4          # mus_pdf_module.f90:128-129
    ld.nSolve = nFluid                          # nElems_fluid + nElems_ghostFromCoarser
    ld.total = np.concatenate([first + mloc, first + halo_m]).astype(np.int64)

    # nghElems(QQN, nElems): 1-based positions in the total list, <= 0 = none / -bcid
    ngh = np.zeros((nElems, QQN), dtype=np.int32)
    loc = (ngh_m >= lo) & (ngh_m < hi)
    pos = np.where(loc, ngh_m - lo + 1, 0)
    if nHalo:
        hp = np.searchsorted(halo_m, np.where(remote, ngh_m, halo_m[0]))
        pos = np.where(remote, nFluid + hp + 1, pos)
    pos = np.where(ngh_m < 0, ngh_m, pos)
    ngh[:nFluid] = pos
    if nHalo:  # halos: neighbours that exist locally (fluid or halo), else 0
        hx, hy, hz = coord_of_morton(halo_m)
        for d in range(QQN):
            xn, yn, zn = hx + cx[d, 0], hy + cx[d, 1], hz + cx[d, 2]
            bid = box_boundary_id(xn, yn, zn, dims, kind)
            m = morton_of_coord(np.mod(xn, n), np.mod(yn, n), np.mod(zn, n))
            isloc = (m >= lo) & (m < hi) & (bid == 0)
            hp = np.searchsorted(halo_m, m)
            hp_c = np.minimum(hp, nHalo - 1)
            ishalo = (~isloc) & (bid == 0) & (halo_m[hp_c] == m)
            p = np.where(isloc, m - lo + 1, np.where(ishalo, nFluid + hp_c + 1, 0))
            ngh[nFluid:, d] = np.where(bid > 0, -bid, p)
    ld.nghElems = ngh

    prop = np.zeros(nElems, dtype=np.int64)
    prop[:nFluid] |= (1 << PRP_FLUID)
    hasbnd = (ngh[:nFluid] < 0).any(axis=1)
    prop[:nFluid][hasbnd] |= (1 << PRP_HASBND)
    ld.property = prop

    ld.neigh = np.zeros(QQ * ld.nSize, dtype=np.int32)
    lib().ora_construct_connectivity(_i(ld.neigh), ld.nSize, nElems, QQ, _i(ld.nghElems),
                                     _l(ld.property), nFluid, nFluid)

    # ---------------- halo exchange lists (reduced link set) ----------------
    ld.recv, ld.send, ld.recv_masks = [], [], []
    if nranks > 1:
        owner = np.searchsorted(np.array([r[1] for r in ranges]), halo_m, side="right")
        for p in range(nranks):
            hsel = np.nonzero(owner == p)[0]
            if hsel.size == 0:
                continue
            epos = nFluid + hsel + 1
            posl, mask = _recv_positions(ld, epos, QQ, inv, nFluid, comm_reduced)
            ld.recv.append(dict(proc=p, elemPos=epos.astype(np.int32), pos=posl))
            ld.recv_masks.append(mask)
        if _with_send:
            # init_sendBuffers (mus_construction_module.fpp:1297-1356): the peer's halo
            # list and link bitmask, replayed here instead of being sent over MPI.
            for p in range(nranks):
                if p == rank:
                    continue
                other = _peer_desc(level, QQ, kind, p, nranks, comm_reduced, octants)
                for r, m in zip(other.recv, other.recv_masks):
                    if r["proc"] != rank:
                        continue
                    tids = other.total[r["elemPos"] - 1]
                    epos = (tids - first - lo + 1).astype(np.int32)
                    ei, di = np.nonzero(m)
                    posl = ((epos[ei].astype(np.int64) - 1) * QQ + di + 1).astype(np.int32)
                    ld.send.append(dict(proc=p, elemPos=epos, pos=posl))
            sendmask = np.zeros(nElems, dtype=bool)
            for snd in ld.send:
                sendmask[snd["elemPos"] - 1] = True
            ld.property[sendmask] |= (1 << PRP_SENDHALO)

    # ---------------- boundary lists ----------------------------------------
    ld.bc = []
    bcs = {"cavity": ((1, "wall", "wall"), (2, "lid", "velocity_bounceback")),
           "channel": ((1, "wall", "wall"), (2, "inlet", "velocity_bounceback"),
                       (3, "outlet", "pressure"))}.get(kind, ())
    if bcs:
        ld.bc_elemBuffer = (np.nonzero(hasbnd)[0] + 1).astype(np.int32)   # levelDesc%bc_elemBuffer
        posInBuf = np.zeros(nElems + 1, dtype=np.int32)
        posInBuf[ld.bc_elemBuffer] = np.arange(1, ld.bc_elemBuffer.size + 1)
        # weights of the normal: 4 / 2 / 1 for axis / edge / corner directions (assignBCList,
        # mus_construction_module.fpp:2237-2252); prevailing directions = normalised stencil
        clen = (cx[:QQN] ** 2).sum(axis=1)
        wgt = np.where(clen == 1, 4, np.where(clen == 2, 2, 1)).astype(np.int64)
        prevail = cx[:QQN].astype(np.float64) / np.sqrt(clen.astype(np.float64))[:, None]
        for bid, label, bkind in bcs:
            hit = (ngh[:nFluid] == -bid)                      # [elem][k]: boundary in direction k
            elems = np.nonzero(hit.any(axis=1))[0] + 1
            bitmask = np.zeros((elems.size, QQN), dtype=bool)
            for k in range(QQN):                              # bitmask(cxDirInv(k)) = .true.
                bitmask[:, inv[k] - 1] |= hit[elems - 1, k]
            e_idx, d_idx = np.nonzero(bitmask)                # elem-major, dir ascending
            el = elems[e_idx]
            dirs = d_idx + 1
            links = ld.neigh[(dirs - 1) * ld.nSize + el - 1]  # FETCH(iDir, elem)
            pib = posInBuf[el]
            outPos = inv[dirs - 1] + (pib.astype(np.int64) - 1) * QQ
            bc = dict(id=bid, label=label, kind=bkind, elems=elems.astype(np.int32),
                      bitmask=bitmask, links=links.astype(np.int32),
                      iDir=dirs.astype(np.int32), posInBuffer=pib.astype(np.int32),
                      outPos=outPos.astype(np.int32), elemOfLink=el.astype(np.int32))
            # element normals: -sum_k weight_k c_k over the boundary directions, aligned to the
            # best prevailing direction (normalizeBC :2393-2440, tem_determine_discreteVector)
            nrm = -(hit[elems - 1].astype(np.int64) * wgt[None, :]) @ cx[:QQN]
            nlen = np.sqrt((nrm.astype(np.float64) ** 2).sum(axis=1))
            dots = np.clip((nrm / nlen[:, None]) @ prevail.T, -1.0, 1.0)
            best = np.zeros(elems.size, dtype=np.int64)
            for i in range(elems.size):                       # first strict maximum, exit at 1
                mx = -2.0
                for k in range(QQN):
                    if dots[i, k] > mx:
                        mx, best[i] = dots[i, k], k
                        if abs(mx - 1.0) <= np.finfo(float).eps:
                            break
            bc["normal"] = cx[best].astype(np.int32)
            bc["normalInd"] = (best + 1).astype(np.int32)
            bc["posInBcElemBuf"] = posInBuf[elems].astype(np.int32)
            # iElem / statePos per link (mus_set_outletExpol, mus_bc_header_module.fpp:2256-2319)
            bc["iElemOfLink"] = (e_idx + 1).astype(np.int32)
            bc["statePos"] = (dirs + e_idx * QQ).astype(np.int32)
            # neighbours along the inward normal (setFieldBCNeigh, mus_construction_module.fpp:
            # 1733-1840): position of the element at x + k*normal, k = 1..2; a missing first
            # neighbour -> the element itself, a missing second -> the first
            ex, ey, ez = x[elems - 1], y[elems - 1], z[elems - 1]
            npos = np.zeros((elems.size, 2), dtype=np.int32)
            for k in (1, 2):
                xn, yn, zn = ex + k * bc["normal"][:, 0], ey + k * bc["normal"][:, 1], ez + k * bc["normal"][:, 2]
                inside = ((xn >= 0) & (xn < dims[0]) & (yn >= 0) & (yn < dims[1]) & (zn >= 0)
                          & (zn < dims[2]))
                m = morton_of_coord(np.mod(xn, n), np.mod(yn, n), np.mod(zn, n))
                p = np.where((m >= lo) & (m < hi), m - lo + 1, 0)
                if nHalo:
                    hp = np.minimum(np.searchsorted(halo_m, m), nHalo - 1)
                    p = np.where((p == 0) & (halo_m[hp] == m), nFluid + hp + 1, p)
                p = np.where(inside, p, 0)
                prev = elems if k == 1 else npos[:, 0]
                npos[:, k - 1] = np.where(p > 0, p, prev)
            npos[:, 1] = np.where(npos[:, 0] == elems, elems, npos[:, 1])
            bc["neighPos"] = npos
            ld.bc.append(bc)
    else:
        ld.bc_elemBuffer = np.zeros(0, dtype=np.int32)
    return ld


def _recv_positions(ld, epos, QQ, inv, halo_offset, comm_reduced):
    """init_recvBuffers (mus_construction_module.fpp:1203-1261): elem-major, dir ascending;
    link iDir of halo h is received when the element h pulls inv(iDir) from is local."""
    e = epos.astype(np.int64)
    keep = np.zeros((e.size, QQ), dtype=bool)
    for d in range(1, QQ + 1):
        neighDir = inv[d - 1]
        nghElem = (ld.neigh[(neighDir - 1) * ld.nSize + e - 1].astype(np.int64) - 1) // QQ + 1
        keep[:, d - 1] = (nghElem <= halo_offset) | (not comm_reduced)
    ei, di = np.nonzero(keep)
    return ((e[ei] - 1) * QQ + di + 1).astype(np.int32), keep


_PEER_CACHE = {}


def _peer_desc(level, QQ, kind, p, nranks, comm_reduced, octants=8):
    key = (level, QQ, kind, p, nranks, comm_reduced, octants)
    if key not in _PEER_CACHE:
        _PEER_CACHE[key] = build_level_desc(level, QQ, kind, p, nranks, comm_reduced,
                                            _with_send=False, octants=octants)
    return _PEER_CACHE[key]


# --------------------------------------------------------------------------
# physics / unit conversion (mus_physics_module.f90:511-580)
# --------------------------------------------------------------------------
class Physics:
    def __init__(self, dx, dt, rho0=1.0):
        self.dx, self.dt, self.rho0 = dx, dt, rho0
        self.fac_vel = dx / dt
        self.fac_visc = dx ** 2 / dt
        self.fac_press = rho0 * dx ** 2 / dt ** 2
        self.fac_strainRate = 1.0 / dt


def barycenters(ld, origin, length):
    """tem_BaryOfId (tem_geometry_module.f90:419-435)."""
    dx = length / float(1 << ld.level)
    x, y, z = coord_of_id(ld.total, ld.level)
    o = np.asarray(origin, dtype=np.float64)
    return np.stack([o[0] + (x.astype(np.float64) + 0.5) * dx,
                     o[1] + (y.astype(np.float64) + 0.5) * dx,
                     o[2] + (z.astype(np.float64) + 0.5) * dx], axis=1)


# --------------------------------------------------------------------------
# the scheme: state arrays + one level step
# --------------------------------------------------------------------------
class Scheme:
    """mus_scheme_type restricted to what the single-level step touches."""

    def __init__(self, ld, relaxation="bgk", kind="fluid", omega=1.0, lambda_=0.25,
                 omega_bulk=None):
        if kind not in ("fluid", "fluid_incompressible"):
            raise ValueError("scheme kind %r is outside the hot path" % kind)
        self.ld = ld
        self.QQ = ld.QQ
        self.relax = RELAX[relaxation]
        self.incomp = 1 if kind == "fluid_incompressible" else 0
        n = ld.nSize * self.QQ
        self.state = [np.full(n, -1.0e6), np.full(n, -1.0e6)]   # poison as in mus_construct :513
        self.aux = np.full(ld.nSize * 4, -1.0e6)
        self.nNow, self.nNext = 0, 1
        self.visc = np.full(ld.nSize, (1.0 / omega - 0.5) / 3.0)
        self.omega = np.full(ld.nSize, float(omega))
        self.omega_uniform = float(omega)
        self.rp = _Relax(lambda_, omega if omega_bulk is None else omega_bulk)
        self.bc_vel = {}      # bc id -> [nLinks][3] lattice velocity per link
        self.bc_rho = {}      # bc id -> [nElems] lattice density of a pressure boundary
        # the mesh only knows the boundary id; 'pressure' ids are bound to a kind here
        # (boundary_condition table of the Lua config): pressure_expol | pressure_antibounceback
        self.bc_kind = {bc["id"]: bc["kind"] for bc in ld.bc}
        self.bcBuffer = np.zeros(max(1, ld.bc_elemBuffer.size) * self.QQ)
        self.exchange = None  # callable(state) for multi-rank runs
        self.force = None     # (order, posInTotal int32 1-based, force [n][3] lattice units)

    # -- source = { force = ... } (mus_source_module.f90:83-310: all nSolve elements whose
    #    barycentre lies in the variable's shape; here: a list of positions or all of them)
    def set_force(self, force, order=2, posInTotal=None):
        ld = self.ld
        pos = (np.arange(1, ld.nSolve + 1, dtype=np.int32) if posInTotal is None
               else np.ascontiguousarray(posInTotal, dtype=np.int32))
        F = np.ascontiguousarray(np.broadcast_to(np.asarray(force, dtype=np.float64), (pos.size, 3)))
        if order not in (1, 2):
            raise ValueError("force source: order must be 1 or 2")
        self.force = (int(order), pos, F)

    def add_src_to_aux(self):
        """field%source%method%addSrcToAuxField (mus_auxField_module.f90:341-375)"""
        if self.force is not None and self.force[0] == 2:
            _, pos, F = self.force
            lib().ora_add_force_to_aux(_d(self.aux), self.incomp, int(pos.size), _i(pos), _d(F))

    def apply_source_terms(self):
        """mus_apply_sourceTerms (mus_source_module.f90:430-512) on state(:, nNext)"""
        if self.force is not None:
            order, pos, F = self.force
            rc = lib().ora_apply_src_force(self.relax, self.QQ, order, _d(self.state[self.nNext]),
                                           _d(self.aux), _d(self.omega), self.rp.omegaBulk,
                                           int(pos.size), _i(pos), _d(F))
            if rc != 0:
                raise RuntimeError("no oracle force source for this configuration")

    # -- initial condition: f = fEq(rho, u) (+ fNeq(S) = 0), mus_init_pdf -----
    def init_equilibrium(self, rho, vel):
        ld, QQ = self.ld, self.QQ
        rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (ld.nElems,)))
        vel = np.ascontiguousarray(np.broadcast_to(np.asarray(vel, dtype=np.float64), (ld.nElems, 3)))
        st = self.state[self.nNext]
        lib().ora_init_equilibrium(QQ, self.incomp, ld.nElems, _d(rho), _d(vel), _d(st))
        self.state[self.nNow][:] = st            # mus_flow_module.fpp:181-185
        self.calc_aux(self.state[self.nNext], local_only=True)

    # -- mus_init_pdf with the strain-rate part: f = fEq(rho, u) + fNeq(omega, S) ----------
    def init_pdf(self, rho, vel, S6):
        """S6[nElems][6] = (Sxx, Syy, Szz, Sxy, Syz, Sxz) in lattice units (mus_flow_module.fpp:484-589)"""
        ld, QQ = self.ld, self.QQ
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        vel = np.ascontiguousarray(vel, dtype=np.float64)
        S6 = np.ascontiguousarray(S6, dtype=np.float64)
        assert rho.shape == (ld.nElems,) and vel.shape == (ld.nElems, 3) and S6.shape == (ld.nElems, 6)
        omega = np.ascontiguousarray(1.0 / (3.0 * self.visc[:ld.nElems] + 0.5))
        st = self.state[self.nNext]
        lib().ora_init_pdf(QQ, self.incomp, ld.nElems, _d(rho), _d(vel), _d(S6), _d(omega), _d(st))
        self.state[self.nNow][:] = st
        self.calc_aux(self.state[self.nNext], local_only=True)

    def calc_aux(self, state, local_only=False):
        L = lib()
        fn = L.ora_calc_aux_incomp if self.incomp else L.ora_calc_aux
        if local_only:
            # initial aux: moments of the element's own PDFs (mus_init_aux ... initAuxField)
            ident = np.zeros(self.QQ * self.ld.nSize, dtype=np.int32)
            e = np.arange(self.ld.nSize, dtype=np.int64)
            for d in range(self.QQ):
                ident[d * self.ld.nSize:(d + 1) * self.ld.nSize] = e * self.QQ + d + 1
            fn(self.QQ, _d(self.aux), _d(state), _i(ident), self.ld.nSize, self.ld.nElems)
        else:
            fn(self.QQ, _d(self.aux), _d(state), _i(self.ld.neigh), self.ld.nSize, self.ld.nSolve)

    def set_boundary(self):
        ld, L = self.ld, lib()
        if not ld.bc:
            return
        st = self.state[self.nNext]
        L.ora_fill_bcBuffer(_d(self.bcBuffer), _d(st), self.QQ, _i(ld.bc_elemBuffer),
                            int(ld.bc_elemBuffer.size))
        for bc in ld.bc:
            kind = self.bc_kind[bc["id"]]
            if kind == "wall":
                continue                        # do_nothing: bounce-back lives in neigh
            if kind in ("pressure_expol", "pressure_antibounceback"):
                nE = int(bc["elems"].size)
                rho = np.ascontiguousarray(np.broadcast_to(self.bc_rho[bc["id"]], (nE,)), dtype=np.float64)
                npos = np.ascontiguousarray(bc["neighPos"], dtype=np.int32)
                if kind == "pressure_expol":
                    nb = np.zeros(2 * nE * self.QQ)      # requireNeighBufPre_nNext, nNeighs = 2
                    L.ora_fill_neighBuffer(_d(nb), _d(st), _i(ld.neigh), ld.nSize, self.QQ, 2, nE,
                                           _i(npos), 0)
                    L.ora_pressure_expol(_d(st), _d(self.bcBuffer), _d(self.aux), _i(ld.neigh), ld.nSize,
                                         self.QQ, self.incomp, nE, _i(bc["elems"]),
                                         _i(bc["posInBcElemBuf"]), _i(bc["normalInd"]), _d(rho),
                                         int(bc["links"].size), _i(bc["links"]), _i(bc["statePos"]),
                                         _d(nb))
                else:
                    n1 = np.ascontiguousarray(npos[:, :1])  # requireNeighBufPost, nNeighs = 1
                    nb = np.zeros(nE * self.QQ)
                    L.ora_fill_neighBuffer(_d(nb), _d(st), _i(ld.neigh), ld.nSize, self.QQ, 1, nE,
                                           _i(n1), 1)
                    L.ora_pressure_antibounceback(_d(st), _d(self.bcBuffer), self.QQ, self.incomp, nE,
                                                  _i(bc["elems"]), _i(bc["posInBcElemBuf"]), _d(rho),
                                                  _d(self.omega), int(bc["links"].size),
                                                  _i(bc["links"]), _i(bc["iElemOfLink"]),
                                                  _i(bc["iDir"]), _d(nb))
                continue
            if kind == "velocity_bounceback":
                v = np.ascontiguousarray(self.bc_vel[bc["id"]], dtype=np.float64)
                L.ora_velocity_bounceback(_d(st), _d(self.bcBuffer), self.QQ,
                                          int(bc["links"].size), _i(bc["links"]), _i(bc["outPos"]),
                                          _i(bc["iDir"]), _i(bc["posInBuffer"]), _d(v),
                                          self.incomp)
            else:
                raise ValueError("boundary kind %r not restated" % kind)

    def step(self):
        """do_fast_singleLevel (mus_control_module.f90:507-701), steps 3-9."""
        L = lib()
        self.set_boundary()                                     # 3 (on state(:,nNext))
        self.nNow, self.nNext = self.nNext, self.nNow           # 4 mus_swap_now_next
        self.calc_aux(self.state[self.nNow])                    # 5
        self.add_src_to_aux()                                   # 5 (source -> auxField)
        L.ora_update_omega(_d(self.omega), _d(self.visc), self.ld.nSolve)   # 6
        rc = L.ora_compute(self.relax, self.QQ, self.incomp, _d(self.state[self.nNow]),
                           _d(self.state[self.nNext]), _d(self.aux), _i(self.ld.neigh),
                           _d(self.omega), self.ld.nSize, self.ld.nSolve,
                           ctypes.byref(self.rp))               # 7
        if rc != 0:
            raise RuntimeError("no oracle kernel for this (relaxation, layout, kind)")
        self.apply_source_terms()                               # 8
        if self.exchange is not None:
            self.exchange(self)                                 # 9 exchange_real(state(:,next))

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()

    def total_mass(self):
        return lib().ora_total_mass(_d(self.state[self.nNext]), self.QQ, self.ld.nFluid)

    def set_omega(self, omega):
        self.visc[:] = (1.0 / omega - 0.5) / 3.0
        self.omega[:] = omega
        self.omega_uniform = float(omega)


def level_of(treeID):
    """tem_LevelOf: the first treeID of level L is (8^L - 1) / 7"""
    level, first, count = 0, 0, 1
    while treeID >= first + count:
        first, count, level = first + count, count * 8, level + 1
    return level


def pdf_serialize(schemes, treeID, levelPointer):
    """mus_pdf_serialize (mus_buffer_module.fpp:80-137): buffer of a chunk of the global treeID
    list, QQ values of state(:, nNext) per element; schemes = {level: Scheme}"""
    QQ = next(iter(schemes.values())).QQ
    buf = np.empty(len(treeID) * QQ)
    for i, (t, p) in enumerate(zip(treeID, levelPointer)):
        s = schemes[level_of(int(t))]
        buf[i * QQ:(i + 1) * QQ] = s.state[s.nNext][(p - 1) * QQ:p * QQ]
    return buf


def pdf_unserialize(schemes, treeID, levelPointer, buf):
    """mus_pdf_unserialize (mus_buffer_module.fpp:147-190): only state(:, nNext) is set"""
    QQ = next(iter(schemes.values())).QQ
    for i, (t, p) in enumerate(zip(treeID, levelPointer)):
        s = schemes[level_of(int(t))]
        s.state[s.nNext][(p - 1) * QQ:p * QQ] = buf[i * QQ:(i + 1) * QQ]


def global_tree(levels):
    """the mesh's global treeID list (all fluid elements, space-filling-curve order) and the
    levelPointer into each level's total list, from per-level descriptors with `total`"""
    maxL = max(levels)
    keys, ids, ptrs = [], [], []
    for l, ld in levels.items():
        tid = np.asarray(ld.total[:ld.nFluid], dtype=np.int64)
        m = tid - first_id_at_level(l)
        keys.append(m << (3 * (maxL - l)))
        ids.append(tid)
        ptrs.append(np.arange(1, ld.nFluid + 1, dtype=np.int32))
    keys, ids, ptrs = np.concatenate(keys), np.concatenate(ids), np.concatenate(ptrs)
    o = np.argsort(keys, kind="stable")
    return ids[o], ptrs[o]


# --------------------------------------------------------------------------
# ghost dependency build (oracle/dependencies.c): independent of the product's generator
# --------------------------------------------------------------------------
def _block_start(ld):
    b = np.zeros(5, dtype=np.int32)
    b[1] = ld.nFluid
    b[2] = b[1] + ld.nGhostFromCoarser
    b[3] = b[2] + ld.nGhostFromFiner
    b[4] = b[3] + ld.nHalo
    return b


def ngh_elems_from_total(ld, QQ, solid_ids=None):
    """levelDesc%neigh(1)%nghElems of a (multi-level) descriptor from its total list alone: the
    stencil neighbour's treeID (periodic wrap at the universe cube, tem_topology_module.f90:
    590-638) looked up block by block (tem_treeIDinTotal); a neighbour that is a solid cell
    (obstacle, boundary id 1) gives -1, a missing one 0.  Returns [nElems][QQ-1], 1-based."""
    level = ld.level
    n = 1 << level
    cx = cx_dir(QQ)
    first = first_id_at_level(level)
    total = np.asarray(ld.total, dtype=np.int64)
    nEl = total.size
    x, y, z = coord_of_morton(total - first)
    blocks = _block_start(ld)
    solid = np.sort(np.asarray(solid_ids if solid_ids is not None else [], dtype=np.int64))
    out = np.zeros((nEl, QQ - 1), dtype=np.int32)
    for d in range(QQ - 1):
        nid = first + morton_of_coord(np.mod(x + cx[d, 0], n), np.mod(y + cx[d, 1], n), np.mod(z + cx[d, 2], n))
        pos = np.zeros(nEl, dtype=np.int64)
        for b in range(4):
            lo, hi = int(blocks[b]), int(blocks[b + 1])
            if hi == lo:
                continue
            blk = total[lo:hi]
            k = np.minimum(np.searchsorted(blk, nid), hi - lo - 1)
            hit = (blk[k] == nid) & (pos == 0)
            pos = np.where(hit, lo + k + 1, pos)
        if solid.size:
            k = np.minimum(np.searchsorted(solid, nid), solid.size - 1)
            pos = np.where((pos == 0) & (solid[k] == nid), -1, pos)
        out[:, d] = pos
    return out


class LsfStore:
    """tem_intpMatrixLSF_type of one interpolation order (shared by all levels)"""

    def __init__(self, order):
        L = lib()
        L.ora_lsf_new.restype = ctypes.c_void_p
        L.ora_lsf_delete.argtypes = [ctypes.c_void_p]
        L.ora_lsf_count.argtypes = [ctypes.c_void_p]
        L.ora_lsf_get.argtypes = [ctypes.c_void_p, ctypes.c_int, _ip, _ip, _ip, _ip, _dp]
        L.ora_lsf_append.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, _ip, _ip]
        self.order = order
        self.h = ctypes.c_void_p(L.ora_lsf_new(order))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_lsf_delete(self.h)
            self.h = None

    def append(self, QQ, neighDir):
        """append_intpMatrixLSF: returns (success, 1-based position)"""
        d = np.ascontiguousarray(neighDir, dtype=np.int32)
        pos = np.zeros(1, dtype=np.int32)
        ok = lib().ora_lsf_append(self.h, QQ, d.size, _i(d), _i(pos))
        return bool(ok), int(pos[0])

    def __len__(self):
        return lib().ora_lsf_count(self.h)

    def get(self, pos):
        """(hashID, invertible, matrix [nCoeffs][nSources]) of the 1-based position"""
        hid, inv, r, c = (np.zeros(1, dtype=np.int32) for _ in range(4))
        lib().ora_lsf_get(self.h, pos, _i(hid), _i(inv), _i(r), _i(c), None)
        A = np.zeros(int(r[0]) * int(c[0]))
        lib().ora_lsf_get(self.h, pos, _i(hid), _i(inv), _i(r), _i(c), _d(A))
        return int(hid[0]), bool(inv[0]), A.reshape(int(r[0]), int(c[0]))


def build_dependencies(levels, QQ, order_max, ngh=None):
    """tem_build_verticalDependencies + mus_intp_update_depFromCoarser over all levels of a
    multi-level mesh given as {level: descriptor with total, nFluid, nGhostFromCoarser,
    nGhostFromFiner, nHalo}; ngh: {level: nghElems} (default: the descriptors' own).
    Returns ({level: dict(fromFiner=(srcOffset, srcPos), fromCoarser=dict(parentPos, childNum,
    coord, order, nSrc, src, dir, posInMat, weights))}, {order: LsfStore})."""
    L = lib()
    L.ora_update_dep_from_coarser.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _ip, _ip, _dp, _ip,
                                              _ip, _ip, _ip, _ip, _ip, _dp, ctypes.c_void_p, ctypes.c_void_p]
    stores = {1: LsfStore(1), 2: LsfStore(2)}
    out = {}
    lv = sorted(levels)
    for l in lv:                                   # source level ascending, as the reference loops
        ld = levels[l]
        dep = {}
        total = np.ascontiguousarray(ld.total, dtype=np.int64)
        nGFC, nGFF = int(ld.nGhostFromCoarser), int(ld.nGhostFromFiner)
        if nGFC:
            C = levels[l - 1]
            ctot = np.ascontiguousarray(C.total, dtype=np.int64)
            cb = _block_start(C)
            gid = np.ascontiguousarray(total[ld.nFluid:ld.nFluid + nGFC])
            parent, child = np.zeros(nGFC, dtype=np.int32), np.zeros(nGFC, dtype=np.int32)
            coord = np.zeros((nGFC, 3))
            L.ora_vertical_dep_from_coarser(nGFC, _l(gid), _l(ctot), _i(cb), _i(parent), _i(child), _d(coord))
            cngh = np.ascontiguousarray(ngh[l - 1] if ngh is not None else C.nghElems, dtype=np.int32)
            order, nSrc, pim = (np.zeros(nGFC, dtype=np.int32) for _ in range(3))
            src, dirs = np.zeros((nGFC, 27), dtype=np.int32), np.zeros((nGFC, 27), dtype=np.int32)
            w = np.zeros((nGFC, 27))
            rc = L.ora_update_dep_from_coarser(QQ, order_max, nGFC, _i(parent), _i(child), _d(coord), _i(cngh),
                                               _i(order), _i(nSrc), _i(src), _i(dirs), _i(pim), _d(w),
                                               stores[1].h, stores[2].h)
            if rc != 0:
                raise RuntimeError("a ghostFromCoarser has no source (the reference aborts)")
            dep["fromCoarser"] = dict(parentPos=parent, childNum=child, coord=coord, order=order, nSrc=nSrc,
                                      src=src, dir=dirs, posInMat=pim, weights=w)
        if nGFF:
            F = levels[l + 1]
            ftot = np.ascontiguousarray(F.total, dtype=np.int64)
            fb = _block_start(F)
            o0 = ld.nFluid + nGFC
            gid = np.ascontiguousarray(total[o0:o0 + nGFF])
            so, sp = np.zeros(nGFF + 1, dtype=np.int32), np.zeros(8 * nGFF, dtype=np.int32)
            L.ora_vertical_dep_from_finer(nGFF, _l(gid), _l(ftot), _i(fb), _i(so), _i(sp))
            dep["fromFiner"] = (so, sp[:so[-1]].copy())
        out[l] = dep
    return out, stores


class PassiveScalarScheme:
    """scheme kind 'passive_scalar' on one level: mus_calcAuxField_zerothMoment +
    mus_advRel_kPS_* (mus_compute_passiveScalar_module.fpp), transport velocity in lattice units."""
    VARIANT = {("bgk", "first"): 1, ("bgk", "second"): 2, ("trt", "standard"): 3}

    def __init__(self, ld, relaxation="bgk", variant="first", diff_coeff=0.01, lambda_=0.25):
        self.ld, self.QQ = ld, ld.QQ
        self.variant = self.VARIANT[(relaxation, variant)]
        self.diff_coeff, self.lambda_ = float(diff_coeff), float(lambda_)
        n = ld.nSize * self.QQ
        self.state = [np.full(n, -1.0e6), np.full(n, -1.0e6)]
        self.aux = np.full(ld.nSize, -1.0e6)
        self.nNow, self.nNext = 0, 1
        self.transVel = np.zeros((ld.nSolve, 3))

    def init_equilibrium(self, rho, vel=None):
        """f = w * rho * (1 + 3 c.u) (first-order equilibrium) for every element of the level"""
        ld, QQ = self.ld, self.QQ
        rho = np.broadcast_to(np.asarray(rho, dtype=np.float64), (ld.nElems,))
        vel = np.zeros((ld.nElems, 3)) if vel is None else np.broadcast_to(np.asarray(vel, float), (ld.nElems, 3))
        c, w = cx_dir(QQ).astype(np.float64), weights(QQ)
        st = self.state[self.nNext]
        st[:] = 0.0
        f = w[None, :] * rho[:, None] * (1.0 + 3.0 * (vel @ c.T))
        st[:ld.nElems * QQ] = f.reshape(-1)
        self.state[self.nNow][:] = st

    def set_transport_velocity(self, vel):
        self.transVel = np.ascontiguousarray(np.broadcast_to(np.asarray(vel, dtype=np.float64),
                                                             (self.ld.nSolve, 3)))

    def step(self):
        L, ld = lib(), self.ld
        self.nNow, self.nNext = self.nNext, self.nNow
        L.ora_calc_aux_zeroth(self.QQ, _d(self.aux), _d(self.state[self.nNow]), _i(ld.neigh), ld.nSize,
                              ld.nSolve)
        rc = L.ora_compute_passive_scalar(self.variant, self.QQ, _d(self.state[self.nNow]),
                                          _d(self.state[self.nNext]), _i(ld.neigh), ld.nSize, ld.nSolve,
                                          _d(self.transVel), self.diff_coeff, self.lambda_)
        if rc != 0:
            raise RuntimeError("no oracle passive-scalar kernel for this configuration")

    def run(self, nsteps):
        for _ in range(nsteps):
            self.step()

    def total_mass(self):
        return lib().ora_total_mass(_d(self.state[self.nNext]), self.QQ, self.ld.nFluid)


def exchange_all(schemes):
    """comm_isend_irecv_real for all ranks held in one process: gather every send
    buffer from state(:,next), then scatter into the receivers."""
    L = lib()
    mail = {}
    for r, s in enumerate(schemes):
        st = s.state[s.nNext]
        for snd in s.ld.send:
            buf = np.empty(snd["pos"].size)
            L.ora_comm_gather(_d(buf), _d(st), _i(snd["pos"]), int(buf.size))
            mail[(r, snd["proc"])] = buf
    for r, s in enumerate(schemes):
        st = s.state[s.nNext]
        for rcv in s.ld.recv:
            buf = mail[(rcv["proc"], r)]
            assert buf.size == rcv["pos"].size
            L.ora_comm_scatter(_d(st), _d(buf), _i(rcv["pos"]), int(buf.size))


def run_multi(schemes, nsteps):
    for _ in range(nsteps):
        for s in schemes:
            s.step()
        exchange_all(schemes)


# --------------------------------------------------------------------------
# multi-level: do_recursive_multiLevel (mus_control_module.f90:242-497)
# --------------------------------------------------------------------------
class MultiLevelScheme:
    """one Scheme per level + the recursive level-sync schedule with ghost interpolation.
    `levels`: {level: descriptor} with the tem_levelDesc_type members (any generator);
    `tables`: {(level, 'fromFiner'| ('fromCoarser', order)): dict of flat dependency arrays}
    acoustic scaling: nu_lat doubles per finer level (mus_physics_module.f90:528)."""

    def __init__(self, levels, tables, relaxation="bgk", kind="fluid", omega_min=1.7, lambda_=0.25,
                 omega_bulk=None, order=1):
        self.levels = dict(levels)
        self.minLevel, self.maxLevel = min(levels), max(levels)
        self.tables = tables
        self.order = order
        self.s = {}
        nu0 = (1.0 / omega_min - 0.5) / 3.0
        for l, ld in self.levels.items():
            nu = nu0 * 2.0 ** (l - self.minLevel)
            om = 1.0 / (3.0 * nu + 0.5)
            self.s[l] = Scheme(ld, relaxation, kind, omega=om, lambda_=lambda_,
                               omega_bulk=om if omega_bulk is None else omega_bulk)
            self.s[l].visc[:] = nu
            self.s[l].omega[:] = om

    def _from_finer(self, l):
        """do_intpFinerAndExchange: my ghostFromFiner <- level l+1"""
        t = self.tables.get((l, "fromFiner"))
        if t is None or len(t["targets"]) == 0:
            return
        c, f = self.s[l], self.s[l + 1]
        lib().ora_fill_my_ghosts_from_finer_avg(
            c.QQ, c.incomp, _d(f.state[f.nNext]), _d(f.aux), _d(c.state[c.nNext]), len(t["targets"]),
            _i(t["targets"]), _i(t["srcOffset"]), _i(t["srcPos"]), _d(c.visc))

    def _aux_from_finer(self, l):
        """mus_intpAuxFieldCoarserAndExchange: auxField of my ghostFromFiner <- level l+1"""
        t = self.tables.get((l, "fromFiner"))
        if t is None or len(t["targets"]) == 0:
            return
        c, f = self.s[l], self.s[l + 1]
        lib().ora_fill_arbi_from_finer_avg(4, _d(f.aux), _d(c.aux), len(t["targets"]), _i(t["targets"]),
                                           _i(t["srcOffset"]), _i(t["srcPos"]))

    def _from_coarser(self, l):
        """do_intpCoarserAndExchange: ghostFromCoarser of level l+1 <- me, orders 0..order"""
        c, f = self.s[l], self.s[l + 1]
        for o in range(0, self.order + 1):
            t = self.tables.get((l + 1, ("fromCoarser", o)))
            if t is None or len(t["targets"]) == 0:
                continue
            coord = np.ascontiguousarray(t["coord"], dtype=np.float64)
            lib().ora_fill_finer_ghosts_from_me(
                o, c.QQ, c.incomp, _d(c.state[c.nNext]), _d(c.aux), _d(f.state[f.nNext]),
                len(t["targets"]), _i(t["targets"]), _i(t["srcOffset"]), _i(t["srcPos"]),
                _d(np.ascontiguousarray(t["weights"])), _i(t["posInMat"]), _i(t["matOffset"]),
                _d(np.ascontiguousarray(t["matrices"])), _d(coord), _d(f.visc))

    def do_computation(self, l=None):
        l = self.minLevel if l is None else l
        if l < self.maxLevel:
            for _ in range(2):                                   # nNesting = 2 (acoustic)
                self.do_computation(l + 1)
        self.advance(l)
        self.interpolate(l)

    def interpolate(self, l):
        if l < self.maxLevel:
            self._from_finer(l)
            self._from_coarser(l)

    def advance(self, l):
        """the level step without the ghost interpolation that closes it"""
        L = lib()
        s = self.s[l]
        s.set_boundary()
        s.nNow, s.nNext = s.nNext, s.nNow
        s.calc_aux(s.state[s.nNow])
        s.add_src_to_aux()
        if l < self.maxLevel:
            self._aux_from_finer(l)
        L.ora_update_omega(_d(s.omega), _d(s.visc), s.ld.nSolve)
        rc = L.ora_compute(s.relax, s.QQ, s.incomp, _d(s.state[s.nNow]), _d(s.state[s.nNext]), _d(s.aux),
                           _i(s.ld.neigh), _d(s.omega), s.ld.nSize, s.ld.nSolve, ctypes.byref(s.rp))
        if rc != 0:
            raise RuntimeError("no oracle kernel for this (relaxation, layout, kind)")
        s.apply_source_terms()

    def fill_helper_elements(self):
        """mus_init_flow once state(:, nNext) of the fluid elements is filled (initial condition or
        mus_readRestart; mus_flow_module.fpp:206-240): mus_initAuxField (:1677-1737) -- auxField of
        the FLUID elements from their own PDFs, auxField of the ghostFromFiner elements by
        averaging, finest level first --, fillHelperElementsFineToCoarse (:1517-1588) and
        fillHelperElementsCoarseToFine (:1601-1673).  (The auxField interpolated for the
        ghostFromCoarser elements, mus_intpAuxFieldFinerAndExchange, is recomputed by the first
        level step before anything reads it and is not restated.)"""
        L = lib()
        for s in self.s.values():
            ident = np.zeros(s.QQ * s.ld.nSize, dtype=np.int32)
            e = np.arange(s.ld.nSize, dtype=np.int64)
            for d in range(s.QQ):
                ident[d * s.ld.nSize:(d + 1) * s.ld.nSize] = e * s.QQ + d + 1
            (L.ora_calc_aux_incomp if s.incomp else L.ora_calc_aux)(
                s.QQ, _d(s.aux), _d(s.state[s.nNext]), _i(ident), s.ld.nSize, s.ld.nFluid)
        for l in range(self.maxLevel - 1, self.minLevel - 1, -1):
            self._aux_from_finer(l)
            self._from_finer(l)
        for l in range(self.minLevel, self.maxLevel):
            self._from_coarser(l)

    def run(self, ncycles):
        for _ in range(ncycles):
            self.do_computation()

    def total_mass(self):
        """sum over levels of fluid PDFs weighted by the cell volume relative to minLevel"""
        tot = 0.0
        for l, s in self.s.items():
            tot += s.total_mass() / 8.0 ** (l - self.minLevel)
        return tot


class CoupledMultiLevel:
    """BASELINE config 5: a flow and a passive scalar transported by it on the same multi-level
    mesh.  The reference has no such run (it aborts for a passive scalar on a multi-level mesh,
    mus_scheme_module.f90:166-190, and holds one scheme per process): this is the documented
    extension, built from reference pieces only -- the recursive schedule of
    do_recursive_multiLevel with both schemes advancing inside every level step (flow first: the
    scalar's transport velocity is the auxField velocity of the same level step), the scalar's
    kernels (mus_compute_passiveScalar_module.fpp) and, for the scalar's ghosts, the reference's
    interpolation of ARBITRARY values applied to the PDFs (fillArbiMyGhostsFromFiner_avg,
    fillArbiFinerGhostsFromMe_weighAvg / _linear / _quad).  Diffusivity scales acoustically like
    the viscosity: the lattice value doubles per finer level."""

    def __init__(self, flow, relaxation="bgk", variant="first", diff_coeff_min=0.01, lambda_=0.25):
        self.flow = flow
        self.minLevel, self.maxLevel = flow.minLevel, flow.maxLevel
        self.diff = {l: diff_coeff_min * 2.0 ** (l - self.minLevel) for l in flow.levels}
        self.ps = {l: PassiveScalarScheme(ld, relaxation, variant, diff_coeff=self.diff[l], lambda_=lambda_)
                   for l, ld in flow.levels.items()}

    def _ps_from_finer(self, l):
        t = self.flow.tables.get((l, "fromFiner"))
        if t is None or len(t["targets"]) == 0:
            return
        c, f = self.ps[l], self.ps[l + 1]
        lib().ora_fill_arbi_from_finer_avg(c.QQ, _d(f.state[f.nNext]), _d(c.state[c.nNext]), len(t["targets"]),
                                           _i(t["targets"]), _i(t["srcOffset"]), _i(t["srcPos"]))

    def _ps_from_coarser(self, l):
        c, f = self.ps[l], self.ps[l + 1]
        for o in range(0, self.flow.order + 1):
            t = self.flow.tables.get((l + 1, ("fromCoarser", o)))
            if t is None or len(t["targets"]) == 0:
                continue
            coord = np.ascontiguousarray(t["coord"], dtype=np.float64)
            lib().ora_fill_arbi_finer_from_me(
                o, c.QQ, _d(c.state[c.nNext]), _d(f.state[f.nNext]), len(t["targets"]), _i(t["targets"]),
                _i(t["srcOffset"]), _i(t["srcPos"]), _d(np.ascontiguousarray(t["weights"])), _i(t["posInMat"]),
                _i(t["matOffset"]), _d(np.ascontiguousarray(t["matrices"])), _d(coord))

    def do_computation(self, l=None):
        l = self.minLevel if l is None else l
        if l < self.maxLevel:
            for _ in range(2):
                self.do_computation(l + 1)
        self.flow.advance(l)
        p, ld = self.ps[l], self.flow.levels[l]
        p.set_transport_velocity(self.flow.s[l].aux[:ld.nSolve * 4].reshape(-1, 4)[:, 1:])
        p.step()
        self.flow.interpolate(l)
        if l < self.maxLevel:
            self._ps_from_finer(l)
            self._ps_from_coarser(l)

    def run(self, ncycles):
        for _ in range(ncycles):
            self.do_computation()

    def scalar_mass(self):
        return sum(p.total_mass() / 8.0 ** (l - self.minLevel) for l, p in self.ps.items())


# --------------------------------------------------------------------------
# multi-level on several ranks held in one process (lockstep recursion)
# --------------------------------------------------------------------------
class MultiRankMultiLevel:
    """one MultiLevelScheme per rank on the descriptors of a partitioned multi-level mesh; after
    every level step the ranks exchange the halo elements of that level: all QQ links of
    state(:, next) (comm_isend_irecv_real through sendBuffer / recvBuffer) and their auxField
    entries (auxField%sendBuffer, mus_auxField_module.f90:377-396), then interpolate their own
    ghosts.  Test infrastructure: the truth is the single-domain MultiLevelScheme."""

    def __init__(self, rank_levels, rank_tables, ghost_comm=None, **kw):
        """ghost_comm: None, or comm[rank][level][kind]['send' | 'recv'] lists (kind 'fromCoarser' /
        'fromFiner') of ghosts that one rank interpolates for others -- the reference's
        sendBufferFromCoarser / FromFiner (exchanged after the interpolation that fills them)"""
        self.r = [MultiLevelScheme(lv, tb, **kw) for lv, tb in zip(rank_levels, rank_tables)]
        self.minLevel, self.maxLevel = self.r[0].minLevel, self.r[0].maxLevel
        self.ghost_comm = ghost_comm

    def _exchange_ghosts(self, l, kind, with_aux):
        """comm_isend_irecv_real on the level's FromCoarser / FromFiner buffers (state(:, next),
        all QQ links per element) and, for ghostFromFiner elements, their auxField entries
        (mus_intpAuxFieldCoarserAndExchange, mus_auxField_module.f90:404-444)"""
        if self.ghost_comm is None:
            return
        L = lib()
        mail_s, mail_a = {}, {}
        for r, m in enumerate(self.r):
            s = m.s[l]
            st = s.state[s.nNext]
            for snd in self.ghost_comm[r][l][kind]["send"]:
                buf = np.empty(snd["pos"].size)
                L.ora_comm_gather(_d(buf), _d(st), _i(snd["pos"]), int(buf.size))
                mail_s[(r, snd["proc"])] = buf
                mail_a[(r, snd["proc"])] = s.aux.reshape(-1, 4)[snd["elemPos"].astype(np.int64) - 1].copy()
        for r, m in enumerate(self.r):
            s = m.s[l]
            st = s.state[s.nNext]
            for rcv in self.ghost_comm[r][l][kind]["recv"]:
                buf = mail_s[(rcv["proc"], r)]
                assert buf.size == rcv["pos"].size
                L.ora_comm_scatter(_d(st), _d(buf), _i(rcv["pos"]), int(buf.size))
                if with_aux:
                    s.aux.reshape(-1, 4)[rcv["elemPos"].astype(np.int64) - 1] = mail_a[(rcv["proc"], r)]

    def _exchange(self, l):
        L = lib()
        mail_s, mail_a = {}, {}
        for r, m in enumerate(self.r):
            s = m.s[l]
            st = s.state[s.nNext]
            for snd in s.ld.send:
                buf = np.empty(snd["pos"].size)
                L.ora_comm_gather(_d(buf), _d(st), _i(snd["pos"]), int(buf.size))
                mail_s[(r, snd["proc"])] = buf
                e = snd["elemPos"].astype(np.int64)
                mail_a[(r, snd["proc"])] = s.aux.reshape(-1, 4)[e - 1].copy()
        for r, m in enumerate(self.r):
            s = m.s[l]
            st = s.state[s.nNext]
            for rcv in s.ld.recv:
                buf = mail_s[(rcv["proc"], r)]
                assert buf.size == rcv["pos"].size
                L.ora_comm_scatter(_d(st), _d(buf), _i(rcv["pos"]), int(buf.size))
                e = rcv["elemPos"].astype(np.int64)
                s.aux.reshape(-1, 4)[e - 1] = mail_a[(rcv["proc"], r)]

    def _level_step(self, m, l):
        L = lib()
        s = m.s[l]
        s.set_boundary()
        s.nNow, s.nNext = s.nNext, s.nNow
        s.calc_aux(s.state[s.nNow])
        s.add_src_to_aux()
        if l < m.maxLevel:
            m._aux_from_finer(l)
        L.ora_update_omega(_d(s.omega), _d(s.visc), s.ld.nSolve)
        rc = L.ora_compute(s.relax, s.QQ, s.incomp, _d(s.state[s.nNow]), _d(s.state[s.nNext]), _d(s.aux),
                           _i(s.ld.neigh), _d(s.omega), s.ld.nSize, s.ld.nSolve, ctypes.byref(s.rp))
        if rc != 0:
            raise RuntimeError("no oracle kernel for this (relaxation, layout, kind)")
        s.apply_source_terms()

    def do_computation(self, l=None):
        l = self.minLevel if l is None else l
        if l < self.maxLevel:
            for _ in range(2):
                self.do_computation(l + 1)
        for m in self.r:
            self._level_step(m, l)
        self._exchange(l)
        if l > self.minLevel:                  # recvBufferFromCoarser after the level's own step
            self._exchange_ghosts(l, "fromCoarser", with_aux=False)    # (mus_control_module.f90:434-465)
        if l < self.maxLevel:
            for m in self.r:
                m._from_finer(l)
            self._exchange_ghosts(l, "fromFiner", with_aux=True)       # do_intpFinerAndExchange
            for m in self.r:
                m._from_coarser(l)
            self._exchange_ghosts(l + 1, "fromCoarser", with_aux=False)  # do_intpCoarserAndExchange

    def run(self, ncycles):
        for _ in range(ncycles):
            self.do_computation()
