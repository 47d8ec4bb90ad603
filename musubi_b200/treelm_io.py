"""treelm mesh files: read and write the on-disk mesh format of the
reference and build a single-level descriptor from it.

Format (tem/doc_pages/fileformat.md; writers tem_global_module.f90:844-943 dump_tem_global,
treelmesh_module.f90 dump_treelmesh, tem_bc_prop_module.f90:764-879):

  <dir>/header.lua     Lua script: version, label, comment, boundingbox{origin{x,y,z}, length},
                       nElems, minLevel, maxLevel, nProperties, effBoundingbox, property{...}
  <dir>/elemlist.lsb   per element two little-endian int64: treeID, property bit mask
  <dir>/bnd.lua        nSides, nBCtypes, bclabel{...}
  <dir>/bnd.lsb        int64 boundary_ID(nSides, nBndElems): one row of nSides = 26 ids (treelm
                       direction order, tem_param_module.f90:133-146) per element with prp_hasBnd

The header is a Lua script; the keys written by dump_tem_global are plain assignments, which
is what the reader below accepts (no Lua interpreter is embedded in the product)."""
import os
import re

import numpy as np

from .treelm_multilevel import (PRP_FLUID, PRP_HASBND, construct_connectivity, coords, first_id,
                                morton, stencil_tables)

# treelm's 26 directions (qOffset, tem_param_module.f90:133-137): W,S,B,E,N,T, BS,TS,BN,TN,
# BW,BE,TW,TE, SW,NW,SE,NE, BSW,BSE,BNW,BNE,TSW,TSE,TNW,TNE
Q_OFFSET = np.array([
    [-1, 0, 0], [0, -1, 0], [0, 0, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1],
    [0, -1, -1], [0, -1, 1], [0, 1, -1], [0, 1, 1], [-1, 0, -1], [1, 0, -1], [-1, 0, 1], [1, 0, 1],
    [-1, -1, 0], [-1, 1, 0], [1, -1, 0], [1, 1, 0],
    [-1, -1, -1], [1, -1, -1], [-1, 1, -1], [1, 1, -1], [-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]],
    dtype=np.int64)
N_SIDES = 26


def stencil_to_treelm(QQ):
    """tem_stencil_map_toTreelmDef: stencil direction (0-based, without the rest) -> treelm side"""
    cx, _ = stencil_tables(QQ)
    out = np.zeros(QQ - 1, dtype=np.int64)
    for q in range(QQ - 1):
        out[q] = int(np.nonzero((Q_OFFSET == cx[q]).all(axis=1))[0][0])
    return out


def level_of(treeID):
    level, first, count = 0, 0, 1
    while treeID >= first + count:
        first, count, level = first + count, count * 8, level + 1
    return level


def dump_treelmesh(dirname, treeID, prop, origin=(0.0, 0.0, 0.0), length=1.0, bc_labels=(),
                   boundary_ID=None, label="musb200", comment=""):
    """write header.lua / elemlist.lsb (/ bnd.lua / bnd.lsb).  boundary_ID: [nBndElems][26] int64
    for the elements carrying prp_hasBnd, in element order."""
    os.makedirs(dirname, exist_ok=True)
    treeID = np.ascontiguousarray(treeID, dtype="<i8")
    prop = np.ascontiguousarray(prop, dtype="<i8")
    if treeID.shape != prop.shape or treeID.ndim != 1:
        raise ValueError("treeID and property must be 1-d arrays of equal length")
    levels = [level_of(int(treeID.min())), level_of(int(treeID.max()))] if treeID.size else [0, 0]
    np.stack([treeID, prop], axis=1).tofile(os.path.join(dirname, "elemlist.lsb"))
    has_bnd = boundary_ID is not None and len(boundary_ID) > 0
    o = [float(v) for v in origin]
    with open(os.path.join(dirname, "header.lua"), "w") as fh:
        fh.write("version = 1\nlabel = '%s'\ncomment = '%s'\n" % (label, comment))
        fh.write("boundingbox = {\n    origin = { %.17g, %.17g, %.17g },\n    length = %.17g\n}\n"
                 % (o[0], o[1], o[2], float(length)))
        fh.write("nElems = %d\nminLevel = %d\nmaxLevel = %d\nnProperties = %d\n"
                 % (treeID.size, min(levels), max(levels), 1 if has_bnd else 0))
        fh.write("effBoundingbox = {\n    origin = { %.17g, %.17g, %.17g },\n"
                 "    effLength = { %.17g, %.17g, %.17g }\n}\n" % (o[0], o[1], o[2], length, length, length))
        if has_bnd:
            fh.write("property = {\n    {\n        label = 'has boundaries',\n        bitpos = %d,\n"
                     "        nElems = %d\n    }\n}\n" % (PRP_HASBND, len(boundary_ID)))
    if has_bnd:
        b = np.ascontiguousarray(boundary_ID, dtype="<i8")
        if b.shape != (int(((prop >> PRP_HASBND) & 1).sum()), N_SIDES):
            raise ValueError("boundary_ID needs one row of 26 ids per element with prp_hasBnd")
        b.tofile(os.path.join(dirname, "bnd.lsb"))
        with open(os.path.join(dirname, "bnd.lua"), "w") as fh:
            fh.write("nSides = %d\nnBCtypes = %d\nbclabel = { %s }\n"
                     % (N_SIDES, len(bc_labels), ", ".join("'%s'" % s for s in bc_labels)))


def _lua_number(text, key, cast=float, default=None):
    m = re.search(r"(?<![\w.])" + re.escape(key) + r"\s*=\s*([-+0-9.eE]+)", text)
    if not m:
        if default is not None:
            return default
        raise ValueError("treelm header: key %r not found" % key)
    return cast(float(m.group(1)))


def load_treelmesh(dirname):
    """read a treelm mesh directory -> dict(treeID, property, origin, length, nElems, minLevel,
    maxLevel, bc_labels, boundary_ID)"""
    hdr = open(os.path.join(dirname, "header.lua")).read()
    n = _lua_number(hdr, "nElems", int)
    m = re.search(r"boundingbox\s*=\s*\{.*?origin\s*=\s*\{([^}]*)\}.*?length\s*=\s*([-+0-9.eE]+)", hdr, re.S)
    if not m:
        raise ValueError("treelm header: boundingbox missing")
    origin = tuple(float(v) for v in m.group(1).replace(",", " ").split())
    raw = np.fromfile(os.path.join(dirname, "elemlist.lsb"), dtype="<i8")
    if raw.size != 2 * n:
        raise ValueError("elemlist.lsb holds %d int64, header says nElems = %d" % (raw.size, n))
    raw = raw.reshape(n, 2)
    out = dict(treeID=raw[:, 0].astype(np.int64), property=raw[:, 1].astype(np.int64), origin=origin,
               length=float(m.group(2)), nElems=n, minLevel=_lua_number(hdr, "minLevel", int),
               maxLevel=_lua_number(hdr, "maxLevel", int), bc_labels=[], boundary_ID=None)
    bl = os.path.join(dirname, "bnd.lua")
    if os.path.exists(bl):
        b = open(bl).read()
        ns = _lua_number(b, "nSides", int)
        mm = re.search(r"bclabel\s*=\s*\{([^}]*)\}", b, re.S)
        out["bc_labels"] = re.findall(r"['\"]([^'\"]*)['\"]", mm.group(1)) if mm else []
        nb = int(((out["property"] >> PRP_HASBND) & 1).sum())
        bid = np.fromfile(os.path.join(dirname, "bnd.lsb"), dtype="<i8")
        if bid.size != nb * ns:
            raise ValueError("bnd.lsb: %d ids for %d boundary elements x %d sides" % (bid.size, nb, ns))
        out["boundary_ID"] = bid.reshape(nb, ns).astype(np.int64)
    return out


def dump_weights(filename, weights, elem_offset=0, nElems_global=None):
    """tem_dump_weights (treelmesh_module.f90:2169-2233): one native double per element of the
    mesh in tree order (the file the restart header's `weights` key and mesh.weights name, read
    back by tem_load_weights and fed to tem_balance_sparta); this rank's share lands at byte
    elem_offset * 8 of a file sized for the whole mesh."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    n = w.size + elem_offset if nElems_global is None else int(nElems_global)
    if elem_offset < 0 or elem_offset + w.size > n:
        raise ValueError("weights: elements %d..%d outside the mesh of %d" % (elem_offset, elem_offset + w.size, n))
    # all ranks write into one file at once: never truncate what another rank has written
    fd = os.open(filename, os.O_RDWR | os.O_CREAT, 0o644)
    try:
        if os.fstat(fd).st_size != n * 8:
            os.ftruncate(fd, n * 8)           # sizes the file; existing bytes below n * 8 are kept
        buf, off = memoryview(w).cast("B"), int(elem_offset) * 8
        while len(buf):
            k = os.pwrite(fd, buf[:1 << 30], off)
            buf, off = buf[k:], off + k
    finally:
        os.close(fd)


def load_weights(filename, elem_offset=0, nElems=None):
    """this rank's weights from a file written by tem_dump_weights"""
    size = os.path.getsize(filename) // 8
    n = size - elem_offset if nElems is None else int(nElems)
    if elem_offset < 0 or n < 0 or elem_offset + n > size:
        raise ValueError("weights file %s holds %d elements, asked for %d..%d" % (filename, size, elem_offset, elem_offset + n))
    return np.fromfile(filename, dtype=np.float64, count=n, offset=elem_offset * 8)


class FileLevelDesc:
    """tem_levelDesc_type + pdf_data_type of a single-level treelm mesh read from disk, on one
    rank: total list [fluid | halo], property, nghElems (neighbour position, or -boundary id where
    the mesh file names a boundary), neigh (mus_construct_connectivity), halo send / recv lists.
    Neighbours across the universe cube wrap periodically (tem_IdOfCoord); boundaries of kind
    'wall' need no link lists (the bounce-back lives in neigh); for the other kinds of the hot path
    (velocity_bounceback, pressure_expol / pressure_antibounceback, bound through bc_kind =
    {label: kind}) the link and neighbour lists of mus_init_boundary are built: assignBCList
    (mus_construction_module.fpp:2203-2440), mus_set_bcLinks / mus_set_inletUbb / outletExpol
    (mus_bc_header_module.fpp:1702-1967, 2256-2319), setFieldBCNeigh (:1733-1900)."""

    def __init__(self, mesh, QQ, rank=0, nranks=1, bc_kind=None):
        if mesh["minLevel"] != mesh["maxLevel"]:
            raise ValueError("FileLevelDesc handles single-level meshes; use the multi-level generator")
        bc_kind = bc_kind or {}
        for lab in mesh["bc_labels"]:
            if bc_kind.get(lab, "wall") not in ("wall", "velocity_bounceback", "pressure", "pressure_expol",
                                               "pressure_antibounceback"):
                raise ValueError("boundary %r: kind %r is outside the hot path" % (lab, bc_kind[lab]))
        self.level, self.QQ, self.rank, self.nranks = mesh["minLevel"], QQ, rank, nranks
        L, QQN = self.level, QQ - 1
        n1 = 1 << L
        tid = mesh["treeID"]
        if np.any(np.diff(tid) <= 0):
            raise ValueError("treeID list must be strictly increasing")
        N = tid.size
        # treelmesh_module.f90:1276-1296: equal contiguous ranges of the space-filling curve
        base, rem = divmod(N, nranks)
        cnt = np.array([base + (1 if r < rem else 0) for r in range(nranks)], dtype=np.int64)
        off = np.concatenate([[0], np.cumsum(cnt)])
        lo, hi = int(off[rank]), int(off[rank + 1])
        self.nFluid = hi - lo
        cx, inv = stencil_tables(QQ)
        side_of = stencil_to_treelm(QQ)
        my = tid[lo:hi]
        hasb = ((mesh["property"] >> PRP_HASBND) & 1).astype(bool)
        brow = np.full(N, -1, dtype=np.int64)
        brow[hasb] = np.arange(int(hasb.sum()))

        def neighbours(gsel):
            """global neighbour index into tid (>= 0), -(bcid) - 1 for a boundary, per direction"""
            x, y, z = coords(tid[gsel] - first_id(L))
            out = np.zeros((gsel.size, QQN), dtype=np.int64)
            for q in range(QQN):
                nid = first_id(L) + morton((x + cx[q, 0]) % n1, (y + cx[q, 1]) % n1, (z + cx[q, 2]) % n1)
                j = np.searchsorted(tid, nid)
                found = (j < N) & (tid[np.minimum(j, N - 1)] == nid)
                g = np.where(found, j, -1)
                if mesh["boundary_ID"] is not None:
                    rows = brow[gsel]
                    bid = np.where(rows >= 0, mesh["boundary_ID"][np.maximum(rows, 0), side_of[q]], 0)
                    g = np.where(bid > 0, -bid - 1, g)
                if np.any(g == -1):
                    raise ValueError("mesh is not closed: a neighbour is neither an element nor a "
                                     "boundary (a d3q19 mesh file read with d3q27?)")
                out[:, q] = g
            return out

        gidx = neighbours(np.arange(lo, hi))
        # halos: remote neighbours, appended in ascending treeID order
        remote = (gidx >= 0) & ((gidx < lo) | (gidx >= hi))
        halo_g = np.unique(gidx[remote])
        self.nHalo = int(halo_g.size)
        self.nElems = self.nFluid + self.nHalo
        self.nSize = (self.nElems + 3) // 4 * 4
        self.nGhostFromCoarser = self.nGhostFromFiner = 0
        self.nSolve = self.nFluid
        self.total = np.concatenate([my, tid[halo_g]]).astype(np.int64)
        self.property = np.zeros(self.nElems, dtype=np.int64)
        self.property[:self.nFluid] = mesh["property"][lo:hi]

        def positions(g):
            """global index -> 1-based position in the total list (0 = not on this rank),
            boundaries -> -(bcid)"""
            local = (g >= lo) & (g < hi)
            h = np.searchsorted(halo_g, np.maximum(g, 0))
            is_halo = (g >= 0) & ~local & (h < self.nHalo) & (halo_g[np.minimum(h, max(self.nHalo - 1, 0))] == g) \
                if self.nHalo else np.zeros_like(local)
            return np.where(local, g - lo + 1, np.where(is_halo, self.nFluid + h + 1,
                                                        np.where(g < 0, g + 1, 0)))

        self.nghElems = np.zeros((self.nElems, QQN), dtype=np.int32)
        self.nghElems[:self.nFluid] = positions(gidx)
        if self.nHalo:
            self.nghElems[self.nFluid:] = positions(neighbours(halo_g))
        self.neigh = construct_connectivity(QQ, self.nghElems, self.property, self.nFluid,
                                            self.nFluid, self.nSize)
        self.bc_labels = list(mesh["bc_labels"])
        self._tid, self._lo, self._hi, self._positions = tid, lo, hi, positions
        self._build_bc(bc_kind, cx, inv)
        # halo exchange lists (init_recvBuffers / init_sendBuffers, comm_reduced): per remote rank
        # the links a local element pulls from a halo, element-major, direction ascending
        owner = np.searchsorted(off, halo_g, side="right") - 1
        self.recv, self.send = [], []
        need = np.zeros((self.nHalo, QQ), dtype=bool)          # need[h, d]: some local pulls d from h
        for d in range(QQN):
            p = self.nghElems[:self.nFluid, inv[d] - 1]        # element at x - c_d
            h = p[p > self.nFluid] - self.nFluid - 1
            need[h, d] = True
        self._need, self._halo_g, self._owner, self._off = need, halo_g, owner, off
        for r in np.unique(owner):
            hs = np.nonzero(owner == r)[0]
            e, d = np.nonzero(need[hs])
            self.recv.append(dict(proc=int(r), pos=((self.nFluid + hs[e]) * QQ + d + 1).astype(np.int32),
                                  elemPos=(self.nFluid + hs + 1).astype(np.int32)))
        self._mesh, self._bc_kind = mesh, bc_kind

    def _build_bc(self, bc_kind, cx, inv):
        """boundary element buffer and per-boundary lists, all 1-based as the Fortran arrays"""
        QQ, QQN, nF = self.QQ, self.QQ - 1, self.nFluid
        ngh = self.nghElems[:nF]
        hasb = ((self.property[:nF] >> PRP_HASBND) & 1).astype(bool)
        self.bc_elemBuffer = (np.nonzero(hasb)[0] + 1).astype(np.int32)
        posInBuf = np.zeros(nF + 1, dtype=np.int64)
        posInBuf[self.bc_elemBuffer] = np.arange(1, self.bc_elemBuffer.size + 1)
        self.bc = []
        if not self.bc_labels:
            return
        length = (cx[:QQN] ** 2).sum(axis=1)
        wgt = np.where(length == 1, 4, np.where(length == 2, 2, 1))
        prevail = cx[:QQN] / np.sqrt(length)[:, None]
        x, y, z = coords(self.total[:nF] - first_id(self.level))
        n1 = 1 << self.level
        for bid, lab in enumerate(self.bc_labels, start=1):
            kind = bc_kind.get(lab, "wall")
            hit = ngh == -bid                                   # [nF][QQN]: stencil dir k sees boundary bid
            sel = np.nonzero(hit.any(axis=1))[0]
            bc = dict(id=bid, kind=kind, label=lab, elems=(sel + 1).astype(np.int32))
            links, iDir, pib, outPos, iEl, sPos, nInd, pbe, nPos = [], [], [], [], [], [], [], [], []
            for iElem, e in enumerate(sel, start=1):
                ks = np.nonzero(hit[e])[0]
                ds = np.sort(inv[ks])                              # bitmask(cxDirInv(k)) = true, 1-based dirs
                nrm = -(wgt[ks, None] * cx[ks]).sum(axis=0).astype(np.float64)
                for d in ds:
                    links.append(self.neigh[(d - 1) * self.nSize + e])
                    iDir.append(d)
                    pib.append(posInBuf[e + 1])
                    outPos.append(inv[d - 1] + (posInBuf[e + 1] - 1) * QQ)
                    iEl.append(iElem)
                    sPos.append(d + (iElem - 1) * QQ)
                # tem_determine_discreteVector: first strict maximum of the projection, exit at 1
                nl = np.sqrt((nrm * nrm).sum())
                best, mx = 0, -2.0
                for k in range(QQN):
                    dp = min(max(float((nrm / nl) @ prevail[k]), -1.0), 1.0)
                    if dp > mx:
                        mx, best = dp, k
                        if abs(mx - 1.0) <= np.finfo(float).eps:
                            break
                nInd.append(best + 1)
                pbe.append(posInBuf[e + 1])
                # setFieldBCNeigh: the elements at x + k * normal, k = 1, 2 (last valid one repeated)
                np2 = [e + 1, e + 1]
                for k in (1, 2):
                    xn, yn, zn = x[e] + k * cx[best, 0], y[e] + k * cx[best, 1], z[e] + k * cx[best, 2]
                    p = 0
                    if 0 <= xn < n1 and 0 <= yn < n1 and 0 <= zn < n1:
                        nid = first_id(self.level) + morton(np.array([xn]), np.array([yn]), np.array([zn]))[0]
                        j = int(np.searchsorted(self._tid, nid))
                        if j < self._tid.size and self._tid[j] == nid:
                            p = int(self._positions(np.array([j]))[0])
                    if p <= 0:
                        if k == 1:
                            np2 = [e + 1, e + 1]
                        else:
                            np2[1] = np2[0]
                        break
                    np2[k - 1] = p
                    if k == 1:
                        np2[1] = p
                nPos.append(np2)
            i32 = lambda v: np.array(v, dtype=np.int32)  # noqa: E731
            bc.update(links=i32(links), iDir=i32(iDir), posInBuffer=i32(pib), outPos=i32(outPos),
                      iElemOfLink=i32(iEl), statePos=i32(sPos), normalInd=i32(nInd), posInBcElemBuf=i32(pbe),
                      neighPos=np.array(nPos, dtype=np.int32).reshape(-1, 2))
            self.bc.append(bc)

    def build_send(self, peers):
        """send lists = the receivers' recv lists seen from here.  peers: {rank: FileLevelDesc};
        what MPI does for the reference (tem_comm: the receiver communicates its requests)."""
        self.send = []
        lo = int(self._off[self.rank])
        for r, other in sorted(peers.items()):
            if r == self.rank:
                continue
            hs = np.nonzero(other._owner == self.rank)[0]
            if hs.size == 0:
                continue
            e, d = np.nonzero(other._need[hs])
            local = other._halo_g[hs[e]] - lo
            self.send.append(dict(proc=int(r), pos=(local * self.QQ + d + 1).astype(np.int32),
                                  elemPos=(other._halo_g[hs] - lo + 1).astype(np.int32)))

    def barycenters(self):
        x, y, z = coords(self.total - first_id(self.level))
        dx = self._mesh["length"] / (1 << self.level)
        o = self._mesh["origin"]
        return np.stack([o[0] + (x + 0.5) * dx, o[1] + (y + 0.5) * dx, o[2] + (z + 0.5) * dx], axis=1)


def mesh_from_level_desc(ld, length=1.0, origin=(0.0, 0.0, 0.0)):
    """the inverse direction, for round trips and for handing synthetic meshes to the reference:
    single-rank descriptor of the box generator -> (treeID, property, boundary_ID, labels)"""
    if ld.nranks != 1:
        raise ValueError("dump the mesh from a single-rank descriptor")
    side_of = stencil_to_treelm(ld.QQ)
    tid = np.asarray(ld.total[:ld.nFluid], dtype=np.int64)
    ngh = np.asarray(ld.nghElems[:ld.nFluid], dtype=np.int64)
    hasb = (ngh <= 0).any(axis=1)
    prop = np.full(ld.nFluid, 1 << PRP_FLUID, dtype=np.int64)
    prop[hasb] |= 1 << PRP_HASBND
    bid = np.zeros((int(hasb.sum()), N_SIDES), dtype=np.int64)
    rows = ngh[hasb]
    for q in range(ld.QQ - 1):
        bid[:, side_of[q]] = np.where(rows[:, q] <= 0, -rows[:, q], 0)
    labels = [b["label"] for b in sorted(ld.bc, key=lambda b: b["id"])]
    return dict(treeID=tid, property=prop, boundary_ID=bid if hasb.any() else None, bc_labels=labels,
                origin=origin, length=length)
