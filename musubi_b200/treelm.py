"""Host-side level descriptor of a single-level box mesh, produced by the C++
generator (csrc/host/treelm_box.cpp).  The attribute names follow the
reference's tem_levelDesc_type / pdf_data_type / boundary_type members so that
what is handed to libmusb200 reads like what the Fortran shim hands over."""
import numpy as np

from . import _lib  # noqa: F401  (musubi_b200._lib is part of the package surface: bench.py, tests)
from ._lib import P_I32, P_I64, P_DBL, mesh, ptr

KIND = {"periodic": 0, "cavity": 1, "channel": 2}
_BC_LABEL = {"cavity": ("wall", "lid"), "channel": ("wall", "inlet", "outlet")}


class LevelDesc:
    """tem_levelDesc_type + pdf_data_type of one level on one rank."""

    def __init__(self, level, QQ, kind="periodic", rank=0, nranks=1, comm_reduced=True, octants=8):
        """octants: the domain is the first 1, 2, 4 or 8 octants of the level-`level` universe cube
        in Morton order (extent doubles in x, then y, then z); walled kinds only when < 8."""
        h = mesh.musb200_mesh_box_create(level, QQ, KIND[kind], rank, nranks, 1 if comm_reduced else 0,
                                         octants)
        self.octants = octants
        if not h:
            raise ValueError("bad box-mesh parameters")
        try:
            info = np.zeros(16, dtype=np.int64)
            mesh.musb200_mesh_info(h, ptr(info, P_I64))
            (self.nFluid, self.nHalo, self.nElems, self.nSize, nBcE, nBc, nRp, nSp, nRv, nSv, nRe,
             nSe) = [int(v) for v in info[:12]]
            self.level, self.QQ, self.kind, self.rank, self.nranks = level, QQ, kind, rank, nranks
            self.nGhostFromCoarser = self.nGhostFromFiner = 0
            self.nSolve = self.nFluid
            self.total = np.zeros(self.nElems, dtype=np.int64)
            self.property = np.zeros(self.nElems, dtype=np.int64)
            self.nghElems = np.zeros((self.nElems, QQ - 1), dtype=np.int32)
            self.neigh = np.zeros(QQ * self.nSize, dtype=np.int32)
            mesh.musb200_mesh_total(h, ptr(self.total, P_I64))
            mesh.musb200_mesh_property(h, ptr(self.property, P_I64))
            mesh.musb200_mesh_nghelems(h, ptr(self.nghElems, P_I32))
            mesh.musb200_mesh_neigh(h, ptr(self.neigh, P_I32))
            self.bc_elemBuffer = np.zeros(nBcE, dtype=np.int32)
            mesh.musb200_mesh_bc_elembuffer(h, ptr(self.bc_elemBuffer, P_I32))
            self.recv, self.send = [], []
            for d, (np_, nv, ne, out) in enumerate(((nSp, nSv, nSe, self.send), (nRp, nRv, nRe, self.recv))):
                proc = np.zeros(np_, dtype=np.int32)
                nVals = np.zeros(np_, dtype=np.int32)
                pos = np.zeros(nv, dtype=np.int32)
                ecnt = np.zeros(np_, dtype=np.int32)
                epos = np.zeros(ne, dtype=np.int32)
                mesh.musb200_mesh_comm(h, d, ptr(proc, P_I32), ptr(nVals, P_I32), ptr(pos, P_I32),
                                       ptr(ecnt, P_I32), ptr(epos, P_I32))
                o = oe = 0
                for i in range(np_):
                    out.append(dict(proc=int(proc[i]), pos=pos[o:o + nVals[i]].copy(),
                                    elemPos=epos[oe:oe + ecnt[i]].copy()))
                    o += int(nVals[i])
                    oe += int(ecnt[i])
            self.bc = []
            for i in range(nBc):
                sz = np.zeros(4, dtype=np.int32)
                mesh.musb200_mesh_bc_info(h, i, ptr(sz, P_I32))
                bc = dict(id=int(sz[0]), kind=("wall", "velocity_bounceback", "pressure")[int(sz[1])],
                          label=_BC_LABEL[kind][int(sz[0]) - 1],
                          elems=np.zeros(sz[2], dtype=np.int32), links=np.zeros(sz[3], dtype=np.int32),
                          outPos=np.zeros(sz[3], dtype=np.int32),
                          posInBuffer=np.zeros(sz[3], dtype=np.int32),
                          iDir=np.zeros(sz[3], dtype=np.int32))
                mesh.musb200_mesh_bc_lists(h, i, ptr(bc["elems"], P_I32), ptr(bc["links"], P_I32),
                                           ptr(bc["outPos"], P_I32), ptr(bc["posInBuffer"], P_I32),
                                           ptr(bc["iDir"], P_I32))
                # boundary_type%elemLvl / %neigh(level)%posInState / outletExpol (pressure boundaries)
                bc.update(normalInd=np.zeros(sz[2], dtype=np.int32),
                          posInBcElemBuf=np.zeros(sz[2], dtype=np.int32),
                          neighPos=np.zeros((int(sz[2]), 2), dtype=np.int32),
                          iElemOfLink=np.zeros(sz[3], dtype=np.int32),
                          statePos=np.zeros(sz[3], dtype=np.int32))
                mesh.musb200_mesh_bc_elem_lists(h, i, ptr(bc["normalInd"], P_I32),
                                                ptr(bc["posInBcElemBuf"], P_I32), ptr(bc["neighPos"], P_I32),
                                                ptr(bc["iElemOfLink"], P_I32), ptr(bc["statePos"], P_I32))
                self.bc.append(bc)
            self._bary_args = None
            self._h_for_bary = None
        finally:
            self._handle = h

    def barycenters(self, origin=(0.0, 0.0, 0.0), length=1.0):
        out = np.zeros((self.nElems, 3))
        mesh.musb200_mesh_bary(self._handle, float(origin[0]), float(origin[1]), float(origin[2]),
                               float(length), ptr(out, P_DBL))
        return out

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h:
            mesh.musb200_mesh_destroy(h)
            self._handle = None


class DeviceCube:
    """treelm's predefined cube (all 8^level elements in Morton order) on ONE rank whose
    connectivity libmusb200 generates on the device (musb200_level_create_cube): periodic, or
    closed by walls.  No host index lists exist, so the 32-bit limit nSize*QQ < 2^31 of pdf%neigh
    does not apply (BASELINE config 3, 512^3 d3q27, on one GPU)."""
    device_generated = True

    def __init__(self, level, QQ, kind="periodic"):
        if kind not in ("periodic", "walls"):
            raise ValueError("DeviceCube: kind must be 'periodic' or 'walls'")
        n = 8 ** level
        self.level, self.QQ, self.kind, self.rank, self.nranks = level, QQ, kind, 0, 1
        self.nFluid = self.nElems = self.nSize = self.nSolve = n
        self.nGhostFromCoarser = self.nGhostFromFiner = self.nHalo = 0
        self.bc, self.send, self.recv = [], [], []
        self.bc_elemBuffer = np.zeros(0, dtype=np.int32)

    def barycenters(self, origin=(0.0, 0.0, 0.0), length=1.0, chunk=1 << 24):
        """tem_BaryOfId: origin + (coord + 0.5) * dx, dx = length / 2^level"""
        from .treelm_multilevel import coords
        out = np.empty((self.nElems, 3))
        dx = float(length) / (1 << self.level)
        for s in range(0, self.nElems, chunk):
            x, y, z = coords(np.arange(s, min(self.nElems, s + chunk), dtype=np.int64))
            for k, c in enumerate((x, y, z)):
                out[s:s + chunk, k] = origin[k] + (c + 0.5) * dx
        return out
