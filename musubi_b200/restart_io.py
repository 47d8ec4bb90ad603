"""Restart files of the reference: the on-disk form of the state that
mus_writeRestart / mus_readRestart exchange with tem_restart (mus_restart_module.f90:57-248,
tem_restart_module.f90) -- so that a site running the Fortran reference can continue from a dump
of the device state and this library can continue from a dump of the reference.

Files (tem_restart_openWrite / _writeHeader / _closeWrite, tem_restart_module.f90:1070-1321,
1424-1468; prefixes tem_init_restart_alloc :609-612):

  <prefix><simName>_<stamp>.lsb          raw native doubles: for every element of the mesh in global
                                         tree order nScalars*nDofs values (rank r writes at byte
                                         elemOffset * nScalars * nDofs * 8 -- the MPI file view of
                                         tem_init_restart_create_types, :636-757), per element in
                                         the order of mus_pdf_serialize (mus_buffer_module.fpp:80-137);
                                         '.msb' on a big-endian machine (env_module.f90:406-417)
  <prefix><simName>_header_<stamp>.lua   Lua script: binary_name{...}, solver_configFile, mesh,
                                         weights, time_point{sim, iter, clock}, nElems, nDofs,
                                         solver, varsys{systemname, variable{{name, ncomponents,
                                         state_varpos}}, nScalars, nStateVars, nAuxScalars, nAuxVars}
                                         (tem_time_out tem_time_module.f90:272-309, tem_mesh_out
                                         tem_global_module.f90:136-215, tem_varSys_out_single
                                         tem_varSys_module.fpp:1201-1293)
  <prefix><simName>_lastHeader.lua       the same header, rewritten after every dump

<stamp> = the simulation time in Fortran's EN12.3 edit descriptor, left-adjusted
(tem_timeformatter_module.f90:48, 214-231): 10.0006 -> '10.001E+00'.

The reader takes what tem_restart_readHeader (:1323-1420) takes from the header and accepts the
Lua subset aot_out writes (assignments of numbers, strings, booleans and nested table
constructors); no Lua interpreter is embedded in the product."""
import os
import sys
from decimal import ROUND_HALF_EVEN, Decimal

import numpy as np

ENDIAN_SUFFIX = ".lsb" if sys.byteorder == "little" else ".msb"


# ------------------------------------------------------------------------------------------
# Fortran ENw.d
def fortran_en(x, d=3, width=None):
    """the value in Fortran's ENw.d form: exponent a multiple of three, 1 <= |significand| < 1000,
    d decimals, two exponent digits ('E+00'); zero -> 0.000E+00.  Left-adjusted unless width."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        raise ValueError("no EN form for %r" % x)
    if x == 0.0:
        e3, m = 0, Decimal(0)
    else:
        v = Decimal(x)                                   # exact
        e3 = (v.adjusted() // 3) * 3                     # floor to a multiple of three
        q = Decimal(1).scaleb(-d)
        m = v.scaleb(-e3).quantize(q, rounding=ROUND_HALF_EVEN)
        if abs(m) >= 1000:                               # rounded up to the next power of 1000
            e3 += 3
            m = v.scaleb(-e3).quantize(q, rounding=ROUND_HALF_EVEN)
    s = "%sE%s%02d" % (format(m, "." + str(d) + "f"), "+" if e3 >= 0 else "-", abs(e3))
    return s if width is None else s.rjust(width)


def time_stamp(sim):
    """tem_timeformatter_sim_stamp with the default form '(EN12.3)'"""
    return fortran_en(sim, 3)


# ------------------------------------------------------------------------------------------
# the Lua subset of aot_out
class _Lua:
    def __init__(self, text):
        self.t, self.i, self.n = text, 0, len(text)

    def _skip(self):
        t = self.t
        while self.i < self.n:
            c = t[self.i]
            if c in " \t\r\n":
                self.i += 1
            elif t.startswith("--", self.i):
                if t.startswith("--[[", self.i):
                    j = t.find("]]", self.i)
                    self.i = self.n if j < 0 else j + 2
                else:
                    j = t.find("\n", self.i)
                    self.i = self.n if j < 0 else j + 1
            else:
                break

    def _peek(self):
        self._skip()
        return self.t[self.i] if self.i < self.n else ""

    def _name(self):
        self._skip()
        j = self.i
        while j < self.n and (self.t[j].isalnum() or self.t[j] == "_"):
            j += 1
        if j == self.i or self.t[self.i].isdigit():
            raise ValueError("restart header: name expected at offset %d" % self.i)
        s, self.i = self.t[self.i:j], j
        return s

    def _expect(self, ch):
        if self._peek() != ch:
            raise ValueError("restart header: %r expected at offset %d" % (ch, self.i))
        self.i += 1

    def value(self):
        c = self._peek()
        if c == "{":
            return self._table()
        if c in "'\"":
            j = self.i + 1
            out = []
            while j < self.n and self.t[j] != c:
                if self.t[j] == "\\" and j + 1 < self.n:
                    j += 1
                out.append(self.t[j])
                j += 1
            if j >= self.n:
                raise ValueError("restart header: unterminated string")
            self.i = j + 1
            return "".join(out)
        if c.isdigit() or c in "+-.":
            j = self.i + 1
            while j < self.n and (self.t[j].isalnum() or self.t[j] in "."
                                  or (self.t[j] in "+-" and self.t[j - 1] in "eE")):
                j += 1
            tok, self.i = self.t[self.i:j], j
            try:
                return int(tok)
            except ValueError:
                return float(tok)
        w = self._name()
        if w in ("true", "false"):
            return w == "true"
        if w == "nil":
            return None
        raise ValueError("restart header: unsupported expression %r (only literals are read)" % w)

    def _table(self):
        self._expect("{")
        keyed, items = {}, []
        while True:
            c = self._peek()
            if c == "}":
                self.i += 1
                break
            if c in ",;":
                self.i += 1
                continue
            save = self.i
            if c.isalpha() or c == "_":
                name = self._name()
                if self._peek() == "=":
                    self.i += 1
                    keyed[name] = self.value()
                    continue
                self.i = save
            items.append(self.value())
        if keyed and items:
            keyed.update({k + 1: v for k, v in enumerate(items)})
            return keyed
        return keyed if keyed else items

    def chunk(self):
        out = {}
        while self._peek():
            if self._peek() == ";":
                self.i += 1
                continue
            name = self._name()
            self._expect("=")
            out[name] = self.value()
        return out


def parse_lua_assignments(text):
    """{name: value} of a script made of `name = literal` statements"""
    return _Lua(text).chunk()


# ------------------------------------------------------------------------------------------
def _real(x):
    return fortran_en(x, 15, 24)       # how aot_out_val prints a double


def _lua_str(s):
    return "'%s'" % str(s).replace("\\", "\\\\").replace("'", "\\'")


def _int_list(vals, indent):
    rows = [", ".join(str(int(v)) for v in vals[i:i + 8]) for i in range(0, len(vals), 8)]
    return "{ " + (",\n" + " " * indent).join(rows) + " }"


def header_text(binary_name, time, nElems, varsys, mesh="./mesh/", weights="", nDofs=1,
                solver="Musubi_v2.0", solver_configFile="musubi.lua", solver_spec=""):
    """the header script as tem_restart_writeHeader composes it.  time = dict(sim[, iter, clock]);
    varsys = dict(systemname, variable=[dict(name, ncomponents[, state_varpos])], nAuxScalars,
    nAuxVars); mesh = the mesh directory or dict(predefined, origin, length, refinementLevel);
    solver_spec = text appended verbatim (mus_writeSolverSpecInfo's scratch file)."""
    L = [" binary_name = {", "    %s" % _lua_str(binary_name), "}",
         " solver_configFile = %s" % _lua_str(solver_configFile)]
    if isinstance(mesh, dict):
        L += [" mesh = {", "    predefined = %s," % _lua_str(mesh["predefined"]),
              "    origin = { %s }," % ", ".join(_real(v).strip() for v in mesh["origin"]),
              "    length = %s," % _real(mesh["length"]),
              "    refinementLevel = %d" % int(mesh["refinementLevel"]), "}"]
    else:
        L.append(" mesh = %s" % _lua_str(mesh))
    L.append(" weights = %s" % _lua_str(weights))
    tp = ["    sim = %s" % _real(time["sim"])]
    if time.get("iter") is not None:
        tp.append("    iter = %d" % int(time["iter"]))
    if time.get("clock") is not None:
        tp.append("    clock = %s" % _real(time["clock"]))
    L += [" time_point = {", ",\n".join(tp), "}", " nElems = %d" % int(nElems), " nDofs = %d" % int(nDofs),
          " solver = %s" % _lua_str(solver), " varsys = {",
          "    systemname = %s," % _lua_str(varsys["systemname"]), "    variable = {"]
    vs = []
    for v in varsys["variable"]:
        e = ["            name = %s" % _lua_str(v["name"]), "            ncomponents = %d" % int(v["ncomponents"])]
        if v.get("state_varpos") is not None and len(v["state_varpos"]):
            e.append("            state_varpos = %s" % _int_list(list(v["state_varpos"]), 16))
        vs.append("        {\n" + ",\n".join(e) + "\n        }")
    nScalars = sum(int(v["ncomponents"]) for v in varsys["variable"])
    L += [",\n".join(vs), "    },", "    nScalars = %d," % nScalars,
          "    nStateVars = %d," % len(varsys["variable"]),
          "    nAuxScalars = %d," % int(varsys.get("nAuxScalars", 0)),
          "    nAuxVars = %d" % int(varsys.get("nAuxVars", 0)), "}"]
    text = "\n".join(L) + "\n"
    if solver_spec:
        text += solver_spec if solver_spec.endswith("\n") else solver_spec + "\n"
    return text


def fluid_varsys(kind, QQ):
    """the variable system a single-field flow scheme dumps: the state variable 'pdf' with QQ
    components; auxField = density + velocity (mus_scheme_module.f90:222, the tutorial's header
    mus/examples/tutorials/tut_05_restart.md:53-68)"""
    return dict(systemname=kind, variable=[dict(name="pdf", ncomponents=QQ, state_varpos=list(range(1, QQ + 1)))],
                nAuxScalars=4, nAuxVars=2)


def write_restart(prefix, sim_name, data, time, varsys, mesh="./mesh/", elem_offset=0, nElems_global=None,
                  write_header=True, **header_kw):
    """one dump.  data: this rank's elements in tree order, nScalars values each (what
    mus_pdf_serialize / musb200_pdf_serialize produced); elem_offset / nElems_global: this rank's
    position in the global tree (tree%elemOffset, tree%global%nElems) -- every rank writes its
    share into the one file, rank 0 (write_header) the two header scripts.
    Returns (binary path, header path)."""
    nScalars = sum(int(v["ncomponents"]) for v in varsys["variable"]) * int(header_kw.get("nDofs", 1))
    d = np.ascontiguousarray(data, dtype=np.float64).ravel()
    if d.size % nScalars:
        raise ValueError("restart dump: %d values are not a multiple of nScalars = %d" % (d.size, nScalars))
    nLoc = d.size // nScalars
    nGlob = nLoc if nElems_global is None else int(nElems_global)
    if elem_offset < 0 or elem_offset + nLoc > nGlob:
        raise ValueError("restart dump: elements %d..%d outside the mesh of %d" % (elem_offset, elem_offset + nLoc, nGlob))
    stamp = time_stamp(time["sim"])
    base = prefix + sim_name
    if os.path.dirname(base):
        os.makedirs(os.path.dirname(base), exist_ok=True)
    bin_name = base + "_" + stamp + ENDIAN_SUFFIX
    # every rank of a dump arrives here at once: open WITHOUT truncating (O_CREAT is atomic, two
    # ranks that both find the file missing still share one file) and write this rank's share
    # at its offset -- the MPI_File_write_all of tem_restart_writeData in POSIX terms
    fd = os.open(bin_name, os.O_RDWR | os.O_CREAT, 0o644)
    try:
        buf, off = memoryview(d).cast("B"), int(elem_offset) * nScalars * 8
        while len(buf):
            k = os.pwrite(fd, buf[:1 << 30], off)
            buf, off = buf[k:], off + k
    finally:
        os.close(fd)
    hdr_name = base + "_header_" + stamp + ".lua"
    if write_header:
        text = header_text(bin_name, time, nGlob, varsys, mesh=mesh, **header_kw)
        for name in (hdr_name, base + "_lastHeader.lua"):
            with open(name, "w") as fh:
                fh.write(text)
    return bin_name, hdr_name


def tree_order(levelDescs):
    """tree%treeID and the levelPointer of the fluid elements of all levels on this rank:
    the leaves sorted along the space-filling curve (an element of level l precedes whatever
    follows its last descendant; treelm stores the mesh this way, treelmesh_module.f90), with
    levelPointer(i) = 1-based position of leaf i in its level's total list
    (mus_construction_module.fpp, levelPointer).  levelDescs: {level: descriptor with total, nFluid}."""
    finest = max(levelDescs)
    first, lp, ids = [], [], []
    for l, ld in levelDescs.items():
        t = np.asarray(ld.total[:ld.nFluid], dtype=np.int64)
        offset = (8 ** l - 1) // 7                                 # tem_firstIdAtLevel
        first.append((t - offset) * 8 ** (finest - l))             # Morton index of the first finest descendant
        ids.append(t)
        lp.append(np.arange(1, ld.nFluid + 1, dtype=np.int32))
    first, ids, lp = np.concatenate(first), np.concatenate(ids), np.concatenate(lp)
    order = np.argsort(first, kind="stable")
    return ids[order], lp[order]


class RestartFile:
    """what tem_restart_readHeader keeps of a header, plus access to the binary file"""

    def __init__(self, header_path, base_dir=None):
        self.header_path = header_path
        self.header = h = parse_lua_assignments(open(header_path).read())
        for key in ("binary_name", "time_point", "varsys"):
            if key not in h:
                raise ValueError("restart header %s: %s missing" % (header_path, key))
        self.nElems = int(h.get("nElems", 1))
        self.nDofs = int(h.get("nDofs", 1))
        self.solver = h.get("solver", "")
        self.solver_configFile = h.get("solver_configFile", "")
        self.mesh = h.get("mesh")
        tp = h["time_point"]
        self.time = dict(sim=float(tp["sim"]), iter=tp.get("iter"), clock=tp.get("clock")) \
            if isinstance(tp, dict) else dict(sim=float(tp), iter=None, clock=None)
        vs = h["varsys"]
        self.systemname = vs["systemname"]
        self.variables = vs["variable"]
        self.nScalars = int(vs.get("nScalars", sum(int(v["ncomponents"]) for v in self.variables)))
        names = h["binary_name"]
        name = names[0] if isinstance(names, list) else names
        # the name is relative to the directory the solver ran in
        cands = [name] if os.path.isabs(name) else \
            [os.path.join(base_dir or os.getcwd(), name),
             os.path.join(os.path.dirname(os.path.abspath(header_path)), os.path.basename(name))]
        self.binary_path = next((c for c in cands if os.path.exists(c)), None)
        if self.binary_path is None:
            raise FileNotFoundError("restart binary %r of %s not found" % (name, header_path))
        self.byteorder = ">" if self.binary_path.endswith(".msb") else "<"
        want = self.nElems * self.nScalars * self.nDofs * 8
        if os.path.getsize(self.binary_path) != want:
            raise ValueError("restart binary %s holds %d bytes, header says %d elements x %d scalars"
                             % (self.binary_path, os.path.getsize(self.binary_path), self.nElems,
                                self.nScalars * self.nDofs))

    def part(self, rank, nranks):
        """(elemOffset, nElems) of a rank: treelm's equal distribution, the first `remainder`
        ranks hold one element more (treelmesh_module.f90:1276-1296)"""
        share, rem = divmod(self.nElems, int(nranks))
        return rank * share + min(rank, rem), share + (1 if rank < rem else 0)

    def read(self, elem_offset=0, nElems=None):
        """[nElems][nScalars*nDofs] doubles of the elements elem_offset .. (tem_restart_readData)"""
        n = self.nElems - elem_offset if nElems is None else int(nElems)
        if elem_offset < 0 or n < 0 or elem_offset + n > self.nElems:
            raise ValueError("restart read: elements %d..%d outside the file's %d" % (elem_offset, elem_offset + n, self.nElems))
        w = self.nScalars * self.nDofs
        d = np.fromfile(self.binary_path, dtype=self.byteorder + "f8", count=n * w, offset=elem_offset * w * 8)
        return d.astype(np.float64, copy=False).reshape(n, w)


def read_restart(header_path, rank=0, nranks=1, base_dir=None):
    """-> (RestartFile, elemOffset, this rank's [nElems_local][nScalars] block)"""
    rf = RestartFile(header_path, base_dir)
    off, n = rf.part(rank, nranks)
    return rf, off, rf.read(off, n)
