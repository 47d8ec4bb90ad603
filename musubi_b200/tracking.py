"""Tracking output of the reference, read side of the hot path (SURVEY.md section 8f, n2 "restart /
tracking bridge"): the auxField moments the device step produces -> the derived *_phy variables ->
the reference's ascii / asciiSpatial result files, so that the reference's own regression check
(pysys-extensions/apes/apeshelper.py:90-123: numpy.loadtxt of the produced and the reference
file + numpy.allclose(rtol=1e-10, atol=1e-5)) runs on files written by this library.

  shape  canoND point / line -> elements: tem_cano_initSubTree (tem_canonical_module.f90:826-916),
         tem_CoordOfReal (tem_geometry_module.f90:149-185), tem_lineCubeOverlap / rayCubeOverlap
         (shapes/tem_line_module.fpp:69-174): a cube is half open, [origin, endPnt)
  vars   density_phy, pressure_phy, velocity_phy, vel_mag_phy, kinetic_energy_phy
         (mus_derQuan_module.fpp / mus_derQuanIncomp_module.fpp, factors mus_physics_module.f90:522-543)
  files  hvs_asciiSpatial_open / _dump_elem_data and hvs_ascii_open / _dump_elem_data
         (libharvesting/hvs_ascii_module.f90:393-640, 912-1030), header getHeader (:1251-1320):
         <folder><simName>_<label>_p<rank 5 digits>[_t<EN12.3 stamp>].res, values in e24.16e3"""
import os
from decimal import ROUND_HALF_EVEN, Decimal

import numpy as np

from .restart_io import time_stamp
from .treelm_multilevel import coords, first_id, morton


# ------------------------------------------------------------------------------------------
def fortran_e(x, w=24, d=16, e=3):
    """Fortran's Ew.dEe edit descriptor: 0.dddd with d digits and an e-digit exponent, right
    adjusted in w columns ('  0.3125000000000000E+000')."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        return ("NaN" if x != x else ("Infinity" if x > 0 else "-Infinity")).rjust(w)
    if x == 0.0:
        mant, ex = "0" * d, 0
    else:
        v = abs(Decimal(x))
        ex = v.adjusted() + 1                                # 0.1 <= v / 10^ex < 1
        q = v.scaleb(d - ex).quantize(Decimal(1), rounding=ROUND_HALF_EVEN)
        if q >= Decimal(10) ** d:                            # rounded up to 1.000...
            ex += 1
            q = v.scaleb(d - ex).quantize(Decimal(1), rounding=ROUND_HALF_EVEN)
        mant = str(int(q)).rjust(d, "0")
    sign = "-" if (x < 0.0 or (x == 0.0 and str(x)[0] == "-")) else ""
    return ("%s0.%sE%s%0*d" % (sign, mant, "+" if ex >= 0 else "-", e, abs(ex))).rjust(w)


# ------------------------------------------------------------------------------------------
class Physics:
    """mus_physics_type: the conversion factors between lattice and physical units of one level
    (mus_physics_module.f90:522-543)."""

    def __init__(self, dx, dt, rho0=1.0):
        self.dx, self.dt, self.rho0 = float(dx), float(dt), float(rho0)
        self.fac_vel = dx / dt
        self.fac_press = rho0 * dx ** 2 / dt ** 2
        self.fac_energy = rho0 * dx ** 5 / dt ** 2


def derive(name, aux, phys, incompressible=False):
    """[n][ncomp] of a derived variable from auxField rows (rho, ux, uy, uz) in lattice units"""
    aux = np.asarray(aux, dtype=np.float64).reshape(-1, 4)
    rho, u = aux[:, 0], aux[:, 1:4]
    if name == "density_phy":
        return (rho * phys.rho0)[:, None]
    if name == "pressure_phy":
        return (rho * (1.0 / 3.0) * phys.fac_press)[:, None]
    if name == "velocity_phy":
        return u * phys.fac_vel
    if name == "vel_mag_phy":
        return (np.sqrt(u[:, 0] * u[:, 0] + u[:, 1] * u[:, 1] + u[:, 2] * u[:, 2]) * phys.fac_vel)[:, None]
    if name == "kinetic_energy_phy":
        dens = 1.0 if incompressible else rho
        return ((u[:, 0] * u[:, 0] + u[:, 1] * u[:, 1] + u[:, 2] * u[:, 2]) * 0.5 * dens * phys.fac_energy)[:, None]
    raise ValueError("tracking variable %r is not derived on this path" % name)


NCOMP = {"density_phy": 1, "pressure_phy": 1, "velocity_phy": 3, "vel_mag_phy": 1, "kinetic_energy_phy": 1}


# ------------------------------------------------------------------------------------------
def _level_of(tid):
    level, first, count = 0, 0, 1
    while tid >= first + count:
        first, count, level = first + count, count * 8, level + 1
    return level


def barycenters_of(treeID, origin, length):
    """tem_BaryOfId for a list of treeIDs of one level"""
    t = np.asarray(treeID, dtype=np.int64)
    if t.size == 0:
        return np.zeros((0, 3))
    level = _level_of(int(t[0]))
    x, y, z = coords(t - first_id(level))
    dx = float(length) / float(1 << level)
    o = np.asarray(origin, dtype=np.float64)
    return np.stack([o[0] + (x + 0.5) * dx, o[1] + (y + 0.5) * dx, o[2] + (z + 0.5) * dx], axis=1)


def select_point(treeID, point, origin, length, max_level=None):
    """0-based position of the leaf holding the point (ascending treeID list of one level or the
    leaves of several), or -1: the point's cell at max_level, then its ancestors"""
    t = np.asarray(treeID, dtype=np.int64)
    maxL = max_level if max_level is not None else _level_of(int(t.max()))
    n = 1 << maxL
    c = [max(min(int((float(point[i]) - float(origin[i])) * (float(n) / float(length))), n - 1), 0) for i in range(3)]
    m = int(morton(np.array([c[0]]), np.array([c[1]]), np.array([c[2]]))[0])
    order = np.argsort(t, kind="stable")
    for level in range(maxL, -1, -1):
        tid = first_id(level) + (m >> (3 * (maxL - level)))
        k = int(np.searchsorted(t[order], tid))
        if k < t.size and t[order][k] == tid:
            return int(order[k])
    return -1


def select_line(treeID, line_origin, vec, origin, length):
    """0-based positions (in list order) of the elements a canoND line overlaps"""
    t = np.asarray(treeID, dtype=np.int64)
    out = np.zeros(t.size, dtype=bool)
    lo, v = np.asarray(line_origin, dtype=np.float64), np.asarray(vec, dtype=np.float64)
    lv_of = np.zeros(t.size, dtype=np.int64)
    for level in range(0, 21):                       # level of every element from the id ranges
        sel = (t >= first_id(level)) & (t < first_id(level + 1))
        lv_of[sel] = level
    for level in np.unique(lv_of):
        sel = np.nonzero(lv_of == level)[0]
        x, y, z = coords(t[sel] - first_id(int(level)))
        dx = float(length) / float(1 << int(level))
        cmin = np.stack([origin[0] + x * dx, origin[1] + y * dx, origin[2] + z * dx], axis=1)
        cmax = cmin + dx
        ok = np.ones(sel.size, dtype=bool)
        t_near = np.zeros(sel.size)
        t_far = np.full(sel.size, np.finfo(np.float64).max)
        for i in range(3):
            if abs(v[i]) <= np.finfo(np.float64).eps:            # .feq. 0
                ok &= ~((lo[i] < cmin[:, i]) | (lo[i] >= cmax[:, i]))
            else:
                t1, t2 = (cmin[:, i] - lo[i]) / v[i], (cmax[:, i] - lo[i]) / v[i]
                t1, t2 = np.minimum(t1, t2), np.maximum(t1, t2)
                t_near, t_far = np.maximum(t_near, t1), np.minimum(t_far, t2)
                ok &= ~((t_near > t_far) | (t_far < 0.0))
        p = lo[None, :] + t_near[:, None] * v[None, :]
        proj = ((p - lo[None, :]) @ v) / float(v @ v)
        ok &= (proj >= 0.0) & (proj < 1.0)
        out[sel[ok]] = True
    return np.nonzero(out)[0]


# ------------------------------------------------------------------------------------------
def _header(variables, reduced=False):
    cols, red = [], ("_red" if reduced else "")
    for name in variables:
        n = NCOMP[name]
        cols += [name + red] if n == 1 else ["%s%s_%02d" % (name, red, c) for c in range(1, n + 1)]
    return "".join(" " + c.rjust(24) for c in cols)


def _basename(folder, sim_name, label, rank):
    return "%s%s_%s_p%05d" % (folder, sim_name, label, rank)


def write_ascii_spatial(folder, sim_name, label, sim_time, bary, values, variables, rank=0):
    """one asciiSpatial file: a row per element, barycentre + the variables' components.
    values: [n][sum of ncomp].  Returns the file name."""
    bary, values = np.asarray(bary, dtype=np.float64), np.asarray(values, dtype=np.float64)
    if bary.shape[0] != values.shape[0] or values.shape[1] != sum(NCOMP[v] for v in variables):
        raise ValueError("tracking: %r rows of %r values for %r" % (bary.shape, values.shape, variables))
    if os.path.dirname(folder):
        os.makedirs(os.path.dirname(folder), exist_ok=True)
    name = _basename(folder, sim_name, label, rank) + "_t" + time_stamp(sim_time) + ".res"
    with open(name, "w") as fh:
        fh.write("# Rank of the process: %7d\n" % rank)
        fh.write("#" + "".join(" " + c.rjust(24) for c in ("coordX", "coordY", "coordZ")) + _header(variables) + "\n")
        for b, row in zip(bary, values):
            fh.write("".join(" " + fortran_e(x) for x in b) + "".join(" " + fortran_e(x) for x in row) + "\n")
    return name


class AsciiTracker:
    """the `ascii` format: one file, one row per dump -- the time and the variables of the
    tracked element(s) (point tracking)"""

    def __init__(self, folder, sim_name, label, variables, rank=0, reduced=False):
        """reduced: the rows hold spatial reductions (reduction = 'sum' ...) of the variables,
        computed by the caller; the column names get the reference's '_red' suffix"""
        if os.path.dirname(folder):
            os.makedirs(os.path.dirname(folder), exist_ok=True)
        self.name = _basename(folder, sim_name, label, rank) + ".res"
        self.variables = list(variables)
        new = not os.path.exists(self.name)
        self.fh = open(self.name, "a")
        if new:                       # an existing file is appended to, as after a restart
            self.fh.write("# Rank of the process: %7d\n" % rank)
            self.fh.write("#" + "time".rjust(23) + _header(self.variables, reduced) + "\n")

    def dump(self, sim_time, values):
        v = np.asarray(values, dtype=np.float64).ravel()
        self.fh.write(fortran_e(sim_time) + "".join(" " + fortran_e(x) for x in v) + "\n")

    def close(self):
        self.fh.close()


def reduce_spatial(values, op, volumes=None):
    """tem_reduction_spatial (tem_reduction_spatial_module.f90:419-667) over the tracked
    elements of one rank: values [n][ncomp] -> [ncomp].  volumes: dx^3 per element (the
    reference weights the squares of l2norm / l2normalized with the element volume).  Across
    ranks the partial results combine as the reference's mpi_reduce does (sum / max / min)."""
    v = np.asarray(values, dtype=np.float64)
    if v.ndim == 1:
        v = v[:, None]
    vol = np.ones(v.shape[0]) if volumes is None else np.asarray(volumes, dtype=np.float64)
    if op == "sum":
        return v.sum(axis=0)
    if op == "average":
        return v.sum(axis=0) / float(v.shape[0])
    if op in ("l2norm", "l2_norm"):
        return np.sqrt((v * v * vol[:, None]).sum(axis=0))
    if op == "l2normalized":
        return np.sqrt((v * v * vol[:, None]).sum(axis=0) / vol.sum())
    if op in ("linfnorm", "linf_norm", "l_inf_norm"):
        return np.abs(v).max(axis=0)
    if op in ("max", "maximum"):
        return v.max(axis=0)
    if op in ("min", "minimum"):
        return v.min(axis=0)
    raise ValueError("spatial reduction %r is not one of the reference's" % op)


def track(variables, aux, phys, incompressible=False):
    """the variables' components side by side: [n][sum ncomp]"""
    return np.concatenate([derive(v, aux, phys, incompressible) for v in variables], axis=1)
