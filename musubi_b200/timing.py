"""Performance report of the reference (mus_perf_measure / dump_timing / calc_MLUPS,
mus_tools_module.f90:474-691): the MLUPS figure as the reference counts it and the one-line
record it appends to `timing_file` (musubi.lua: timing_file = 'mus_timing.res'), from the device
timers of libmusb200 (musb200_timers) -- so that runs of this path land in the same table as
runs of the Fortran solver.

Note the reference's count for multi-level runs: per iteration a level enters with
nElems(l) / sf^(maxLevel - l) (integer division: the COARSE levels are scaled down, i.e. the
unit is one finest-level step), and the iteration count of the main loop -- which advances once
per coarsest-level step -- is divided by sf^(maxLevel - minLevel) once more
(mus_tools_module.f90:505-507).  For nLevels > 1 the figure is therefore
4^(maxLevel - minLevel) times smaller than the number of element updates actually performed per
second; bench.py reports SURVEY section 8d's count and this one side by side."""
import os

from .restart_io import fortran_en


def calc_mlups(nElems, iters, time_s, scale_factor=2):
    """calc_MLUPS (mus_tools_module.f90:658-691).  nElems: {level: global fluid element count}"""
    max_level = max(nElems)
    updates = sum(int(n) // int(scale_factor ** (max_level - l)) for l, n in nElems.items())
    return float(updates * int(iters)) / (float(time_s) * 1000000.0)


def perf_measure(nElems, main_loop_iters, t_mainloop, t_compute, scale_factor=2):
    """mus_perf_measure's two figures (:498-531): iter = (now%iter - min%iter) /
    sf^(maxLevel - minLevel) with the main loop's iteration counter (one per coarsest-level step);
    MLUPs over the main-loop time, MLUPs_kernel over the summed compute timers"""
    iters = int(main_loop_iters) // int(scale_factor ** (max(nElems) - min(nElems)))
    return calc_mlups(nElems, iters, t_mainloop, scale_factor), calc_mlups(nElems, iters, t_compute, scale_factor)


def dump_timing(filename, revision, sim_name, dom_size, n_procs, mlups, mlups_kernel, imbalance, t_musubi,
                max_iter, total_dens, timers, t_aux=0.0, t_relax=0.0, ratios=None):
    """dump_timing (:547-650): appends one record, writing the header line when the file is new.
    timers: [(name, seconds)] in the order of mus_timerHandles; ratios: dict with the keys
    Comp, Comm, BCbuffer, BC, Intp in percent."""
    ratios = ratios or {}
    head = "#" + "Revision".rjust(15) + "SimName".rjust(21) + "DomSize".rjust(15) + "nProcs".rjust(10) \
        + "MLUPs".rjust(14) + "MLUPs_kernel".rjust(14) + "imbalance(%)".rjust(14) + "timeMusubi".rjust(14) \
        + "maxIter".rjust(10) + "totalDens".rjust(19)
    out = str(revision).rjust(16) + (" " + str(sim_name)).rjust(21) + ("%d" % dom_size).rjust(15) \
        + ("%d" % n_procs).rjust(10) + fortran_en(mlups, 2, 14) + fortran_en(mlups_kernel, 2, 14) \
        + ("%.2f" % imbalance).rjust(14) + fortran_en(t_musubi, 4, 14) + ("%d" % max_iter).rjust(10) \
        + fortran_en(total_dens, 9, 19)
    for name, val in timers:
        head += ("time" + name).rjust(16)
        out += ("%.4f" % val).rjust(16)
    head += "timeAux".rjust(12) + "timeRelax".rjust(12)
    out += ("%.2f" % t_aux).rjust(12) + ("%.2f" % t_relax).rjust(12)
    for key, width in (("Comp", 12), ("Comm", 12), ("BCbuffer", 15), ("BC", 12), ("Intp", 12)):
        head += (key + "(%)").rjust(width)
        out += ("%.2f" % float(ratios.get(key, 0.0))).rjust(width)
    new = not os.path.exists(filename)
    with open(filename, "a") as fh:
        if new:
            fh.write(head + "\n")
        fh.write(out + "\n")
    return head, out
