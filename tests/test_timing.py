"""calc_MLUPS / dump_timing of the reference (mus_tools_module.f90:474-691) as restated in
musubi_b200/timing.py."""
import numpy as np

from musubi_b200 import timing


def test_calc_mlups_counts_like_the_reference():
    # single level: plain updates per second
    assert timing.calc_mlups({8: 256 ** 3}, 500, 0.5) == 256 ** 3 * 500 / 0.5e6
    # two levels: per iteration the COARSE level enters with nElems / 2 (integer division) ...
    n = {7: 1835009, 8: 2013265}
    assert timing.calc_mlups(n, 100, 1.0) == (1835009 // 2 + 2013265) * 100 / 1e6
    # ... and the main loop's iteration count is halved once more
    full, kernel = timing.perf_measure(n, main_loop_iters=201, t_mainloop=1.0, t_compute=0.5)
    assert full == timing.calc_mlups(n, 100, 1.0) and kernel == 2.0 * full
    # the updates actually performed per coarse cycle are 4x that count (2 levels)
    performed = (n[7] + 2 * n[8]) * 200 / 1e6
    assert abs(performed / timing.perf_measure(n, 200, 1.0, 1.0)[0] - 4.0) < 1e-6
    # three levels
    assert timing.calc_mlups({4: 9, 5: 9, 6: 9}, 1, 1e-6) == 9 // 4 + 9 // 2 + 9


def test_dump_timing_record_layout(tmp_path):
    f = str(tmp_path / "mus_timing.res")
    args = dict(revision="b200-r01", sim_name="cavity", dom_size=16777216, n_procs=1, mlups=17079.06,
                mlups_kernel=16825.5, imbalance=0.0, t_musubi=12.3456, max_iter=500, total_dens=16777274.173347898,
                timers=[("MainLoop", 0.4912), ("L08_compute", 0.4821)],
                ratios={"Comp": 97.7, "BC": 1.7})
    head, out = timing.dump_timing(f, **args)
    timing.dump_timing(f, **args)
    lines = open(f).read().splitlines()
    assert len(lines) == 3 and lines[0] == head and lines[1] == lines[2] == out
    assert len(head) == len(out)                       # the columns line up
    cols, vals = head[1:].split(), out.split()
    assert cols[:10] == ["Revision", "SimName", "DomSize", "nProcs", "MLUPs", "MLUPs_kernel", "imbalance(%)",
                         "timeMusubi", "maxIter", "totalDens"]
    assert cols[10:12] == ["timeMainLoop", "timeL08_compute"] and cols[-5:] == ["Comp(%)", "Comm(%)", "BCbuffer(%)",
                                                                              "BC(%)", "Intp(%)"]
    rec = dict(zip(cols, vals))
    assert rec["MLUPs"] == "17.08E+03" and rec["totalDens"] == "16.777274173E+06" and rec["timeMusubi"] == "12.3456E+00"
    assert float(rec["Comp(%)"]) == 97.7 and int(rec["DomSize"]) == 16777216
    assert np.isclose(float(rec["MLUPs_kernel"]), 16825.5, rtol=1e-3)
