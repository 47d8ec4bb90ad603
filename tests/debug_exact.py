"""per-kernel bit-exactness report of the CUDA path vs the oracle after 1 and 20 steps"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import musubi_b200 as mb
from oracle import musoracle as mo
from helpers import make_pair
mb.mus_init(0, 1, 0)
for relax in ("bgk", "trt", "mrt"):
    for QQ in (19, 27):
        ident = {"kind": "fluid", "relaxation": relax, "layout": "d3q%d" % QQ}
        ld, old, ref, sch = make_pair(mb, mo, 4, ident, 1.7, omega_bulk=1.2)
        n = ld.nFluid * QQ
        for steps in (1, 19):
            sch.do_computation(steps); ref.run(steps)
            got = sch.download_state(4)[:n].reshape(-1, QQ); exp = ref.state[ref.nNext][:n].reshape(-1, QQ)
            diff = got != exp
            aux = sch.download_aux(4)[:ld.nFluid*4]; 
            print(relax, QQ, "steps", steps, "ndiff", int(diff.sum()), "maxrel %.2e" % np.max(np.abs(got-exp)/np.abs(exp)),
                  "per-dir", diff.sum(axis=0).tolist() if diff.any() else "", "aux ndiff", int((aux != ref.aux[:ld.nFluid*4]).sum()))
        sch.destroy()
mb.mus_finalize()
