"""Edge cases and size-independent properties on the device:
tiny and ragged meshes against the oracle, the error behaviour of the C ABI (the reference
aborts through tem_abort; the library returns a code + message), and round-trip / conservation
properties at BASELINE config 2's full size (256^3)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import musubi_b200
    musubi_b200.mus_init(0, 1, 0)
    yield musubi_b200
    musubi_b200.mus_finalize()


KERNELS = [(r, q) for q in (19, 27) for r in ("bgk", "trt", "mrt")]


@pytest.mark.parametrize("relax,QQ", KERNELS)
@pytest.mark.parametrize("level", [1, 2])
def test_tiny_periodic_meshes_match_oracle(mb, oracle, relax, QQ, level):
    """2^3 and 4^3 periodic cubes: every neighbour wraps around, at level 1 an element is its own
    second neighbour; far fewer elements than one CTA"""
    from musubi_b200 import cases
    ident = {"kind": "fluid", "relaxation": relax, "layout": "d3q%d" % QQ}
    ld = mb.LevelDesc(level, QQ, "periodic")
    old = oracle.build_level_desc(level, QQ, "periodic")
    assert np.array_equal(ld.neigh, old.neigh)
    ref = oracle.Scheme(old, relax, "fluid", omega=1.6, lambda_=0.25, omega_bulk=1.2)
    rng = np.random.default_rng(level * 100 + QQ)
    ref.init_equilibrium(1.0 + 0.02 * rng.standard_normal(old.nElems), 0.03 * rng.standard_normal((old.nElems, 3)))
    sch = mb.Scheme(ident, ld, float(1.0 / (3.0 * ref.visc[0] + 0.5)), lambda_=0.25, omega_bulk=1.2)
    sch.upload_state(level, ref.state[ref.nNow], ref.state[ref.nNext])
    ref.run(25)
    sch.do_computation(25)
    n = ld.nFluid * QQ
    assert np.array_equal(sch.download_state(level)[:n], ref.state[ref.nNext][:n])
    sch.destroy()


def test_abi_error_behaviour(mb):
    """wrong calls return the documented codes with a message instead of aborting"""
    from musubi_b200._lib import P_I32, P_I64, last_error, lib, ptr
    QQ, n = 19, 8
    ld = mb.LevelDesc(1, QQ, "periodic")
    assert lib.musb200_step(7, 7, 1) == 1 and "level 7" in last_error()                  # ERR_ARG
    assert lib.musb200_level_create(1, 15, 15, 4, n, n, 0, 0, 0, ptr(ld.neigh, P_I32), None, None) == 4
    assert lib.musb200_level_create(1, QQ, QQ, 3, n, n, 0, 0, 0, ptr(ld.neigh, P_I32), None, None) == 4
    assert lib.musb200_level_create(1, QQ, QQ, 4, 4, n, 0, 0, 0, ptr(ld.neigh, P_I32), None, None) == 1
    bad = ld.neigh.copy()
    bad[3] = 2 * QQ + 5          # direction 1 of element 4 pulls direction 5 of element 3
    assert lib.musb200_level_create(1, QQ, QQ, 4, ld.nSize, n, 0, 0, 0, ptr(bad, P_I32), None, None) == 5
    assert "neither a plain pull nor a bounce-back" in last_error()                         # ERR_CONNECTIVITY
    assert lib.musb200_level_create(1, QQ, QQ, 4, ld.nSize, n, 0, 0, 0, ptr(ld.neigh, P_I32),
                                    ptr(ld.property, P_I64), ptr(ld.total, P_I64)) == 0
    assert lib.musb200_step(1, 1, 1) == 6 and "set_relaxation" in last_error()             # ERR_STATE
    assert lib.musb200_set_now_next(1, 1, 1) == 1
    assert lib.musb200_set_relaxation(1, 7, 0, None, 1.5, 0.25, 1.5) == 1
    assert lib.musb200_set_relaxation(1, 0, 0, None, 1.5, 0.25, 1.5) == 0
    links = np.array([1], dtype=np.int32)
    assert lib.musb200_bc_register(1, 1, 1, 1, ptr(links, P_I32), ptr(links, P_I32), ptr(links, P_I32),
                                   ptr(links, P_I32)) == 6                                  # no bc_elemBuffer yet
    assert lib.musb200_bc_register(1, 1, 9, 0, None, None, None, None) == 4                 # unknown BC kind
    assert lib.musb200_source_force(1, 3, n, None, ptr(np.zeros(3), ctypes.POINTER(ctypes.c_double)), 1) == 1
    assert lib.musb200_set_species(1, 0, 1, 0.1, 0.25) == 1                                 # needs nAuxScalars = 1
    assert lib.musb200_scheme_bind(99) == 1
    assert lib.musb200_step(1, 1, 3) == 0
    assert lib.musb200_level_destroy(1) == 0
    # scheme selection outside the hot path: the reference tem_aborts, the library reports code 4
    r, k, q = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    for kind, rel, var, lay in (("multispecies_liquid", "bgk", "standard", "d3q19"),
                                ("fluid", "cumulant", "standard", "d3q27"),
                                ("fluid", "bgk", "standard", "d2q9"),
                                ("fluid_incompressible", "trt", "standard", "d3q27"),
                                ("passive_scalar", "mrt", "standard", "d3q19")):
        assert lib.musb200_scheme_select(kind.encode(), rel.encode(), var.encode(), lay.encode(),
                                         ctypes.byref(r), ctypes.byref(k), ctypes.byref(q)) == 4


def test_full_size_cavity_properties(mb):
    """BASELINE config 2 at its full size (256^3, D3Q19 TRT): index lists and state survive the
    device round trip bit for bit; with the lid at rest (pure bounce-back box, swirling initial
    flow) mass is conserved to 1e-13; with the lid moving nothing turns NaN and the only mass
    source is the well-known one of velocity bounce-back at the lid's edges (< 1e-6 relative)"""
    from musubi_b200 import cases
    level, QQ = 8, 19
    ld = mb.LevelDesc(level, QQ, "cavity")
    assert ld.nFluid == 256 ** 3
    sch = mb.Scheme({"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}, ld, 1.7, lambda_=3.0 / 16.0)
    assert np.array_equal(sch.download_neigh(level), ld.neigh)            # bit-exact index lists
    rho, vel = cases.taylor_green(ld)
    st = cases.equilibrium_state(QQ, rho, vel, ld.nSize)
    del rho, vel
    sch.upload_state(level, st)
    back = sch.download_state(level)
    assert np.array_equal(back, st)                                       # AOS -> SoA -> AOS identity
    del back, st
    sch.set_bc_values(level, 2, cases.lid_values(ld, (0.0, 0.0, 0.0)))    # closed box
    m0 = sch.reduce(level)[0]
    sch.do_computation(40)
    mass, vmax, nan = sch.reduce(level)
    assert nan == 0 and vmax > 0.0
    assert abs(mass / m0 - 1.0) < 1e-13
    sch.set_bc_values(level, 2, cases.lid_values(ld))                     # lid at (0.05, 0, 0)
    sch.do_computation(40)
    mass2, vmax2, nan = sch.reduce(level)
    assert nan == 0 and 0.0 < vmax2 < 0.2
    assert abs(mass2 / mass - 1.0) < 1e-6
    # restart round trip at full size: dump in treeID order, perturb, restore, dump again
    tid = np.asarray(ld.total[:ld.nFluid], dtype=np.int64)
    lp = np.arange(1, ld.nFluid + 1, dtype=np.int32)
    half = ld.nFluid // 2
    a = sch.pdf_serialize(tid[:half], lp[:half])
    sch.pdf_unserialize(tid[:half], lp[:half], a * 0.5)
    sch.pdf_unserialize(tid[:half], lp[:half], a)
    assert np.array_equal(sch.pdf_serialize(tid[:half], lp[:half]), a)
    sch.destroy()
