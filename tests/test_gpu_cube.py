"""treelm's predefined cube generated on the device (musb200_level_create_cube) and the
device-side equilibrium initial state: both must equal what the host path delivers -- the index
list bit for bit (musb200_neigh_download against the host generator's and the oracle's list),
the initial PDFs bit for bit against the oracle's mus_init_pdf restatement -- and a run on the
device-generated mesh must be bit-identical to the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mb():
    import musubi_b200
    musubi_b200.mus_init(0, 1, 0)
    yield musubi_b200
    musubi_b200.mus_finalize()


@pytest.mark.parametrize("QQ", [19, 27])
@pytest.mark.parametrize("level", [1, 2, 4, 6])
@pytest.mark.parametrize("kind", ["periodic", "walls"])
def test_device_generated_connectivity_is_the_host_list(mb, oracle, QQ, level, kind):
    ident = {"kind": "fluid", "relaxation": "bgk", "layout": "d3q%d" % QQ}
    host_kind = "periodic" if kind == "periodic" else "cavity"     # walls + lid: same bounce-back list
    ld = mb.LevelDesc(level, QQ, host_kind)
    old = oracle.build_level_desc(level, QQ, host_kind)
    cube = mb.DeviceCube(level, QQ, kind)
    assert (cube.nFluid, cube.nSize) == (ld.nFluid, ld.nSize)
    sch = mb.Scheme(ident, cube, 1.7)
    got = sch.download_neigh(level)
    assert np.array_equal(got, ld.neigh)
    assert np.array_equal(got, old.neigh)
    assert np.allclose(cube.barycenters((0.0, 0.0, 0.0), 2.0), ld.barycenters((0.0, 0.0, 0.0), 2.0), rtol=0, atol=0)
    sch.destroy()


@pytest.mark.parametrize("relax,QQ,kindname", [("trt", 19, "fluid"), ("mrt", 27, "fluid"), ("bgk", 27, "fluid"),
                                               ("mrt", 19, "fluid_incompressible"),
                                               ("bgk", 27, "fluid_incompressible")])
def test_device_cube_run_with_device_initial_state_matches_oracle(mb, oracle, relax, QQ, kindname):
    from musubi_b200 import cases
    level = 4
    ident = {"kind": kindname, "relaxation": relax, "layout": "d3q%d" % QQ}
    cube = mb.DeviceCube(level, QQ, "periodic")
    old = oracle.build_level_desc(level, QQ, "periodic")
    ref = oracle.Scheme(old, relax, kindname, omega=1.8, lambda_=0.25, omega_bulk=1.3)

    class _B:        # taylor_green needs only barycenters()
        barycenters = cube.barycenters
    rho, vel = cases.taylor_green(_B, mean=(0.02, -0.01, 0.015))
    ref.init_equilibrium(rho, vel)
    sch = mb.Scheme(ident, cube, float(1.0 / (3.0 * ref.visc[0] + 0.5)), lambda_=0.25, omega_bulk=1.3)
    sch.init_equilibrium(level, rho, vel)
    n = cube.nFluid * QQ
    for which in (1, 2):                                   # initial PDFs, both buffers
        assert np.array_equal(sch.download_state(level, which)[:n], ref.state[ref.nNext][:n])
    sch.do_computation(30)
    ref.run(30)
    assert np.array_equal(sch.download_state(level)[:n], ref.state[ref.nNext][:n])
    sch.destroy()
