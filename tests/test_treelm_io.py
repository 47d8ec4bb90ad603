"""treelm mesh files: write / read round trip in the reference's on-disk format
and the single-level descriptor built from a mesh file against the box generator and the oracle."""
import numpy as np
import pytest


def test_file_format_layout(tmp_path):
    from musubi_b200 import treelm_io as tio
    tid = np.array([73, 74, 80, 99], dtype=np.int64)
    prop = np.array([2, 2 | 8, 2, 2 | 8], dtype=np.int64)
    bid = np.zeros((2, 26), dtype=np.int64)
    bid[0, 0], bid[1, 25] = 1, 2
    tio.dump_treelmesh(str(tmp_path), tid, prop, origin=(0.5, -1.0, 2.0), length=4.0,
                       bc_labels=("wall", "lid"), boundary_ID=bid)
    raw = (tmp_path / "elemlist.lsb").read_bytes()
    assert len(raw) == 4 * 16                              # 16 bytes per element
    assert np.frombuffer(raw, dtype="<i8").tolist() == [73, 2, 74, 10, 80, 2, 99, 10]
    assert len((tmp_path / "bnd.lsb").read_bytes()) == 2 * 26 * 8
    m = tio.load_treelmesh(str(tmp_path))
    assert m["nElems"] == 4 and m["minLevel"] == 3 and m["maxLevel"] == 3
    assert m["origin"] == (0.5, -1.0, 2.0) and m["length"] == 4.0
    assert np.array_equal(m["treeID"], tid) and np.array_equal(m["property"], prop)
    assert m["bc_labels"] == ["wall", "lid"] and np.array_equal(m["boundary_ID"], bid)


def test_stencil_to_treelm_map():
    from musubi_b200 import treelm_io as tio
    from musubi_b200.treelm_multilevel import stencil_tables
    for QQ in (19, 27):
        cx, _ = stencil_tables(QQ)
        m = tio.stencil_to_treelm(QQ)
        assert sorted(m.tolist()) == list(range(QQ - 1))
        assert np.array_equal(tio.Q_OFFSET[m], cx[:QQ - 1])
    assert tio.stencil_to_treelm(19).tolist() == list(range(18))   # the first 18 sides coincide


@pytest.mark.parametrize("kind,QQ", [("periodic", 19), ("cavity", 19), ("cavity", 27), ("periodic", 27)])
def test_descriptor_from_file_equals_box_generator(tmp_path, oracle, kind, QQ):
    """dump the generator's mesh in treelm format, read it back, rebuild the descriptor:
    total / property / nghElems / neigh bit-identical (walls are implicit bounce-back), and neigh
    equals the oracle's restatement of mus_construct_connectivity on the same nghElems."""
    import musubi_b200 as mb
    from musubi_b200 import treelm_io as tio
    ld = mb.LevelDesc(4, QQ, kind)
    m = tio.mesh_from_level_desc(ld, length=2.0)
    tio.dump_treelmesh(str(tmp_path), m["treeID"], m["property"], m["origin"], m["length"],
                       m["bc_labels"], m["boundary_ID"])
    fd = tio.FileLevelDesc(tio.load_treelmesh(str(tmp_path)), QQ)
    assert (fd.nFluid, fd.nHalo, fd.nSize) == (ld.nFluid, 0, ld.nSize)
    assert np.array_equal(fd.total, ld.total)
    assert np.array_equal(fd.property, ld.property)
    assert np.array_equal(fd.nghElems, ld.nghElems)
    assert np.array_equal(fd.neigh, ld.neigh)
    ng = np.zeros(QQ * fd.nSize, dtype=np.int32)
    oracle.lib().ora_construct_connectivity(oracle._i(ng), fd.nSize, fd.nElems, QQ,
                                            oracle._i(np.ascontiguousarray(fd.nghElems)),
                                            oracle._l(fd.property), fd.nFluid, fd.nFluid)
    assert np.array_equal(fd.neigh, ng)


@pytest.mark.parametrize("nranks", [2, 4])
def test_partitioned_descriptor_from_file_equals_box_generator(tmp_path, nranks):
    import musubi_b200 as mb
    from musubi_b200 import treelm_io as tio
    QQ = 19
    whole = mb.LevelDesc(4, QQ, "periodic")
    m = tio.mesh_from_level_desc(whole)
    tio.dump_treelmesh(str(tmp_path), m["treeID"], m["property"])
    mesh = tio.load_treelmesh(str(tmp_path))
    fds = {r: tio.FileLevelDesc(mesh, QQ, r, nranks) for r in range(nranks)}
    for r, fd in fds.items():
        fd.build_send(fds)
        ld = mb.LevelDesc(4, QQ, "periodic", r, nranks)
        assert np.array_equal(fd.total, ld.total)
        assert np.array_equal(fd.nghElems, ld.nghElems)
        assert np.array_equal(fd.neigh, ld.neigh)
        for mine, ref in ((fd.recv, ld.recv), (fd.send, ld.send)):
            assert [c["proc"] for c in mine] == [c["proc"] for c in ref]
            for a, b in zip(mine, ref):
                assert np.array_equal(a["pos"], b["pos"])


def test_sphere_obstacle_mesh_runs_through_the_file_path(tmp_path, oracle):
    """a mesh the box generator cannot produce: fluid cells around a solid sphere (cells absent,
    neighbours see boundary 'sphere'); connectivity against the oracle's C restatement."""
    from musubi_b200 import treelm_io as tio
    from musubi_b200.treelm_multilevel import first_id, morton
    L, QQ = 4, 19
    n = 1 << L
    g = np.arange(n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    solid = (X - 7.5) ** 2 + (Y - 7.5) ** 2 + (Z - 7.5) ** 2 < 9.0
    code = morton(X[~solid].ravel(), Y[~solid].ravel(), Z[~solid].ravel())
    order = np.argsort(code)
    tid = first_id(L) + code[order]
    xs, ys, zs = X[~solid].ravel()[order], Y[~solid].ravel()[order], Z[~solid].ravel()[order]
    bid = np.zeros((tid.size, 26), dtype=np.int64)
    for s, c in enumerate(tio.Q_OFFSET):
        bid[:, s] = solid[(xs + c[0]) % n, (ys + c[1]) % n, (zs + c[2]) % n]
    hasb = bid.any(axis=1)
    prop = np.where(hasb, 2 | 8, 2).astype(np.int64)
    tio.dump_treelmesh(str(tmp_path), tid, prop, bc_labels=("sphere",), boundary_ID=bid[hasb])
    fd = tio.FileLevelDesc(tio.load_treelmesh(str(tmp_path)), QQ)
    assert fd.nFluid == int((~solid).sum()) and (fd.nghElems < 0).sum() > 0
    ng = np.zeros(QQ * fd.nSize, dtype=np.int32)
    oracle.lib().ora_construct_connectivity(oracle._i(ng), fd.nSize, fd.nElems, QQ,
                                            oracle._i(np.ascontiguousarray(fd.nghElems)),
                                            oracle._l(fd.property), fd.nFluid, fd.nFluid)
    assert np.array_equal(fd.neigh, ng)


def _sphere_mesh(tmp_path, L=4):
    from musubi_b200 import treelm_io as tio
    from musubi_b200.treelm_multilevel import first_id, morton
    n = 1 << L
    g = np.arange(n)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    c = (n - 1) / 2.0
    solid = (X - c) ** 2 + (Y - c) ** 2 + (Z - c) ** 2 < (n / 5.0) ** 2
    code = morton(X[~solid].ravel(), Y[~solid].ravel(), Z[~solid].ravel())
    order = np.argsort(code)
    tid = first_id(L) + code[order]
    xs, ys, zs = X[~solid].ravel()[order], Y[~solid].ravel()[order], Z[~solid].ravel()[order]
    bid = np.zeros((tid.size, 26), dtype=np.int64)
    for s, d in enumerate(tio.Q_OFFSET):
        bid[:, s] = solid[(xs + d[0]) % n, (ys + d[1]) % n, (zs + d[2]) % n]
    hasb = bid.any(axis=1)
    tio.dump_treelmesh(str(tmp_path), tid, np.where(hasb, 2 | 8, 2).astype(np.int64), length=1.0,
                       bc_labels=("sphere",), boundary_ID=bid[hasb])
    return tio.load_treelmesh(str(tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("QQ,relax", [(19, "trt"), (27, "mrt")])
def test_flow_past_sphere_from_mesh_file_matches_oracle(tmp_path, oracle, QQ, relax):
    """the whole path for a mesh that only exists as treelm files: read, build the descriptor,
    run driven by a body force on the device and in the oracle, compare bit for bit"""
    import musubi_b200 as mb
    from musubi_b200 import treelm_io as tio
    mb.mus_init(0, 1, 0)
    try:
        fd = tio.FileLevelDesc(_sphere_mesh(tmp_path), QQ)
        ref = oracle.Scheme(fd, relax, "fluid", omega=1.5, lambda_=0.25, omega_bulk=1.2)
        ref.init_equilibrium(1.0, np.array([0.02, 0.0, 0.01]))
        ref.set_force([2.0e-5, 0.0, -1.0e-5])
        ident = {"kind": "fluid", "relaxation": relax, "layout": "d3q%d" % QQ}
        sch = mb.Scheme(ident, fd, float(1.0 / (3.0 * ref.visc[0] + 0.5)), lambda_=0.25, omega_bulk=1.2)
        sch.set_force(fd.level, [2.0e-5, 0.0, -1.0e-5])
        sch.upload_state(fd.level, ref.state[ref.nNow], ref.state[ref.nNext])
        assert np.array_equal(sch.download_neigh(fd.level)[:QQ * fd.nSize], fd.neigh)
        ref.run(40)
        sch.do_computation(40)
        k = fd.nFluid * QQ
        assert np.array_equal(sch.download_state(fd.level)[:k], ref.state[ref.nNext][:k])
        m0 = fd.nFluid * 1.0
        assert abs(sch.reduce(fd.level)[0] - m0) < 1e-11 * m0
        sch.destroy()
    finally:
        mb.mus_finalize()
