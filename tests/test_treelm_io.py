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


@pytest.mark.parametrize("kind,QQ,nranks", [("cavity", 19, 1), ("cavity", 27, 1), ("channel", 19, 1),
                                            ("channel", 27, 1), ("cavity", 19, 2), ("channel", 19, 4)])
def test_boundary_lists_from_file_equal_box_generator(tmp_path, kind, QQ, nranks):
    """velocity and pressure boundaries of a mesh file: bc_elemBuffer, links, outPos, posInBuffer,
    iDir, normal index, neighbour positions ... bit-identical to the C++ box generator's"""
    import musubi_b200 as mb
    from musubi_b200 import treelm_io as tio
    whole = mb.LevelDesc(4, QQ, kind)
    m = tio.mesh_from_level_desc(whole)
    tio.dump_treelmesh(str(tmp_path), m["treeID"], m["property"], bc_labels=m["bc_labels"],
                       boundary_ID=m["boundary_ID"])
    mesh = tio.load_treelmesh(str(tmp_path))
    binding = {"lid": "velocity_bounceback", "inlet": "velocity_bounceback", "outlet": "pressure"}
    for r in range(nranks):
        fd = tio.FileLevelDesc(mesh, QQ, r, nranks, bc_kind=binding)
        ld = mb.LevelDesc(4, QQ, kind, r, nranks)
        assert np.array_equal(fd.bc_elemBuffer, ld.bc_elemBuffer)
        assert [b["id"] for b in fd.bc] == [b["id"] for b in ld.bc]
        for a, b in zip(fd.bc, ld.bc):
            assert a["kind"] == b["kind"] and a["label"] == b["label"]
            for key in ("elems", "links", "outPos", "posInBuffer", "iDir", "normalInd", "posInBcElemBuf",
                        "neighPos", "iElemOfLink", "statePos"):
                assert np.array_equal(a[key], b[key]), (r, a["label"], key)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,outlet", [("cavity", None), ("channel", "pressure_expol"),
                                         ("channel", "pressure_antibounceback")])
def test_boundaries_from_mesh_file_run_on_device(tmp_path, oracle, kind, outlet):
    """lid-driven cavity / channel with velocity inlet and pressure outlet, the mesh and its
    boundaries read from treelm files: device against the oracle on the oracle's own descriptor"""
    import musubi_b200 as mb
    from musubi_b200 import cases
    from musubi_b200 import treelm_io as tio
    from musubi_b200._lib import check, lib
    mb.mus_init(0, 1, 0)
    try:
        QQ, level = 19, 4
        whole = mb.LevelDesc(level, QQ, kind)
        m = tio.mesh_from_level_desc(whole)
        tio.dump_treelmesh(str(tmp_path), m["treeID"], m["property"], bc_labels=m["bc_labels"],
                           boundary_ID=m["boundary_ID"])
        binding = {"lid": "velocity_bounceback", "inlet": "velocity_bounceback", "outlet": outlet}
        fd = tio.FileLevelDesc(tio.load_treelmesh(str(tmp_path)), QQ, bc_kind=binding)
        old = oracle.build_level_desc(level, QQ, kind)
        ref = oracle.Scheme(old, "trt", "fluid", omega=1.7, lambda_=3.0 / 16.0)
        ref.init_equilibrium(1.0, np.zeros(3))
        v = cases.lid_values(whole, (0.04, 0.01, 0.0))
        ref.bc_vel[2] = v
        sch = mb.Scheme({"kind": "fluid", "relaxation": "trt", "layout": "d3q19"}, fd,
                        float(1.0 / (3.0 * ref.visc[0] + 0.5)), lambda_=3.0 / 16.0)
        sch.upload_state(level, ref.state[ref.nNow], ref.state[ref.nNext])
        sch.set_bc_values(level, 2, v)
        if kind == "channel":
            nOut = len(fd.bc[2]["elems"])
            ref.bc_kind[3] = outlet
            ref.bc_rho[3] = np.full(nOut, 1.0)
            sch.set_bc_values(level, 3, ref.bc_rho[3])
            check(lib.musb200_aux_upload(level, ref.aux.ctypes.data))
        ref.run(40)
        sch.do_computation(40)
        k = fd.nFluid * QQ
        assert np.array_equal(sch.download_state(level)[:k], ref.state[ref.nNext][:k])
        sch.destroy()
    finally:
        mb.mus_finalize()


@pytest.mark.parametrize("boxes,QQ,method", [([(5, 11)], 19, "linear"), ([(4, 12), (12, 20)], 19, "linear"),
                                             ([(5, 11)], 27, "quadratic")])
def test_multilevel_mesh_file_round_trip(tmp_path, oracle, boxes, QQ, method):
    """a multi-level mesh written as treelm files (leaves of all levels in space-filling-curve
    order) and read back: the descriptors, ghost lists and dependencies rebuilt from the leaf list
    alone equal those of the parametric generator"""
    from musubi_b200 import treelm_io as tio
    from musubi_b200 import treelm_multilevel as tm
    lv, intp = tm.build_multilevel(4, boxes, QQ=QQ, intp_method=method)
    tid, lp = oracle.global_tree(lv)
    prop = np.full(tid.size, 2, dtype=np.int64)
    tio.dump_treelmesh(str(tmp_path), tid, prop, length=1.0)
    m = tio.load_treelmesh(str(tmp_path))
    assert (m["minLevel"], m["maxLevel"]) == (4, 4 + len(boxes)) and np.array_equal(m["treeID"], tid)
    lv2, intp2 = tm.build_from_kinds(tm.kinds_from_leaves(m["treeID"]), QQ=QQ, intp_method=method)
    assert sorted(lv2) == sorted(lv)
    for l in lv:
        A, B = lv[l], lv2[l]
        for key in ("nFluid", "nGhostFromCoarser", "nGhostFromFiner", "nSize"):
            assert getattr(A, key) == getattr(B, key), (l, key)
        for key in ("total", "property", "nghElems", "neigh"):
            assert np.array_equal(getattr(A, key), getattr(B, key)), (l, key)
        assert len(A.depFromFiner) == len(B.depFromFiner)
        assert all(np.array_equal(a, b) for a, b in zip(A.depFromFiner, B.depFromFiner))
        for a, b in zip(A.depFromCoarser, B.depFromCoarser):
            assert a["order"] == b["order"] and a["posInMat"] == b["posInMat"]
            assert np.array_equal(a["sources"], b["sources"])


def test_holes_in_a_multilevel_leaf_list_become_walls():
    """cells that are neither leaves, nor under a leaf, nor above one are solid obstacles"""
    from musubi_b200 import treelm_multilevel as tm
    lv, _ = tm.build_multilevel(5, [(10, 22)], QQ=19, cylinder=(32.0, 32.0, 3.0, 29, 35))
    tid = np.concatenate([L.total[:L.nFluid] for L in lv.values()])
    kinds = tm.kinds_from_leaves(tid)
    fine = kinds[6]
    x, y, z = tm.coords(np.nonzero(fine == 9)[0])
    assert x.size > 0
    assert np.all(((x + 0.5 - 32.0) ** 2 + (y + 0.5 - 32.0) ** 2 < 9.0) & (z >= 29) & (z < 35))
    lv2, _ = tm.build_from_kinds(kinds, QQ=19)
    for l in lv:
        assert lv2[l].nFluid == lv[l].nFluid
        nf = lv[l].nFluid
        assert np.array_equal(lv2[l].neigh.reshape(19, -1)[:, :nf], lv[l].neigh.reshape(19, -1)[:, :nf])


def test_weights_file_round_trip_and_sparta_cut(tmp_path):
    """tem_dump_weights layout (one double per element, ranks write at their element offset) and
    the cut tem_balance_sparta makes from the file"""
    from musubi_b200 import treelm_io as tio
    from musubi_b200 import treelm_multilevel as tm
    lv, _ = tm.build_multilevel(4, [(5, 11)], QQ=19)
    w = tm.level_weights(lv)
    f = str(tmp_path / "sim_weight_t0.000E+00.lsb")
    h = w.size // 3
    tio.dump_weights(f, w[h:], elem_offset=h, nElems_global=w.size)      # rank 1 first
    tio.dump_weights(f, w[:h], elem_offset=0, nElems_global=w.size)
    assert np.fromfile(f).tobytes() == w.tobytes()
    assert np.array_equal(tio.load_weights(f), w)
    assert np.array_equal(tio.load_weights(f, h, 10), w[h:h + 10])
    cnt = tm.sparta_split(tio.load_weights(f), 3)
    ranks = tm.partition_multilevel(lv, 3, weights=tio.load_weights(f))
    assert [sum(rl[l].nFluid for l in lv) for rl in ranks] == list(cnt)
    import pytest
    with pytest.raises(ValueError):
        tio.load_weights(f, w.size - 2, 5)
    with pytest.raises(ValueError):
        tm.partition_multilevel(lv, 3, weights=w[:-1])
