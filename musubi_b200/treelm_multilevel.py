"""Synthetic multi-level treelm meshes (nested refined boxes in a periodic cube,
optionally with a solid cylinder inside the finest box) and their level
descriptors, on one rank.

This plays the role Seeder + treelm's tem_find_allElements + mus_construct play
for the Fortran host: it produces the per-level arrays BOTH the CPU oracle and
libmusb200 consume (total list, property, nghElems, neigh, ghost dependencies,
interpolation source lists / weights / least-square matrices).  Rules followed
(reference file:line):

  * total list per level [fluid | ghostFromCoarser | ghostFromFiner | halo], each
    block ascending treeID           tem_construction_module.f90:2358-2460
  * ghosts are created for the stencil neighbours of fluid elements that live on
    another level, in reqNesting = nNesting + 1 = 3 layers
                                     mus_construction_module.fpp:310-318
  * vertical dependencies: parent of a ghostFromCoarser (childNum, coord =
    0.25*childPosition), the locally present children of a ghostFromFiner
                                     tem_construction_module.f90:2894-2985
  * coarse->fine source selection, order fallback, weights, least-square matrices
                                     mus_interpolate_module.fpp:544-1011,
                                     mus_interpolate_header_module.f90:405-607,
                                     tem_matrix_module.fpp:161-425
  * state connectivity               mus_connectivity_module.fpp:73-179
All lists are 1-based as in Fortran.
"""
import numpy as np

from .cases import _CX27

PRP_FLUID, PRP_SOLID, PRP_HASBND = 1, 2, 3
NO_INTP, WEIGHTED_AVERAGE, LINEAR, QUADRATIC = -1, 0, 1, 2
ORDER_OF = {"average": 0, "weighted_average": 0, "linear": 1, "quadratic": 2}

# childPosition(childNum, xyz), tem_param_module.f90:170-173 (Z-curve order)
CHILD_POSITION = np.array([[-1, -1, -1], [1, -1, -1], [-1, 1, -1], [1, 1, -1],
                           [-1, -1, 1], [1, -1, 1], [-1, 1, 1], [1, 1, 1]], dtype=np.float64)


def _spread3(v):
    v = v.astype(np.uint64)
    v = (v | (v << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x1249249249249249)
    return v


def _compact3(v):
    v = v.astype(np.uint64) & np.uint64(0x1249249249249249)
    v = (v | (v >> np.uint64(2))) & np.uint64(0x10C30C30C30C30C3)
    v = (v | (v >> np.uint64(4))) & np.uint64(0x100F00F00F00F00F)
    v = (v | (v >> np.uint64(8))) & np.uint64(0x1F0000FF0000FF)
    v = (v | (v >> np.uint64(16))) & np.uint64(0x1F00000000FFFF)
    v = (v | (v >> np.uint64(32))) & np.uint64(0x1FFFFF)
    return v


def morton(x, y, z):
    return (_spread3(x) | (_spread3(y) << np.uint64(1)) | (_spread3(z) << np.uint64(2))).astype(np.int64)


def coords(m):
    m = np.asarray(m).astype(np.uint64)
    return (_compact3(m).astype(np.int64), _compact3(m >> np.uint64(1)).astype(np.int64),
            _compact3(m >> np.uint64(2)).astype(np.int64))


def first_id(level):
    return (8 ** level - 1) // 7


def stencil_tables(QQ):
    cx = np.vstack([_CX27[:QQ - 1], np.zeros((1, 3))]).astype(np.int64)
    inv = np.zeros(QQ, dtype=np.int64)
    for i in range(QQ):
        for j in range(QQ):
            if np.array_equal(cx[i], -cx[j]):
                inv[i] = j + 1
    return cx, inv


def weighted_avg_dirs(QQ):
    """init_cxDirWeightedAvg (mus_interpolate_header_module.f90:568-607): per child the
    parent + the face/edge(/corner) neighbours on the child's side, as 1-based directions."""
    cx, _ = stencil_tables(QQ)
    out = []
    for c in range(8):
        side = CHILD_POSITION[c]
        dirs = []
        for q in range(QQ):
            v = cx[q]
            if all(v[k] == 0 or v[k] == side[k] for k in range(3)):
                dirs.append(q + 1)
        out.append(set(dirs))
    return out


class MLLevel:
    """tem_levelDesc_type + pdf_data_type of one level (one rank)."""
    pass


def construct_connectivity(QQ, nghElems, property_, nFluid, haloOffset, nSize):
    """mus_construct_connectivity for AOS + PULL, vectorised (host-side product code;
    the oracle has its own C restatement and the tests compare the two)."""
    _, inv = stencil_tables(QQ)
    nElems = nghElems.shape[0]
    neigh = np.zeros(QQ * nSize, dtype=np.int32)
    e = np.arange(1, nElems + 1, dtype=np.int64)
    neigh[(QQ - 1) * nSize:(QQ - 1) * nSize + nElems] = (e - 1) * QQ + QQ
    solid_e = ((property_ >> PRP_SOLID) & 1).astype(bool)
    for d in range(1, QQ):
        nghDir = inv[d - 1]
        npos = nghElems[:, nghDir - 1].astype(np.int64)
        nprop = np.where(npos > 0, property_[np.maximum(npos, 1) - 1], 0)
        solid = solid_e | (((nprop >> PRP_SOLID) & 1).astype(bool))
        missing_nonghost = (npos <= 0) & ((e <= nFluid) | (e > haloOffset))
        sdir = np.where(missing_nonghost | solid, inv[d - 1], d)
        src = np.where((npos <= 0) | solid, e, npos)
        neigh[(d - 1) * nSize:(d - 1) * nSize + nElems] = (src - 1) * QQ + sdir
    return neigh


def _poly(order, c):
    """polyLinear_3D / polyQuadratic_3D (tem_matrix_module.fpp:464-523)"""
    if order == LINEAR:
        return np.array([1.0, c[0], c[1], c[2]])
    return np.array([1.0, c[0], c[1], c[2], c[0] ** 2, c[1] ** 2, c[2] ** 2,
                     c[0] * c[1], c[1] * c[2], c[2] * c[0]])


def invert_matrix(A):
    """invert_matrix (tem_matrix_module.fpp:610-660) = DGETRF + DGETRI as the reference's vendored
    LAPACK executes them for these 4 x 4 / 10 x 10 matrices (tem/external/lapack: n is below every
    block size, so DGETF2, DTRTI2 and the unblocked DGETRI loop run): partial pivoting on the
    first largest entry, column scaling by the reciprocal pivot, and an EXACT zero pivot as the
    only singularity criterion -- a merely ill-conditioned A^T A is inverted, as the reference
    does.  Returns (inverse, errCode)."""
    a = np.array(A, dtype=np.float64)
    n = a.shape[0]
    piv = np.zeros(n, dtype=np.int64)
    info = 0
    for j in range(n):                                  # DGETF2
        jp = j + int(np.argmax(np.abs(a[j:, j])))       # IDAMAX: first maximum
        piv[j] = jp
        if a[jp, j] != 0.0:
            if jp != j:
                a[[j, jp], :] = a[[jp, j], :]
            if j < n - 1:
                if abs(a[j, j]) >= np.finfo(np.float64).tiny:
                    a[j + 1:, j] = a[j + 1:, j] * (1.0 / a[j, j])
                else:
                    a[j + 1:, j] = a[j + 1:, j] / a[j, j]
        elif info == 0:
            info = j + 1
        if j < n - 1:                                   # DGER, alpha = -1
            for k in range(j + 1, n):
                if a[j, k] != 0.0:
                    a[j + 1:, k] = a[j + 1:, k] + a[j + 1:, j] * (-1.0 * a[j, k])
    if info != 0:
        return None, info
    for j in range(n):                                  # DTRTRI: singularity check, then DTRTI2
        if a[j, j] == 0.0:
            return None, j + 1
    for j in range(n):
        a[j, j] = 1.0 / a[j, j]
        ajj = -a[j, j]
        for k in range(j):                              # DTRMV upper, no transpose, non-unit
            if a[k, j] != 0.0:
                t = a[k, j]
                a[:k, j] = a[:k, j] + t * a[:k, k]
                a[k, j] = a[k, j] * a[k, k]
        a[:j, j] = ajj * a[:j, j]
    for j in range(n - 1, -1, -1):                      # DGETRI: inv(A) * L = inv(U)
        work = a[:, j].copy()
        a[j + 1:, j] = 0.0
        for k in range(j + 1, n):                       # DGEMV, alpha = -1, beta = 1
            if work[k] != 0.0:
                a[:, j] = a[:, j] + (-1.0 * work[k]) * a[:, k]
    for j in range(n - 2, -1, -1):
        if piv[j] != j:
            a[:, [j, piv[j]]] = a[:, [piv[j], j]]
    return a, 0


def new_intp(order_max):
    """the interpolation tables shared by all levels: intp%fillFinerFromMe(order)%intpMat_forLSF"""
    return {"order": order_max, "matrices": {LINEAR: [], QUADRATIC: []},
            "mat_ids": {LINEAR: {}, QUADRATIC: {}}, "mat_ok": {LINEAR: [], QUADRATIC: []}}


def append_intp_matrix_lsf(intp, order, dirs, cxr):
    """append_intpMatrixLSF (tem_matrix_module.fpp:161-244): one matrix ((A^T A)^-1 A^T, rows of A
    = the polynomial basis at the source offsets cxDir) per distinct set of source directions,
    identified by the bit set of the directions; returns its 0-based position, or -1 when A^T A
    is singular (the caller falls back to the next lower order)."""
    key = 0
    for d in dirs:
        key |= 1 << int(d)
    ids = intp["mat_ids"][order]
    if key in ids:
        i = ids[key]
        return (i if intp["mat_ok"][order][i] else -1)
    A = np.array([_poly(order, cxr[d - 1]) for d in dirs])
    inv, err = invert_matrix(A.T @ A)
    ok = err == 0
    M = inv @ A.T if ok else np.zeros((1, 1))
    ids[key] = len(intp["matrices"][order])
    intp["matrices"][order].append(M)
    intp["mat_ok"][order].append(bool(ok))
    return ids[key] if ok else -1


def build_multilevel(min_level, boxes, QQ=19, cylinder=None, intp_method="linear"):
    """nested refined boxes in a periodic cube.

    boxes: list, one entry per level above min_level: (lo, hi) in cells of THAT level's
           parent level (the box [lo,hi)^3 of parent cells is replaced by its children).
    cylinder: None or (cx, cy, r, zlo, zhi) in cells of the finest level: solid cylinder
              along z between zlo and zhi (its cells are absent from the mesh; neighbours
              see a wall, boundary id 1).  Keep it >= 8 finest cells away from the box faces.
    returns {level: MLLevel}, intp dict (matrices per order).
    """
    levels = list(range(min_level, min_level + len(boxes) + 1))
    max_level = levels[-1]
    # kind per cell: 0 none, 1 fluid, 2 ghostFromCoarser, 3 ghostFromFiner, 9 solid
    kind = {}
    region = {}  # region[l] = (lo, hi) box of level-l cells that belong to level >= l
    region[min_level] = (0, 1 << min_level)
    for i, (lo, hi) in enumerate(boxes):
        region[min_level + i + 1] = (2 * lo, 2 * hi)
        plo, phi = region[min_level + i]
        assert plo + 4 <= lo and hi + 4 <= phi, "refined box needs a margin of >= 4 parent cells"
    for l in levels:
        n = 1 << l
        k = np.zeros(n ** 3, dtype=np.int8)
        lo, hi = region[l]
        ax = np.arange(lo, hi, dtype=np.int64)
        X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
        inside = np.ones(X.shape, dtype=bool)
        if l < max_level:
            clo, chi = region[l + 1]
            clo, chi = clo // 2, chi // 2
            inside &= ~((X >= clo) & (X < chi) & (Y >= clo) & (Y < chi) & (Z >= clo) & (Z < chi))
        m = morton(X[inside].ravel(), Y[inside].ravel(), Z[inside].ravel())
        k[m] = 1
        if l == max_level and cylinder is not None:
            ccx, ccy, r, zlo, zhi = cylinder
            xs, ys, zs = coords(m)
            sol = (((xs + 0.5 - ccx) ** 2 + (ys + 0.5 - ccy) ** 2) < r * r) & (zs >= zlo) & (zs < zhi)
            k[m[sol]] = 9
        kind[l] = k

    return build_from_kinds(kind, QQ, intp_method)


def kinds_from_leaves(treeID):
    """dense cell kinds per level from the leaf list of ANY multi-level mesh (e.g. a treelm mesh
    file): 1 = leaf (fluid), 9 = solid: a cell that is no leaf, lies under no leaf and contains
    none -- the holes Seeder leaves for obstacles; their faces act as walls."""
    treeID = np.asarray(treeID, dtype=np.int64)
    lvl = np.zeros(treeID.size, dtype=np.int64)
    for l in range(1, 21):
        lvl[treeID >= first_id(l)] = l
    levels = list(range(int(lvl.min()), int(lvl.max()) + 1))
    kind, leaf = {}, {}
    for l in levels:
        k = np.zeros(8 ** l, dtype=np.int8)
        k[treeID[lvl == l] - first_id(l)] = 1
        kind[l], leaf[l] = k, k == 1
    under = {levels[0]: np.zeros(8 ** levels[0], dtype=bool)}      # some ancestor is a leaf
    for l in levels[1:]:
        under[l] = np.repeat(under[l - 1] | leaf[l - 1], 8)
    holds = {levels[-1]: np.zeros(8 ** levels[-1], dtype=bool)}    # some descendant is a leaf
    for l in reversed(levels[:-1]):
        holds[l] = (holds[l + 1] | leaf[l + 1]).reshape(-1, 8).any(axis=1)
    for l in levels:
        kind[l][~leaf[l] & ~under[l] & ~holds[l]] = 9
    return kind


def build_from_kinds(kind, QQ=19, intp_method="linear"):
    """level descriptors, ghost layers and vertical dependencies of a multi-level mesh given as
    dense cell kinds per level ({level: int8 array over the Morton codes}: 1 fluid leaf, 9 solid,
    0 elsewhere).  Solids are honoured on every level."""
    cx, inv = stencil_tables(QQ)
    QQN = QQ - 1
    levels = sorted(kind)
    min_level, max_level = levels[0], levels[-1]
    kind = {l: np.array(k, dtype=np.int8, copy=True) for l, k in kind.items()}
    pos = {}

    # ---- ghost layers: reqNesting = 3 rounds of stencil neighbours -----------------
    for l in levels:
        n = 1 << l
        k = kind[l]
        frontier = np.nonzero(k == 1)[0]
        for _ in range(3):
            if frontier.size == 0:
                break
            fx, fy, fz = coords(frontier)
            new = []
            for q in range(QQN):
                xn, yn, zn = (fx + cx[q, 0]) % n, (fy + cx[q, 1]) % n, (fz + cx[q, 2]) % n
                mn = morton(xn, yn, zn)
                empty = k[mn] == 0
                if not empty.any():
                    continue
                mn, xn, yn, zn = mn[empty], xn[empty], yn[empty], zn[empty]
                is_gfc = np.zeros(mn.size, dtype=bool)
                is_gff = np.zeros(mn.size, dtype=bool)
                if l > min_level:
                    is_gfc = kind[l - 1][morton(xn >> 1, yn >> 1, zn >> 1)] == 1
                if l < max_level:
                    base = morton(xn << 1, yn << 1, zn << 1)
                    for c in range(8):   # a ghostFromFiner needs at least one fluid child
                        is_gff |= kind[l + 1][base + c] == 1
                k[mn[is_gfc]] = 2
                k[mn[is_gff & ~is_gfc]] = 3
                new.append(mn[is_gfc | is_gff])
            frontier = np.unique(np.concatenate(new)) if new else np.zeros(0, dtype=np.int64)

    # ---- total lists + positions ----------------------------------------------------
    out = {}
    for l in levels:
        L = MLLevel()
        k = kind[l]
        fl, gc, gf = np.nonzero(k == 1)[0], np.nonzero(k == 2)[0], np.nonzero(k == 3)[0]
        codes = np.concatenate([fl, gc, gf])
        L.level, L.QQ = l, QQ
        L.nFluid, L.nGhostFromCoarser, L.nGhostFromFiner, L.nHalo = fl.size, gc.size, gf.size, 0
        L.nElems = codes.size
        L.nSize = (L.nElems + 3) // 4 * 4
        L.nSolve = L.nFluid + L.nGhostFromCoarser
        L.total = (first_id(l) + codes).astype(np.int64)
        L.codes = codes
        L.solid_ids = (first_id(l) + np.nonzero(k == 9)[0]).astype(np.int64)   # obstacle cells (walls)
        p = np.zeros(k.size, dtype=np.int32)
        p[codes] = np.arange(1, codes.size + 1, dtype=np.int32)
        pos[l] = p
        out[l] = L
    for l in levels:
        L = out[l]
        n = 1 << l
        x, y, z = coords(L.codes)
        ngh = np.zeros((L.nElems, QQN), dtype=np.int32)
        for q in range(QQN):
            mn = morton((x + cx[q, 0]) % n, (y + cx[q, 1]) % n, (z + cx[q, 2]) % n)
            pq = pos[l][mn]
            wall = (kind[l][mn] == 9)
            ngh[:, q] = np.where(wall, -1, pq)
        L.nghElems = ngh
        prop = np.zeros(L.nElems, dtype=np.int64)
        prop[:L.nFluid] = 1 << PRP_FLUID
        hasbnd = (ngh[:L.nFluid] < 0).any(axis=1)
        prop[:L.nFluid][hasbnd] |= 1 << PRP_HASBND
        L.property = prop
        L.neigh = construct_connectivity(QQ, ngh, prop, L.nFluid, L.nElems, L.nSize)
        L.bc_elemBuffer = np.zeros(0, dtype=np.int32)   # only 'wall' (do_nothing) boundaries
        L.bc, L.recv, L.send = [], [], []
        L.bary_unit = np.stack([(x + 0.5) / n, (y + 0.5) / n, (z + 0.5) / n], axis=1)

    # ---- vertical dependencies --------------------------------------------------------
    order_max = ORDER_OF[intp_method]
    wavg = weighted_avg_dirs(QQ)
    nmax_wavg = 7 if QQ == 19 else 8
    nmin = {LINEAR: 4, QUADRATIC: 10}
    intp = new_intp(order_max)
    cxr = cx.astype(np.float64)

    def lsf_matrix(order, dirs):
        return append_intp_matrix_lsf(intp, order, dirs, cxr)

    for l in levels:
        L = out[l]
        # ghostFromFiner <- children (tem_build_verticalDependencies second loop)
        L.depFromFiner = []
        if L.nGhostFromFiner:
            gcodes = L.codes[L.nFluid + L.nGhostFromCoarser:]
            for g in gcodes:
                ch = pos[l + 1][8 * g + np.arange(8)]
                ch = ch[ch > 0]
                assert ch.size > 0, "ghostFromFiner without any child present"
                L.depFromFiner.append(ch.astype(np.int32))
        L.intpFromFiner = np.arange(1, L.nGhostFromFiner + 1, dtype=np.int32)
        # ghostFromCoarser <- parent + its stencil neighbours
        L.depFromCoarser = []
        L.intpFromCoarser = {o: [] for o in range(0, order_max + 1)}
        if L.nGhostFromCoarser:
            C = out[l - 1]
            gcodes = L.codes[L.nFluid:L.nFluid + L.nGhostFromCoarser]
            for i, g in enumerate(gcodes):
                parentPos = int(pos[l - 1][g >> 3])
                assert parentPos > 0, "ghostFromCoarser without parent"
                childNum = int(g & 7) + 1
                coord = 0.25 * CHILD_POSITION[childNum - 1]
                srcs, dirs = [], []
                for iNeigh in range(1, QQ + 1):
                    p = parentPos if iNeigh == QQ else int(C.nghElems[parentPos - 1, iNeigh - 1])
                    if p > 0:
                        srcs.append(p)
                        dirs.append(iNeigh)
                # find_possIntpOrderAndUpdateMySources
                order = WEIGHTED_AVERAGE
                for o in range(order_max, LINEAR - 1, -1):
                    if len(srcs) >= nmin[o]:
                        order = o
                        break
                if order in (WEIGHTED_AVERAGE, LINEAR):
                    keep = [j for j, d in enumerate(dirs) if d in wavg[childNum - 1]]
                    if len(keep) == nmax_wavg:
                        srcs, dirs = [srcs[j] for j in keep], [dirs[j] for j in keep]
                dep = dict(childNum=childNum, coord=coord, posInMat=-1, weights=None)
                if order == QUADRATIC:
                    mi = lsf_matrix(QUADRATIC, dirs)
                    if mi >= 0:
                        dep["posInMat"] = mi
                    else:
                        order = LINEAR
                if order == LINEAR:
                    mi = lsf_matrix(LINEAR, dirs)
                    if mi >= 0:
                        dep["posInMat"] = mi
                    else:
                        order = WEIGHTED_AVERAGE
                if order == WEIGHTED_AVERAGE:
                    # compute_weight, 'linear_distance'
                    dist = np.abs(cxr[np.array(dirs) - 1] - coord[None, :])
                    w = (1.0 - dist[:, 0]) * (1.0 - dist[:, 1]) * (1.0 - dist[:, 2])
                    dep["weights"] = w / w.sum()
                dep["order"] = order
                dep["sources"] = np.array(srcs, dtype=np.int32)
                dep["dirs"] = np.array(dirs, dtype=np.int32)
                L.depFromCoarser.append(dep)
                L.intpFromCoarser[order].append(i + 1)
        for o in L.intpFromCoarser:
            L.intpFromCoarser[o] = np.array(L.intpFromCoarser[o], dtype=np.int32)
    return out, intp


def intp_tables(L, intp, direction, order=None):
    """flattens the dependency lists of one level into the arrays musb200_intp_register
    (and the oracle) take: targets (positions in the total list), CSR sources, weights,
    posInMat, concatenated row-major matrices, child coordinates."""
    if direction == "fromFiner":
        off0 = L.nFluid + L.nGhostFromCoarser
        tg = (L.intpFromFiner + off0).astype(np.int32)
        srcOff = np.zeros(len(tg) + 1, dtype=np.int32)
        src = []
        for i, t in enumerate(L.intpFromFiner):
            s = L.depFromFiner[t - 1]
            src.append(s)
            srcOff[i + 1] = srcOff[i] + len(s)
        src = np.concatenate(src).astype(np.int32) if src else np.zeros(0, dtype=np.int32)
        return dict(targets=tg, srcOffset=srcOff, srcPos=src, weights=np.zeros(0), posInMat=np.zeros(0, np.int32),
                    matOffset=np.zeros(1, np.int32), matrices=np.zeros(0), coord=np.zeros((0, 3)), nMat=0)
    lst = L.intpFromCoarser[order]
    tg = (lst + L.nFluid).astype(np.int32)
    srcOff = np.zeros(len(tg) + 1, dtype=np.int32)
    src, wts, pim, crd = [], [], [], []
    for i, t in enumerate(lst):
        dep = L.depFromCoarser[t - 1]
        src.append(dep["sources"])
        srcOff[i + 1] = srcOff[i] + len(dep["sources"])
        wts.append(dep["weights"] if dep["weights"] is not None else np.zeros(len(dep["sources"])))
        pim.append(dep["posInMat"])
        crd.append(dep["coord"])
    mats = intp["matrices"].get(order, []) if order > 0 else []
    matOff = np.zeros(len(mats) + 1, dtype=np.int32)
    for i, M in enumerate(mats):
        matOff[i + 1] = matOff[i] + M.size
    return dict(targets=tg, srcOffset=srcOff,
                srcPos=np.concatenate(src).astype(np.int32) if src else np.zeros(0, np.int32),
                weights=np.concatenate(wts) if wts else np.zeros(0),
                posInMat=np.array(pim, dtype=np.int32), matOffset=matOff,
                matrices=np.concatenate([M.ravel() for M in mats]) if mats else np.zeros(0),
                coord=np.array(crd).reshape(-1, 3) if crd else np.zeros((0, 3)), nMat=len(mats))


# ------------------------------------------------------------------------------------------------
# several ranks: the global multi-level mesh cut along the space-filling curve
# ------------------------------------------------------------------------------------------------
def _sfc_keys(lv):
    """position of every fluid element on the global space-filling curve: (level, index, key)"""
    maxL = max(lv)
    lvl, idx, key = [], [], []
    for l, L in lv.items():
        c = L.codes[:L.nFluid].astype(np.int64)
        lvl.append(np.full(c.size, l, dtype=np.int64))
        idx.append(np.arange(c.size, dtype=np.int64))
        key.append(c << (3 * (maxL - l)))
    lvl, idx, key = np.concatenate(lvl), np.concatenate(idx), np.concatenate(key)
    o = np.argsort(key, kind="stable")
    return lvl[o], idx[o]


def _grow(mask, ngh, hops):
    """elements within `hops` stencil neighbours of the masked ones (global positions)"""
    out = mask.copy()
    front = mask
    for _ in range(hops):
        rows = ngh[front]
        nxt = np.zeros_like(out)
        p = rows[rows > 0] - 1
        nxt[p] = True
        nxt &= ~out
        out |= nxt
        front = nxt
    return out


def sparta_split(weights, nParts):
    """tem_balance_sparta (tem_sparta_module.f90:112-226) as ONE rank holding the whole weight
    list sees it: the splitter of part k sits after the element whose weight prefix sum is
    closest to (k+1) * W / nParts -- binary search on the prefix sums, then the comparison with
    the neighbouring elements (:182-196).  weights: per element along the space-filling curve.
    Returns the element count of every part (send_count)."""
    w = np.asarray(weights, dtype=np.float64)
    n = w.size
    presum = np.cumsum(w)                       # sequential sum, as the reference's loop
    w_opt = presum[-1] / float(nParts)
    upper = presum[-1]
    count = np.zeros(nParts, dtype=np.int64)
    left_off = 1                                # 1-based, as in the reference
    for iProc in range(nParts):
        lb, ub = left_off, n
        opt_split = (iProc + 1) * w_opt
        if not iProc * w_opt < upper:
            continue
        while True:
            mid = (lb + ub) // 2
            wsplit = presum[mid - 1]
            if abs(wsplit - opt_split) <= np.finfo(np.float64).eps * max(abs(wsplit), abs(opt_split)):
                break                           # .feq.
            if wsplit < opt_split:
                lb = mid
            else:
                ub = mid
            if lb >= ub - 1:
                break
        if abs(wsplit - opt_split) > abs(wsplit - opt_split - w[mid - 1]):
            mid -= 1
        elif mid + 1 <= n:
            if abs(wsplit - opt_split) > abs(wsplit - opt_split + w[mid]):
                mid += 1
        elif opt_split > upper:
            mid = n
        if iProc == nParts - 1:
            mid = n                             # the last part ends with the last element
        count[iProc] = mid - left_off + 1
        left_off = mid + 1
    return count


def level_weights(lv):
    """per-leaf cost along the global space-filling curve: level steps per coarse cycle,
    2^(level - minLevel) -- what mus_getWeights (mus_weights_module.f90:62-143: measured level
    time / nFluid of the level) converges to for a bandwidth-bound sweep of uniform cost per
    element update"""
    glvl, _ = _sfc_keys(lv)
    return 2.0 ** (glvl - min(lv)).astype(np.float64)


def partition_multilevel(lv, nranks, weights=None):
    """cut a single-rank multi-level mesh (build_multilevel) into `nranks` parts the way treelm
    does -- equal contiguous ranges of the global space-filling curve over ALL levels
    (treelmesh_module.f90:1276-1296), or, with weights (one per leaf in curve order, e.g.
    level_weights(lv)), the ranges tem_balance_sparta cuts (mus_dynLoadBal_module.f90:513-577)
    -- and build every rank's level descriptors.

    Per rank and level the total list is [own fluid | ghostFromCoarser | ghostFromFiner | halo]:
      * ghosts are LOCAL and recomputed by interpolation on every rank that needs them: the
        ghostFromCoarser elements within two stencil hops of an own fluid element (nNesting = 2
        sub-steps between interpolations), the ghostFromFiner ones within one hop, plus the ghosts
        that serve as interpolation sources of those;
      * halos are the remote FLUID elements that an own or ghost element pulls from or
        interpolates from; they arrive through the level's halo buffer with all QQ links (and
        their auxField entries), element-major, after every level step.
    Only fluid elements travel; interpolation results never do.  The fluid elements of every
    rank evolve bit-identically to the single-rank run.
    returns [ {level: MLLevel} for each rank ]"""
    levels = sorted(lv)
    QQ = lv[levels[0]].QQ
    glvl, gidx = _sfc_keys(lv)
    N = glvl.size
    if weights is None:
        base, rem = divmod(N, nranks)
        cnt = np.array([base + (1 if r < rem else 0) for r in range(nranks)], dtype=np.int64)
    else:
        if len(weights) != N:
            raise ValueError("partition_multilevel: one weight per leaf (%d), got %d" % (N, len(weights)))
        cnt = sparta_split(weights, nranks)
        if cnt.sum() != N or np.any(cnt <= 0):
            raise ValueError("weighted partition leaves a rank without elements: %r" % (cnt,))
    off = np.concatenate([[0], np.cumsum(cnt)])
    owner = {l: np.full(lv[l].nFluid, -1, dtype=np.int64) for l in levels}
    for r in range(nranks):
        sl = slice(int(off[r]), int(off[r + 1]))
        for l in levels:
            sel = glvl[sl] == l
            owner[l][gidx[sl][sel]] = r

    # global source tables as arrays of global positions (1-based) per ghost
    def gfc_sources(l):
        L = lv[l]
        return [d["sources"] for d in L.depFromCoarser]

    ranks = []
    keep_all = []
    for r in range(nranks):
        need = {}
        for l in levels:
            L = lv[l]
            own = np.zeros(L.nElems, dtype=bool)
            own[:L.nFluid] = owner[l] == r
            kindv = np.zeros(L.nElems, dtype=np.int8)
            kindv[:L.nFluid] = 1
            kindv[L.nFluid:L.nFluid + L.nGhostFromCoarser] = 2
            kindv[L.nFluid + L.nGhostFromCoarser:] = 3
            near2 = _grow(own, L.nghElems, 2)
            near1 = _grow(own, L.nghElems, 1)
            keep = own | (near2 & (kindv == 2)) | (near1 & (kindv == 3))
            need[l] = dict(own=own, kind=kindv, keep=keep)
        # interpolation sources, two passes (a ghost source pulls in its own sources)
        for _ in range(2):
            for l in reversed(levels):
                L, nd = lv[l], need[l]
                if L.nGhostFromCoarser:
                    g0 = L.nFluid
                    for i in np.nonzero(nd["keep"][g0:g0 + L.nGhostFromCoarser])[0]:
                        need[l - 1]["keep"][L.depFromCoarser[i]["sources"] - 1] = True
            for l in levels:
                L, nd = lv[l], need[l]
                if L.nGhostFromFiner:
                    g0 = L.nFluid + L.nGhostFromCoarser
                    for i in np.nonzero(nd["keep"][g0:])[0]:
                        need[l + 1]["keep"][L.depFromFiner[i] - 1] = True
        # everything a solved element (own fluid, kept ghostFromCoarser) pulls from must be present
        for l in levels:
            L, nd = lv[l], need[l]
            solved = nd["keep"] & ((nd["own"]) | (nd["kind"] == 2))
            rows = L.nghElems[solved]
            p = rows[rows > 0] - 1
            fluid_nb = p[nd["kind"][p] == 1]
            nd["keep"][fluid_nb] = True          # remote fluid -> halo; ghosts beyond two hops are
            # not needed (their values never reach an own fluid element before re-interpolation)
        keep_all.append(need)

    for r in range(nranks):
        need = keep_all[r]
        out = {}
        maps = {}
        for l in levels:
            L, nd = lv[l], need[l]
            keep, kindv, own = nd["keep"], nd["kind"], nd["own"]
            fl = np.nonzero(own)[0]
            gc = np.nonzero(keep & (kindv == 2))[0]
            gf = np.nonzero(keep & (kindv == 3))[0]
            ha = np.nonzero(keep & (kindv == 1) & ~own)[0]
            sel = np.concatenate([fl, gc, gf, ha])          # each block ascending treeID already
            g2l = np.zeros(L.nElems + 1, dtype=np.int32)    # global position -> local position
            g2l[sel + 1] = np.arange(1, sel.size + 1, dtype=np.int32)
            maps[l] = (sel, g2l, fl, gc, gf, ha)
            M = MLLevel()
            M.level, M.QQ = l, QQ
            M.nFluid, M.nGhostFromCoarser, M.nGhostFromFiner, M.nHalo = fl.size, gc.size, gf.size, ha.size
            M.nElems = sel.size
            M.nSize = (M.nElems + 3) // 4 * 4
            M.nSolve = M.nFluid + M.nGhostFromCoarser
            M.total = L.total[sel]
            M.codes = L.codes[sel]
            M.globalPos = (sel + 1).astype(np.int64)
            ng = L.nghElems[sel]
            M.nghElems = np.where(ng > 0, g2l[np.maximum(ng, 0)], ng).astype(np.int32)
            M.property = L.property[sel].copy()
            haloOffset = M.nFluid + M.nGhostFromCoarser + M.nGhostFromFiner
            M.neigh = construct_connectivity(QQ, M.nghElems, M.property, M.nFluid, haloOffset, M.nSize)
            M.bc_elemBuffer = np.zeros(0, dtype=np.int32)
            M.bc = []
            M.bary_unit = L.bary_unit[sel]
            M.haloOwner = owner[l][ha]
            out[l] = M
        # dependencies re-expressed in local positions
        for l in levels:
            L, M = lv[l], out[l]
            sel, g2l, fl, gc, gf, ha = maps[l]
            M.depFromFiner, M.depFromCoarser = [], []
            M.intpFromFiner = np.arange(1, M.nGhostFromFiner + 1, dtype=np.int32)
            M.intpFromCoarser = {o: [] for o in L.intpFromCoarser}
            g0 = L.nFluid + L.nGhostFromCoarser
            for g in gf:
                src = maps[l + 1][1][L.depFromFiner[g - g0]]
                assert np.all(src > 0), "a child of a kept ghostFromFiner is missing on this rank"
                M.depFromFiner.append(src.astype(np.int32))
            for i, g in enumerate(gc):
                d = dict(L.depFromCoarser[g - L.nFluid])
                src = maps[l - 1][1][d["sources"]]
                assert np.all(src > 0), "a source of a kept ghostFromCoarser is missing on this rank"
                d["sources"] = src.astype(np.int32)
                M.depFromCoarser.append(d)
                M.intpFromCoarser[d["order"]].append(i + 1)
            for o in M.intpFromCoarser:
                M.intpFromCoarser[o] = np.array(M.intpFromCoarser[o], dtype=np.int32)
        ranks.append(out)

    # halo exchange lists: all QQ links of every halo element, element-major (the order of the
    # receiver's halo block = ascending treeID); the sender's list mirrors it
    for r in range(nranks):
        for l in levels:
            M = ranks[r][l]
            M.recv, M.send = [], []
            h0 = M.nFluid + M.nGhostFromCoarser + M.nGhostFromFiner
            for p in np.unique(M.haloOwner):
                hs = np.nonzero(M.haloOwner == p)[0]
                epos = (h0 + hs + 1).astype(np.int64)
                pos = ((epos[:, None] - 1) * QQ + np.arange(1, QQ + 1)[None, :]).ravel().astype(np.int32)
                M.recv.append(dict(proc=int(p), pos=pos, elemPos=epos.astype(np.int32),
                                   globalPos=M.globalPos[h0 + hs]))
    for r in range(nranks):
        for l in levels:
            M = ranks[r][l]
            g2l = np.zeros(lv[l].nElems + 1, dtype=np.int64)
            g2l[M.globalPos] = np.arange(1, M.nElems + 1)
            for p in range(nranks):
                if p == r:
                    continue
                for rc in ranks[p][l].recv:
                    if rc["proc"] != r:
                        continue
                    epos = g2l[rc["globalPos"]]
                    assert np.all((epos >= 1) & (epos <= M.nFluid)), "a halo is not fluid on its owner"
                    pos = ((epos[:, None] - 1) * QQ + np.arange(1, QQ + 1)[None, :]).ravel().astype(np.int32)
                    M.send.append(dict(proc=int(p), pos=pos, elemPos=epos.astype(np.int32)))
            M.send.sort(key=lambda c: c["proc"])
    return ranks


# ------------------------------------------------------------------------------------------------
# reference-style ghost buffers (sendBufferFromCoarser / FromFiner) from the partition above
# ------------------------------------------------------------------------------------------------
def _filter_table(t, keep):
    """an interpolation table (intp_tables) restricted to the targets with keep[i]"""
    keep = np.asarray(keep, dtype=bool)
    off = np.asarray(t["srcOffset"], dtype=np.int64)
    n = np.diff(off)
    src_keep = np.repeat(keep, n)
    new_off = np.concatenate([[0], np.cumsum(n[keep])]).astype(np.int32)
    out = dict(t)
    out["targets"] = np.asarray(t["targets"])[keep]
    out["srcOffset"] = new_off
    out["srcPos"] = np.asarray(t["srcPos"])[src_keep]
    if len(t["weights"]):
        out["weights"] = np.asarray(t["weights"])[src_keep]
    if len(t["posInMat"]):
        out["posInMat"] = np.asarray(t["posInMat"])[keep]
    if len(t["coord"]):
        out["coord"] = np.asarray(t["coord"]).reshape(-1, 3)[keep]
    return out


def delegate_shared_ghosts(ranks, rtables, lv):
    """Turns the locally-recomputed ghosts of partition_multilevel into the reference's form of
    the same run: a ghost element that several ranks hold is interpolated by ONE of them and
    travels to the others through the level's sendBufferFromCoarser / sendBufferFromFiner
    (tem_construction_module.f90: the six buffer kinds of a level; exchanged by
    do_intpCoarserAndExchange / do_intpFinerAndExchange, mus_control_module.f90:861-1051, and for
    the auxField of ghostFromFiner elements by mus_intpAuxFieldCoarserAndExchange,
    mus_auxField_module.f90:404-444).  The provider of a ghost is chosen by treeID (so that
    messages flow both ways) among the holders that can stand in for the others: a
    ghostFromCoarser element is swept like a fluid element and shipped after every level step
    (recvBufferFromCoarser, mus_control_module.f90:434-465), so its provider must hold the
    element's complete neighbourhood -- as many neighbours as the single-domain mesh `lv` gives
    it; a ghost no holder can provide stays locally interpolated on every rank.

    returns (tables, comm): tables = rtables with the delegated targets removed on the receiving
    ranks; comm[rank][level][kind]['send' | 'recv'] = [dict(proc, pos, elemPos)], kind in
    ('fromCoarser', 'fromFiner'), pos = all QQ state positions of every element, element-major."""
    nranks = len(ranks)
    levels = sorted(ranks[0])
    QQ = ranks[0][levels[0]].QQ
    tables = [{k: dict(v) for k, v in rt.items()} for rt in rtables]
    comm = [{l: {k: {"send": {}, "recv": {}} for k in ("fromCoarser", "fromFiner")} for l in levels}
            for _ in range(nranks)]
    for l in levels:
        for kind in ("fromCoarser", "fromFiner"):
            holders = {}                                   # treeID -> [(rank, 1-based position)]
            for r in range(nranks):
                M = ranks[r][l]
                lo = M.nFluid if kind == "fromCoarser" else M.nFluid + M.nGhostFromCoarser
                hi = lo + (M.nGhostFromCoarser if kind == "fromCoarser" else M.nGhostFromFiner)
                for p in range(lo, hi):
                    holders.setdefault(int(M.total[p]), []).append((r, p + 1))
            drop = [set() for _ in range(nranks)]          # positions a rank no longer interpolates
            for tid in sorted(holders):
                h = holders[tid]
                if len(h) < 2:
                    continue
                able = h
                if kind == "fromCoarser":
                    able = []
                    for r, pos in h:
                        M = ranks[r][l]
                        full = int((lv[l].nghElems[M.globalPos[pos - 1] - 1] > 0).sum())
                        if int((M.nghElems[pos - 1] > 0).sum()) == full:
                            able.append((r, pos))
                    if not able:
                        continue
                prov, ppos = able[tid % len(able)]
                for r, pos in h:
                    if r == prov:
                        continue
                    drop[r].add(pos)
                    comm[prov][l][kind]["send"].setdefault(r, []).append(ppos)
                    comm[r][l][kind]["recv"].setdefault(prov, []).append(pos)
            for r in range(nranks):
                if not drop[r]:
                    continue
                gone = np.array(sorted(drop[r]), dtype=np.int64)
                keys = [(l, "fromFiner")] if kind == "fromFiner" else \
                    [k for k in tables[r] if k[0] == l and isinstance(k[1], tuple) and k[1][0] == "fromCoarser"]
                for k in keys:
                    if k in tables[r]:
                        t = tables[r][k]
                        tables[r][k] = _filter_table(t, ~np.isin(np.asarray(t["targets"], dtype=np.int64), gone))
    d = np.arange(1, QQ + 1, dtype=np.int64)
    out = [{l: {k: {"send": [], "recv": []} for k in ("fromCoarser", "fromFiner")} for l in levels}
           for _ in range(nranks)]
    for r in range(nranks):
        for l in levels:
            for kind in ("fromCoarser", "fromFiner"):
                for way in ("send", "recv"):
                    for p in sorted(comm[r][l][kind][way]):
                        e = np.array(comm[r][l][kind][way][p], dtype=np.int64)
                        pos = ((e[:, None] - 1) * QQ + d[None, :]).ravel().astype(np.int32)
                        out[r][l][kind][way].append(dict(proc=int(p), pos=pos, elemPos=e.astype(np.int32)))
    return tables, out
