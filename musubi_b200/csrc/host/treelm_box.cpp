// treelm_box.cpp -- host-side generator of the per-rank level descriptor of a
// single-level box mesh (libmusb200_mesh.so).  It stands in for the Fortran
// host (treelm + mus_construct) when the library is driven without Musubi:
// bench.py, the tests and the C test driver feed its arrays through the very
// same C ABI (include/musb200.h) the Fortran shim uses.
//
// What it reproduces (reference file:line):
//   Morton treeIDs, periodic wrap     tem/source/tem_topology_module.f90:88-108, 590-638
//   predefined cube, SFC partition    tem/source/treelmesh_module.f90:1224-1318 (:1276-1296)
//   total list [fluid | halo]         tem/source/tem_construction_module.f90:2358-2460
//   neigh (pull list, bounce-back)    mus/source/mus_connectivity_module.fpp:73-179
//   reduced halo link lists           mus/source/mus_construction_module.fpp:1162-1356
//   BC element / link lists           mus/source/mus_construction_module.fpp:2203-2376,
//                                     mus/source/bc/mus_bc_header_module.fpp:1702-1739, 1876-1967
// All lists are produced 1-based, exactly as the Fortran arrays hold them.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

inline uint64_t spread3(uint64_t v) {
  v &= 0x1FFFFF;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
inline uint64_t compact3(uint64_t v) {
  v &= 0x1249249249249249ull;
  v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
  v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
  v = (v | (v >> 8)) & 0x1F0000FF0000FFull;
  v = (v | (v >> 16)) & 0x1F00000000FFFFull;
  v = (v | (v >> 32)) & 0x1FFFFF;
  return v;
}
inline int64_t mortonOf(int x, int y, int z) {
  return (int64_t)(spread3((uint64_t)x) | (spread3((uint64_t)y) << 1) | (spread3((uint64_t)z) << 2));
}

const int kCx[26][3] = {
    {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
    {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1},
    {-1, 0, -1}, {1, 0, -1}, {-1, 0, 1}, {1, 0, 1},
    {-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0},
    {-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1},
    {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};

struct Bc {
  int id, kind;  // kind: 0 wall, 1 velocity_bounceback, 2 pressure (expol / anti-bounce-back)
  std::vector<int32_t> elems, links, outPos, posInBuffer, iDir;
  // per element: discretised inward normal index, slot in bc_elemBuffer, the two neighbours
  // along the normal; per link: element counter and outletExpol%statePos
  std::vector<int32_t> normalInd, posInBcElemBuf, neighPos, iElemOfLink, statePos;
};

struct Comm {
  std::vector<int32_t> proc, nVals, pos, elemCount, elemPos;
};

struct Box {
  int level, QQ, kind, rank, nranks;
  int octants = 8;       // the domain = the first 1, 2, 4 or 8 octants of the universe cube
  int ext[3] = {0, 0, 0};  // its extent in cells: x doubles first, then y, then z (Morton order)
  int64_t lo, hi, firstId;
  int nFluid, nHalo, nElems, nSize;
  std::vector<int64_t> total, property;
  std::vector<int32_t> ngh;    // [nElems][QQN]
  std::vector<int32_t> neigh;  // [QQ][nSize]
  std::vector<int32_t> bcElemBuffer;
  std::vector<Bc> bcs;
  Comm recv, send;
  std::vector<int64_t> partEnd;  // exclusive end (morton index) per rank
  int inv[27];

  int owner(int64_t m) const {
    return (int)(std::upper_bound(partEnd.begin(), partEnd.end(), m) - partEnd.begin());
  }
  // boundary id met when stepping from inside to (x,y,z); 0 = none
  int bid(int x, int y, int z, int) const {
    if (kind == 0) return 0;
    if (kind == 2) {  // channel: walls on the y/z faces, inlet at x < 0, outlet at x >= extent
      if (y < 0 || y >= ext[1] || z < 0 || z >= ext[2]) return 1;
      if (x < 0) return 2;
      if (x >= ext[0]) return 3;
      return 0;
    }
    const bool outxy = x < 0 || x >= ext[0] || y < 0 || y >= ext[1];
    if (outxy || z < 0) return 1;  // 'wall'
    if (z >= ext[2]) return 2;     // 'lid'
    return 0;
  }
};

void build(Box &b, int commReduced) {
  const int QQ = b.QQ, QQN = QQ - 1, n = 1 << b.level;
  b.ext[0] = b.octants >= 2 ? n : n / 2;
  b.ext[1] = b.octants >= 4 ? n : n / 2;
  b.ext[2] = b.octants >= 8 ? n : n / 2;
  // whole octants in Morton order: the element list is the contiguous Morton range [0, nGlob)
  const int64_t nGlob = (int64_t)b.octants * (n / 2) * (n / 2) * (n / 2);
  b.firstId = 0;
  for (int l = 0; l < b.level; ++l) b.firstId = b.firstId * 8 + 1;  // (8^L - 1)/7
  for (int q = 0; q < QQN; ++q)
    for (int j = 0; j < QQN; ++j)
      if (kCx[j][0] == -kCx[q][0] && kCx[j][1] == -kCx[q][1] && kCx[j][2] == -kCx[q][2]) b.inv[q] = j;
  b.inv[QQ - 1] = QQ - 1;
  // contiguous shares, the first `remainder` parts get one more element
  const int64_t share = nGlob / b.nranks, rem = nGlob % b.nranks;
  int64_t first = 0;
  b.partEnd.resize(b.nranks);
  for (int p = 0; p < b.nranks; ++p) {
    const int64_t cnt = share + (p < rem ? 1 : 0);
    if (p == b.rank) { b.lo = first; b.hi = first + cnt; }
    first += cnt;
    b.partEnd[p] = first;
  }
  b.nFluid = (int)(b.hi - b.lo);
  const int nF = b.nFluid;

  // pass 1: neighbour morton index (or -bcid) of every local element
  std::vector<int64_t> nm((size_t)nF * QQN);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nF; ++e) {
    const uint64_t m = (uint64_t)(b.lo + e);
    const int x = (int)compact3(m), y = (int)compact3(m >> 1), z = (int)compact3(m >> 2);
    for (int q = 0; q < QQN; ++q) {
      const int xn = x + kCx[q][0], yn = y + kCx[q][1], zn = z + kCx[q][2];
      const int id = b.bid(xn, yn, zn, n);
      nm[(size_t)e * QQN + q] =
          id > 0 ? -(int64_t)id : mortonOf((xn + n) % n, (yn + n) % n, (zn + n) % n);
    }
  }
  // halos: remote neighbours, unique, ascending treeID
  std::vector<int64_t> halo;
  if (b.nranks > 1) {
    for (size_t i = 0; i < nm.size(); ++i)
      if (nm[i] >= 0 && (nm[i] < b.lo || nm[i] >= b.hi)) halo.push_back(nm[i]);
    std::sort(halo.begin(), halo.end());
    halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
  }
  b.nHalo = (int)halo.size();
  b.nElems = nF + b.nHalo;
  b.nSize = (b.nElems + 3) / 4 * 4;
  b.total.resize(b.nElems);
  for (int e = 0; e < nF; ++e) b.total[e] = b.firstId + b.lo + e;
  for (int h = 0; h < b.nHalo; ++h) b.total[nF + h] = b.firstId + halo[h];

  auto posOf = [&](int64_t m) -> int32_t {  // 1-based position in the total list, 0 = absent
    if (m >= b.lo && m < b.hi) return (int32_t)(m - b.lo + 1);
    auto it = std::lower_bound(halo.begin(), halo.end(), m);
    if (it != halo.end() && *it == m) return (int32_t)(nF + (it - halo.begin()) + 1);
    return 0;
  };
  b.ngh.assign((size_t)b.nElems * QQN, 0);
  b.property.assign(b.nElems, 0);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < nF; ++e) {
    bool hasBnd = false;
    for (int q = 0; q < QQN; ++q) {
      const int64_t m = nm[(size_t)e * QQN + q];
      if (m < 0) { b.ngh[(size_t)e * QQN + q] = (int32_t)m; hasBnd = true; }
      else b.ngh[(size_t)e * QQN + q] = posOf(m);
    }
    b.property[e] = (1ll << 1) | (hasBnd ? (1ll << 3) : 0);  // prp_fluid, prp_hasBnd
  }
#pragma omp parallel for schedule(static)
  for (int h = 0; h < b.nHalo; ++h) {
    const uint64_t m = (uint64_t)halo[h];
    const int x = (int)compact3(m), y = (int)compact3(m >> 1), z = (int)compact3(m >> 2);
    for (int q = 0; q < QQN; ++q) {
      const int xn = x + kCx[q][0], yn = y + kCx[q][1], zn = z + kCx[q][2];
      const int id = b.bid(xn, yn, zn, n);
      b.ngh[(size_t)(nF + h) * QQN + q] =
          id > 0 ? -id : posOf(mortonOf((xn + n) % n, (yn + n) % n, (zn + n) % n));
    }
  }

  // mus_construct_connectivity (AOS + PULL)
  b.neigh.assign((size_t)QQ * b.nSize, 0);
#pragma omp parallel for schedule(static)
  for (int e = 1; e <= b.nElems; ++e) {
    b.neigh[(size_t)(QQ - 1) * b.nSize + e - 1] = (e - 1) * QQ + QQ;
    for (int d = 1; d <= QQN; ++d) {
      const int nghDir = b.inv[d - 1] + 1;
      const int neighPos = b.ngh[(size_t)(e - 1) * QQN + nghDir - 1];
      const bool missingNonGhost = neighPos <= 0;  // single level: every element is fluid or halo
      int sourceDir = d;
      if (missingNonGhost) sourceDir = b.inv[d - 1] + 1;
      const int getFromPos = neighPos <= 0 ? e : neighPos;
      b.neigh[(size_t)(d - 1) * b.nSize + e - 1] = (getFromPos - 1) * QQ + sourceDir;
    }
  }

  // halo exchange lists
  if (b.nranks > 1) {
    // recv: per source rank, halo elements ascending, links dir-ascending
    std::vector<std::vector<int32_t>> rpos(b.nranks), relem(b.nranks);
    for (int h = 0; h < b.nHalo; ++h) {
      const int p = b.owner(halo[h]);
      const int e = nF + h + 1;
      relem[p].push_back(e);
      for (int d = 1; d <= QQ; ++d) {
        const int neighDir = b.inv[d - 1] + 1;
        const int nghElem = (b.neigh[(size_t)(neighDir - 1) * b.nSize + e - 1] - 1) / QQ + 1;
        if (nghElem <= nF || !commReduced) rpos[p].push_back((e - 1) * QQ + d);
      }
    }
    // send: my elements with a stencil neighbour owned by p; link d goes out when
    // the element at (me + c_d) is owned by p  (mirror of the peer's recv rule)
    std::vector<std::vector<int32_t>> spos(b.nranks), selem(b.nranks);
    std::vector<int> ownerOf(QQN);
    for (int e = 0; e < nF; ++e) {
      bool any = false;
      for (int q = 0; q < QQN; ++q) {
        const int64_t m = nm[(size_t)e * QQN + q];
        ownerOf[q] = (m >= 0 && (m < b.lo || m >= b.hi)) ? b.owner(m) : -1;
        any = any || ownerOf[q] >= 0;
      }
      if (!any) continue;
      b.property[e] |= (1ll << 12);  // prp_sendHalo
      for (int p = 0; p < b.nranks; ++p) {
        bool toP = false;
        for (int q = 0; q < QQN; ++q) toP = toP || ownerOf[q] == p;
        if (!toP) continue;
        selem[p].push_back(e + 1);
        for (int d = 1; d <= QQ; ++d) {
          const bool wanted = (d <= QQN && ownerOf[d - 1] == p) || !commReduced;
          if (wanted) spos[p].push_back(e * QQ + d);
        }
      }
    }
    for (int p = 0; p < b.nranks; ++p) {
      if (!relem[p].empty()) {
        b.recv.proc.push_back(p);
        b.recv.nVals.push_back((int32_t)rpos[p].size());
        b.recv.pos.insert(b.recv.pos.end(), rpos[p].begin(), rpos[p].end());
        b.recv.elemCount.push_back((int32_t)relem[p].size());
        b.recv.elemPos.insert(b.recv.elemPos.end(), relem[p].begin(), relem[p].end());
      }
      if (!selem[p].empty()) {
        b.send.proc.push_back(p);
        b.send.nVals.push_back((int32_t)spos[p].size());
        b.send.pos.insert(b.send.pos.end(), spos[p].begin(), spos[p].end());
        b.send.elemCount.push_back((int32_t)selem[p].size());
        b.send.elemPos.insert(b.send.elemPos.end(), selem[p].begin(), selem[p].end());
      }
    }
  }

  // boundary lists.  cavity: id 1 'wall' (do_nothing), id 2 'lid' (velocity_bounceback);
  // channel: id 1 'wall', id 2 'inlet' (velocity_bounceback), id 3 'outlet' (pressure)
  if (b.kind >= 1) {
    std::vector<int32_t> posInBuf(b.nElems + 1, 0);
    for (int e = 0; e < nF; ++e)
      if (b.property[e] & (1ll << 3)) {
        b.bcElemBuffer.push_back(e + 1);
        posInBuf[e + 1] = (int32_t)b.bcElemBuffer.size();
      }
    // prevailing directions = normalised stencil directions; weights 4/2/1 (assignBCList)
    double prevail[26][3];
    int wgt[26];
    for (int k = 0; k < QQN; ++k) {
      const int len = kCx[k][0] * kCx[k][0] + kCx[k][1] * kCx[k][1] + kCx[k][2] * kCx[k][2];
      wgt[k] = len == 1 ? 4 : (len == 2 ? 2 : 1);
      const double r = std::sqrt((double)len);
      for (int c = 0; c < 3; ++c) prevail[k][c] = (double)kCx[k][c] / r;
    }
    const int nIds = b.kind == 1 ? 2 : 3;
    for (int id = 1; id <= nIds; ++id) {
      Bc bc;
      bc.id = id;
      bc.kind = id - 1;
      for (int e = 0; e < nF; ++e) {
        if (!(b.property[e] & (1ll << 3))) continue;
        bool mask[26] = {false};
        bool any = false;
        long long nrm[3] = {0, 0, 0};
        for (int k = 0; k < QQN; ++k)
          if (b.ngh[(size_t)e * QQN + k] == -id) {
            mask[b.inv[k]] = true;
            any = true;
            for (int c = 0; c < 3; ++c) nrm[c] -= (long long)wgt[k] * kCx[k][c];
          }
        if (!any) continue;
        bc.elems.push_back(e + 1);
        const int iElem = (int)bc.elems.size();
        for (int d = 1; d <= QQN; ++d) {
          if (!mask[d - 1]) continue;
          bc.links.push_back(b.neigh[(size_t)(d - 1) * b.nSize + e]);  // FETCH(iDir, elem)
          bc.iDir.push_back(d);
          bc.posInBuffer.push_back(posInBuf[e + 1]);
          bc.outPos.push_back((b.inv[d - 1] + 1) + (posInBuf[e + 1] - 1) * QQ);
          bc.iElemOfLink.push_back(iElem);
          bc.statePos.push_back(d + (iElem - 1) * QQ);
        }
        // tem_determine_discreteVector: first strict maximum of the projection, exit at 1
        const double len = std::sqrt((double)(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]));
        int best = 0;
        double mx = -2.0;
        for (int k = 0; k < QQN; ++k) {
          double dp = (nrm[0] / len) * prevail[k][0] + (nrm[1] / len) * prevail[k][1] +
                      (nrm[2] / len) * prevail[k][2];
          dp = std::min(std::max(dp, -1.0), 1.0);
          if (dp > mx) {
            mx = dp;
            best = k;
            if (std::fabs(mx - 1.0) <= 2.220446049250313e-16) break;
          }
        }
        bc.normalInd.push_back(best + 1);
        bc.posInBcElemBuf.push_back(posInBuf[e + 1]);
        // setFieldBCNeigh: the elements at x + k*normal, k = 1, 2
        const uint64_t m = (uint64_t)(b.lo + e);
        const int x = (int)compact3(m), y = (int)compact3(m >> 1), z = (int)compact3(m >> 2);
        int32_t np[2] = {e + 1, e + 1};
        for (int k = 1; k <= 2; ++k) {
          const int xn = x + k * kCx[best][0], yn = y + k * kCx[best][1], zn = z + k * kCx[best][2];
          int32_t p = 0;
          if (xn >= 0 && xn < b.ext[0] && yn >= 0 && yn < b.ext[1] && zn >= 0 && zn < b.ext[2])
            p = posOf(mortonOf(xn, yn, zn));
          if (p <= 0) {                       // no valid neighbour: keep the last valid one
            if (k == 1) { np[0] = np[1] = e + 1; }
            else np[1] = np[0];
            break;
          }
          np[k - 1] = p;
          if (k == 1) np[1] = p;
        }
        bc.neighPos.push_back(np[0]);
        bc.neighPos.push_back(np[1]);
      }
      b.bcs.push_back(std::move(bc));
    }
  }
}

template <class T>
int copyOut(const std::vector<T> &v, T *out) {
  if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(T));
  return (int)v.size();
}

}  // namespace

extern "C" {

// kind: 0 = fully periodic cube, 1 = cavity (5 walls + moving lid at z = top),
//       2 = channel (4 walls, velocity inlet at x = 0, pressure outlet at x = top)
// octants: the domain is the first 1, 2, 4 or 8 octants of the level-L cube (a periodic mesh
//          must fill the cube: treelm wraps at the universe)
void *musb200_mesh_box_create(int level, int QQ, int kind, int rank, int nranks, int comm_reduced,
                              int octants) {
  if (level < 1 || level > 10 || (QQ != 19 && QQ != 27) || kind < 0 || kind > 2 || nranks < 1 ||
      rank < 0 || rank >= nranks)
    return nullptr;
  if ((octants != 1 && octants != 2 && octants != 4 && octants != 8) || (kind == 0 && octants != 8))
    return nullptr;
  Box *b = new Box();
  b->level = level; b->QQ = QQ; b->kind = kind; b->rank = rank; b->nranks = nranks;
  b->octants = octants;
  build(*b, comm_reduced);
  return b;
}
void musb200_mesh_destroy(void *h) { delete static_cast<Box *>(h); }

// info: nFluid, nHalo, nElems, nSize, nBcElems, nBCs, nRecvProcs, nSendProcs, nRecvVals, nSendVals
int musb200_mesh_info(void *h, int64_t *info) {
  Box *b = static_cast<Box *>(h);
  if (!b || !info) return 1;
  info[0] = b->nFluid; info[1] = b->nHalo; info[2] = b->nElems; info[3] = b->nSize;
  info[4] = (int64_t)b->bcElemBuffer.size(); info[5] = (int64_t)b->bcs.size();
  info[6] = (int64_t)b->recv.proc.size(); info[7] = (int64_t)b->send.proc.size();
  info[8] = (int64_t)b->recv.pos.size(); info[9] = (int64_t)b->send.pos.size();
  info[10] = (int64_t)b->recv.elemPos.size(); info[11] = (int64_t)b->send.elemPos.size();
  return 0;
}
int musb200_mesh_total(void *h, int64_t *out) { return copyOut(static_cast<Box *>(h)->total, out); }
int musb200_mesh_property(void *h, int64_t *out) { return copyOut(static_cast<Box *>(h)->property, out); }
int musb200_mesh_nghelems(void *h, int32_t *out) { return copyOut(static_cast<Box *>(h)->ngh, out); }
int musb200_mesh_neigh(void *h, int32_t *out) { return copyOut(static_cast<Box *>(h)->neigh, out); }
int musb200_mesh_bc_elembuffer(void *h, int32_t *out) { return copyOut(static_cast<Box *>(h)->bcElemBuffer, out); }
// dir 0 = send, 1 = recv
int musb200_mesh_comm(void *h, int dir, int32_t *proc, int32_t *nVals, int32_t *pos,
                      int32_t *elemCount, int32_t *elemPos) {
  Box *b = static_cast<Box *>(h);
  const Comm &c = dir == 0 ? b->send : b->recv;
  copyOut(c.proc, proc); copyOut(c.nVals, nVals); copyOut(c.pos, pos);
  copyOut(c.elemCount, elemCount); copyOut(c.elemPos, elemPos);
  return (int)c.proc.size();
}
// sizes: id, kind, nElems, nLinks
int musb200_mesh_bc_info(void *h, int i, int32_t *sizes) {
  Box *b = static_cast<Box *>(h);
  if (i < 0 || i >= (int)b->bcs.size()) return 1;
  sizes[0] = b->bcs[i].id; sizes[1] = b->bcs[i].kind;
  sizes[2] = (int32_t)b->bcs[i].elems.size(); sizes[3] = (int32_t)b->bcs[i].links.size();
  return 0;
}
int musb200_mesh_bc_lists(void *h, int i, int32_t *elems, int32_t *links, int32_t *outPos,
                          int32_t *posInBuffer, int32_t *iDir) {
  Box *b = static_cast<Box *>(h);
  if (i < 0 || i >= (int)b->bcs.size()) return 1;
  const Bc &bc = b->bcs[i];
  copyOut(bc.elems, elems); copyOut(bc.links, links); copyOut(bc.outPos, outPos);
  copyOut(bc.posInBuffer, posInBuffer); copyOut(bc.iDir, iDir);
  return 0;
}
// per element: normalInd, posInBcElemBuf [nElems], neighPos [nElems][2]; per link: iElem, statePos
int musb200_mesh_bc_elem_lists(void *h, int i, int32_t *normalInd, int32_t *posInBcElemBuf,
                               int32_t *neighPos, int32_t *iElemOfLink, int32_t *statePos) {
  Box *b = static_cast<Box *>(h);
  if (i < 0 || i >= (int)b->bcs.size()) return 1;
  const Bc &bc = b->bcs[i];
  copyOut(bc.normalInd, normalInd); copyOut(bc.posInBcElemBuf, posInBcElemBuf);
  copyOut(bc.neighPos, neighPos); copyOut(bc.iElemOfLink, iElemOfLink);
  copyOut(bc.statePos, statePos);
  return 0;
}
// barycentres (tem_BaryOfId, tem_geometry_module.f90:419-435): out[nElems][3]
int musb200_mesh_bary(void *h, double ox, double oy, double oz, double length, double *out) {
  Box *b = static_cast<Box *>(h);
  const double dx = length / (double)(1 << b->level);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < b->nElems; ++e) {
    const uint64_t m = (uint64_t)(b->total[e] - b->firstId);
    out[3 * (size_t)e + 0] = ox + ((double)compact3(m) + 0.5) * dx;
    out[3 * (size_t)e + 1] = oy + ((double)compact3(m >> 1) + 0.5) * dx;
    out[3 * (size_t)e + 2] = oz + ((double)compact3(m >> 2) + 0.5) * dx;
  }
  return 0;
}

}  // extern "C"
