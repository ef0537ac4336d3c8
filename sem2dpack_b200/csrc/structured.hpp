// Recognises a structured box in an arbitrary ibool table.
//
// The Fortran host hands over what init_main built: for MESH_CART that is a box of nx*nz Q4 elements whose
// ELEMENTS ARE IN RCM ORDER (OPT_RENUMBER, SRC/constants.f90:11, mesh_structured.f90:204-269) and whose GLL
// numbering follows that order (SE_init_numbering, spec_grid.f90:198-314).  Nothing in ibool says "box", but
// the topology does: every element of CART_build has the same orientation (knods = SW,SE,NE,NW,
// mesh_structured.f90:24-35), so the right neighbour of e is the element whose i = 1 column is e's i = N
// column, and the element above is the one whose j = 1 row is e's j = N row.  If following those links
// arranges all elements in one nx*nz rectangle -- or in two rectangles of equal width, which is what the
// split-node fault row of `ezflt` leaves (mesh_cartesian.f90:234-261) -- the mesh is routed to the z-marching
// strip kernel (strip_kernels.cuh) and the caller's node ids are mapped to GLL lattice positions.
// Anything else (unstructured meshes, elements of mixed orientation, holes) keeps the any-mesh kernels.
#pragma once
#include <cstdint>
#include <vector>

namespace s2d {

struct StructuredBox {
  bool ok = false;
  int nx = 0, nz = 0, ezflt = 0;     // ezflt = element rows below the split-node row (0 = one block)
  std::vector<int32_t> ex, ez;       // (nelem) position of every element in the box
  std::vector<int32_t> gx, gz;       // (npoin) lattice position of every node (gz counts the duplicated fault row)
};

// ibool(N,N,nelem), 1-based node ids.  lower_hint: a node (1-based) known to lie in the lower block of a
// two-block mesh (node1 of a two-sided fault), or 0.
inline StructuredBox detect_structured(const int32_t* ibool, int N, int nelem, size_t npoin, int32_t lower_hint) {
  StructuredBox B;
  const size_t n2 = (size_t)N * N;
  auto ib = [&](int e, int i, int j) { return ibool[(size_t)e * n2 + (size_t)j * N + i] - 1; };  // 0-based i, j
  // element whose left / bottom edge carries a given edge-interior node
  std::vector<int32_t> by_left(npoin, -1), by_down(npoin, -1);
  for (int e = 0; e < nelem; ++e) {
    int32_t& l = by_left[ib(e, 0, 1)];
    int32_t& d = by_down[ib(e, 1, 0)];
    if (l != -1 || d != -1) return B;  // two elements with the same left / bottom edge: not a box of one orientation
    l = e;
    d = e;
  }
  std::vector<int32_t> right(nelem), up(nelem), has_left(nelem, 0), has_down(nelem, 0);
  for (int e = 0; e < nelem; ++e) {
    const int32_t r = by_left[ib(e, N - 1, 1)], u = by_down[ib(e, 1, N - 1)];
    right[e] = (r != e) ? r : -1;
    up[e] = (u != e) ? u : -1;
    if (right[e] >= 0) {
      for (int j = 0; j < N; ++j)
        if (ib(e, N - 1, j) != ib(right[e], 0, j)) return B;
      if (has_left[right[e]]) return B;
      has_left[right[e]] = 1;
    }
    if (up[e] >= 0) {
      for (int i = 0; i < N; ++i)
        if (ib(e, i, N - 1) != ib(up[e], i, 0)) return B;
      if (has_down[up[e]]) return B;
      has_down[up[e]] = 1;
    }
  }
  // blocks: start at every element without left and bottom neighbours
  struct Block {
    int nx, nz;
    std::vector<int32_t> elems;  // row-major
  };
  std::vector<Block> blocks;
  size_t covered = 0;
  for (int e0 = 0; e0 < nelem; ++e0) {
    if (has_left[e0] || has_down[e0]) continue;
    Block b;
    b.nx = 0;
    b.nz = 0;
    for (int rs = e0; rs >= 0; rs = up[rs]) {  // first element of every row
      if (has_left[rs]) return B;
      int cnt = 0;
      for (int e = rs; e >= 0; e = right[e]) {
        b.elems.push_back(e);
        if (++cnt > nelem) return B;
      }
      if (b.nz == 0) b.nx = cnt;
      else if (cnt != b.nx) return B;
      if (++b.nz > nelem) return B;
    }
    // the k-th element of a row sits below the k-th element of the next row
    for (int r = 0; r + 1 < b.nz; ++r)
      for (int k = 0; k < b.nx; ++k)
        if (up[b.elems[(size_t)r * b.nx + k]] != b.elems[(size_t)(r + 1) * b.nx + k]) return B;
    for (int k = 0; k < b.nx; ++k)
      if (up[b.elems[(size_t)(b.nz - 1) * b.nx + k]] != -1) return B;
    covered += b.elems.size();
    blocks.push_back(std::move(b));
    if (blocks.size() > 2) return B;
  }
  if (covered != (size_t)nelem || blocks.empty()) return B;
  if (blocks.size() == 2) {
    if (blocks[0].nx != blocks[1].nx) return B;
    // which block is below the split-node row?  The one that holds the hint node; without a hint, the one found
    // first (lowest element id at its origin).
    if (lower_hint > 0) {
      bool in0 = false;
      for (int32_t e : blocks[0].elems) {
        for (size_t k = 0; k < n2 && !in0; ++k) in0 = ibool[(size_t)e * n2 + k] == lower_hint;
        if (in0) break;
      }
      if (!in0) std::swap(blocks[0], blocks[1]);
    }
  }
  B.nx = blocks[0].nx;
  B.nz = blocks[0].nz + (blocks.size() == 2 ? blocks[1].nz : 0);
  B.ezflt = blocks.size() == 2 ? blocks[0].nz : 0;
  B.ex.assign(nelem, -1);
  B.ez.assign(nelem, -1);
  B.gx.assign(npoin, -1);
  B.gz.assign(npoin, -1);
  int row0 = 0;
  for (size_t bi = 0; bi < blocks.size(); ++bi) {
    const Block& b = blocks[bi];
    for (int r = 0; r < b.nz; ++r)
      for (int k = 0; k < b.nx; ++k) {
        const int e = b.elems[(size_t)r * b.nx + k];
        const int iz = row0 + r;
        B.ex[e] = k;
        B.ez[e] = iz;
        for (int j = 0; j < N; ++j)
          for (int i = 0; i < N; ++i) {
            const int32_t nd = ib(e, i, j);
            const int32_t x = k * (N - 1) + i, z = iz * (N - 1) + j + (int)bi;  // the upper block starts one lattice row higher
            if (B.gx[nd] == -1) {
              B.gx[nd] = x;
              B.gz[nd] = z;
            } else if (B.gx[nd] != x || B.gz[nd] != z) {
              return B;  // a node shared in a way a box does not share it (e.g. a merged periodic pair)
            }
          }
      }
    row0 += b.nz;
  }
  for (size_t k = 0; k < npoin; ++k)
    if (B.gx[k] < 0) return B;  // a node no element holds
  const size_t LX = (size_t)B.nx * (N - 1) + 1, LZ = (size_t)B.nz * (N - 1) + 1 + (B.ezflt > 0 ? 1 : 0);
  if (LX * LZ != npoin) return B;
  B.ok = true;
  return B;
}

}  // namespace s2d
