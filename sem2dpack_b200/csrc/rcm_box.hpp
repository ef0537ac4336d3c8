// Reverse Cuthill-McKee order of the elements of an nx x nz box, as the reference renumbers every structured mesh
// by default (OPT_RENUMBER = .true., SRC/constants.f90:10-15): MESH_STRUCTURED_renumber (SRC/mesh_structured.f90:
// 204-269) builds the 8-neighbour element graph and calls genrcm (SRC/rcm.f90, SPARSPAK): a pseudo-peripheral root
// (root_find: repeated level structures, restarting from the minimum-degree node of the last level, first minimum
// wins), the Cuthill-McKee sweep (rcm: every node's newly found neighbours are insertion-sorted by ascending degree
// -- except the first of them, which the loop bound `fnbr < l` never moves), then the reversal.
// The order is integer data that must match the reference bit for bit: it decides the element order of
// ibool_sem2d.dat and, through SE_init_numbering, the id of every GLL node.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace s2d {

// perm[new] = old, both 0-based; natural order: element (i,j) (0-based) at i + nx*j
inline std::vector<int32_t> rcm_box_perm(int nx, int nz) {
  const int n = nx * nz;
  // adjacency in the reference's order: (i-1,j-1) (i,j-1) (i+1,j-1) (i-1,j) (i+1,j) (i-1,j+1) (i,j+1) (i+1,j+1)
  std::vector<int32_t> row(n + 1, 0), adj;
  adj.reserve((size_t)8 * n);
  static const int di[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, dj[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
  for (int j = 0; j < nz; ++j)
    for (int i = 0; i < nx; ++i) {
      const int e = i + nx * j;
      row[e] = (int32_t)adj.size();
      for (int k = 0; k < 8; ++k) {
        const int a = i + di[k], b = j + dj[k];
        if (a >= 0 && a < nx && b >= 0 && b < nz) adj.push_back(a + nx * b);
      }
    }
  row[n] = (int32_t)adj.size();
  std::vector<uint8_t> mask(n, 1);
  std::vector<int32_t> level(n), lrow, perm(n);
  // level structure rooted at `root` over the unmasked nodes; returns the number of levels, lrow[l] = first index
  // of level l in `level`, lrow[nlev] = component size; the mask is left as it was found
  auto level_set = [&](int root) {
    lrow.clear();
    mask[root] = 0;
    level[0] = root;
    int lvlend = 0, size = 1, nlev = 0;
    for (;;) {
      const int lbegin = lvlend;
      lvlend = size;
      ++nlev;
      lrow.push_back(lbegin);
      for (int q = lbegin; q < lvlend; ++q) {
        const int node = level[q];
        for (int p = row[node]; p < row[node + 1]; ++p) {
          const int nb = adj[p];
          if (mask[nb]) {
            level[size++] = nb;
            mask[nb] = 0;
          }
        }
      }
      if (size - lvlend <= 0) break;
    }
    lrow.push_back(lvlend);
    for (int q = 0; q < size; ++q) mask[level[q]] = 1;
    return nlev;
  };
  int num = 0;
  for (int start = 0; start < n && num < n; ++start) {
    if (!mask[start]) continue;
    // ---- root_find
    int root = start;
    int nlev = level_set(root);
    int size = lrow[nlev];
    if (!(nlev == 1 || nlev == size)) {
      for (;;) {
        int mindeg = size;
        const int jstrt = lrow[nlev - 1];
        root = level[jstrt];
        if (jstrt < size - 1) {
          for (int q = jstrt; q < size; ++q) {
            const int node = level[q];
            int ndeg = 0;
            for (int p = row[node]; p < row[node + 1]; ++p) ndeg += mask[adj[p]] ? 1 : 0;
            if (ndeg < mindeg) {
              root = node;
              mindeg = ndeg;
            }
          }
        }
        const int nlev2 = level_set(root);
        if (nlev2 <= nlev) break;
        nlev = nlev2;
        if (size <= nlev) break;
      }
    }
    // ---- degrees within the component (all of its nodes are unmasked here)
    std::vector<int32_t> deg(n, 0);
    level_set(root);
    size = lrow.back();
    for (int q = 0; q < size; ++q) {
      const int node = level[q];
      int d = 0;
      for (int p = row[node]; p < row[node + 1]; ++p) d += mask[adj[p]] ? 1 : 0;
      deg[node] = d;
    }
    // ---- Cuthill-McKee sweep into perm[num ...]
    int32_t* pm = perm.data() + num;
    pm[0] = root;
    mask[root] = 0;
    if (size > 1) {
      int lvlend = 0, lnbr = 1;  // counts, i.e. 1-based positions of the reference
      while (lvlend < lnbr) {
        const int lbegin = lvlend + 1;
        lvlend = lnbr;
        for (int q = lbegin; q <= lvlend; ++q) {
          const int node = pm[q - 1];
          const int fnbr = lnbr + 1;
          for (int p = row[node]; p < row[node + 1]; ++p) {
            const int nb = adj[p];
            if (mask[nb]) {
              ++lnbr;
              mask[nb] = 0;
              pm[lnbr - 1] = nb;
            }
          }
          if (lnbr <= fnbr) continue;
          int k = fnbr;
          while (k < lnbr) {
            int l = k;
            ++k;
            const int nb = pm[k - 1];
            while (fnbr < l) {
              const int lp = pm[l - 1];
              if (deg[lp] <= deg[nb]) break;
              pm[l] = lp;
              --l;
            }
            pm[l] = nb;
          }
        }
      }
      std::reverse(pm, pm + size);
    }
    num += size;
  }
  return perm;
}

}  // namespace s2d
