// ORACLE (test infrastructure only -- never linked into the product library).
// CPU restatement of the reference's reverse Cuthill-McKee element renumbering.
// Follows /root/reference/SRC/rcm.f90 (degree, genrcm, level_set, rcm, root_find,
// perm_inverse; SPARSPAK as packaged by Burkardt) and the adjacency builder in
// /root/reference/SRC/mesh_structured.f90:204-269.  All arrays are 1-based like the
// Fortran (index 0 unused) so the control flow can be checked line by line.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>

namespace orc {
namespace rcmlib {

typedef std::vector<int> ivec;

// level_set (rcm.f90): BFS level structure of the masked component containing root.
// level[] is written starting at offset `lo` (the Fortran passes perm(num)).
inline void level_set(int root, const ivec& adj_row, const ivec& adj, ivec& mask, int& level_num,
                      ivec& level_row, int* level /*1-based view*/) {
  mask[root] = 0;
  level[1] = root;
  level_num = 0;
  int lvlend = 0, iccsze = 1;
  for (;;) {
    int lbegin = lvlend + 1;
    lvlend = iccsze;
    level_num = level_num + 1;
    level_row[level_num] = lbegin;
    for (int i = lbegin; i <= lvlend; ++i) {
      int node = level[i];
      int jstrt = adj_row[node], jstop = adj_row[node + 1] - 1;
      for (int j = jstrt; j <= jstop; ++j) {
        int nbr = adj[j];
        if (mask[nbr] != 0) {
          iccsze = iccsze + 1;
          level[iccsze] = nbr;
          mask[nbr] = 0;
        }
      }
    }
    int lvsize = iccsze - lvlend;
    if (lvsize <= 0) break;
  }
  level_row[level_num + 1] = lvlend + 1;
  for (int i = 1; i <= iccsze; ++i) mask[level[i]] = 1;
}

// root_find (rcm.f90): pseudo-peripheral node
inline void root_find(int& root, const ivec& adj_row, const ivec& adj, ivec& mask, int& level_num,
                      ivec& level_row, int* level) {
  level_set(root, adj_row, adj, mask, level_num, level_row, level);
  int iccsze = level_row[level_num + 1] - 1;
  if (level_num == 1) return;
  if (level_num == iccsze) return;
  for (;;) {
    int mindeg = iccsze;
    int jstrt = level_row[level_num];
    root = level[jstrt];
    if (jstrt < iccsze) {
      for (int j = jstrt; j <= iccsze; ++j) {
        int node = level[j];
        int ndeg = 0;
        int kstrt = adj_row[node], kstop = adj_row[node + 1] - 1;
        for (int k = kstrt; k <= kstop; ++k) {
          int nabor = adj[k];
          if (0 < mask[nabor]) ndeg = ndeg + 1;
        }
        if (ndeg < mindeg) {
          root = node;
          mindeg = ndeg;
        }
      }
    }
    int level_num2;
    level_set(root, adj_row, adj, mask, level_num2, level_row, level);
    if (level_num2 <= level_num) break;
    level_num = level_num2;
    if (iccsze <= level_num) break;
  }
}

// degree (rcm.f90): degrees in the masked component; ls = BFS order
inline void degree(int root, ivec& adj_row, const ivec& adj, const ivec& mask, ivec& deg,
                   int& iccsze, int* ls) {
  ls[1] = root;
  adj_row[root] = -adj_row[root];
  int lvlend = 0;
  iccsze = 1;
  for (;;) {
    int lbegin = lvlend + 1;
    lvlend = iccsze;
    for (int i = lbegin; i <= lvlend; ++i) {
      int node = ls[i];
      int jstrt = -adj_row[node];
      int jstop = std::abs(adj_row[node + 1]) - 1;
      int ideg = 0;
      for (int j = jstrt; j <= jstop; ++j) {
        int nbr = adj[j];
        if (mask[nbr] != 0) {
          ideg = ideg + 1;
          if (0 <= adj_row[nbr]) {
            adj_row[nbr] = -adj_row[nbr];
            iccsze = iccsze + 1;
            ls[iccsze] = nbr;
          }
        }
      }
      deg[node] = ideg;
    }
    int lvsize = iccsze - lvlend;
    if (lvsize == 0) break;
  }
  for (int i = 1; i <= iccsze; ++i) {
    int node = ls[i];
    adj_row[node] = -adj_row[node];
  }
}

// rcm (rcm.f90): RCM ordering of the component containing root
inline void rcm(int root, ivec& adj_row, const ivec& adj, ivec& mask, int* perm, int& iccsze,
                ivec& deg) {
  degree(root, adj_row, adj, mask, deg, iccsze, perm);
  mask[root] = 0;
  if (iccsze <= 1) return;
  int lvlend = 0, lnbr = 1;
  while (lvlend < lnbr) {
    int lbegin = lvlend + 1;
    lvlend = lnbr;
    for (int i = lbegin; i <= lvlend; ++i) {
      int node = perm[i];
      int jstrt = adj_row[node], jstop = adj_row[node + 1] - 1;
      int fnbr = lnbr + 1;
      for (int j = jstrt; j <= jstop; ++j) {
        int nbr = adj[j];
        if (mask[nbr] != 0) {
          lnbr = lnbr + 1;
          mask[nbr] = 0;
          perm[lnbr] = nbr;
        }
      }
      if (lnbr <= fnbr) continue;
      int k = fnbr;
      while (k < lnbr) {
        int l = k;
        k = k + 1;
        int nbr = perm[k];
        while (fnbr < l) {
          int lperm = perm[l];
          if (deg[lperm] <= deg[nbr]) break;
          perm[l + 1] = lperm;
          l = l - 1;
        }
        perm[l + 1] = nbr;
      }
    }
  }
  for (int i = 1; i <= iccsze / 2; ++i) {  // ivec_reverse
    int t = perm[i];
    perm[i] = perm[iccsze + 1 - i];
    perm[iccsze + 1 - i] = t;
  }
}

// genrcm (rcm.f90): perm(new) = old, 1-based values, perm sized node_num+1
inline void genrcm(int node_num, ivec& adj_row, const ivec& adj, ivec& perm) {
  ivec mask(node_num + 2, 1), level_row(node_num + 2, 0), deg(node_num + 2, 0);
  int num = 1;
  for (int i = 1; i <= node_num; ++i) {
    if (mask[i] != 0) {
      int root = i, level_num = 0, iccsze = 0;
      int* pview = perm.data() + (num - 1);  // perm(num) passed as array start
      root_find(root, adj_row, adj, mask, level_num, level_row, pview);
      rcm(root, adj_row, adj, mask, pview, iccsze, deg);
      num = num + iccsze;
      if (node_num < num) return;
    }
  }
}

// mesh_structured.f90:204-269 -- 8-neighbour element graph of an nx*nz box, then RCM.
// Returns perm (new->old) and perm_inv (old->new), both 1-based, index 0 unused.
inline void structured_rcm(int nx, int nz, ivec& perm, ivec& perm_inv) {
  int nelem = nx * nz;
  ivec adj_row(nelem + 2, 0);
  ivec adj;
  adj.reserve((size_t)8 * nelem + 1);
  adj.push_back(0);
  int e = 0, nadj = 0;
  for (int j = 1; j <= nz; ++j)
    for (int i = 1; i <= nx; ++i) {
      e = e + 1;
      adj_row[e] = nadj + 1;
      const int di[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
      const int dj[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
      for (int k = 0; k < 8; ++k) {
        int ii = i + di[k], jj = j + dj[k];
        if (ii <= nx && ii >= 1 && jj <= nz && jj >= 1) {
          nadj = nadj + 1;
          adj.push_back((jj - 1) * nx + ii);
        }
      }
    }
  adj_row[e + 1] = nadj + 1;
  perm.assign(nelem + 1, 0);
  genrcm(nelem, adj_row, adj, perm);
  perm_inv.assign(nelem + 1, 0);
  for (int i = 1; i <= nelem; ++i) perm_inv[perm[i]] = i;
}

}  // namespace rcmlib
}  // namespace orc
