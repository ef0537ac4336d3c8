// Host-side planning for the assembly stage: greedy mesh colouring and CTA patches.
// The reference has no colouring (SRC/solver.f90:310-318 is only a design comment); the rule here
// is first-fit in ascending element id, two elements conflicting when they share any GLL node.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace s2d {

// Greedy first-fit colouring.  ibool is (n2, nelem) with 1-based node ids.  Returns the number of
// colours; color[e] is 0-based.  A per-node bit mask of the colours already incident to the node
// makes this O(nelem * n2).
inline int greedy_coloring(const int32_t* ibool, int n2, int nelem, size_t npoin,
                           std::vector<int32_t>& color) {
  std::vector<uint64_t> used(npoin, 0);
  color.assign(nelem, 0);
  int ncol = 0;
  for (int e = 0; e < nelem; ++e) {
    const int32_t* ib = ibool + (size_t)e * n2;
    uint64_t m = 0;
    for (int k = 0; k < n2; ++k) m |= used[(size_t)ib[k] - 1];
    int c = 0;
    while (c < 64 && ((m >> c) & 1ull)) ++c;
    if (c >= 64) c = 63;  // cannot happen on 2-D quad meshes (valence is bounded)
    color[e] = c;
    const uint64_t bit = 1ull << c;
    for (int k = 0; k < n2; ++k) used[(size_t)ib[k] - 1] |= bit;
    ncol = std::max(ncol, c + 1);
  }
  return ncol;
}

// Patch plan for the CTA-patch kernel: elements are grouped into patches of at most EP spatially
// clustered elements; each patch carries its unique node list (interior-to-patch nodes first, then
// nodes shared with other patches), 16-bit local indices per element node, and in-patch colours.
// Forces of nodes private to a patch are stored directly; partial sums of shared nodes go to halo
// slots that a second kernel adds in ascending patch order (deterministic, no atomics).
struct PatchPlan {
  int EP = 0;          // max elements per patch
  int n2 = 0;          // ngll*ngll
  int npatch = 0;
  int max_nloc = 0;    // max nodes per patch
  int max_colors = 0;  // max in-patch colours
  std::vector<int32_t> pelem_start;  // (npatch+1) range in `elems`
  std::vector<int32_t> elems;        // element ids (0-based) in patch-major order
  std::vector<int32_t> ecolor;       // (nelem, patch-major) in-patch colour
  std::vector<int32_t> pnode_start;  // (npatch+1) range in `pnode`
  std::vector<int32_t> pnint;        // (npatch) number of private nodes (first in the list)
  std::vector<int32_t> pnode;        // global node id (0-based) per patch-local node
  std::vector<int32_t> pslot;        // per patch-local node: -1 private, else halo slot
  std::vector<uint16_t> lidx;        // (n2, nelem patch-major) patch-local node index
  // shared nodes
  std::vector<int32_t> snode;        // (nshared) global node (0-based)
  std::vector<int32_t> sstart;       // (nshared+1) slot range, slots of one node are contiguous
  size_t nslots = 0;
};

// Cluster elements by breadth-first growth over node-adjacency, seeded at the lowest unassigned
// element id; candidates sharing more nodes with the growing patch are taken first.
inline void build_patch_plan(const int32_t* ibool, int n2, int nelem, size_t npoin, int EP,
                             PatchPlan& P) {
  P.EP = EP;
  P.n2 = n2;
  // node -> elements CSR
  std::vector<int32_t> nstart(npoin + 1, 0);
  for (size_t q = 0; q < (size_t)nelem * n2; ++q) nstart[(size_t)ibool[q]]++;
  for (size_t k = 0; k < npoin; ++k) nstart[k + 1] += nstart[k];
  std::vector<int32_t> nelems(nstart[npoin]);
  {
    std::vector<int32_t> fill(nstart.begin(), nstart.end() - 1);
    for (int e = 0; e < nelem; ++e)
      for (int k = 0; k < n2; ++k) nelems[fill[(size_t)ibool[(size_t)e * n2 + k] - 1]++] = e;
  }
  std::vector<int32_t> patch_of(nelem, -1);
  std::vector<int32_t> score(nelem, 0);  // nodes shared with the current patch
  P.pelem_start.assign(1, 0);
  P.elems.clear();
  P.elems.reserve(nelem);
  int next_seed = 0;
  std::vector<int32_t> frontier;
  while (true) {
    while (next_seed < nelem && patch_of[next_seed] >= 0) ++next_seed;
    if (next_seed >= nelem) break;
    const int p = (int)P.pelem_start.size() - 1;
    frontier.clear();
    int count = 0;
    int cur = next_seed;
    while (cur >= 0 && count < EP) {
      patch_of[cur] = p;
      P.elems.push_back(cur);
      ++count;
      // raise the score of unassigned neighbours
      const int32_t* ib = ibool + (size_t)cur * n2;
      for (int k = 0; k < n2; ++k) {
        const size_t nd = (size_t)ib[k] - 1;
        if (nstart[nd + 1] - nstart[nd] < 2) continue;
        for (int q = nstart[nd]; q < nstart[nd + 1]; ++q) {
          const int o = nelems[q];
          if (patch_of[o] >= 0) continue;
          if (score[o] == 0) frontier.push_back(o);
          score[o]++;
        }
      }
      // best candidate: highest score, ties -> lowest id
      cur = -1;
      int best = 0;
      for (size_t q = 0; q < frontier.size();) {
        const int o = frontier[q];
        if (patch_of[o] >= 0) {
          frontier[q] = frontier.back();
          frontier.pop_back();
          continue;
        }
        if (score[o] > best || (score[o] == best && o < cur)) {
          best = score[o];
          cur = o;
        }
        ++q;
      }
    }
    for (int o : frontier) score[o] = 0;
    P.pelem_start.push_back((int32_t)P.elems.size());
  }
  P.npatch = (int)P.pelem_start.size() - 1;
  // per-node: number of distinct patches touching it
  std::vector<int32_t> node_np(npoin, 0), node_last(npoin, -1);
  for (int p = 0; p < P.npatch; ++p)
    for (int q = P.pelem_start[p]; q < P.pelem_start[p + 1]; ++q) {
      const int32_t* ib = ibool + (size_t)P.elems[q] * n2;
      for (int k = 0; k < n2; ++k) {
        const size_t nd = (size_t)ib[k] - 1;
        if (node_last[nd] != p) {
          node_last[nd] = p;
          node_np[nd]++;
        }
      }
    }
  // shared nodes and their slot ranges (ascending node id; within a node ascending patch id
  // because patches are visited in ascending order below)
  std::vector<int32_t> sidx(npoin, -1);
  P.snode.clear();
  P.sstart.assign(1, 0);
  for (size_t nd = 0; nd < npoin; ++nd)
    if (node_np[nd] > 1) {
      sidx[nd] = (int32_t)P.snode.size();
      P.snode.push_back((int32_t)nd);
      P.sstart.push_back(P.sstart.back() + node_np[nd]);
    }
  P.nslots = (size_t)P.sstart.back();
  std::vector<int32_t> sfill(P.snode.size(), 0);
  // patch-local node lists, local indices, in-patch colours
  P.pnode_start.assign(1, 0);
  P.pnint.assign(P.npatch, 0);
  P.pnode.clear();
  P.pslot.clear();
  P.lidx.assign((size_t)nelem * n2, 0);
  P.ecolor.assign(nelem, 0);
  std::vector<int32_t> local_of(npoin, -1);
  std::vector<int32_t> touched;
  std::vector<uint32_t> cmask;
  P.max_nloc = 0;
  P.max_colors = 0;
  std::fill(node_last.begin(), node_last.end(), -1);
  for (int p = 0; p < P.npatch; ++p) {
    touched.clear();
    for (int q = P.pelem_start[p]; q < P.pelem_start[p + 1]; ++q) {
      const int32_t* ib = ibool + (size_t)P.elems[q] * n2;
      for (int k = 0; k < n2; ++k) {
        const size_t nd = (size_t)ib[k] - 1;
        if (node_last[nd] != p) {
          node_last[nd] = p;
          touched.push_back((int32_t)nd);
        }
      }
    }
    // private nodes first (in first-touch order, which follows the element numbering and keeps
    // runs of consecutive global ids together), then shared ones
    int nl = 0;
    for (int32_t nd : touched)
      if (node_np[nd] == 1) {
        local_of[nd] = nl++;
        P.pnode.push_back(nd);
        P.pslot.push_back(-1);
      }
    P.pnint[p] = nl;
    for (int32_t nd : touched)
      if (node_np[nd] > 1) {
        local_of[nd] = nl++;
        P.pnode.push_back(nd);
        const int s = sidx[nd];
        P.pslot.push_back(P.sstart[s] + sfill[s]++);
      }
    P.pnode_start.push_back((int32_t)P.pnode.size());
    P.max_nloc = std::max(P.max_nloc, nl);
    cmask.assign(nl, 0u);
    for (int q = P.pelem_start[p]; q < P.pelem_start[p + 1]; ++q) {
      const int32_t* ib = ibool + (size_t)P.elems[q] * n2;
      uint32_t m = 0;
      for (int k = 0; k < n2; ++k) {
        const int l = local_of[(size_t)ib[k] - 1];
        P.lidx[(size_t)q * n2 + k] = (uint16_t)l;
        m |= cmask[l];
      }
      int c = 0;
      while (c < 31 && ((m >> c) & 1u)) ++c;
      P.ecolor[q] = c;
      for (int k = 0; k < n2; ++k) cmask[local_of[(size_t)ib[k] - 1]] |= (1u << c);
      P.max_colors = std::max(P.max_colors, c + 1);
    }
  }
}

}  // namespace s2d
