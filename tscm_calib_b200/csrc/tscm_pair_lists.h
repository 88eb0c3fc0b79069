// tscm_pair_lists.h — host-side work lists of the per-camera-pair Schur update
// (tscm_schur_pairs.cuh).  Plain C++ (no CUDA types) so that the CPU test-suite can check it.
//
// For every camera pair (a <= b), in pair order (0,0), (0,1), ... the frames seen by BOTH
// cameras give one entry (view of a, view of b) each; entries are cut into work items of at
// most `chunk` consecutive common frames; items are then ordered by the first frame they
// touch (stable, so a pair's items keep their frame order) — the warps running at the same
// time work on the same stretch of frames and share its per-view blocks through L2.
#pragma once

#include <algorithm>
#include <cstddef>
#include <vector>

namespace tscm {

struct PairEntry { int view_a, view_b; };   // layout of int2
struct PairRange { int begin, end; };       // layout of int2: entries [begin, end) of an item

struct PairLists {
  std::vector<PairEntry> ent;        // all entries, pair-major, frames increasing within a pair
  std::vector<PairRange> item_range; // [nitems] in launch order
  std::vector<int> pair_item;        // [npairs + 1] range of a pair in pair_items
  std::vector<int> pair_items;       // launch positions of a pair's items, in frame order
  std::vector<short> pair_a, pair_b; // [npairs]
};

inline PairLists build_pair_lists(int C, int F, int V, const int* view_camera, const int* view_frame,
                                  int chunk) {
  PairLists L;
  std::vector<PairRange> raw_range;
  std::vector<int> raw_pair, raw_first;
  std::vector<int> view_of((size_t)C * F, -1);   // view of camera m in frame f, or -1
  for (int v = 0; v < V; ++v) view_of[(size_t)view_camera[v] * F + view_frame[v]] = v;
  for (int a = 0; a < C; ++a)
    for (int b = a; b < C; ++b) {
      const int pr = (int)L.pair_a.size();
      L.pair_a.push_back((short)a);
      L.pair_b.push_back((short)b);
      int in_item = 0;
      const int* va = &view_of[(size_t)a * F];
      const int* vb = &view_of[(size_t)b * F];
      for (int f = 0; f < F; ++f) {
        if (va[f] < 0 || vb[f] < 0) continue;
        if (in_item == 0) {
          raw_range.push_back(PairRange{(int)L.ent.size(), (int)L.ent.size()});
          raw_pair.push_back(pr);
          raw_first.push_back(f);
        }
        L.ent.push_back(PairEntry{va[f], vb[f]});
        raw_range.back().end = (int)L.ent.size();
        if (++in_item == chunk) in_item = 0;
      }
    }
  const int nitems = (int)raw_range.size();
  std::vector<int> order(nitems);
  for (int k = 0; k < nitems; ++k) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return raw_first[x] < raw_first[y]; });
  L.item_range.resize(nitems);
  std::vector<std::vector<int>> items_of(L.pair_a.size());
  for (int pos = 0; pos < nitems; ++pos) {
    L.item_range[pos] = raw_range[order[pos]];
    items_of[raw_pair[order[pos]]].push_back(pos);
  }
  for (size_t pr = 0; pr < L.pair_a.size(); ++pr) {
    L.pair_item.push_back((int)L.pair_items.size());
    L.pair_items.insert(L.pair_items.end(), items_of[pr].begin(), items_of[pr].end());
  }
  L.pair_item.push_back((int)L.pair_items.size());
  return L;
}

}  // namespace tscm
