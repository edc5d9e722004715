// sort_replay.h — replays what libstdc++'s std::sort does to a sequence, step for step.
//
// pcl::EuclideanClusterExtraction::extract (called at reference src:229 and src:276) ends with
//     std::sort(clusters.rbegin(), clusters.rend(), comparePointClusters)   // a.size() < b.size()
// which is NOT a stable sort: for more than 16 clusters the order of equal-size clusters is
// whatever libstdc++'s introsort leaves (SURVEY.md §7 hard-part 2).  Keypoint order is
// observable (it selects the RNG draws of the 3DSC azimuth origin), so the device code runs the
// same algorithm: introsort loop (threshold 16, median-of-3 moved to first, unguarded Hoare
// partition, depth limit 2*floor(log2 n) then heap sort) followed by the final insertion sort.
// Written from the published algorithm; compiled for host and device.
#ifndef FE_SORT_REPLAY_H_
#define FE_SORT_REPLAY_H_

#if defined(__CUDACC__)
#define FE_HD __host__ __device__ __forceinline__
#else
#define FE_HD inline
#endif
#if defined(__CUDACC__)
#pragma nv_diag_suppress 20013, 20015  // host lambdas are only ever called from the host instantiation
#endif

namespace fe {

// Acc: struct with  V get(int i) const;  void set(int i, V v);  bool less(V a, V b) const;
// where index i addresses the sequence in the order std::sort sees it.
template <class Acc, class V>
struct SortReplay {
  Acc& a;
  FE_HD explicit SortReplay(Acc& acc) : a(acc) {}

  FE_HD void swap_at(int i, int j) { V t = a.get(i); a.set(i, a.get(j)); a.set(j, t); }

  FE_HD void unguarded_linear_insert(int last) {
    V val = a.get(last);
    int next = last - 1;
    while (a.less(val, a.get(next))) { a.set(last, a.get(next)); last = next; --next; }
    a.set(last, val);
  }
  FE_HD void insertion_sort(int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
      if (a.less(a.get(i), a.get(first))) {
        V val = a.get(i);
        for (int k = i; k > first; --k) a.set(k, a.get(k - 1));
        a.set(first, val);
      } else {
        unguarded_linear_insert(i);
      }
    }
  }
  FE_HD void final_insertion_sort(int first, int last) {
    if (last - first > 16) {
      insertion_sort(first, first + 16);
      for (int i = first + 16; i != last; ++i) unguarded_linear_insert(i);
    } else {
      insertion_sort(first, last);
    }
  }
  FE_HD void move_median_to_first(int result, int x, int y, int z) {
    V vx = a.get(x), vy = a.get(y), vz = a.get(z);
    if (a.less(vx, vy)) {
      if (a.less(vy, vz)) swap_at(result, y);
      else if (a.less(vx, vz)) swap_at(result, z);
      else swap_at(result, x);
    } else if (a.less(vx, vz)) swap_at(result, x);
    else if (a.less(vy, vz)) swap_at(result, z);
    else swap_at(result, y);
  }
  FE_HD int unguarded_partition(int first, int last, int pivot) {
    for (;;) {
      while (a.less(a.get(first), a.get(pivot))) ++first;
      --last;
      while (a.less(a.get(pivot), a.get(last))) --last;
      if (!(first < last)) return first;
      swap_at(first, last);
      ++first;
    }
  }
  FE_HD int unguarded_partition_pivot(int first, int last) {
    int mid = first + (last - first) / 2;
    move_median_to_first(first, first + 1, mid, last - 1);
    return unguarded_partition(first + 1, last, first);
  }
  // heap sort of [first, first+len) (std::partial_sort(first, last, last))
  FE_HD void push_heap(int first, int hole, int top, V value) {
    int parent = (hole - 1) / 2;
    while (hole > top && a.less(a.get(first + parent), value)) {
      a.set(first + hole, a.get(first + parent));
      hole = parent;
      parent = (hole - 1) / 2;
    }
    a.set(first + hole, value);
  }
  FE_HD void adjust_heap(int first, int hole, int len, V value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (a.less(a.get(first + child), a.get(first + (child - 1)))) child--;
      a.set(first + hole, a.get(first + child));
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      a.set(first + hole, a.get(first + (child - 1)));
      hole = child - 1;
    }
    push_heap(first, hole, top, value);
  }
  FE_HD void heap_sort(int first, int last) {
    const int len = last - first;
    if (len >= 2) {
      int parent = (len - 2) / 2;
      for (;;) {
        V value = a.get(first + parent);
        adjust_heap(first, parent, len, value);
        if (parent == 0) break;
        parent--;
      }
    }
    while (last - first > 1) {
      --last;
      V value = a.get(last);
      a.set(last, a.get(first));
      adjust_heap(first, 0, last - first, value);
    }
  }
  FE_HD static int floor_log2(int n) { int l = 0; while (n > 1) { n >>= 1; l++; } return l; }

  // std::sort(first, last) over positions [0, n)
  FE_HD void sort(int n) {
    if (n <= 0) return;
    // explicit stack instead of the recursion on the right part; the parts are disjoint, so the
    // order in which they are finished does not change the outcome
    int st_first[64], st_last[64], st_depth[64];
    int sp = 0;
    st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * floor_log2(n); sp = 1;
    while (sp > 0) {
      --sp;
      int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
      while (last - first > 16) {
        if (depth == 0) { heap_sort(first, last); break; }
        --depth;
        int cut = unguarded_partition_pivot(first, last);
        if (sp < 64) { st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; ++sp; }
        last = cut;
      }
    }
    final_insertion_sort(0, n);
  }
};

// The cluster order PCL ends with: `ids[0..n)` are clusters in discovery order, `size_of(id)`
// their sizes.  std::sort runs over the REVERSED sequence with size<.
template <class IdT, class SizeFn>
struct ReversedClusterAcc {
  IdT* ids;
  int n;
  SizeFn size_of;
  FE_HD IdT get(int i) const { return ids[n - 1 - i]; }
  FE_HD void set(int i, IdT v) { ids[n - 1 - i] = v; }
  FE_HD bool less(IdT x, IdT y) const { return size_of(x) < size_of(y); }
};

template <class IdT, class SizeFn>
FE_HD void pcl_cluster_order(IdT* ids, int n, SizeFn size_of) {
  ReversedClusterAcc<IdT, SizeFn> acc = {ids, n, size_of};
  SortReplay<ReversedClusterAcc<IdT, SizeFn>, IdT> s(acc);
  s.sort(n);
}

}  // namespace fe
#endif
