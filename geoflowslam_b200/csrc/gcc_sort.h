// Restatement of libstdc++'s std::sort (introsort: median-of-3 quicksort to depth 2*floor(log2 n),
// heapsort fallback, final insertion sort with a 16-element threshold) as one host/device
// function.  std::sort is NOT stable, and the reference's quadtree (DistributeOctTree,
// src/ORBextractor.cc:686-688) sorts (size, node) pairs whose comparator (compareNodes, :552-565)
// ties constantly -- the processing order of tied nodes decides which nodes get split before the
// feature budget is hit.  Bit-exact keypoint selection therefore needs the exact permutation the
// reference's GCC build produces; this file reproduces that algorithm step for step.
//
// Elements are (first, second) int pairs held in two parallel arrays.  `less(a1, a2, b1, b2)`
// compares element a with element b.
#pragma once

#if defined(__CUDACC__)
#define GFS_HD __host__ __device__ __forceinline__
#else
#define GFS_HD inline
#endif

namespace gfs {

struct PairArr {
  int* f;  // pair.first  (node size)
  int* s;  // pair.second (node handle)
};

template <class Less>
struct GccSort {
  PairArr a;
  Less less;

  GFS_HD bool lt(int i, int j) const { return less(a.f[i], a.s[i], a.f[j], a.s[j]); }
  GFS_HD bool lt_val(int vf, int vs, int j) const { return less(vf, vs, a.f[j], a.s[j]); }
  GFS_HD bool lt_idx_val(int i, int vf, int vs) const { return less(a.f[i], a.s[i], vf, vs); }
  GFS_HD void swp(int i, int j) {
    int t = a.f[i]; a.f[i] = a.f[j]; a.f[j] = t;
    t = a.s[i]; a.s[i] = a.s[j]; a.s[j] = t;
  }
  GFS_HD void mov(int dst, int src) { a.f[dst] = a.f[src]; a.s[dst] = a.s[src]; }

  // std::__move_median_to_first(result, a, b, c)
  GFS_HD void median_to_first(int r, int x, int y, int z) {
    if (lt(x, y)) {
      if (lt(y, z)) swp(r, y);
      else if (lt(x, z)) swp(r, z);
      else swp(r, x);
    } else if (lt(x, z)) swp(r, x);
    else if (lt(y, z)) swp(r, z);
    else swp(r, y);
  }
  // std::__unguarded_partition(first, last, pivot)
  GFS_HD int unguarded_partition(int first, int last, int pivot) {
    while (true) {
      while (lt(first, pivot)) ++first;
      --last;
      while (lt(pivot, last)) --last;
      if (!(first < last)) return first;
      swp(first, last);
      ++first;
    }
  }
  // std::__push_heap on [first, ...)
  GFS_HD void push_heap(int first, int hole, int top, int vf, int vs) {
    int parent = (hole - 1) / 2;
    while (hole > top && lt_idx_val(first + parent, vf, vs)) {
      mov(first + hole, first + parent);
      hole = parent;
      parent = (hole - 1) / 2;
    }
    a.f[first + hole] = vf;
    a.s[first + hole] = vs;
  }
  // std::__adjust_heap
  GFS_HD void adjust_heap(int first, int hole, int len, int vf, int vs) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
      child = 2 * (child + 1);
      if (lt(first + child, first + (child - 1))) child--;
      mov(first + hole, first + child);
      hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
      child = 2 * (child + 1);
      mov(first + hole, first + (child - 1));
      hole = child - 1;
    }
    push_heap(first, hole, top, vf, vs);
  }
  // std::__partial_sort(first, last, last) == make_heap + sort_heap
  GFS_HD void heap_sort(int first, int last) {
    const int len = last - first;
    if (len >= 2) {
      int parent = (len - 2) / 2;
      while (true) {
        int vf = a.f[first + parent], vs = a.s[first + parent];
        adjust_heap(first, parent, len, vf, vs);
        if (parent == 0) break;
        parent--;
      }
    }
    while (last - first > 1) {
      --last;
      int vf = a.f[last], vs = a.s[last];
      mov(last, first);
      adjust_heap(first, 0, last - first, vf, vs);
    }
  }
  // std::__unguarded_linear_insert
  GFS_HD void unguarded_linear_insert(int last) {
    int vf = a.f[last], vs = a.s[last];
    int next = last - 1;
    while (lt_val(vf, vs, next)) {
      mov(last, next);
      last = next;
      --next;
    }
    a.f[last] = vf;
    a.s[last] = vs;
  }
  // std::__insertion_sort
  GFS_HD void insertion_sort(int first, int last) {
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
      if (lt(i, first)) {
        int vf = a.f[i], vs = a.s[i];
        for (int k = i; k > first; --k) mov(k, k - 1);
        a.f[first] = vf;
        a.s[first] = vs;
      } else {
        unguarded_linear_insert(i);
      }
    }
  }

  // std::sort(first, first + n).  `heapsorted` (optional) is set when the depth limit fired.
  GFS_HD void sort(int n, int* heapsorted = nullptr) {
    if (n <= 0) return;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;  // std::__lg
    // __introsort_loop with an explicit stack (the recursion only ever handles disjoint
    // ranges, so the order in which pending ranges are processed does not change the result)
    int stF[64], stL[64], stD[64];
    int sp = 0;
    stF[0] = 0; stL[0] = n; stD[0] = 2 * lg; sp = 1;
    while (sp > 0) {
      --sp;
      int first = stF[sp], last = stL[sp], depth = stD[sp];
      while (last - first > 16) {
        if (depth == 0) {
          heap_sort(first, last);
          if (heapsorted) *heapsorted = 1;
          break;
        }
        --depth;
        int mid = first + (last - first) / 2;
        median_to_first(first, first + 1, mid, last - 1);
        int cut = unguarded_partition(first + 1, last, first);
        stF[sp] = cut; stL[sp] = last; stD[sp] = depth; sp++;
        last = cut;
      }
    }
    // __final_insertion_sort
    if (n > 16) {
      insertion_sort(0, 16);
      for (int i = 16; i != n; ++i) unguarded_linear_insert(i);
    } else {
      insertion_sort(0, n);
    }
  }
};

}  // namespace gfs
