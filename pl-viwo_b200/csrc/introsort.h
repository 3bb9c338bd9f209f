// libstdc++'s std::sort, restated so that it can run inside a kernel.
//
// Grider_GRID::perform_griding sorts every cell's FAST corners with std::sort(begin, end, compare_response)
// (open_vins/ov_core/src/track/Grider_GRID.h:128, comparator Grider_FAST.h:57: first.response > second.response) and keeps
// the first num_features_grid.  FAST responses are small integers, so the cut usually falls inside a group of equal
// responses and WHICH of them survive is decided by the exact sequence of swaps of libstdc++'s introsort (the sort is
// not stable).  A different sort — stable, bitonic, radix — keeps different corners and every feature id after that
// differs from the reference.  This header therefore follows bits/stl_algo.h / bits/stl_heap.h step by step:
// introsort loop (median-of-3 moved to the front, unguarded Hoare partition, depth limit 2 lg n, heap sort fallback),
// then the final insertion sort with the 16-element threshold.  The permutation depends only on the comparator's
// answers, so sorting 4-byte packed corners instead of 28-byte cv::KeyPoints gives the same order.
// tests/test_host_cpu.py checks the host instantiation against the real std::sort on tie-heavy inputs; the device
// instantiation is checked against the host one on the GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PLVIWO_HD __host__ __device__ __forceinline__
#else
#define PLVIWO_HD inline
#endif

namespace plviwo {
namespace isort {

constexpr int kThreshold = 16;   // std::_S_threshold

// packed corner: x | y << 12 | score << 24; compare_response(first, second) == first.response > second.response
PLVIWO_HD bool comp(unsigned a, unsigned b) { return (a >> 24) > (b >> 24); }

PLVIWO_HD void swap_at(unsigned *v, int a, int b) {
  unsigned t = v[a];
  v[a] = v[b];
  v[b] = t;
}

// std::__move_median_to_first(result, a, b, c)
PLVIWO_HD void move_median_to_first(unsigned *v, int result, int a, int b, int c) {
  if (comp(v[a], v[b])) {
    if (comp(v[b], v[c])) swap_at(v, result, b);
    else if (comp(v[a], v[c])) swap_at(v, result, c);
    else swap_at(v, result, a);
  } else if (comp(v[a], v[c])) swap_at(v, result, a);
  else if (comp(v[b], v[c])) swap_at(v, result, c);
  else swap_at(v, result, b);
}

// std::__unguarded_partition(first, last, pivot)
PLVIWO_HD int unguarded_partition(unsigned *v, int first, int last, int pivot) {
  while (true) {
    while (comp(v[first], v[pivot])) ++first;
    --last;
    while (comp(v[pivot], v[last])) --last;
    if (!(first < last)) return first;
    swap_at(v, first, last);
    ++first;
  }
}

// std::__adjust_heap followed by std::__push_heap (bits/stl_heap.h), on the range starting at v
PLVIWO_HD void adjust_heap(unsigned *v, int hole, int len, unsigned value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (comp(v[child], v[child - 1])) child--;
    v[hole] = v[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    v[hole] = v[child - 1];
    hole = child - 1;
  }
  int parent = (hole - 1) / 2;
  while (hole > top && comp(v[parent], value)) {
    v[hole] = v[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  v[hole] = value;
}

// std::__partial_sort(first, last, last): make_heap + sort_heap
PLVIWO_HD void heap_sort(unsigned *v, int len) {
  if (len >= 2) {
    int parent = (len - 2) / 2;
    while (true) {
      adjust_heap(v, parent, len, v[parent]);
      if (parent == 0) break;
      parent--;
    }
  }
  int last = len;
  while (last > 1) {
    --last;
    unsigned value = v[last];
    v[last] = v[0];
    adjust_heap(v, 0, last, value);
  }
}

// std::__unguarded_linear_insert(last)
PLVIWO_HD void unguarded_linear_insert(unsigned *v, int last) {
  unsigned val = v[last];
  int next = last - 1;
  while (comp(val, v[next])) {
    v[last] = v[next];
    last = next;
    --next;
  }
  v[last] = val;
}

// std::__insertion_sort(first, last)
PLVIWO_HD void insertion_sort(unsigned *v, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (comp(v[i], v[first])) {
      unsigned val = v[i];
      for (int k = i; k > first; --k) v[k] = v[k - 1];   // std::move_backward(first, i, i + 1)
      v[first] = val;
    } else {
      unguarded_linear_insert(v, i);
    }
  }
}

PLVIWO_HD int lg(int n) {   // std::__lg
  int k = 0;
  while (n > 1) { n >>= 1; k++; }
  return k;
}

// std::sort(v, v + n, compare_response).  The recursion of __introsort_loop (recurse into the right part, loop on the
// left part) is unrolled with an explicit stack; the parts are disjoint, so the order in which they are finished does
// not change the result.
PLVIWO_HD void sort(unsigned *v, int n) {
  if (n <= 0) return;
  int stack_first[64], stack_last[64], stack_depth[64];
  int sp = 0;
  stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = 2 * lg(n);
  sp = 1;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > kThreshold) {
      if (depth == 0) {
        heap_sort(v + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      move_median_to_first(v, first, first + 1, mid, last - 1);
      const int cut = unguarded_partition(v, first + 1, last, first);
      stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth;   // the recursive call
      sp++;
      last = cut;
    }
  }
  if (n > kThreshold) {   // std::__final_insertion_sort
    insertion_sort(v, 0, kThreshold);
    for (int i = kThreshold; i != n; ++i) unguarded_linear_insert(v, i);
  } else {
    insertion_sort(v, 0, n);
  }
}

// The first `keep` elements of std::sort(v, v + n, compare_response), without sorting the rest (v[keep ..] is left in an
// unspecified order).  Exactly the permutation std::sort produces for those positions, at the cost of a selection instead
// of a sort: a partition step leaves every element of the left part "not less" than every element of the right part, and
// neither the later partition steps of one part nor the final insertion sort (which moves an element left only past
// STRICTLY smaller ones) ever carry an element across that boundary.  So a right part that starts at or beyond position
// `keep` cannot influence the first `keep` outputs and its recursive call is skipped; the final insertion sort stops at the
// first skipped boundary.  The depth counter runs as in the full sort (it only depends on the path to the left parts); if
// it hits zero the part is heap-sorted as a whole, as std::sort does.
PLVIWO_HD void sort_prefix(unsigned *v, int n, int keep) {
  if (n <= 0) return;
  if (keep >= n) {
    sort(v, n);
    return;
  }
  int stack_first[64], stack_last[64], stack_depth[64];
  int sp = 0;
  int bound = n;   // [0, bound) has gone through the whole introsort loop
  stack_first[0] = 0; stack_last[0] = n; stack_depth[0] = 2 * lg(n);
  sp = 1;
  while (sp > 0) {
    --sp;
    int first = stack_first[sp], last = stack_last[sp], depth = stack_depth[sp];
    while (last - first > kThreshold) {
      if (depth == 0) {
        heap_sort(v + first, last - first);
        break;
      }
      --depth;
      const int mid = first + (last - first) / 2;
      move_median_to_first(v, first, first + 1, mid, last - 1);
      const int cut = unguarded_partition(v, first + 1, last, first);
      if (cut < keep) {   // the right part holds wanted positions: the recursive call
        stack_first[sp] = cut; stack_last[sp] = last; stack_depth[sp] = depth;
        sp++;
      } else if (cut < bound) {
        bound = cut;
      }
      last = cut;
    }
  }
  if (bound > kThreshold) {   // std::__final_insertion_sort, restricted to [0, bound)
    insertion_sort(v, 0, kThreshold);
    for (int i = kThreshold; i != bound; ++i) unguarded_linear_insert(v, i);
  } else {
    insertion_sort(v, 0, bound);
  }
}

}  // namespace isort
}  // namespace plviwo
