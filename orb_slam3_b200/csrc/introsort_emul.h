// Re-implementation of libstdc++'s std::sort control flow (introsort: median-of-3
// unguarded partition down to 16-element runs, heapsort when the depth limit is
// hit, then one final insertion sort) for use inside the GPU octree cull.
//
// Why: ORBextractor::DistributeOctTree sorts its expandable nodes with an
// unstable std::sort whose comparator ties on (count, UL.x)
// (reference src/ORBextractor.cc:538-553, :700).  Which of several tied nodes is
// split last decides the surviving keypoint set, so bit-exact parity needs the
// same permutation the reference's libstdc++ produces.  The sequence of
// comparisons and moves below is that of GCC's bits/stl_algo.h / stl_heap.h
// (__introsort_loop, __unguarded_partition_pivot, __final_insertion_sort,
// __heap_select + __sort_heap); it is pinned against the host's real std::sort
// in tests/test_introsort.py.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORB_HD __host__ __device__ __forceinline__
#else
#define ORB_HD inline
#endif

namespace orbb200 {

// One sortable record: the ordering key (count, UL.x) packed into one word (count in the upper 20 bits, UL.x
// < 4096 in the lower 12: one integer compare, and records that tie in the reference's comparator have equal
// keys here too); `id` is the payload.  Eight bytes: one 64-bit load / store per move.
struct __attribute__((aligned(8))) SortNode {
  uint32_t key;
  int id;
};

ORB_HD SortNode make_sort_node(int count, int ulx, int id) {
  SortNode s;
  s.key = ((uint32_t)count << 12) | (uint32_t)(ulx & 0xfff);
  s.id = id;
  return s;
}

ORB_HD bool node_less(const SortNode& a, const SortNode& b) { return a.key < b.key; }

ORB_HD void sn_swap(SortNode& a, SortNode& b) {
  SortNode t = a;
  a = b;
  b = t;
}

// ---- heap helpers (stl_heap.h: __push_heap / __adjust_heap) ----
ORB_HD void sn_push_heap(SortNode* first, int hole, int top, SortNode value) {
  int parent = (hole - 1) / 2;
  while (hole > top && node_less(first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

ORB_HD void sn_adjust_heap(SortNode* first, int hole, int len, SortNode value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (node_less(first[child], first[child - 1])) child--;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  sn_push_heap(first, hole, top, value);
}

// __partial_sort(first, last, last) == __heap_select(first,last,last) + __sort_heap
ORB_HD void sn_heap_sort(SortNode* first, int len) {
  if (len < 2) return;
  // __make_heap
  for (int parent = (len - 2) / 2;; parent--) {
    SortNode v = first[parent];
    sn_adjust_heap(first, parent, len, v);
    if (parent == 0) break;
  }
  // __heap_select's scan over [middle,last) is empty (middle == last)
  // __sort_heap
  for (int last = len; last > 1;) {
    --last;
    SortNode v = first[last];  // __pop_heap(first, last, last)
    first[last] = first[0];
    sn_adjust_heap(first, 0, last, v);
  }
}

ORB_HD void sn_move_median_to_first(SortNode* a, int result, int ia, int ib, int ic) {
  if (node_less(a[ia], a[ib])) {
    if (node_less(a[ib], a[ic])) sn_swap(a[result], a[ib]);
    else if (node_less(a[ia], a[ic])) sn_swap(a[result], a[ic]);
    else sn_swap(a[result], a[ia]);
  } else if (node_less(a[ia], a[ic])) sn_swap(a[result], a[ia]);
  else if (node_less(a[ib], a[ic])) sn_swap(a[result], a[ic]);
  else sn_swap(a[result], a[ib]);
}

ORB_HD int sn_unguarded_partition(SortNode* a, int first, int last, int pivot) {
  while (true) {
    while (node_less(a[first], a[pivot])) ++first;
    --last;
    while (node_less(a[pivot], a[last])) --last;
    if (!(first < last)) return first;
    sn_swap(a[first], a[last]);
    ++first;
  }
}

ORB_HD void sn_unguarded_linear_insert(SortNode* a, int last) {
  SortNode val = a[last];
  int next = last - 1;
  while (node_less(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

ORB_HD void sn_insertion_sort(SortNode* a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (node_less(a[i], a[first])) {
      SortNode val = a[i];
      for (int k = i; k > first; --k) a[k] = a[k - 1];  // move_backward
      a[first] = val;
    } else {
      sn_unguarded_linear_insert(a, i);
    }
  }
}

// std::sort(a, a+n, compareNodes).  `stack` needs 2*64 ints of scratch.
ORB_HD void introsort_emul(SortNode* a, int n) {
  if (n <= 1) return;
  int lg = 0;
  for (int t = n; t > 1; t >>= 1) lg++;
  // explicit stack replaces the recursion __introsort_loop(cut, last, depth)
  int stk_first[64], stk_last[64], stk_depth[64];
  int sp = 0;
  stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = 2 * lg; sp = 1;
  while (sp > 0) {
    --sp;
    int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
    // libstdc++ recurses on the right part first and loops on the left part;
    // the two parts are disjoint so finishing the right part before continuing
    // with the left one is the same sequence of operations per sub-range.
    while (last - first > 16) {
      if (depth == 0) {
        sn_heap_sort(a + first, last - first);
        break;
      }
      --depth;
      int mid = first + (last - first) / 2;
      sn_move_median_to_first(a, first, first + 1, mid, last - 1);
      int cut = sn_unguarded_partition(a, first + 1, last, first);
      // "recurse" on [cut,last): push the left remainder, continue with right
      stk_first[sp] = first; stk_last[sp] = cut; stk_depth[sp] = depth; ++sp;
      first = cut;
    }
  }
  // __final_insertion_sort
  if (n > 16) {
    sn_insertion_sort(a, 0, 16);
    for (int i = 16; i < n; ++i) sn_unguarded_linear_insert(a, i);
  } else {
    sn_insertion_sort(a, 0, n);
  }
}

// ---- the same sort, level-synchronous ------------------------------------------------------------------------
// introsort's recursion works on disjoint sub-ranges, and so does the final insertion pass: unguarded_partition
// leaves every element left of a cut <= every element right of it, so an insertion never moves an element across
// a cut (the scan stops at the first element that is not greater -- the left neighbour of the range at the
// latest).  The operations of different sub-ranges therefore commute, and the permutation std::sort produces is
// reproduced by ANY schedule that applies, per sub-range, the same operations in the same order.  Here all
// sub-ranges of one recursion depth are processed together, one thread per sub-range: a range of more than 16
// elements is partitioned exactly like __introsort_loop does (median of three to the front, unguarded partition,
// heapsort at the depth limit) and hands its two halves to the next level; a range of at most 16 elements gets
// its share of __final_insertion_sort at once (guarded for the range that starts the array, bounded by the
// range start otherwise, which is where the unguarded scan stops anyway).  The critical path is one partition
// per level (n, n/2, n/4, ... elements) instead of every partition and every insertion of the array in turn.
// `work` = 6 * cap ints (two lists of (first, last, depth)), cap >= n / 8 + 4; all threads of the backend call it.
ORB_HD void sn_bounded_linear_insert(SortNode* a, int last, int lo) {
  SortNode val = a[last];
  int next = last - 1;
  while (next >= lo && node_less(val, a[next])) {
    a[last] = a[next];
    last = next;
    --next;
  }
  a[last] = val;
}

#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class BE>
#ifdef __CUDACC__
__host__ __device__
#endif
void introsort_levels(BE& be, SortNode* a, int n, int* work, int cap) {
  if (n <= 1) return;
  const int tid = be.tid(), nt = be.nthreads();
  int* cnt[2] = {be.shared_int(1), be.shared_int(2)};
  int* list[2] = {work, work + 3 * cap};
  if (tid == 0) {
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    list[0][0] = 0; list[0][1] = n; list[0][2] = 2 * lg;
    *cnt[0] = 1; *cnt[1] = 0;
  }
  be.sync();
  for (int level = 0;; level++) {
    const int c = level & 1;
    const int count = *cnt[c];
    if (count == 0) break;
    const int* cur = list[c];
    int* nxt = list[c ^ 1];
    for (int r = tid; r < count; r += nt) {
      const int first = cur[3 * r], last = cur[3 * r + 1];
      int depth = cur[3 * r + 2];
      if (last - first > 16) {
        if (depth == 0) {
          sn_heap_sort(a + first, last - first);
        } else {
          --depth;
          const int mid = first + (last - first) / 2;
          sn_move_median_to_first(a, first, first + 1, mid, last - 1);
          const int cut = sn_unguarded_partition(a, first + 1, last, first);
          const int slot = be.atomic_add(cnt[c ^ 1], 2);
          nxt[3 * slot] = first;   nxt[3 * slot + 1] = cut;  nxt[3 * slot + 2] = depth;
          nxt[3 * slot + 3] = cut; nxt[3 * slot + 4] = last; nxt[3 * slot + 5] = depth;
        }
      } else if (first == 0) {
        sn_insertion_sort(a, 0, last);
      } else {
        for (int i = first; i < last; ++i) sn_bounded_linear_insert(a, i, first);
      }
    }
    be.sync();
    if (tid == 0) *cnt[c] = 0;
    be.sync();
  }
}

}  // namespace orbb200
