// CPU check of the partition restatement used by k_match_filter (csrc/track.cu): for one quicksort step of libstdc++'s
// std::sort — __move_median_to_first + __unguarded_partition — the swaps are (Lo[k], Ro[k]) for k < K and the cut is
// min(Lo[K], Ro[K-1]) (Lo[0] when K = 0), where Lo / Ro list the positions of elements >= / <= pivot from the left /
// right in the segment BEFORE the step and K = #{k : Lo[k] < Ro[k]}.  Here that formula runs next to the real
// std::__unguarded_partition_pivot (bits/stl_algo.h) on the same segments, level by level like the kernel, and both the
// resulting arrays and the cuts must be identical.  Also reports how many levels the quicksort phase takes (the kernel
// declines beyond 2*floor(log2 n), where libstdc++ switches to heapsort).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
using namespace std;
typedef uint32_t E;                                    // (key << 16) | original position, compared on the key only
static inline uint32_t key(E e) { return e >> 16; }
struct Cmp { bool operator()(E a, E b) const { return key(a) < key(b); } };

static int restated_partition(vector<E> &arr, int first, int last) {
  const int ia = first + 1, ib = first + (last - first) / 2, ic = last - 1;
  const uint32_t ka = key(arr[ia]), kb = key(arr[ib]), kc = key(arr[ic]);
  int pick;
  if (ka < kb) pick = (kb < kc) ? ib : ((ka < kc) ? ic : ia);
  else pick = (ka < kc) ? ia : ((kb < kc) ? ic : ib);
  swap(arr[first], arr[pick]);
  const uint32_t pivot = key(arr[first]);
  const int lo = first + 1, len = last - lo;
  vector<int> Lo, Ro;
  for (int i = 0; i < len; ++i) {
    if (key(arr[lo + i]) >= pivot) Lo.push_back(lo + i);
    if (key(arr[last - 1 - i]) <= pivot) Ro.push_back(last - 1 - i);
  }
  const int nmin = (int)min(Lo.size(), Ro.size());
  int K = 0;
  while (K < nmin && Lo[K] < Ro[K]) ++K;
  for (int j = 0; j < K; ++j) swap(arr[Lo[j]], arr[Ro[j]]);
  if (K == 0) return Lo.empty() ? last : Lo[0];
  int cut = Ro[K - 1];
  if (K < (int)Lo.size()) cut = min(cut, Lo[K]);
  return cut;
}

int main() {
  mt19937 rng(3);
  int bad = 0, over_limit = 0, trials = 0;
  for (int trial = 0; trial < 4000; ++trial, ++trials) {
    const int n = 17 + rng() % 3000, mode = trial % 7;
    vector<E> a(n);
    for (int i = 0; i < n; ++i) {
      uint32_t k;
      switch (mode) {
        case 0: k = rng() % 2001; break;                       // random, few duplicates
        case 1: k = rng() % 50; break;                         // heavy duplication
        case 2: k = i / 2 + (rng() % 5); break;                // nearly sorted with duplicates
        case 3: k = (n - i) / 3; break;                        // descending runs of equal keys
        case 4: k = min(i, n - 1 - i); break;                  // organ pipe (median-of-3 killer)
        case 5: k = 7; break;                                  // constant
        default: k = (i % (n / 4 + 1)) * 3 + rng() % 3; break; // four concatenated ascending runs (level-major map order)
      }
      a[i] = (k << 16) | (uint32_t)i;
    }
    vector<E> b = a;
    vector<pair<int, int>> segs = {{0, n}};
    int level = 0;
    while (!segs.empty()) {
      vector<pair<int, int>> nxt;
      for (auto [f, l] : segs) {
        const int c1 = restated_partition(a, f, l);
        const int c2 = (int)(std::__unguarded_partition_pivot(b.begin() + f, b.begin() + l, __gnu_cxx::__ops::__iter_comp_iter(Cmp())) - b.begin());
        if (c1 != c2 || !equal(a.begin() + f, a.begin() + l, b.begin() + f)) {
          if (bad < 5) printf("MISMATCH trial %d mode %d n %d segment [%d,%d): cut %d vs %d\n", trial, mode, n, f, l, c1, c2);
          ++bad;
          a = b;
        }
        if (l - c2 > 16) nxt.push_back({c2, l});
        if (c2 - f > 16) nxt.push_back({f, c2});
      }
      segs.swap(nxt);
      ++level;
    }
    int lg = 0;
    for (int m = n; m > 1; m >>= 1) ++lg;
    if (level > 2 * lg) ++over_limit;
  }
  printf("trials %d mismatches %d beyond_depth_limit %d\n", trials, bad, over_limit);
  return bad ? 1 : 0;
}
