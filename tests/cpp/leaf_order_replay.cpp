// tests/cpp/leaf_order_replay.cpp - selection_sort_replay (csrc/scene_builder.cpp: the O(n log n) replay of the reference's selection sort,
// bvh.cuh:46-81) against the plain loop on random and tie-heavy keys, sub-ranges included. Built and run by tests/test_host.py.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "scene_builder.cpp"
using namespace rt;
int main() {
  std::mt19937 rng(7);
  int bad = 0;
  for (int trial = 0; trial < 300; ++trial) {
    const int n = 2 + (int)(rng() % 3000);
    SceneDesc sd; sd.obj.resize(n);
    const int mode = trial % 4;
    for (int i = 0; i < n; ++i) {
      float k = mode == 0 ? (float)(rng() % 1000000) / 7.0f : mode == 1 ? (float)(rng() % 17) : mode == 2 ? (float)(rng() % 3) - 1.0f : (float)((rng() % 50) * 0.25f);
      if (mode == 2 && rng() % 5 == 0) k = -0.0f;
      for (int a = 0; a < 3; ++a) sd.obj[i].box_min[a] = k;
    }
    std::vector<int> a(n), b(n);
    for (int i = 0; i < n; ++i) a[i] = b[i] = i;
    std::shuffle(a.begin(), a.end(), rng); b = a;
    const int s0 = (int)(rng() % (n / 2 + 1)), e0 = n - (int)(rng() % (n / 4 + 1));
    selection_sort_replay(a, sd, s0, e0, 1);
    for (int i = s0; i < e0 - 1; ++i) { int best = i; for (int j = i + 1; j < e0; ++j) if (sd.obj[b[j]].box_min[1] < sd.obj[b[best]].box_min[1]) best = j; if (best != i) std::swap(b[i], b[best]); }
    if (a != b) { ++bad; printf("mismatch trial %d n %d mode %d\n", trial, n, mode); }
  }
  printf("sorttest: %d mismatches\n", bad);
  return bad != 0;
}
