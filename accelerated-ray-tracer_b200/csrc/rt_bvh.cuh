// rt_bvh.cuh — device-side construction of the 4-wide BVH over the top-level objects.
//
// Replaces bvh_node's constructor (bvh.cuh:29-84): one GPU thread, recursive `new`, O(n^2) selection
// sort per level. Here: 63-bit Morton codes of the box centroids, radix sort (CUB), then PLOC
// (parallel locally-ordered clustering, Meister & Bittner 2018): every cluster looks RT_PLOC_RADIUS
// places left and right along the Morton order for the partner with the smallest merged surface
// area, mutual nearest neighbours merge, the array is compacted, repeat until one cluster is left.
// Unlike a plain radix tree over the codes this keeps oversized objects (the r = 5000 fog sphere of
// the Book-2 final scene, Cornell walls) out of the deep levels: nothing wants to merge with them until
// the end, so they end up next to the root instead of inflating every ancestor box on some deep path.
// Finally a greedy surface-area collapse of the binary tree into 4-wide nodes. The topology differs
// from the reference's on purpose; what must match is which objects a ray can hit (leaf boxes are the
// objects' own boxes, interior boxes exact unions) — see closest_hit in rt_intersect.cuh.
// (The Karras radix-tree kernels below are kept as the RT_BVH_LBVH build variant for A/B runs.)
#pragma once
#include <cub/cub.cuh>
#include "rt_scene_dev.h"

namespace rt {

struct BuildBox { float mn[3], mx[3]; };

RT_D unsigned int f2ord(float f) {  // order-preserving float -> uint
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
RT_D float ord2f(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void k_bvh_bounds(const BuildBox* boxes, int n, unsigned int* cb /*[6] ord-encoded centroid min/max*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int a = 0; a < 3; ++a) {
    const float c = 0.5f * boxes[i].mn[a] + 0.5f * boxes[i].mx[a];
    atomicMin(&cb[a], f2ord(c));
    atomicMax(&cb[3 + a], f2ord(c));
  }
}

RT_D unsigned long long expand21(unsigned long long v) {  // spread 21 bits to every third bit
  v &= 0x1FFFFFull;
  v = (v | v << 32) & 0x1F00000000FFFFull;
  v = (v | v << 16) & 0x1F0000FF0000FFull;
  v = (v | v << 8) & 0x100F00F00F00F00Full;
  v = (v | v << 4) & 0x10C30C30C30C30C3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

__global__ void k_bvh_morton(const BuildBox* boxes, int n, const unsigned int* cb, unsigned long long* keys, int* vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long code = 0;
  for (int a = 0; a < 3; ++a) {
    const float lo = ord2f(cb[a]), hi = ord2f(cb[3 + a]);
    const float c = 0.5f * boxes[i].mn[a] + 0.5f * boxes[i].mx[a];
    const float ext = hi - lo;
    float f = ext > 0.f ? (c - lo) / ext : 0.f;
    f = fminf(fmaxf(f, 0.f), 1.f);
    const unsigned long long q = (unsigned long long)fminf(f * 2097152.f, 2097151.f);
    code |= expand21(q) << (2 - a);
  }
  keys[i] = code;
  vals[i] = i;
}

// delta(i, j): common prefix length of keys i and j (ties broken by index), -1 outside the array
RT_D int bvh_delta(const unsigned long long* keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(i ^ j);
  return __clzll(a ^ b);
}

// Binary radix tree: interior nodes 0..n-2, leaves n-1..2n-2 (leaf k = sorted object k).
__global__ void k_bvh_karras(const unsigned long long* keys, int n, int* left, int* right, int* parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (bvh_delta(keys, n, i, i + 1) - bvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = bvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (bvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (bvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = bvh_delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) / 2; ; t = (t + 1) / 2) {
    if (bvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  const int lc = (lo == gamma) ? (n - 1 + gamma) : gamma;
  const int rc = (hi == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
  left[i] = lc; right[i] = rc;
  parent[lc] = i; parent[rc] = i;
  if (i == 0) parent[0] = -1;
}

// Bottom-up box fit: second arrival at a node computes its box (exact fminf/fmaxf unions).
__global__ void k_bvh_fit(const BuildBox* boxes, const int* sorted, int n, const int* left, const int* right,
                          const int* parent, BuildBox* nbox /*[2n-1]*/, int* flags) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int node = n - 1 + k;
  nbox[node] = boxes[sorted[k]];
  __threadfence();
  node = parent[node];
  while (node >= 0) {
    if (atomicAdd(&flags[node], 1) == 0) return;  // first arrival leaves
    __threadfence();
    const BuildBox a = nbox[left[node]], b = nbox[right[node]];
    BuildBox u;
    for (int x = 0; x < 3; ++x) { u.mn[x] = fminf(a.mn[x], b.mn[x]); u.mx[x] = fmaxf(a.mx[x], b.mx[x]); }
    nbox[node] = u;
    __threadfence();
    node = parent[node];
  }
}

RT_D float box_area(const BuildBox& b) {
  const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
  return dx * dy + dy * dz + dz * dx;
}

// ---- PLOC ----
#ifndef RT_PLOC_RADIUS
#define RT_PLOC_RADIUS 16
#endif
#define RT_PLOC_BLOCK 256
// Binary tree ids: interior nodes 0..n-2 (in creation order), leaves n-1..2n-2 (leaf k = sorted object k).
__global__ void k_ploc_init(const BuildBox* boxes, const int* sorted, int n, int* cl, BuildBox* nbox) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  cl[k] = n - 1 + k;
  nbox[n - 1 + k] = boxes[sorted[k]];
}
// nn[i] = the cluster within RT_PLOC_RADIUS places of i whose union with i has the smallest surface area.
// Ties are broken by a total order on PAIRS (even left position first, then the smaller left position), which
// (a) guarantees a mutual pair every round and (b) pairs a run of equidistant clusters as (0,1)(2,3)(4,5)...
// instead of merging one pair per round into a chain: regular grids (the C5 sphere grid) are full of such runs.
// The search of one cluster i over its neighbours' boxes (box_of(j) for j in [i - R, i + R]).
template <class BoxOf>
RT_D int ploc_nearest(int i, int m, BoxOf box_of) {
  const BuildBox a = box_of(i);
  float best = FLT_MAX; int bj = -1; unsigned bkey = 0xFFFFFFFFu;
  for (int d = -RT_PLOC_RADIUS; d <= RT_PLOC_RADIUS; ++d) {
    const int j = i + d;
    if (d == 0 || j < 0 || j >= m) continue;
    const BuildBox b = box_of(j);
    BuildBox u;
    for (int x = 0; x < 3; ++x) { u.mn[x] = fminf(a.mn[x], b.mn[x]); u.mx[x] = fmaxf(a.mx[x], b.mx[x]); }
    const float ar = box_area(u);
    const unsigned lft = (unsigned)min(i, j);
    const unsigned key = ((lft & 1u) << 31) | lft;  // the right partner is then fixed by the distance |d|: ascending |d| below
    if (ar < best || (ar == best && (key < bkey || (key == bkey && abs(d) < abs(bj - i))))) { best = ar; bj = j; bkey = key; }
  }
  return bj;
}
__global__ void __launch_bounds__(RT_PLOC_BLOCK) k_ploc_nn(const int* cl, int m, const BuildBox* nbox, int* nn) {
  __shared__ BuildBox sb[RT_PLOC_BLOCK + 2 * RT_PLOC_RADIUS];
  const int start = blockIdx.x * RT_PLOC_BLOCK - RT_PLOC_RADIUS;
  for (int t = threadIdx.x; t < RT_PLOC_BLOCK + 2 * RT_PLOC_RADIUS; t += RT_PLOC_BLOCK) {
    const int g = start + t;
    if (g >= 0 && g < m) sb[t] = nbox[cl[g]];
  }
  __syncthreads();
  const int i = blockIdx.x * RT_PLOC_BLOCK + threadIdx.x;
  if (i >= m) return;
  nn[i] = ploc_nearest(i, m, [&](int j) { return sb[j - start]; });
}
// Mutual nearest neighbours merge into a new interior node that takes the place of the LEFT partner.
__global__ void k_ploc_merge(const int* cl, int m, const int* nn, BuildBox* nbox, int* left, int* right, int* next_id, int* cl_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int j = nn[i];
  int out = cl[i];
  if (j >= 0 && nn[j] == i) {
    if (i < j) {
      const int id = atomicAdd(next_id, 1);
      const int l = cl[i], r = cl[j];
      left[id] = l; right[id] = r;
      const BuildBox a = nbox[l], b = nbox[r];
      BuildBox u;
      for (int x = 0; x < 3; ++x) { u.mn[x] = fminf(a.mn[x], b.mn[x]); u.mx[x] = fmaxf(a.mx[x], b.mx[x]); }
      nbox[id] = u;
      out = id;
    } else {
      out = -1;  // absorbed by its partner
    }
  }
  cl_out[i] = out;
}
struct PlocAlive { __device__ bool operator()(int v) const { return v >= 0; } };

// Scenes of up to RT_PLOC_SMALL objects (every scene function of the reference: 8 .. 1409): ALL rounds in one CTA, the
// cluster array in shared memory, block-wide scan for the compaction - one launch instead of ~40 rounds x (3 kernels + a
// host read of the cluster count). Same search, same tie rules, same merge as the kernels above.
#define RT_PLOC_SMALL 2048
#define RT_PLOC_SMALL_THREADS 1024
__global__ void __launch_bounds__(RT_PLOC_SMALL_THREADS) k_ploc_small(const BuildBox* boxes, const int* sorted, int n, BuildBox* nbox,
                                                                      int* left, int* right, int* root_out) {
  __shared__ int cl[2][RT_PLOC_SMALL];
  __shared__ int nn[RT_PLOC_SMALL];
  __shared__ int s_next, s_m;
  typedef cub::BlockScan<int, RT_PLOC_SMALL_THREADS> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  const int tid = threadIdx.x;
  for (int k = tid; k < n; k += RT_PLOC_SMALL_THREADS) { cl[0][k] = n - 1 + k; nbox[n - 1 + k] = boxes[sorted[k]]; }
  if (tid == 0) { s_next = 0; s_m = n; }
  __syncthreads();
  int cur = 0;
  while (true) {
    const int m = s_m;
    if (m <= 1) break;
    for (int i = tid; i < m; i += RT_PLOC_SMALL_THREADS) nn[i] = ploc_nearest(i, m, [&](int j) { return nbox[cl[cur][j]]; });
    __syncthreads();
    int out[2] = {-1, -1};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = 2 * tid + q;  // two consecutive clusters per thread: the scan below keeps their order
      if (i >= m) continue;
      const int j = nn[i];
      out[q] = cl[cur][i];
      if (j >= 0 && nn[j] == i) {
        if (i < j) {
          const int id = atomicAdd(&s_next, 1);
          const int l = cl[cur][i], r = cl[cur][j];
          left[id] = l; right[id] = r;
          const BuildBox a = nbox[l], b = nbox[r];
          BuildBox u;
          for (int x = 0; x < 3; ++x) { u.mn[x] = fminf(a.mn[x], b.mn[x]); u.mx[x] = fmaxf(a.mx[x], b.mx[x]); }
          nbox[id] = u;
          out[q] = id;
        } else {
          out[q] = -1;  // absorbed by its partner
        }
      }
    }
    const int alive = (out[0] >= 0) + (out[1] >= 0);
    int pos, total;
    Scan(scan_tmp).ExclusiveSum(alive, pos, total);
    if (out[0] >= 0) cl[cur ^ 1][pos++] = out[0];
    if (out[1] >= 0) cl[cur ^ 1][pos] = out[1];
    __syncthreads();  // also orders the nbox / left / right writes before the next round's reads
    if (tid == 0) s_m = total;
    cur ^= 1;
    __syncthreads();
  }
  if (tid == 0) *root_out = cl[cur][0];
}

// Collapse to 4-wide nodes. One CTA, level-synchronous work queue: task = (binary node, output node).
// Each task opens the interior child with the largest surface area until it has 4 children.
__global__ void __launch_bounds__(1024) k_bvh_collapse(int n, int root, const int* root_dev /* if not null: *root_dev instead of root */,
                                                       const int* left, const int* right, const BuildBox* nbox,
                                                       const int* sorted, const DTlp* tlps, BVH4Node* out,
                                                       int* n_out, int2* qa, int2* qb) {
  __shared__ int s_count, s_next;
  if (threadIdx.x == 0) { qa[0] = make_int2(root_dev ? *root_dev : root, 0); s_count = 1; s_next = 0; *n_out = 1; }
  __syncthreads();
  int2* cur = qa; int2* nxt = qb;
  while (true) {
    const int count = s_count;
    if (count == 0) break;
    for (int t = threadIdx.x; t < count; t += blockDim.x) {
      const int2 task = cur[t];
      int kids[4]; int nk = 2;
      kids[0] = left[task.x]; kids[1] = right[task.x];
      while (nk < 4) {
        int best = -1; float ba = -1.f;
        for (int c = 0; c < nk; ++c)
          if (kids[c] < n - 1) { const float a = box_area(nbox[kids[c]]); if (a > ba) { ba = a; best = c; } }
        if (best < 0) break;
        const int open = kids[best];
        kids[best] = left[open];
        kids[nk++] = right[open];
      }
      BVH4Node node;
      int n_int = 0;  // interior children get consecutive output nodes (and queue entries): siblings share cache lines
      for (int c = 0; c < nk; ++c) n_int += kids[c] < n - 1;
      int o_next = n_int ? atomicAdd(n_out, n_int) : 0;
      int q_next = n_int ? atomicAdd(&s_next, n_int) : 0;
      for (int c = 0; c < 4; ++c) {
        if (c < nk) {
          const BuildBox b = nbox[kids[c]];
          node.lox[c] = b.mn[0]; node.loy[c] = b.mn[1]; node.loz[c] = b.mn[2];
          node.hix[c] = b.mx[0]; node.hiy[c] = b.mx[1]; node.hiz[c] = b.mx[2];
          if (kids[c] >= n - 1) {
            const int obj = sorted[kids[c] - (n - 1)];
            node.child[c] = tlps[obj].ref; node.tlp[c] = (uint32_t)obj | ((uint32_t)tlps[obj].queue << 28);
          } else {
            const int o = o_next++;
            node.child[c] = RT_NODE_FLAG | (uint32_t)o; node.tlp[c] = 0;
            nxt[q_next++] = make_int2(kids[c], o);
          }
        } else {
          node.lox[c] = node.loy[c] = node.loz[c] = FLT_MAX;
          node.hix[c] = node.hiy[c] = node.hiz[c] = -FLT_MAX;
          node.child[c] = RT_NODE_EMPTY; node.tlp[c] = 0;
        }
      }
      out[task.y] = node;
    }
    __syncthreads();
    if (threadIdx.x == 0) { s_count = s_next; s_next = 0; }
    __syncthreads();
    int2* tmp = cur; cur = nxt; nxt = tmp;
  }
}

// n == 1 (and n == 0): a root with one (no) leaf child.
__global__ void k_bvh_trivial(int n, const BuildBox* boxes, const DTlp* tlps, BVH4Node* out, int* n_out) {
  BVH4Node node;
  for (int c = 0; c < 4; ++c) {
    node.lox[c] = node.loy[c] = node.loz[c] = FLT_MAX;
    node.hix[c] = node.hiy[c] = node.hiz[c] = -FLT_MAX;
    node.child[c] = RT_NODE_EMPTY; node.tlp[c] = 0;
  }
  if (n == 1) {
    node.lox[0] = boxes[0].mn[0]; node.loy[0] = boxes[0].mn[1]; node.loz[0] = boxes[0].mn[2];
    node.hix[0] = boxes[0].mx[0]; node.hiy[0] = boxes[0].mx[1]; node.hiz[0] = boxes[0].mx[2];
    node.child[0] = tlps[0].ref; node.tlp[0] = (uint32_t)tlps[0].queue << 28;
  }
  out[0] = node;
  *n_out = 1;
}

}  // namespace rt
