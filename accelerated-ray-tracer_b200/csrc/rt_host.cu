// rt_host.cu — host side of librt_b200.so: scene flattening + upload, device BVH build, the wave
// loop, and the C ABI of include/rt_api.h.
//
// Replaces the body of the reference's scene functions (e.g. final_scene, main.cu:1178-1237):
// cudaDeviceSetLimit(stack/heap) + cudaMallocManaged fb + curandState[nx*ny] + create_world<<<1,1>>>
// + render_init + render + host PPM loop. No device heap, no device recursion, no managed memory.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <algorithm>
#include <cuda_runtime.h>
#include "rt_api.h"
#include "scene_builder.h"
#include "rt_kernels.cuh"
#include "rt_bvh.cuh"

using namespace rt;

static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  return fail(std::string("CUDA error: ") + cudaGetErrorString(e_) + " in " #x); } } while (0)

// ---- libdevice math service ----
__global__ void k_devmath(int op, float x, float* out) {
  *out = op == 0 ? sinf(x) : op == 1 ? cosf(x) : tanf(x);
}
struct CudaDevMath : DevMath {
  float* d = nullptr; bool ok = true;
  CudaDevMath() { if (cudaMalloc(&d, sizeof(float)) != cudaSuccess) ok = false; }
  ~CudaDevMath() { if (d) cudaFree(d); }
  std::map<std::pair<int, uint32_t>, float> memo;  // generators ask for the same few angles over and over
  float call(int op, float x) {
    uint32_t bits; memcpy(&bits, &x, 4);
    auto it = memo.find(std::make_pair(op, bits));
    if (it != memo.end()) return it->second;
    const float v = eval(op, x);
    memo[std::make_pair(op, bits)] = v;
    return v;
  }
  float eval(int op, float x) {
    float h = 0.f;
    k_devmath<<<1, 1>>>(op, x, d);
    if (cudaMemcpy(&h, d, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) ok = false;
    return h;
  }
  float sinf_(float x) override { return call(0, x); }
  float cosf_(float x) override { return call(1, x); }
  float tanf_(float x) override { return call(2, x); }
};
struct HostDevMath : DevMath {  // rt_scene_export_host only
  float sinf_(float x) override { return sinf(x); }
  float cosf_(float x) override { return cosf(x); }
  float tanf_(float x) override { return tanf(x); }
};

// ---- device block cache ----
// cudaMalloc / cudaFree (and cudaMallocHost / cudaStreamCreate) cost tens to hundreds of milliseconds per scene
// build/destroy cycle - more than the scene build itself. Device blocks go back to a per-process free list, keyed by
// (device, canonical size), instead of to the driver, and are handed out again to the next scene on the same device;
// pinned counters and streams are recycled the same way. rt_trim_device_cache() releases everything.
#include <mutex>
#include <unordered_map>
namespace {
std::mutex g_cache_mu;
std::map<std::pair<int, size_t>, std::vector<void*>> g_cache;  // (device, bytes) -> free blocks
std::map<int, size_t> g_cache_bytes;  // per device
size_t cache_cap_bytes() {  // per-device cap of the free list (RT_CACHE_MAX_MB overrides the 8 GiB default)
  static const size_t cap = [] { const char* e = getenv("RT_CACHE_MAX_MB"); return e ? (size_t)atoll(e) << 20 : (size_t)8 << 30; }();
  return cap;
}
void trim_device_locked(int dev) {  // caller holds g_cache_mu and has made `dev` current
  for (auto it = g_cache.begin(); it != g_cache.end();) {
    if (it->first.first == dev) { for (void* p : it->second) cudaFree(p); it = g_cache.erase(it); } else ++it;
  }
  g_cache_bytes[dev] = 0;
}

size_t canonical_bytes(size_t want) {  // sizes are rounded so that a slightly different scene finds the same block sizes
  if (want < 512) return 512;
  if (want < ((size_t)1 << 20)) { size_t p2 = 512; while (p2 < want) p2 <<= 1; return p2; }
  return (want + ((size_t)2 << 20) - 1) / ((size_t)2 << 20) * ((size_t)2 << 20);
}
cudaError_t cached_malloc(void** out, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    auto it = g_cache.find(std::make_pair(dev, bytes));
    if (it != g_cache.end() && !it->second.empty()) {
      *out = it->second.back();
      it->second.pop_back();
      g_cache_bytes[dev] -= bytes;
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {  // the memory may be sitting in our own free list under other size classes
    cudaGetLastError();
    { std::lock_guard<std::mutex> lk(g_cache_mu); trim_device_locked(dev); }
    e = cudaMalloc(out, bytes);
  }
  return e;
}
void cached_free(void* p, size_t bytes, int dev) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    if (g_cache_bytes[dev] + bytes <= cache_cap_bytes()) {
      g_cache[std::make_pair(dev, bytes)].push_back(p);
      g_cache_bytes[dev] += bytes;
      return;
    }
  }
  cudaFree(p);
}
}  // namespace

// CUDA events that are destroyed on every exit path (the CU macro returns early on errors).
struct Events {
  std::vector<cudaEvent_t> v;
  cudaError_t make(cudaEvent_t* e, unsigned flags = cudaEventDefault) { cudaError_t r = cudaEventCreateWithFlags(e, flags); if (r == cudaSuccess) v.push_back(*e); return r; }
  ~Events() { for (auto e : v) cudaEventDestroy(e); }
};

template <class T> struct DBuf {
  T* p = nullptr; size_t n = 0; size_t bytes = 0; int dev = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), n(o.n), bytes(o.bytes), dev(o.dev) { o.p = nullptr; o.n = 0; o.bytes = 0; }
  DBuf& operator=(DBuf&& o) noexcept { if (this != &o) { free(); p = o.p; n = o.n; bytes = o.bytes; dev = o.dev; o.p = nullptr; o.n = 0; o.bytes = 0; } return *this; }
  cudaError_t alloc(size_t count) {  // n / bytes are committed only when the allocation succeeded
    free();
    const size_t want = canonical_bytes(std::max<size_t>(count, 1) * sizeof(T));
    cudaGetDevice(&dev);
    void* q = nullptr;
    const cudaError_t e = cached_malloc(&q, want);
    if (e != cudaSuccess) return e;
    p = (T*)q; n = count; bytes = want;
    return cudaSuccess;
  }
  cudaError_t upload(const std::vector<T>& h) {
    cudaError_t e = alloc(h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  }
  void free() { if (p) cached_free(p, bytes, dev); p = nullptr; n = 0; bytes = 0; }
  ~DBuf() { free(); }
};

#define RT_MAX_POOLS 4
// Per-scene host-side CUDA objects (streams, pinned counter block), recycled across scenes of one device.
struct HostCtx { int dev = 0; cudaStream_t streams[RT_MAX_POOLS] = {nullptr}; void* pinned = nullptr; };
namespace {
std::vector<HostCtx> g_ctx_pool;
bool ctx_acquire(int dev, size_t pinned_bytes, HostCtx& out) {
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = 0; i < g_ctx_pool.size(); ++i)
      if (g_ctx_pool[i].dev == dev) { out = g_ctx_pool[i]; g_ctx_pool.erase(g_ctx_pool.begin() + i); return true; }
  }
  out = HostCtx();
  out.dev = dev;
  for (int k = 0; k < RT_MAX_POOLS; ++k) if (cudaStreamCreate(&out.streams[k]) != cudaSuccess) return false;
  return cudaMallocHost(&out.pinned, pinned_bytes) == cudaSuccess;
}
void ctx_release(const HostCtx& c) { std::lock_guard<std::mutex> lk(g_cache_mu); g_ctx_pool.push_back(c); }
}  // namespace

extern "C" void rt_trim_device_cache(void) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  int cur = -1;
  cudaGetDevice(&cur);  // restored below: the caller's current device is not ours to change
  for (auto& kv : g_cache) { cudaSetDevice(kv.first.first); for (void* p : kv.second) cudaFree(p); }
  g_cache.clear();
  g_cache_bytes.clear();
  for (auto& c : g_ctx_pool) {
    cudaSetDevice(c.dev);
    for (int k = 0; k < RT_MAX_POOLS; ++k) if (c.streams[k]) cudaStreamDestroy(c.streams[k]);
    if (c.pinned) cudaFreeHost(c.pinned);
  }
  g_ctx_pool.clear();
  if (cur >= 0) cudaSetDevice(cur);
}

struct rt_scene {
  int device = 0;
  SceneDesc sd;
  std::vector<int> rank;
  // groups (RT_OBJ_BVH) resolved into instanced members: what the device tables are built from (= sd when there is no group)
  bool grouped = false;
  SceneDesc sdx;
  std::vector<int> rank_x, origin;   // per expanded top-level entry: tie-break rank, index of the entry of sd.top it came from
  const SceneDesc& flat() const { return grouped ? sdx : sd; }
  const std::vector<int>& flat_rank() const { return grouped ? rank_x : rank; }
  // flattened scene
  DBuf<DSphere> spheres; DBuf<DQuad> quads; DBuf<DXform> xforms; DBuf<DMedium> media;
  DBuf<DMat> mats; DBuf<DTex> texs; DBuf<DImage> images; DBuf<DTlp> tlp; DBuf<BVH4Node> nodes; DBuf<float4> qplanes;
  std::vector<DBuf<unsigned char>> image_px;
  HostCtx ctx; bool has_ctx = false;
  DScene dscene;
  int n_nodes = 0; float bvh_ms = 0.f; uint64_t h2d_bytes = 0;
  // render state
  cudaStream_t stream = nullptr, pool_stream[RT_MAX_POOLS] = {nullptr};  // pool 0 runs on `stream`
  DBuf<float4> ray_o[2], ray_d[2], thr[2], col; DBuf<float2> hit; DBuf<uint32_t> rng;
  DBuf<int> queues;
  DBuf<WaveCounters> counters;          // one per slot pool
  DBuf<unsigned long long> next_work;
  WaveCounters* h_counters = nullptr;  // pinned, one per slot pool
  DBuf<float> accum, fb, aov_t, reduce_tmp; DBuf<int> aov_obj, aov_mat; DBuf<unsigned long long> acc64;
  size_t slots_cap = 0, pix_cap = 0, accum_valid_pix = 0;
  // adaptive sampling (rt_render_adaptive): state of the pass rt_render is asked to run
  struct Adaptive {
    bool on = false, first = false;   // first: zero the fixed-point sums
    int sample_base = 0, sample_count = 0, n_active = 0;
    DBuf<int> active; DBuf<unsigned long long> acc64_odd;
    std::vector<int> tile_spp;        // per tile of the last adaptive render (rt_readback_spp)
    int tile = 0, tiles_x = 0, tiles_y = 0;
  } adapt;
  RenderParams last{}; rt_render_stats stats{}; bool has_aov = false; float last_gamma = 2.2f; int last_spp_total = 0;
  ~rt_scene() {
    // kernels of a failed rt_render may still be in flight on the pool streams: nothing goes back to the block cache
    // (where the next scene could pick it up) before the device is idle
    cudaDeviceSynchronize();
    if (has_ctx) ctx_release(ctx);
  }
};

// ---- SD -> flattened device tables ----
struct Flattener {
  const SceneDesc& sd;
  std::vector<DSphere> spheres; std::vector<DQuad> quads; std::vector<DXform> xforms; std::vector<DMedium> media;
  bool bad = false;  // met a group (RT_OBJ_BVH) where only geometry can be (e.g. a medium boundary)
  explicit Flattener(const SceneDesc& s) : sd(s) {}
  static DQuad mkquad(const rt_object_desc& o) {
    DQuad q;
    q.nx = o.n[0]; q.ny = o.n[1]; q.nz = o.n[2]; q.D = o.D;
    q.Qx = o.Q[0]; q.Qy = o.Q[1]; q.Qz = o.Q[2]; q.mat = o.mat;
    q.ux = o.u[0]; q.uy = o.u[1]; q.uz = o.u[2]; q.pad0 = 0;
    q.vx = o.v[0]; q.vy = o.v[1]; q.vz = o.v[2]; q.pad1 = 0;
    q.wx = o.w[0]; q.wy = o.w[1]; q.wz = o.w[2]; q.pad2 = 0;
    return q;
  }
  uint32_t flatten(int id) {
    const rt_object_desc& o = sd.obj[id];
    switch (o.kind) {
      case RT_OBJ_SPHERE: {
        DSphere s; s.cx = o.c0[0]; s.cy = o.c0[1]; s.cz = o.c0[2]; s.radius = o.radius;
        s.dx = o.dc[0]; s.dy = o.dc[1]; s.dz = o.dc[2]; s.mat = o.mat;
        spheres.push_back(s);
        return make_ref(G_SPHERE, (uint32_t)spheres.size() - 1);
      }
      case RT_OBJ_QUAD:
        quads.push_back(mkquad(o));
        return make_ref(G_QUAD, (uint32_t)quads.size() - 1);
      case RT_OBJ_BOX: {
        uint32_t first = (uint32_t)quads.size();
        for (int i = 0; i < 6; ++i) quads.push_back(mkquad(sd.obj[o.child + i]));
        return make_ref(G_BOX, first);
      }
      case RT_OBJ_TRANSLATE: {
        DXform x; memset(&x, 0, sizeof(x));
        x.kind = X_TRANSLATE; x.child = flatten(o.child); x.a = o.offset[0]; x.b = o.offset[1]; x.c = o.offset[2];
        xforms.push_back(x);
        return make_ref(G_XFORM, (uint32_t)xforms.size() - 1);
      }
      case RT_OBJ_ROTATE_Y: {
        DXform x; memset(&x, 0, sizeof(x));
        x.kind = X_ROTATE_Y; x.child = flatten(o.child); x.a = o.sin_t; x.b = o.cos_t;
        xforms.push_back(x);
        return make_ref(G_XFORM, (uint32_t)xforms.size() - 1);
      }
      case RT_OBJ_WITH_MATERIAL:  // geometry of the child; the override lives in the top-level entry (upload_scene)
        return flatten(o.child);
      case RT_OBJ_BVH:  // resolved by expand_groups before the scene gets here; what is left is a group in a place it cannot be
        bad = true;
        return make_ref(G_SPHERE, 0);
      case RT_OBJ_MEDIUM: {
        DMedium m; m.boundary = flatten(o.child); m.neg_inv_density = o.neg_inv_density; m.mat = o.mat; m.pad = 0;
        media.push_back(m);
        return make_ref(G_MEDIUM, (uint32_t)media.size() - 1);
      }
    }
    return make_ref(G_SPHERE, 0);
  }
};

static bool tex_needs_uv(const SceneDesc& sd, int t, int depth = 0) {
  if (t < 0 || depth > 8) return false;
  const rt_texture_desc& d = sd.tex[t];
  if (d.kind == RT_TEX_IMAGE || d.kind == RT_TEX_UV_OFFSET) return true;
  if (d.kind == RT_TEX_CHECKER) return tex_needs_uv(sd, d.even, depth + 1) || tex_needs_uv(sd, d.odd, depth + 1);
  return false;
}

static bool tex_procedural(const SceneDesc& sd, int t, int depth = 0) {
  if (t < 0 || depth > 8) return false;
  const rt_texture_desc& d = sd.tex[t];
  if (d.kind == RT_TEX_NOISE || d.kind == RT_TEX_NOODLE || d.kind == RT_TEX_FELT) return true;
  if (d.kind == RT_TEX_CHECKER) return tex_procedural(sd, d.even, depth + 1) || tex_procedural(sd, d.odd, depth + 1);
  if (d.kind == RT_TEX_UV_OFFSET) return tex_procedural(sd, d.even, depth + 1);
  return false;
}

static int queue_of(const SceneDesc& sd, const rt_material_desc& m) {
  if ((m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_ISOTROPIC) && tex_procedural(sd, m.tex)) return Q_PROCEDURAL;
  switch (m.kind) {
    case RT_MAT_LAMBERTIAN: return Q_LAMBERTIAN;
    case RT_MAT_METAL: return Q_METAL;
    case RT_MAT_DIELECTRIC: return Q_DIELECTRIC;
    case RT_MAT_DIFFUSE_LIGHT: return Q_LIGHT;
    default: return Q_ISOTROPIC;
  }
}

static int upload_scene(rt_scene* s) {
  const SceneDesc& sd = s->flat();
  Flattener F(sd);
  const int n = (int)sd.top.size();
  {
    size_t n_sph = 0;
    for (const auto& o : sd.obj) n_sph += o.kind == RT_OBJ_SPHERE;
    big_reserve(F.spheres, n_sph);
  }
  std::vector<DTlp> tlp; std::vector<BuildBox> boxes; std::vector<uint32_t> refs;
  big_resize(tlp, (size_t)n); big_resize(boxes, (size_t)n); big_resize(refs, (size_t)n);
  for (int k = 0; k < n; ++k) {
    const rt_object_desc& o = sd.obj[sd.top[k]];
    tlp[k].ref = F.flatten(sd.top[k]);
    // instance wrappers carry no material of their own: the wrapped object's, unless a with_material sits on the way down
    // (hittable.cuh:170-174 sets rec.mat_ptr after the child's hit, so the outermost override wins)
    int mid = sd.top[k];
    while (sd.obj[mid].kind == RT_OBJ_TRANSLATE || sd.obj[mid].kind == RT_OBJ_ROTATE_Y) mid = sd.obj[mid].child;
    const int mat = sd.obj[mid].mat;
    if (mat < 0 || mat >= (int)sd.mat.size()) return fail("upload_scene: top-level object without a material");
    tlp[k].mat = mat;
    tlp[k].queue = queue_of(sd, sd.mat[mat]);
    tlp[k].rank = s->flat_rank()[k];
    refs[k] = tlp[k].ref;
    for (int a = 0; a < 3; ++a) { boxes[k].mn[a] = o.box_min[a]; boxes[k].mx[a] = o.box_max[a]; }
  }
  if (F.bad) return fail("upload_scene: a group (bvh_node) where only geometry can be");
  std::vector<DMat> mats; big_resize(mats, sd.mat.size());
  for (size_t i = 0; i < sd.mat.size(); ++i) {
    const rt_material_desc& m = sd.mat[i];
    DMat d; memset(&d, 0, sizeof(d));
    d.kind = m.kind; d.tex = m.tex; d.ax = m.albedo[0]; d.ay = m.albedo[1]; d.az = m.albedo[2]; d.param = m.param;
    d.needs_uv = tex_needs_uv(sd, m.tex) ? 1 : 0;
    mats[i] = d;
  }
  std::vector<DTex> texs; big_resize(texs, sd.tex.size());
  for (size_t i = 0; i < sd.tex.size(); ++i) {
    const rt_texture_desc& t = sd.tex[i];
    DTex d; memset(&d, 0, sizeof(d));
    d.kind = t.kind; d.even = t.even; d.odd = t.odd; d.image = t.image;
    d.cx = t.color[0]; d.cy = t.color[1]; d.cz = t.color[2]; d.scale = t.scale;
    for (int k = 0; k < 13; ++k) d.p[k] = t.p[k];
    texs[i] = d;
  }
  std::vector<DImage> images(sd.img.size());
  uint64_t bytes = 0;
  for (size_t i = 0; i < sd.img.size(); ++i) {
    const HostImage& im = sd.img_data[i];
    s->image_px.emplace_back();
    DBuf<unsigned char>& ib = s->image_px.back();
    unsigned char* d = nullptr;  // no pixels: the texture renders (0,1,1) like the reference's invalid DeviceImage
    if (!im.px.empty()) {
      CU(ib.alloc(im.px.size()));
      CU(cudaMemcpy(ib.p, im.px.data(), im.px.size(), cudaMemcpyHostToDevice));
      d = ib.p;
    }
    images[i].data = d; images[i].width = im.width; images[i].height = im.height; images[i].bpp = im.bpp; images[i].pad = 0;
    bytes += im.px.size();
  }
  {
    std::vector<float4> qp(F.quads.size());
    for (size_t i = 0; i < qp.size(); ++i) qp[i] = make_float4(F.quads[i].nx, F.quads[i].ny, F.quads[i].nz, F.quads[i].D);
    CU(s->qplanes.upload(qp));
  }
  CU(s->spheres.upload(F.spheres)); CU(s->quads.upload(F.quads)); CU(s->xforms.upload(F.xforms)); CU(s->media.upload(F.media));
  CU(s->mats.upload(mats)); CU(s->texs.upload(texs)); CU(s->images.upload(images)); CU(s->tlp.upload(tlp));
  bytes += F.spheres.size() * sizeof(DSphere) + F.quads.size() * sizeof(DQuad) + F.xforms.size() * sizeof(DXform) +
           F.media.size() * sizeof(DMedium) + mats.size() * sizeof(DMat) + texs.size() * sizeof(DTex) +
           images.size() * sizeof(DImage) + tlp.size() * sizeof(DTlp) + boxes.size() * sizeof(BuildBox) + refs.size() * 4;

  // ---- BVH build on the device ----
  DBuf<BuildBox> d_boxes; DBuf<uint32_t> d_refs;
  CU(d_boxes.upload(boxes)); CU(d_refs.upload(refs));
  CU(s->nodes.alloc(std::max(n, 1)));
  DBuf<int> d_nout; CU(d_nout.alloc(1));
  cudaEvent_t e0, e1;
  Events evh;
  CU(evh.make(&e0)); CU(evh.make(&e1));
  CU(cudaEventRecord(e0));
  if (n <= 1) {
    k_bvh_trivial<<<1, 1>>>(n, d_boxes.p, s->tlp.p, s->nodes.p, d_nout.p);
  } else {
    const int B = 256, G = (n + B - 1) / B;
    DBuf<unsigned int> cb; CU(cb.alloc(6));
    unsigned int init[6] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u};
    CU(cudaMemcpy(cb.p, init, sizeof(init), cudaMemcpyHostToDevice));
    DBuf<unsigned long long> k_in, k_out; DBuf<int> v_in, v_out;
    CU(k_in.alloc(n)); CU(k_out.alloc(n)); CU(v_in.alloc(n)); CU(v_out.alloc(n));
    k_bvh_bounds<<<G, B>>>(d_boxes.p, n, cb.p);
    k_bvh_morton<<<G, B>>>(d_boxes.p, n, cb.p, k_in.p, v_in.p);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, n, 0, 63);
    DBuf<unsigned char> tmp; CU(tmp.alloc(tmp_bytes));
    cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, n, 0, 63);
    DBuf<int> left, right; DBuf<BuildBox> nbox; DBuf<int2> qa, qb;
    CU(left.alloc(n)); CU(right.alloc(n)); CU(nbox.alloc(2 * n)); CU(qa.alloc(n)); CU(qb.alloc(n));
    int root = 0;
    DBuf<int> d_root;  // small scenes: the root stays on the device
    // RT_BVH_BUILDER=lbvh: the plain radix tree over the Morton codes (Karras 2012) instead of PLOC - a different
    // topology over the same leaves, kept for A/B runs and for the test that both return the same hits
    const char* builder = getenv("RT_BVH_BUILDER");
    if (builder && !strcmp(builder, "lbvh")) {
    DBuf<int> parent, flags;
    CU(parent.alloc(2 * n)); CU(flags.alloc(n));
    CU(cudaMemset(flags.p, 0, n * sizeof(int)));
    k_bvh_karras<<<G, B>>>(k_out.p, n, left.p, right.p, parent.p);
    k_bvh_fit<<<G, B>>>(d_boxes.p, v_out.p, n, left.p, right.p, parent.p, nbox.p, flags.p);
    } else {
    if (n <= RT_PLOC_SMALL) {
      // every scene function of the reference: all PLOC rounds in one launch of one CTA
      CU(d_root.alloc(1));
      k_ploc_small<<<1, RT_PLOC_SMALL_THREADS>>>(d_boxes.p, v_out.p, n, nbox.p, left.p, right.p, d_root.p);
    } else {
    // PLOC rounds: nearest neighbour search -> merge -> compaction, until one cluster (the root) is left
    DBuf<int> cl_a, cl_b, nn, next_id, d_m;
    CU(cl_a.alloc(n)); CU(cl_b.alloc(n)); CU(nn.alloc(n)); CU(next_id.alloc(1)); CU(d_m.alloc(1));
    CU(cudaMemset(next_id.p, 0, sizeof(int)));
    size_t sel_bytes = 0;
    cub::DeviceSelect::If(nullptr, sel_bytes, cl_b.p, cl_a.p, d_m.p, n, PlocAlive());
    DBuf<unsigned char> sel_tmp; CU(sel_tmp.alloc(sel_bytes));
    k_ploc_init<<<G, B>>>(d_boxes.p, v_out.p, n, cl_a.p, nbox.p);
    int m = n, rounds = 0;
    while (m > 1) {
      const int Gm = (m + RT_PLOC_BLOCK - 1) / RT_PLOC_BLOCK;
      k_ploc_nn<<<Gm, RT_PLOC_BLOCK>>>(cl_a.p, m, nbox.p, nn.p);
      k_ploc_merge<<<Gm, RT_PLOC_BLOCK>>>(cl_a.p, m, nn.p, nbox.p, left.p, right.p, next_id.p, cl_b.p);
      cub::DeviceSelect::If(sel_tmp.p, sel_bytes, cl_b.p, cl_a.p, d_m.p, m, PlocAlive());
      int m_new = 0;
      CU(cudaMemcpy(&m_new, d_m.p, sizeof(int), cudaMemcpyDeviceToHost));
      if (m_new >= m || ++rounds > 4096) return fail("upload_scene: PLOC made no progress");
      m = m_new;
    }
    CU(cudaMemcpy(&root, cl_a.p, sizeof(int), cudaMemcpyDeviceToHost));
    }
    }
    k_bvh_collapse<<<1, 1024>>>(n, root, d_root.p, left.p, right.p, nbox.p, v_out.p, s->tlp.p, s->nodes.p, d_nout.p, qa.p, qb.p);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
  }
  CU(cudaEventRecord(e1));
  CU(cudaEventSynchronize(e1));
  CU(cudaGetLastError());
  CU(cudaEventElapsedTime(&s->bvh_ms, e0, e1));
  CU(cudaMemcpy(&s->n_nodes, d_nout.p, sizeof(int), cudaMemcpyDeviceToHost));
  s->h2d_bytes = bytes;

  DScene& D = s->dscene;
  D.spheres = s->spheres.p; D.quads = s->quads.p; D.xforms = s->xforms.p; D.media = s->media.p;
  D.mats = s->mats.p; D.texs = s->texs.p; D.images = s->images.p; D.tlp = s->tlp.p; D.nodes = s->nodes.p; D.qplanes = s->qplanes.p;
  D.n_tlp = n; D.n_nodes = s->n_nodes;
  const rt_camera_desc& c = sd.cam;
  D.cam.origin = v3(c.origin[0], c.origin[1], c.origin[2]);
  D.cam.llc = v3(c.lower_left_corner[0], c.lower_left_corner[1], c.lower_left_corner[2]);
  D.cam.horizontal = v3(c.horizontal[0], c.horizontal[1], c.horizontal[2]);
  D.cam.vertical = v3(c.vertical[0], c.vertical[1], c.vertical[2]);
  D.cam.u = v3(c.u[0], c.u[1], c.u[2]);
  D.cam.v = v3(c.v[0], c.v[1], c.v[2]);
  D.cam.lens_radius = c.lens_radius; D.cam.time0 = c.time0; D.cam.time1 = c.time1;
  return 0;
}

// ---- C ABI ----
extern "C" const char* rt_last_error(void) { return g_err.c_str(); }

static int finish_build(rt_scene* s, int dev, rt_scene** out) {
  s->rank = reference_leaf_order(s->sd);
  for (const auto& o : s->sd.obj) s->grouped = s->grouped || o.kind == RT_OBJ_BVH;
  if (s->grouped) {
    const std::string err = expand_groups(s->sd, s->sdx, s->origin);
    if (!err.empty()) { delete s; return fail("rt_build_scene: " + err); }
    // ties between members of different entries follow the reference order of their entries, then the member order
    const int n = (int)s->sdx.top.size();
    std::vector<int> order(n);
    for (int k = 0; k < n; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return s->rank[s->origin[a]] < s->rank[s->origin[b]]; });
    s->rank_x.assign(n, 0);
    for (int pos = 0; pos < n; ++pos) s->rank_x[order[pos]] = pos;
  }
  if (upload_scene(s)) { delete s; return 1; }
  if (!ctx_acquire(dev, RT_MAX_POOLS * sizeof(WaveCounters), s->ctx)) { delete s; return fail("cudaStreamCreate / cudaMallocHost failed"); }
  s->has_ctx = true;
  for (int k = 0; k < RT_MAX_POOLS; ++k) s->pool_stream[k] = s->ctx.streams[k];
  s->stream = s->pool_stream[0];
  s->h_counters = (WaveCounters*)s->ctx.pinned;
  *out = s;
  return 0;
}

extern "C" int rt_build_scene(const rt_scene_desc* desc, rt_scene** out) {
  if (!desc || !out) return fail("rt_build_scene: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("rt_build_scene: no CUDA device (this library has no CPU path)");
  int dev = desc->device;
  if (dev < 0) { CU(cudaGetDevice(&dev)); }
  CU(cudaSetDevice(dev));
  rt_scene* s = nullptr;
  try {
    s = new rt_scene();
    s->device = dev;
    {
      CudaDevMath dm;
      std::string err = generate_scene(s->sd, dm, desc->scene_id, desc->nx, desc->ny, desc->grid_half,
                                       desc->texture_dir ? desc->texture_dir : "");
      if (!err.empty()) { delete s; return fail("rt_build_scene: " + err); }
      if (!dm.ok) { delete s; return fail("rt_build_scene: device math service failed"); }
    }
    return finish_build(s, dev, out);
  } catch (const std::exception& e) {
    delete s;
    *out = nullptr;
    return fail(std::string("rt_build_scene: ") + e.what());
  }
}

// Build from a caller-made scene description (the flat SD format of rt_scene_desc.h): the generic path behind the
// scene vocabulary. A scene built this way from rt_scene_export's output has the same geometry, materials and camera;
// the scene function's host parameters (background, gradient, default spp) are not part of an SD and come from
// rt_render_params (test_scene_from_exported_description_renders_identically passes them).
extern "C" int rt_build_scene_sd(const void* sd, size_t sd_bytes, const unsigned char* const* image_pixels, int32_t n_images,
                                 int32_t device, rt_scene** out) {
  if (!sd || !out) return fail("rt_build_scene_sd: null argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("rt_build_scene_sd: no CUDA device (this library has no CPU path)");
  int dev = device;
  if (dev < 0) { CU(cudaGetDevice(&dev)); }
  CU(cudaSetDevice(dev));
  rt_scene* s = nullptr;
  try {  // nothing may throw across the C boundary (bad_alloc on a hostile size, ...)
    s = new rt_scene();
    s->device = dev;
    std::string err = sd_deserialize(sd, sd_bytes, image_pixels, n_images, s->sd);
    if (!err.empty()) { delete s; return fail("rt_build_scene_sd: " + err); }
    return finish_build(s, dev, out);
  } catch (const std::exception& e) {
    delete s;  // finish_build deletes the scene itself only on the error paths it returns from
    *out = nullptr;
    return fail(std::string("rt_build_scene_sd: ") + e.what());
  }
}

extern "C" int rt_sd_flatten(const void* sd, size_t sd_bytes, void* buf, size_t cap, size_t* needed, int32_t* origin, int32_t origin_cap,
                             int32_t* n_top_out) {
  if (!sd) return fail("rt_sd_flatten: null argument");
  try {
    SceneDesc in, out;
    std::vector<int> og;
    std::string err = sd_deserialize(sd, sd_bytes, nullptr, 0, in);
    if (err.empty()) err = expand_groups(in, out, og);
    if (!err.empty()) return fail("rt_sd_flatten: " + err);
    const std::string bin = sd_serialize(out);
    if (needed) *needed = bin.size();
    if (buf && cap >= bin.size()) memcpy(buf, bin.data(), bin.size());
    if (n_top_out) *n_top_out = (int32_t)og.size();
    if (origin) for (int k = 0; k < (int)og.size() && k < origin_cap; ++k) origin[k] = og[k];
    return 0;
  } catch (const std::exception& e) {
    return fail(std::string("rt_sd_flatten: ") + e.what());
  }
}

extern "C" void rt_destroy(rt_scene* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  delete s;
}

extern "C" int rt_scene_info_get(rt_scene* s, rt_scene_info* o) {
  if (!s || !o) return fail("rt_scene_info_get: null argument");
  memset(o, 0, sizeof(*o));
  o->scene_id = s->sd.scene_id; o->nx = s->sd.nx; o->ny = s->sd.ny;
  o->n_top = (int)s->sd.top.size(); o->n_obj = (int)s->sd.obj.size(); o->n_mat = (int)s->sd.mat.size();
  o->n_tex = (int)s->sd.tex.size(); o->n_img = (int)s->sd.img.size(); o->n_bvh_nodes = s->n_nodes;
  o->default_nx = s->sd.default_nx; o->default_ny = s->sd.default_ny; o->default_spp = s->sd.default_spp;
  o->gradient_bg = s->sd.gradient_bg;
  for (int k = 0; k < 3; ++k) o->background[k] = s->sd.background[k];
  o->bvh_build_ms = s->bvh_ms; o->h2d_bytes = s->h2d_bytes;
  return 0;
}

static int export_sd(const SceneDesc& sd, const std::vector<int>& rank, void* buf, size_t cap, size_t* needed,
                     int32_t* rank_out, int rank_cap) {
  std::string bin = sd_serialize(sd);
  if (needed) *needed = bin.size();
  if (buf && cap >= bin.size()) memcpy(buf, bin.data(), bin.size());
  if (rank_out) for (int k = 0; k < (int)rank.size() && k < rank_cap; ++k) rank_out[k] = rank[k];
  return 0;
}
extern "C" int rt_scene_export(rt_scene* s, void* buf, size_t cap, size_t* needed, int32_t* rank) {
  if (!s) return fail("rt_scene_export: null scene");
  return export_sd(s->sd, s->rank, buf, cap, needed, rank, (int)s->rank.size());
}
extern "C" int rt_scene_export_host(const rt_scene_desc* desc, void* buf, size_t cap, size_t* needed, int32_t* rank,
                                    int32_t rank_cap) {
  if (!desc) return fail("rt_scene_export_host: null argument");
  HostDevMath dm;
  SceneDesc sd;
  std::string err = generate_scene(sd, dm, desc->scene_id, desc->nx, desc->ny, desc->grid_half,
                                   desc->texture_dir ? desc->texture_dir : "");
  if (!err.empty()) return fail("rt_scene_export_host: " + err);
  std::vector<int> rk = reference_leaf_order(sd);
  return export_sd(sd, rk, buf, cap, needed, rank, rank_cap);
}

// Capacity of a pool's dense layout: its paths plus the warp padding between the RT_NQ queues, in whole trace ranges.
static size_t pool_cap(size_t n) { return (n + 32 * RT_NQ + RT_RANGE - 1) / RT_RANGE * RT_RANGE; }
static size_t pool_subcap(size_t n) { return ((pool_cap(n) / RT_RANGE + RT_NSUB - 1) / RT_NSUB + 1) * RT_RANGE; }

static int ensure_buffers(rt_scene* s, size_t n_slots, size_t n_pix, bool ref_rng, bool aov) {
  if (n_slots > s->slots_cap) {
    s->slots_cap = 0;  // a failure below must not leave a capacity that the freed / null buffers do not back
    s->rng.free();
    // worst case over the pool counts: RT_MAX_POOLS pools, each with its own padding
    const size_t cap = pool_cap(n_slots) + RT_MAX_POOLS * pool_cap(0);
    for (int k = 0; k < 2; ++k) { CU(s->ray_o[k].alloc(cap)); CU(s->ray_d[k].alloc(cap)); CU(s->thr[k].alloc(cap)); }
    CU(s->hit.alloc(cap));
    CU(s->queues.alloc((size_t)RT_NQ * (pool_subcap(n_slots) + RT_MAX_POOLS * pool_subcap(0))));
    s->slots_cap = n_slots;
  }
  if (n_pix > s->pix_cap) {
    s->pix_cap = 0; s->accum_valid_pix = 0;
    s->aov_obj.free(); s->aov_mat.free(); s->aov_t.free(); s->rng.free();
    CU(s->accum.alloc(3 * n_pix)); CU(s->fb.alloc(3 * n_pix)); CU(s->acc64.alloc(3 * n_pix)); CU(s->col.alloc(n_pix));
    s->pix_cap = n_pix;
  }
  if (ref_rng && s->rng.n < 6 * n_pix) CU(s->rng.alloc(6 * s->pix_cap));
  if (aov && s->aov_obj.n < n_pix) { CU(s->aov_obj.alloc(s->pix_cap)); CU(s->aov_mat.alloc(s->pix_cap)); CU(s->aov_t.alloc(s->pix_cap)); }
  if (!s->counters.p) { CU(s->counters.alloc(RT_MAX_POOLS)); CU(s->next_work.alloc(1)); }
  return 0;
}

extern "C" int rt_render(rt_scene* s, const rt_render_params* p, double* device_ms, uint64_t* rays) {
  if (!s || !p) return fail("rt_render: null argument");
  CU(cudaSetDevice(s->device));
  const SceneDesc& sd = s->sd;
  const int spp_total = p->spp > 0 ? p->spp : sd.default_spp;
  const int world = p->world > 0 ? p->world : 1, rank = p->rank;
  if (rank < 0 || rank >= world) return fail("rt_render: rank out of range");
  const bool ref_rng = p->rng_mode == 1;
  if (ref_rng && p->split_mode == 1 && world > 1)
    return fail("rt_render: reference-RNG mode cannot split a pixel's samples across ranks (one sequential stream per pixel)");
  RenderParams P; memset(&P, 0, sizeof(P));
  P.nx = sd.nx; P.ny = sd.ny; P.inv_nx = 1.0f / (float)sd.nx;
  if (p->split_mode == 1) {  // spp split: all pixels, a share of the samples
    P.rank = 0; P.world = 1; P.rows_local = sd.ny;
    const int base = (int)((long long)spp_total * rank / world), end = (int)((long long)spp_total * (rank + 1) / world);
    P.sample_base = base; P.sample_count = end - base;
  } else {  // tile split: scanlines j = rank (mod world), all samples
    P.rank = rank; P.world = world; P.rows_local = (sd.ny - rank + world - 1) / world;
    P.sample_base = 0; P.sample_count = spp_total;
  }
  const size_t n_pix = (size_t)P.rows_local * P.nx;
  const bool adaptive = s->adapt.on;
  if (adaptive) {  // one pass of rt_render_adaptive: its own sample numbers, over the pixels of the tiles still active
    if (ref_rng) return fail("rt_render: adaptive passes need the counter-based RNG (rng_mode 0)");
    P.sample_base = s->adapt.sample_base; P.sample_count = s->adapt.sample_count;
    P.active = s->adapt.active.p; P.n_active = s->adapt.n_active;
  }
  P.work_total = (long long)(adaptive ? (size_t)P.n_active : n_pix) * P.sample_count;
  if (ref_rng) {
    P.n_slots = (int)n_pix;  // a slot is a pixel: one sequential XORWOW stream each
  } else {
    // paths in flight. Sweep on C4, 1000 spp (4 pools): 4 Mi 4080, 8 Mi 4280, 16 Mi 4388, 32 Mi 4486, 64 Mi 4537 Mrays/s (longer
    // kernels, fewer kernel tails); the same order at 30 and 125 spp. 32 Mi paths = 4.2 GB of path state and queues.
    long long target = p->slots > 0 ? p->slots : 32 * 1024 * 1024;
    if (p->slots <= 0) if (const char* e = getenv("RT_SLOTS")) target = atoll(e);
    target = (target + 127) / 128 * 128;
    P.n_slots = (int)std::max<long long>(0, std::min<long long>(target, P.work_total));
  }
  P.max_depth = p->max_depth > 0 ? p->max_depth : 50;
  if (P.max_depth > 255) return fail("rt_render: max_depth above 255 (the bounce count shares a state word with the sample number)");
  if ((long long)P.sample_base + P.sample_count >= (1 << 23)) return fail("rt_render: sample numbers above 2^23 (split the job into progressive passes)");
  P.tmin = p->t_min > 0 ? p->t_min : 0.001f;
  if (p->override_background) { P.background = v3(p->background[0], p->background[1], p->background[2]); P.gradient = p->gradient_bg; }
  else { P.background = v3(sd.background[0], sd.background[1], sd.background[2]); P.gradient = sd.gradient_bg; }
  P.seed = p->seed ? p->seed : 1984ull;
  const float gamma = p->gamma > 0 ? p->gamma : 2.2f;
  if (ensure_buffers(s, std::max(P.n_slots, 1), n_pix, ref_rng, p->aov != 0)) return 1;

  // Path pools: the paths are split into independent pools, each running its own trace -> shade chain on its own
  // stream. k_trace is issue-bound and k_shade latency-bound, so letting one pool's shade run under another
  // pool's trace fills issue slots that a single chain leaves idle. The pools share only the work counter and the
  // fixed-point accumulators (atomics). Reference-RNG mode and per-kernel profiling use one pool.
  int n_pools = (ref_rng || p->profile) ? 1 : P.n_slots >= 1024 * 1024 ? 4 : P.n_slots >= 256 * 1024 ? 2 : 1;
  if (const char* e = getenv("RT_POOLS")) if (!ref_rng && !p->profile) n_pools = std::max(1, std::min(RT_MAX_POOLS, atoi(e)));
  n_pools = std::max(1, std::min(n_pools, P.n_slots / (4 * RT_BLOCK)));  // every pool gets at least a few blocks of paths
  PathArrays A;
  memset(&A, 0, sizeof(A));
  A.col = s->col.p; A.rng = s->rng.p; A.acc64 = s->acc64.p; A.next_work = s->next_work.p;
  if (adaptive) {
    if (s->adapt.acc64_odd.n < 3 * n_pix) CU(s->adapt.acc64_odd.alloc(3 * n_pix));
    A.acc64_odd = s->adapt.acc64_odd.p;
  }
  cudaStream_t st = s->stream;
  cudaStream_t* streams = s->pool_stream;
  cudaEvent_t e0, e1, evs[RT_MAX_POOLS];
  Events evh;
  CU(evh.make(&e0)); CU(evh.make(&e1));
  for (int k = 0; k < RT_MAX_POOLS; ++k) CU(evh.make(&evs[k], cudaEventDisableTiming));
  CU(cudaEventRecord(e0, st));
  // pool geometry: pool k owns a contiguous range of every path array and of the queue storage
  RenderParams Pp[RT_MAX_POOLS];
  PathArrays Ap[RT_MAX_POOLS];
  size_t capk[RT_MAX_POOLS] = {0};
  {
    const int per = (P.n_slots / n_pools + RT_BLOCK - 1) / RT_BLOCK * RT_BLOCK;
    int base = 0; size_t abase = 0, qbase = 0;
    for (int k = 0; k < n_pools; ++k) {
      const int n = (k == n_pools - 1) ? P.n_slots - base : std::min(per, P.n_slots - base);
      Pp[k] = P; Pp[k].n_slots = n;
      Ap[k] = A;
      for (int b = 0; b < 2; ++b) { Ap[k].ray_o[b] = s->ray_o[b].p + abase; Ap[k].ray_d[b] = s->ray_d[b].p + abase; Ap[k].thr[b] = s->thr[b].p + abase; }
      Ap[k].hit = s->hit.p + abase;
      Ap[k].queues = s->queues.p + qbase;
      Ap[k].subcap = (int)pool_subcap(n);
      capk[k] = pool_cap(n);
      abase += capk[k]; qbase += (size_t)RT_NQ * pool_subcap(n);
      base += n;
    }
  }
  {
    for (int k = 0; k < RT_MAX_POOLS; ++k) { memset(&s->h_counters[k], 0, sizeof(WaveCounters)); if (k < n_pools) s->h_counters[k].order_len = Pp[k].n_slots; }
    CU(cudaMemcpyAsync(s->counters.p, s->h_counters, RT_MAX_POOLS * sizeof(WaveCounters), cudaMemcpyHostToDevice, st));
    // the first n_slots work items are drawn by k_init; reference-RNG mode does not use the counter
    const unsigned long long first_work = (unsigned long long)P.n_slots;
    CU(cudaMemcpyAsync(s->next_work.p, &first_work, sizeof(first_work), cudaMemcpyHostToDevice, st));
    if (!ref_rng && n_pix > 0 && (!adaptive || s->adapt.first)) {
      CU(cudaMemsetAsync(s->acc64.p, 0, 3 * n_pix * sizeof(unsigned long long), st));
      if (adaptive) CU(cudaMemsetAsync(s->adapt.acc64_odd.p, 0, 3 * n_pix * sizeof(unsigned long long), st));
    }
  }
  const int B = RT_BLOCK;
  int launches = 0, waves = 0, prof_waves = 0;
  double prof_trace_ms = 0.0, prof_shade_ms = 0.0;
  if (P.n_slots > 0 && P.sample_count > 0) {
    // Every wave of a pool runs k_trace over its dense layout and k_shade over the queues that wave filled (a lane
    // whose sample ended takes the next work item there). The host reads the queue fills back every `batch` waves: a
    // wave that traced no ray means the pool has run out of work; the job is done when all pools have.
    // A job whose samples all start in the first wave (no regeneration: C1 exactly is 900 000 samples) only shrinks from
    // wave to wave, so the host looks every 4 waves: the grids follow the ray count and the tail kernel takes over sooner (C1: 1.2 -> 0.94 ms).
    int batch = P.work_total <= (long long)P.n_slots ? 4 : 8;
    if (const char* e = getenv("RT_WAVE_BATCH")) batch = std::max(1, atoi(e));
    int tail_rays = 65536;  // a pool's wave at or below this many rays (work counter dry) is handed to k_finish; 0: never
    if (const char* e = getenv("RT_TAIL_RAYS")) tail_rays = std::max(0, atoi(e));
    FILE* wlog = nullptr;  // diagnostics: one line per wave (rays of the wave, k_trace ms, k_shade ms)
    if (p->profile) if (const char* e = getenv("RT_WAVE_LOG")) { wlog = fopen(e, "a"); batch = 1; }
    std::vector<cudaEvent_t> pev;
    if (p->profile) { pev.resize(3 * batch); for (auto& e : pev) CU(evh.make(&e)); }
    cudaEvent_t eset;
    CU(evh.make(&eset, cudaEventDisableTiming));
    CU(cudaEventRecord(eset, st));
    for (int k = 1; k < n_pools; ++k) CU(cudaStreamWaitEvent(streams[k], eset, 0));  // the other pools start after the counters are set
    int Gs[RT_MAX_POOLS], Gt[RT_MAX_POOLS];
    int finish_grid_max = 148 * 4;  // k_finish blocks the GPU holds at once
    {
      int dev = 0, sms = 0, per_sm = 0;
      if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_finish, 128, 0) == cudaSuccess && sms > 0 && per_sm > 0)
        finish_grid_max = sms * per_sm;
    }
    {
      int work_base = 0;  // pool k draws work items [work_base, work_base + n_k) in k_init
      for (int k = 0; k < n_pools; ++k) {
        const int G = (Pp[k].n_slots + B - 1) / B;
        Gs[k] = (int)((capk[k] + (size_t)B * RT_SHADE_ITEMS - 1) / ((size_t)B * RT_SHADE_ITEMS));
        Gt[k] = (int)((capk[k] / RT_RANGE + RT_TWARPS - 1) / RT_TWARPS);
        if (ref_rng) k_init<RNG_REFERENCE><<<G, B, 0, streams[k]>>>(s->dscene, Pp[k], Ap[k], work_base);
        else k_init<RNG_PHILOX><<<G, B, 0, streams[k]>>>(s->dscene, Pp[k], Ap[k], work_base);
        work_base += Pp[k].n_slots;
        ++launches;
      }
    }
    int parity = 0;
    // A BVH of a handful of nodes (the Cornell boxes: 3) has no stack entries worth dropping: the trace kernel without that check
    // (same hits, same ray count; RT_CULL_MIN_NODES in the environment for A/B and for the test that the image does not depend on it)
    int cull_min_nodes = RT_CULL_MIN_NODES;
    if (const char* e = getenv("RT_CULL_MIN_NODES")) cull_min_nodes = atoi(e);
    const bool small_bvh = s->n_nodes < cull_min_nodes;
    bool done[RT_MAX_POOLS];
    for (int k = 0; k < RT_MAX_POOLS; ++k) done[k] = k >= n_pools;
    auto all_done = [&]() { bool d = true; for (int k = 0; k < n_pools; ++k) d = d && done[k]; return d; };
    while (!all_done()) {
      for (int w = 0; w < batch; ++w) {
        for (int k = 0; k < n_pools; ++k) {
          if (done[k]) continue;
          cudaStream_t sk = streams[k];
          WaveCounters* Ck = s->counters.p + k;
          if (p->profile) CU(cudaEventRecord(pev[3 * w], sk));
          if (small_bvh) k_trace_small<<<Gt[k], RT_TBLOCK, 0, sk>>>(s->dscene, P.tmin, Ap[k].ray_o[parity], Ap[k].ray_d[parity], Ap[k].hit,
                                                                    Ap[k].queues, Ap[k].subcap, Ck, parity);
          else k_trace<<<Gt[k], RT_TBLOCK, 0, sk>>>(s->dscene, P.tmin, Ap[k].ray_o[parity], Ap[k].ray_d[parity], Ap[k].hit, Ap[k].queues,
                                                    Ap[k].subcap, Ck, parity);
          if (p->profile) CU(cudaEventRecord(pev[3 * w + 1], sk));
          if (ref_rng) k_shade<RNG_REFERENCE><<<Gs[k], B, 0, sk>>>(s->dscene, Pp[k], Ap[k], Ck, parity);
          else k_shade<RNG_PHILOX><<<Gs[k], B, 0, sk>>>(s->dscene, Pp[k], Ap[k], Ck, parity);
          if (p->profile) CU(cudaEventRecord(pev[3 * w + 2], sk));
          launches += 2;
        }
        parity ^= 1; ++waves;
      }
      // queue fills of the LAST wave of the batch (k_shade zeroes the other parity for the next k_trace)
      for (int k = 0; k < n_pools; ++k) {
        if (done[k]) continue;
        CU(cudaMemcpyAsync(&s->h_counters[k], s->counters.p + k, sizeof(WaveCounters), cudaMemcpyDeviceToHost, streams[k]));
        CU(cudaEventRecord(evs[k], streams[k]));
      }
      int last_all = 0;
      for (int k = 0; k < n_pools; ++k) {
        if (done[k]) continue;
        CU(cudaEventSynchronize(evs[k]));
        int last = 0;
        for (int q = 0; q < RT_NQ; ++q) last += s->h_counters[k].n_queue[parity ^ 1][q];
        done[k] = last == 0;
        last_all += last;
        // Tail of the job: once a pool's wave is less than half full the work counter has run dry (finished samples are
        // no longer replaced), so its waves only shrink from here on: launch grids that fit what is left instead of
        // grids for the pool's capacity (thousands of blocks that start only to find nothing to do).
        if (2 * (long long)last < Pp[k].n_slots) {
            const size_t cap = pool_cap((size_t)last);
            if (!ref_rng && !wlog && last > 0 && last <= tail_rays) {
              // few enough paths left: every lane keeps its path to the end in ONE kernel instead of ~40 more waves that
              // each cost the latency of their slowest ray plus two launches
              int spread_log2 = 0;  // as thin as one resident grid allows (C1 exactly, hand-over at 39 Ki paths: 1 path per lane 0.95 ms, per 32 lanes 1.31)
              while (spread_log2 < 5 && ((cap << (spread_log2 + 1)) + 127) / 128 <= (size_t)finish_grid_max) ++spread_log2;
              k_finish<<<(int)(((cap << spread_log2) + 127) / 128), 128, 0, streams[k]>>>(s->dscene, Pp[k], Ap[k], s->counters.p + k, parity, spread_log2);
              ++launches;
              done[k] = true;
            }
            Gs[k] = (int)std::max<size_t>(1, std::min<size_t>((size_t)Gs[k], (cap + (size_t)B * RT_SHADE_ITEMS - 1) / ((size_t)B * RT_SHADE_ITEMS)));
            Gt[k] = (int)std::max<size_t>(1, std::min<size_t>((size_t)Gt[k], (cap / RT_RANGE + RT_TWARPS - 1) / RT_TWARPS));
        }
      }
      if (p->profile) {
        for (int w = 0; w < batch; ++w) {
          float a = 0.f, b = 0.f;
          CU(cudaEventElapsedTime(&a, pev[3 * w], pev[3 * w + 1]));
          CU(cudaEventElapsedTime(&b, pev[3 * w + 1], pev[3 * w + 2]));
          prof_trace_ms += a; prof_shade_ms += b; ++prof_waves;
          if (wlog) fprintf(wlog, "%d %d %.4f %.4f\n", prof_waves, last_all, a, b);
        }
      }
    }
    if (wlog) fclose(wlog);
    for (int k = 1; k < n_pools; ++k) {  // pool 0's stream carries on alone: it waits for the other pools' last kernels
      CU(cudaEventRecord(evs[k], streams[k]));
      CU(cudaStreamWaitEvent(st, evs[k], 0));
    }
  }
  {
    const int Gp = (int)((n_pix + 255) / 256);
    if (n_pix > 0) {
      const int add = (p->accumulate != 0 && s->accum_valid_pix == n_pix) ? 1 : 0;
      if (ref_rng) k_accumulate<RNG_REFERENCE><<<Gp, 256, 0, st>>>(P, A, s->accum.p, add);
      else k_accumulate<RNG_PHILOX><<<Gp, 256, 0, st>>>(P, A, s->accum.p, add);
      s->accum_valid_pix = n_pix;
      k_resolve<<<Gp, 256, 0, st>>>((int)n_pix, spp_total, gamma, s->accum.p, s->fb.p);
      launches += 2;
    }
  }
  CU(cudaEventRecord(e1, st));
  if (p->aov && n_pix > 0) {
    k_aov<<<(int)((n_pix + B - 1) / B), B, 0, st>>>(s->dscene, P, s->aov_obj.p, s->aov_mat.p, s->aov_t.p, s->counters.p);
  }
  CU(cudaMemcpyAsync(s->h_counters, s->counters.p, RT_MAX_POOLS * sizeof(WaveCounters), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  unsigned long long rays_all = 0; unsigned int overflow_all = 0, nonfinite_all = 0;
  for (int k = 0; k < RT_MAX_POOLS; ++k) { rays_all += s->h_counters[k].rays; overflow_all |= s->h_counters[k].overflow; nonfinite_all += s->h_counters[k].nonfinite; }
  s->last = P; s->has_aov = p->aov != 0; s->last_gamma = gamma; s->last_spp_total = spp_total;
  rt_render_stats& R = s->stats;
  memset(&R, 0, sizeof(R));
  R.device_ms = ms; R.rays = rays_all; R.samples = (uint64_t)(adaptive ? (size_t)P.n_active : n_pix) * (uint64_t)P.sample_count;
  R.waves = waves; R.kernel_launches = launches; R.rows_local = P.rows_local; R.nx = P.nx;
  R.nonfinite_samples = (int32_t)nonfinite_all; R.n_slots = P.n_slots; R.stack_overflow = overflow_all;
  R.profiled_waves = prof_waves; R.trace_ms = prof_trace_ms; R.shade_ms = prof_shade_ms;
  if (device_ms) *device_ms = ms;
  if (rays) *rays = R.rays;
  if (R.stack_overflow) return fail("rt_render: BVH traversal stack overflow (RT_STACK too small for this scene)");
  return 0;
}

#ifdef RT_STATS
extern "C" int rt_debug_stats(unsigned long long* out8, int reset) {
  CU(cudaMemcpyFromSymbol(out8, rt::g_stats, 8 * sizeof(unsigned long long)));
  CU(cudaMemcpyFromSymbol(out8 + 8, rt::g_stats2, 4 * sizeof(unsigned long long)));
  if (reset) { unsigned long long z[8] = {0}; CU(cudaMemcpyToSymbol(rt::g_stats, z, sizeof(z))); CU(cudaMemcpyToSymbol(rt::g_stats2, z, 4 * sizeof(unsigned long long))); }
  return 0;
}
#endif

extern "C" int rt_render_stats_get(rt_scene* s, rt_render_stats* out) {
  if (!s || !out) return fail("rt_render_stats_get: null argument");
  *out = s->stats;
  return 0;
}

extern "C" int rt_readback(rt_scene* s, float* rgb, int32_t* obj_id, int32_t* mat_id) {
  if (!s) return fail("rt_readback: null scene");
  CU(cudaSetDevice(s->device));
  const size_t n_pix = (size_t)s->last.rows_local * s->last.nx;
  if (n_pix == 0) return 0;
  if (rgb) CU(cudaMemcpy(rgb, s->fb.p, n_pix * 3 * sizeof(float), cudaMemcpyDeviceToHost));
  if (obj_id || mat_id) {
    if (!s->has_aov) return fail("rt_readback: the last rt_render had aov = 0");
    if (obj_id) {
      CU(cudaMemcpy(obj_id, s->aov_obj.p, n_pix * sizeof(int), cudaMemcpyDeviceToHost));
      if (s->grouped)  // a hit on a member of a group reports the top-level entry the group sits in
        for (size_t i = 0; i < n_pix; ++i) if (obj_id[i] >= 0) obj_id[i] = s->origin[obj_id[i]];
    }
    if (mat_id) CU(cudaMemcpy(mat_id, s->aov_mat.p, n_pix * sizeof(int), cudaMemcpyDeviceToHost));
  }
  return 0;
}
extern "C" int rt_readback_t(rt_scene* s, float* t) {
  if (!s || !t) return fail("rt_readback_t: null argument");
  if (!s->has_aov) return fail("rt_readback_t: the last rt_render had aov = 0");
  CU(cudaSetDevice(s->device));
  const size_t n_pix = (size_t)s->last.rows_local * s->last.nx;
  if (n_pix) CU(cudaMemcpy(t, s->aov_t.p, n_pix * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int rt_accum_device_ptr(rt_scene* s, void** dptr, size_t* n_floats) {
  if (!s || !dptr) return fail("rt_accum_device_ptr: null argument");
  *dptr = s->accum.p;
  if (n_floats) *n_floats = (size_t)s->last.rows_local * s->last.nx * 3;
  return 0;
}
extern "C" int rt_fb_device_ptr(rt_scene* s, void** dptr, size_t* n_floats) {
  if (!s || !dptr) return fail("rt_fb_device_ptr: null argument");
  *dptr = s->fb.p;
  if (n_floats) *n_floats = (size_t)s->last.rows_local * s->last.nx * 3;
  return 0;
}
// ---- adaptive per-tile sampling ----
// Passes of pass_spp samples; after every pass the error of every still-active tile is estimated from the two half-buffers
// (k_tile_error) and tiles below the threshold stop taking samples. Sample numbers are global (pass k of a tile draws
// numbers [k * pass_spp, (k + 1) * pass_spp)), the sums are fixed-point: a tile that runs all passes holds bit for bit the
// sums of a plain max_spp render, and with threshold = 0 the whole image equals rt_render(spp = max_spp).
extern "C" int rt_render_adaptive(rt_scene* s, const rt_render_params* p, const rt_adaptive_params* a, rt_adaptive_stats* out) {
  if (!s || !p || !a) return fail("rt_render_adaptive: null argument");
  if (p->rng_mode != 0) return fail("rt_render_adaptive: needs the counter-based RNG (rng_mode 0)");
  if (p->split_mode == 1 && p->world > 1) return fail("rt_render_adaptive: tile split only (every rank adapts its own scanlines)");
  const int tile = a->tile > 0 ? a->tile : 16;
  const int max_spp = a->max_spp > 0 ? a->max_spp : (p->spp > 0 ? p->spp : s->sd.default_spp);
  const int pass_spp = std::max(2, a->pass_spp > 0 ? a->pass_spp : 16) & ~1;  // even: both half-buffers grow in every pass
  const int min_spp = std::max(pass_spp, a->min_spp);
  if (max_spp < pass_spp) return fail("rt_render_adaptive: max_spp below one pass");
  CU(cudaSetDevice(s->device));
  const int world = p->world > 0 ? p->world : 1;
  if (p->rank < 0 || p->rank >= world) return fail("rt_render_adaptive: rank out of range");
  const int nx = s->sd.nx, rows = (s->sd.ny - p->rank + world - 1) / world;
  const size_t n_pix = (size_t)nx * rows;
  const int tiles_x = (nx + tile - 1) / tile, tiles_y = (rows + tile - 1) / tile, n_tiles = tiles_x * tiles_y;
  rt_scene::Adaptive& ad = s->adapt;
  ad.tile = tile; ad.tiles_x = tiles_x; ad.tiles_y = tiles_y;
  ad.tile_spp.assign(n_tiles, 0);
  std::vector<int> n_even(n_tiles, 0), n_odd(n_tiles, 0), list;
  std::vector<char> active(n_tiles, 1), below(n_tiles, 0);
  std::vector<float> err(n_tiles, 0.f);
  DBuf<int> d_even, d_odd; DBuf<float> d_err;
  CU(d_even.alloc(std::max(n_tiles, 1))); CU(d_odd.alloc(std::max(n_tiles, 1))); CU(d_err.alloc(std::max(n_tiles, 1)));
  rt_render_params pp = *p;
  pp.accumulate = 0; pp.aov = 0;
  rt_adaptive_stats R; memset(&R, 0, sizeof(R));
  R.tiles = n_tiles;
  struct Off { rt_scene::Adaptive& a; ~Off() { a.on = false; } } off{ad};
  int done = 0, rc = 0;
  while (n_pix > 0 && done < max_spp) {
    const int count = std::min(pass_spp, max_spp - done);
    list.clear();
    for (int t = 0; t < n_tiles; ++t) {
      if (!active[t]) continue;
      const int tx = t % tiles_x, ty = t / tiles_x;
      for (int j = ty * tile; j < std::min(rows, (ty + 1) * tile); ++j)
        for (int i = tx * tile; i < std::min(nx, (tx + 1) * tile); ++i) list.push_back(j * nx + i);
    }
    if (list.empty()) break;
    if (ad.active.n < n_pix) CU(ad.active.alloc(n_pix));
    CU(cudaMemcpy(ad.active.p, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice));
    ad.on = true; ad.first = done == 0; ad.sample_base = done; ad.sample_count = count; ad.n_active = (int)list.size();
    if ((rc = rt_render(s, &pp, nullptr, nullptr)) != 0) return rc;
    R.samples += s->stats.samples; R.rays += s->stats.rays; R.device_ms += s->stats.device_ms; ++R.passes;
    // sample numbers [done, done + count): how many are odd
    const int odd = (done + count) / 2 - done / 2, even = count - odd;
    for (int t = 0; t < n_tiles; ++t) if (active[t]) { n_even[t] += even; n_odd[t] += odd; ad.tile_spp[t] += count; }
    done += count;
    if (done >= max_spp) break;
    CU(cudaMemcpy(d_even.p, n_even.data(), n_tiles * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_odd.p, n_odd.data(), n_tiles * sizeof(int), cudaMemcpyHostToDevice));
    k_tile_error<<<n_tiles, 256, 0, s->stream>>>(nx, rows, tile, tiles_x, s->acc64.p, ad.acc64_odd.p, d_even.p, d_odd.p, d_err.p);
    CU(cudaMemcpyAsync(err.data(), d_err.p, n_tiles * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    if (const char* e = getenv("RT_ADAPTIVE_LOG")) {  // diagnostics: the error map of every pass, one row of tiles per line
      if (FILE* f = fopen(e, "a")) {
        fprintf(f, "pass %d, %d spp:\n", R.passes, done);
        for (int t = 0; t < n_tiles; ++t) fprintf(f, "%s%.4g%s", active[t] ? "" : "(", err[t], (t + 1) % tiles_x == 0 ? (active[t] ? "\n" : ")\n") : (active[t] ? " " : ") "));
        fclose(f);
      }
    }
    R.err_min = FLT_MAX; R.err_max = 0.f; R.err_spp = done;
    for (int t = 0; t < n_tiles; ++t) if (active[t]) { R.err_min = std::min(R.err_min, err[t]); R.err_max = std::max(R.err_max, err[t]); }
    // a tile stops when its estimate has been below the threshold at TWO checks in a row (a single lucky estimate -
    // the half-buffers agreeing by chance - must not end it), and not before min_spp
    for (int t = 0; t < n_tiles; ++t) {
      if (!active[t]) continue;
      below[t] = err[t] < a->threshold ? below[t] + 1 : 0;
      if (done >= min_spp && below[t] >= 2) active[t] = 0;
    }
  }
  ad.on = false;
  // accum = per-pixel MEAN (every tile divides by its own sample count), framebuffer = gamma(mean)
  if (n_pix > 0) {
    CU(cudaMemcpy(d_even.p, n_even.data(), n_tiles * sizeof(int), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_odd.p, n_odd.data(), n_tiles * sizeof(int), cudaMemcpyHostToDevice));
    const int Gp = (int)((n_pix + 255) / 256);
    PathArrays A; memset(&A, 0, sizeof(A)); A.acc64 = s->acc64.p;
    k_accumulate<RNG_PHILOX><<<Gp, 256, 0, s->stream>>>(s->last, A, s->accum.p, 0);
    k_normalize_tiles<<<Gp, 256, 0, s->stream>>>(nx, rows, tile, tiles_x, d_even.p, d_odd.p, s->accum.p);
    k_resolve<<<Gp, 256, 0, s->stream>>>((int)n_pix, 1, p->gamma > 0 ? p->gamma : 2.2f, s->accum.p, s->fb.p);
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    s->last_spp_total = 1;
  }
  int lo = 1 << 30, hi = 0; double tot = 0.0;
  for (int t = 0; t < n_tiles; ++t) { lo = std::min(lo, ad.tile_spp[t]); hi = std::max(hi, ad.tile_spp[t]); R.tiles_converged += active[t] ? 0 : 1; }
  tot = n_pix ? (double)R.samples / (double)n_pix : 0.0;
  R.min_spp_used = n_tiles ? lo : 0; R.max_spp_used = hi; R.mean_spp = (float)tot;
  s->stats.samples = R.samples; s->stats.rays = R.rays; s->stats.device_ms = R.device_ms;
  if (out) *out = R;
  return 0;
}
extern "C" int rt_readback_spp(rt_scene* s, int32_t* spp) {
  if (!s || !spp) return fail("rt_readback_spp: null argument");
  const rt_scene::Adaptive& ad = s->adapt;
  if (ad.tile <= 0) return fail("rt_readback_spp: no adaptive render on this scene yet");
  const int nx = s->last.nx, rows = s->last.rows_local;
  for (int j = 0; j < rows; ++j)
    for (int i = 0; i < nx; ++i) spp[(size_t)j * nx + i] = ad.tile_spp[(j / ad.tile) * ad.tiles_x + i / ad.tile];
  return 0;
}

// ---- native multi-GPU exchange of the spp split (one process, one rt_scene per device) ----
__global__ void k_accum_add(float* dst, const float* src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = fadd(dst[i], src[i]);
}
extern "C" int rt_accum_reduce(rt_scene* dst, rt_scene* const* src, int32_t n_src) {
  if (!dst || (n_src > 0 && !src)) return fail("rt_accum_reduce: null argument");
  CU(cudaSetDevice(dst->device));
  const size_t n = (size_t)dst->last.rows_local * dst->last.nx * 3;
  if (n == 0 || n_src <= 0) return 0;
  const int G = (int)((n + 255) / 256);
  for (int k = 0; k < n_src; ++k) {
    rt_scene* o = src[k];
    if (!o || o == dst) return fail("rt_accum_reduce: bad source scene");
    if ((size_t)o->last.rows_local * o->last.nx * 3 != n || o->accum_valid_pix * 3 != n)
      return fail("rt_accum_reduce: the scenes do not hold accumulation buffers of the same share (render all of them with the same "
                  "resolution and split first)");
    const float* from = o->accum.p;
    if (o->device != dst->device) {
      // peer memory over NVLink when the devices allow it: the add kernel loads the remote buffer directly (no staging
      // copy); otherwise one cudaMemcpyPeerAsync into a buffer on dst's device
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, dst->device, o->device));
      bool direct = false;
      if (can) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(o->device, 0);
        if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) direct = true;
        cudaGetLastError();
      }
      if (!direct) {
        if (dst->reduce_tmp.n < n) CU(dst->reduce_tmp.alloc(n));
        CU(cudaMemcpyPeerAsync(dst->reduce_tmp.p, dst->device, o->accum.p, o->device, n * sizeof(float), dst->stream));
        from = dst->reduce_tmp.p;
      }
    }
    k_accum_add<<<G, 256, 0, dst->stream>>>(dst->accum.p, from, n);
  }
  CU(cudaStreamSynchronize(dst->stream));
  CU(cudaGetLastError());
  return 0;
}

// ---- dynamic tile queue over the devices of one process ----
// The image is cut into n_chunks shares of interleaved scanlines (share c = rows j = c mod n_chunks, the tile split of
// rt_render_params with world = n_chunks). One host thread per scene replica pulls the next share from a shared counter,
// renders it and copies its rows into the caller's full image, until none is left: a GPU that is slower, busier, or
// got the expensive rows simply takes fewer shares. Which device rendered which share does not change a single bit
// (a share's samples depend on pixel and sample number only).
#include <atomic>
#include <thread>
extern "C" int rt_render_queue(rt_scene* const* scenes, int32_t n_scenes, const rt_render_params* p, int32_t n_chunks, float* rgb_full,
                               rt_queue_stats* out) {
  if (!scenes || n_scenes <= 0 || !p || !rgb_full) return fail("rt_render_queue: null argument");
  if (n_scenes > RT_QUEUE_MAX_DEVICES) return fail("rt_render_queue: too many scene replicas");
  if (p->split_mode != 0) return fail("rt_render_queue: the queue hands out tile shares (split_mode 0)");
  for (int k = 0; k < n_scenes; ++k) {
    if (!scenes[k]) return fail("rt_render_queue: null scene");
    if (scenes[k]->sd.nx != scenes[0]->sd.nx || scenes[k]->sd.ny != scenes[0]->sd.ny) return fail("rt_render_queue: the replicas differ in resolution");
  }
  const int nx = scenes[0]->sd.nx, ny = scenes[0]->sd.ny;
  if (n_chunks <= 0) n_chunks = std::min(ny, 8 * n_scenes);
  n_chunks = std::min(n_chunks, ny);
  std::atomic<int> next(0);
  rt_queue_stats R; memset(&R, 0, sizeof(R));
  R.n_chunks = n_chunks;
  std::vector<std::string> errs(n_scenes);
  std::vector<std::thread> th;
  for (int k = 0; k < n_scenes; ++k)
    th.emplace_back([&, k]() {
      rt_scene* s = scenes[k];
      std::vector<float> share;
      for (;;) {
        const int c = next.fetch_add(1);
        if (c >= n_chunks) break;
        rt_render_params pp = *p;
        pp.rank = c; pp.world = n_chunks; pp.split_mode = 0; pp.accumulate = 0;
        if (rt_render(s, &pp, nullptr, nullptr) != 0) { errs[k] = rt_last_error(); next.store(n_chunks); break; }
        const int rows = s->stats.rows_local;
        share.resize((size_t)rows * nx * 3);
        if (rt_readback(s, share.data(), nullptr, nullptr) != 0) { errs[k] = rt_last_error(); next.store(n_chunks); break; }
        for (int lr = 0; lr < rows; ++lr)
          memcpy(rgb_full + (size_t)(lr * n_chunks + c) * nx * 3, share.data() + (size_t)lr * nx * 3, (size_t)nx * 3 * sizeof(float));
        R.chunks_per_device[k] += 1; R.device_ms[k] += s->stats.device_ms; R.rays_per_device[k] += s->stats.rays;
      }
    });
  for (auto& t : th) t.join();
  for (int k = 0; k < n_scenes; ++k) if (!errs[k].empty()) return fail("rt_render_queue: replica " + std::to_string(k) + ": " + errs[k]);
  for (int k = 0; k < n_scenes; ++k) R.rays += R.rays_per_device[k];
  if (out) *out = R;
  return 0;
}

extern "C" int rt_resolve(rt_scene* s, int32_t total_spp, float gamma) {
  if (!s) return fail("rt_resolve: null scene");
  CU(cudaSetDevice(s->device));
  const size_t n_pix = (size_t)s->last.rows_local * s->last.nx;
  if (n_pix == 0) return 0;
  k_resolve<<<(int)((n_pix + 255) / 256), 256, 0, s->stream>>>((int)n_pix, total_spp > 0 ? total_spp : s->last_spp_total,
                                                              gamma > 0 ? gamma : s->last_gamma, s->accum.p, s->fb.p);
  CU(cudaStreamSynchronize(s->stream));
  CU(cudaGetLastError());
  return 0;
}

extern "C" int rt_load_texture(const char* path, unsigned char* rgb, size_t cap, int32_t* w, int32_t* h) {
  if (!path || !w || !h) return fail("rt_load_texture: null argument");
  HostImage im;
  try {
    const std::string err = load_texture_file(path, im);
    if (!err.empty()) return fail("rt_load_texture: " + err);
  } catch (const std::exception& e) {
    return fail(std::string("rt_load_texture: ") + e.what());
  }
  *w = im.width; *h = im.height;
  if (rgb && cap >= im.px.size()) memcpy(rgb, im.px.data(), im.px.size());
  return 0;
}

// ---- image writers ----
static int to_byte(float c, bool dbl) {  // the reference's int(255.99 * c) (float product; double for bouncing_spheres, main.cu:722-724)
  return dbl ? int(255.99 * c) : int(255.99f * c);
}
static uint32_t crc32_of(const unsigned char* p, size_t n, uint32_t crc) {
  static uint32_t table[256]; static bool init = false;
  if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}
static void png_chunk(std::string& out, const char* type, const std::string& data) {
  auto be32 = [&](uint32_t v) { char b[4] = {(char)(v >> 24), (char)(v >> 16), (char)(v >> 8), (char)v}; out.append(b, 4); };
  be32((uint32_t)data.size());
  std::string td = std::string(type, 4) + data;
  out += td;
  be32(crc32_of((const unsigned char*)td.data(), td.size(), 0));
}

// format 0: ASCII P3 exactly as the reference prints it (no clamp unless asked: main.cu:1212-1221 prints whatever
// int(255.99*c) gives); 1: binary P6; 2: PNG (8-bit RGB, stored deflate blocks - no compression library needed).
// P6 and PNG hold bytes, so they always clamp to [0, 255]. Rows are written top scanline first (j = ny-1..0).
extern "C" long rt_write_image(const char* path, const float* rgb, int32_t nx, int32_t ny, int32_t format, int32_t clamp,
                               int32_t double_scale) {
  if (!rgb || nx <= 0 || ny <= 0 || format < 0 || format > 2) { fail("rt_write_image: bad argument"); return -1; }
  const bool dbl = double_scale != 0, clip = clamp != 0 || format != 0;
  auto chan = [&](float c) { int v = to_byte(c, dbl); if (clip) v = v < 0 ? 0 : v > 255 ? 255 : v; return v; };
  std::string out;
  char line[64];
  if (format == 0) {
    out.reserve((size_t)nx * ny * 12 + 32);
    snprintf(line, sizeof(line), "P3\n%d %d\n255\n", nx, ny);
    out += line;
    for (int j = ny - 1; j >= 0; j--)
      for (int i = 0; i < nx; i++) {
        const float* c = rgb + 3 * ((size_t)j * nx + i);
        snprintf(line, sizeof(line), "%d %d %d\n", chan(c[0]), chan(c[1]), chan(c[2]));
        out += line;
      }
  } else {
    std::string raw;  // top row first; PNG rows carry a filter byte (0 = none)
    raw.reserve(((size_t)nx * 3 + 1) * ny);
    for (int j = ny - 1; j >= 0; j--) {
      if (format == 2) raw.push_back(0);
      for (int i = 0; i < nx; i++) {
        const float* c = rgb + 3 * ((size_t)j * nx + i);
        raw.push_back((char)chan(c[0])); raw.push_back((char)chan(c[1])); raw.push_back((char)chan(c[2]));
      }
    }
    if (format == 1) {
      snprintf(line, sizeof(line), "P6\n%d %d\n255\n", nx, ny);
      out = line + raw;
    } else {
      out.assign("\x89PNG\r\n\x1a\n", 8);
      std::string ihdr(13, 0);
      for (int k = 0; k < 4; ++k) { ihdr[k] = (char)((uint32_t)nx >> (24 - 8 * k)); ihdr[4 + k] = (char)((uint32_t)ny >> (24 - 8 * k)); }
      ihdr[8] = 8; ihdr[9] = 2;  // 8 bits per channel, colour type 2 (RGB)
      png_chunk(out, "IHDR", ihdr);
      std::string z("\x78\x01", 2);  // zlib header, then stored (uncompressed) deflate blocks of <= 65535 bytes
      uint32_t a1 = 1, a2 = 0;  // Adler-32 of the raw data
      for (size_t off = 0; off < raw.size() || off == 0; off += 65535) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        const bool last = off + n >= raw.size();
        const char hdr[5] = {(char)(last ? 1 : 0), (char)(n & 0xFF), (char)(n >> 8), (char)(~n & 0xFF), (char)((~n >> 8) & 0xFF)};
        z.append(hdr, 5);
        z.append(raw, off, n);
        for (size_t i = 0; i < n; ++i) { a1 = (a1 + (unsigned char)raw[off + i]) % 65521u; a2 = (a2 + a1) % 65521u; }
        if (last) break;
      }
      const uint32_t ad = (a2 << 16) | a1;
      const char adl[4] = {(char)(ad >> 24), (char)(ad >> 16), (char)(ad >> 8), (char)ad};
      z.append(adl, 4);
      png_chunk(out, "IDAT", z);
      png_chunk(out, "IEND", "");
    }
  }
  FILE* f = path ? fopen(path, "wb") : stdout;
  if (!f) { fail("rt_write_image: cannot open output"); return -1; }
  const size_t w = fwrite(out.data(), 1, out.size(), f);
  if (path) fclose(f); else fflush(f);
  return (long)w;
}

extern "C" long rt_write_ppm(const char* path, const float* rgb, int32_t nx, int32_t ny, int32_t double_scale) {
  return rt_write_image(path, rgb, nx, ny, 0, 0, double_scale);
}
