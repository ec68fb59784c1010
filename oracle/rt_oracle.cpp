// oracle/rt_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain, single-file CPU restatement of the reference's render path (slbouknight/accelerated-ray-tracer,
// citations are /root/reference/src/<file>:<line>), written the way the reference is shaped: an object
// tree walked by recursion, the reference's own binary BVH (built with its selection sort and median
// split, traversed with its left/right/tie rule), one sequential XORWOW stream per pixel. It reads the
// flat scene description (include/rt_scene_desc.h) — either the product's export or the dump the
// reference's own CUDA build wrote into tests/golden — and answers:
//   * primary-hit object / material / t per pixel            (bvh_node::hit, bvh.cuh:95-106)
//   * the rendered float framebuffer in reference-RNG mode   (render/color, main.cu:44-133)
//   * the reference BVH's leaf order                         (bvh_node ctor, bvh.cuh:29-84)
//   * XORWOW known answers                                   (curand_kernel.h:772-874)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
// product (librt_b200.so) never does and has no CPU path.
//
// PINNED: tests/test_oracle.py checks it against the outputs of the reference's own sm_100 CUDA build
// (tests/golden/ref_gpu/*.npz): primary-hit ids and t bit for bit on every golden scene, leaf order,
// and framebuffers within a stated tolerance (host libm vs CUDA libdevice for sinf/acosf/atan2f/
// logf/powf; everything made of + - * / sqrt fma is bit-exact because the fusion pattern of the
// reference's SASS is written out explicitly below with fmaf()).
//
// Float contraction of the reference build (nvcc -fmad=true; read off its sm_100 SASS, see
// csrc/rt_math.h for the list): a*b + c*d -> fma(a,b,c*d); x +- a*b -> fma(+-a,b,x);
// dot = fma(a2,b2,fma(a0,b0,a1*b1)). Compile with -ffp-contract=off so nothing else fuses.
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <vector>
#include "rt_scene_desc.h"

namespace {

struct vec3 { float x, y, z; };
inline vec3 V(float x, float y, float z) { vec3 r = {x, y, z}; return r; }
inline vec3 operator+(vec3 a, vec3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }          // vec3.cuh:57
inline vec3 operator-(vec3 a, vec3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }          // vec3.cuh:62
inline vec3 operator-(vec3 a) { return V(-a.x, -a.y, -a.z); }
inline vec3 operator*(vec3 a, vec3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }          // vec3.cuh:67
inline vec3 operator*(float t, vec3 a) { return V(t * a.x, t * a.y, t * a.z); }               // vec3.cuh:77
inline vec3 operator/(vec3 a, float t) { return V(a.x / t, a.y / t, a.z / t); }               // vec3.cuh:82
inline float dot(vec3 a, vec3 b) { return fmaf(a.z, b.z, fmaf(a.x, b.x, a.y * b.y)); }        // vec3.cuh:92 (fused as in SASS)
inline vec3 cross(vec3 a, vec3 b) {                                                           // vec3.cuh:97
  return V(fmaf(a.y, b.z, -(a.z * b.y)), -fmaf(a.x, b.z, -(a.z * b.x)), fmaf(a.x, b.y, -(a.y * b.x)));
}
inline float length(vec3 a) { return sqrtf(dot(a, a)); }                                      // vec3.cuh:32
inline vec3 unit_vector(vec3 a) { return a / length(a); }                                     // vec3.cuh:155
inline vec3 madd(float t, vec3 b, vec3 a) { return V(fmaf(t, b.x, a.x), fmaf(t, b.y, a.y), fmaf(t, b.z, a.z)); }  // a + t*b
inline vec3 V3(const float* p) { return V(p[0], p[1], p[2]); }
inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

struct ray { vec3 A, B; float tm; };  // ray.cuh:5-21 (time is double there; every consumer rounds it to float)
inline vec3 point_at(const ray& r, float t) { return madd(t, r.B, r.A); }  // ray.cuh:16

// ---- cuRAND XORWOW, curand_init(seed, 0, 0) / curand / curand_uniform (curand_kernel.h:772-874, curand_uniform.h:69-72)
struct xorwow {
  uint32_t d, v[5];
  void init(unsigned long long seed) {
    uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u, s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    d = 6615241u + t1 + t0;
    v[0] = 123456789u + t0; v[1] = 362436069u ^ t0; v[2] = 521288629u + t1; v[3] = 88675123u ^ t1; v[4] = 5783321u + t0;
  }
  uint32_t next() {
    uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
    d += 362437u;
    return v[4] + d;
  }
  float uniform() { return fmaf((float)next(), 2.3283064e-10f, 1.1641532e-10f); }
};

struct hit_record { float t; vec3 p, normal; int mat; float u, v; int top; };  // hittable.cuh:13-21 (+ which d_list entry)

struct aabb { vec3 mn, mx; };
// aabb::hit, aabb.cuh:45-61
inline bool aabb_hit(const aabb& b, const ray& r, float tmin, float tmax) {
  const float o[3] = {r.A.x, r.A.y, r.A.z}, d[3] = {r.B.x, r.B.y, r.B.z};
  const float lo[3] = {b.mn.x, b.mn.y, b.mn.z}, hi[3] = {b.mx.x, b.mx.y, b.mx.z};
  for (int a = 0; a < 3; a++) {
    float invD = 1.0f / d[a];
    float t0 = (lo[a] - o[a]) * invD;
    float t1 = (hi[a] - o[a]) * invD;
    if (invD < 0.0f) { float tmp = t0; t0 = t1; t1 = tmp; }
    tmin = t0 > tmin ? t0 : tmin;
    tmax = t1 < tmax ? t1 : tmax;
    if (tmax <= tmin) return false;
  }
  return true;
}

struct bvh_node { int left, right; int leaf; aabb box; };  // leaf >= 0: left == right == that d_list entry (bvh.cuh:38-43)

struct Scene {
  rt_sd_header h;
  std::vector<rt_texture_desc> tex;
  std::vector<rt_material_desc> mat;
  std::vector<rt_object_desc> obj;
  std::vector<int> top;
  std::vector<rt_image_desc> img;
  std::vector<const unsigned char*> img_px;
  std::vector<bvh_node> nodes;  // nodes[0] = root (when top is non-empty)
  std::vector<int> order;       // d_list after the in-place selection sorts: order[pos] = index into top
};

inline aabb obj_box(const Scene& S, int o) { aabb b = {V3(S.obj[o].box_min), V3(S.obj[o].box_max)}; return b; }

// bvh_node::bvh_node(objects, start, end), bvh.cuh:29-84. `list` holds indices into S.top.
int build_bvh(Scene& S, std::vector<int>& list, int start, int end) {
  const int n = end - start;
  const int me = (int)S.nodes.size();
  S.nodes.push_back(bvh_node());
  if (n == 1) {
    S.nodes[me].left = S.nodes[me].right = -1;
    S.nodes[me].leaf = list[start];
    S.nodes[me].box = obj_box(S, S.top[list[start]]);
    return me;
  }
  float minx = 1e30f, maxx = -1e30f, miny = 1e30f, maxy = -1e30f, minz = 1e30f, maxz = -1e30f;
  for (int i = start; i < end; ++i) {
    const vec3 mn = obj_box(S, S.top[list[i]]).mn;
    if (mn.x < minx) minx = mn.x; if (mn.x > maxx) maxx = mn.x;
    if (mn.y < miny) miny = mn.y; if (mn.y > maxy) maxy = mn.y;
    if (mn.z < minz) minz = mn.z; if (mn.z > maxz) maxz = mn.z;
  }
  const float sx = maxx - minx, sy = maxy - miny, sz = maxz - minz;
  int axis = 0;
  if (sy > sx && sy >= sz) axis = 1;
  else if (sz > sx && sz >= sy) axis = 2;
  // the reference's selection sort (bvh.cuh:66-77), on a key array so that n = 10^4 stays fast
  std::vector<float> key(n);
  for (int i = 0; i < n; ++i) key[i] = S.obj[S.top[list[start + i]]].box_min[axis];
  for (int i = 0; i < n - 1; ++i) {
    int best = i;
    for (int j = i + 1; j < n; ++j) if (key[j] < key[best]) best = j;
    if (best != i) {
      float tk = key[i]; key[i] = key[best]; key[best] = tk;
      int tl = list[start + i]; list[start + i] = list[start + best]; list[start + best] = tl;
    }
  }
  const int mid = start + (n >> 1);
  const int l = build_bvh(S, list, start, mid);
  const int r = build_bvh(S, list, mid, end);
  S.nodes[me].left = l; S.nodes[me].right = r; S.nodes[me].leaf = -1;
  const aabb a = S.nodes[l].box, b = S.nodes[r].box;  // aabb::surrounding_box, aabb.cuh:34-43
  S.nodes[me].box.mn = V(fminf(a.mn.x, b.mn.x), fminf(a.mn.y, b.mn.y), fminf(a.mn.z, b.mn.z));
  S.nodes[me].box.mx = V(fmaxf(a.mx.x, b.mx.x), fmaxf(a.mx.y, b.mx.y), fmaxf(a.mx.z, b.mx.z));
  return me;
}

bool object_hit(const Scene& S, int o, const ray& r, float tmin, float tmax, hit_record& rec);

// sphere::hit, sphere.cuh:51-89; get_sphere_uv 42-49
bool sphere_hit(const rt_object_desc& s, const ray& r, float t_min, float t_max, hit_record& rec) {
  const vec3 cc = madd(r.tm, V3(s.dc), V3(s.c0));  // center.point_at_parameter(r.time())
  const vec3 oc = r.A - cc;
  const float a = dot(r.B, r.B);
  const float b = dot(oc, r.B);
  const float c = fmaf(-s.radius, s.radius, dot(oc, oc));
  const float disc = fmaf(b, b, -(a * c));
  if (disc <= 0.0f) return false;
  const float sq = sqrtf(disc);
  float t = (-b - sq) / a;
  for (int k = 0; k < 2; ++k) {
    if (t > t_min && t < t_max) {
      rec.t = t;
      rec.p = point_at(r, t);
      rec.normal = (rec.p - cc) / s.radius;
      const float theta = acosf(-rec.normal.y);
      const float phi = atan2f(-rec.normal.z, rec.normal.x) + 3.141592654f;
      rec.u = phi / (2 * 3.141592654f);
      rec.v = theta / 3.141592654f;
      rec.mat = s.mat;
      return true;
    }
    t = (-b + sq) / a;
  }
  return false;
}

// quad::hit, quad.cuh:60-90
bool quad_hit(const rt_object_desc& q, const ray& r, float t_min, float t_max, hit_record& rec) {
  const vec3 normal = V3(q.n);
  const float denom = dot(normal, r.B);
  if (fabsf(denom) < 1e-8f) return false;
  const float t = (q.D - dot(normal, r.A)) / denom;
  if (t < t_min || t > t_max) return false;
  const vec3 P = point_at(r, t);
  const vec3 pl = P - V3(q.Q);
  const float alpha = dot(V3(q.w), cross(pl, V3(q.v)));
  const float beta = dot(V3(q.w), cross(V3(q.u), pl));
  if (alpha < 0.f || alpha > 1.f || beta < 0.f || beta > 1.f) return false;
  rec.t = t; rec.p = P; rec.u = alpha; rec.v = beta;
  vec3 n = normal;
  if (dot(n, r.B) > 0.f) n = -n;
  rec.normal = n;
  rec.mat = q.mat;
  return true;
}

// constant_medium::hit(4 args) -> hit(5 args), constant_medium.cuh:67-76, 36-64. The world is always a
// bvh_node, which drops the caller's RNG (bvh.cuh:109-112), so this overload is the only one reached.
bool medium_hit(const Scene& S, const rt_object_desc& m, const ray& r, float tmin, float tmax, hit_record& rec) {
  xorwow fake;
  const uint32_t seed = 1337u ^ fbits(r.A.x) ^ fbits(r.A.y * 3.1f) ^ fbits(r.B.z * 5.7f);
  fake.init(seed);
  hit_record rec1, rec2;
  if (!object_hit(S, m.child, r, -FLT_MAX, FLT_MAX, rec1)) return false;
  if (!object_hit(S, m.child, r, rec1.t + 1e-4f, FLT_MAX, rec2)) return false;
  if (rec1.t < tmin) rec1.t = tmin;
  if (rec2.t > tmax) rec2.t = tmax;
  if (rec1.t >= rec2.t) return false;
  if (rec1.t < 0) rec1.t = 0;
  const float ray_len = length(r.B);
  if (ray_len <= 0.0f || !isfinite(ray_len)) return false;
  const float distance_inside = (rec2.t - rec1.t) * ray_len;
  const float U = fmaxf(1e-6f, fake.uniform());
  const float hit_distance = m.neg_inv_density * logf(U);
  if (hit_distance > distance_inside) return false;
  rec.t = rec1.t + hit_distance / ray_len;
  rec.p = point_at(r, rec.t);
  rec.normal = V(1, 0, 0);
  rec.u = rec.v = 0.0f;
  rec.mat = m.mat;
  return true;
}

bool object_hit(const Scene& S, int o, const ray& r, float tmin, float tmax, hit_record& rec) {
  const rt_object_desc& d = S.obj[o];
  switch (d.kind) {
    case RT_OBJ_SPHERE: return sphere_hit(d, r, tmin, tmax, rec);
    case RT_OBJ_QUAD: return quad_hit(d, r, tmin, tmax, rec);
    case RT_OBJ_BOX: {  // compound6::hit, quad.cuh:124-139
      bool hit_any = false;
      float closest = tmax;
      for (int i = 0; i < 6; ++i) {
        hit_record tmp;
        if (quad_hit(S.obj[d.child + i], r, tmin, closest, tmp)) { hit_any = true; closest = tmp.t; rec = tmp; }
      }
      return hit_any;
    }
    case RT_OBJ_TRANSLATE: {  // hittable.cuh:56-65
      ray moved = {r.A - V3(d.offset), r.B, r.tm};
      if (!object_hit(S, d.child, moved, tmin, tmax, rec)) return false;
      rec.p = rec.p + V3(d.offset);
      return true;
    }
    case RT_OBJ_ROTATE_Y: {  // hittable.cuh:118-145
      const float s = d.sin_t, c = d.cos_t;
      const float ox = fmaf(c, r.A.x, -(s * r.A.z)), oz = fmaf(s, r.A.x, c * r.A.z);
      const float dx = fmaf(c, r.B.x, -(s * r.B.z)), dz = fmaf(s, r.B.x, c * r.B.z);
      ray rot = {V(ox, r.A.y, oz), V(dx, r.B.y, dz), r.tm};
      if (!object_hit(S, d.child, rot, tmin, tmax, rec)) return false;
      const float px = fmaf(c, rec.p.x, s * rec.p.z), pz = fmaf(c, rec.p.z, -(s * rec.p.x));
      const float nx = fmaf(c, rec.normal.x, s * rec.normal.z), nz = fmaf(c, rec.normal.z, -(s * rec.normal.x));
      rec.p = V(px, rec.p.y, pz);
      rec.normal = unit_vector(V(nx, rec.normal.y, nz));
      if (dot(rec.normal, r.B) > 0.f) rec.normal = -rec.normal;
      return true;
    }
    case RT_OBJ_MEDIUM: return medium_hit(S, d, r, tmin, tmax, rec);
    case RT_OBJ_WITH_MATERIAL:  // hittable.cuh:170-174
      if (!object_hit(S, d.child, r, tmin, tmax, rec)) return false;
      rec.mat = d.mat;
      return true;
    case RT_OBJ_BVH: {
      // A bvh_node used as an object (bvh.cuh:95-106 entered from a wrapper's hit()): its own box, then the closest hit
      // over the members, each behind its own box like every leaf of a bvh_node. The nodes between the root and the
      // leaves hold unions of member boxes and the slab test is monotone in the box, so they reject nothing a member's
      // box accepts; a plain walk over the member list returns what the tree returns (exact-t ties aside).
      aabb gb = {V3(d.box_min), V3(d.box_max)};
      if (!aabb_hit(gb, r, tmin, tmax)) return false;
      std::vector<int> members;
      for (int c = o; c >= 0; c = S.obj[c].inward) members.push_back(S.obj[c].child);
      bool any = false;
      float closest = tmax;
      for (size_t i = members.size(); i-- > 0;) {
        const int m = members[i];
        aabb mb = {V3(S.obj[m].box_min), V3(S.obj[m].box_max)};
        if (!aabb_hit(mb, r, tmin, closest)) continue;
        hit_record tmp;
        if (object_hit(S, m, r, tmin, closest, tmp)) { any = true; closest = tmp.t; rec = tmp; }
      }
      return any;
    }
  }
  return false;
}

// bvh_node::hit, bvh.cuh:95-106. A leaf wrapper has left == right == the object, so the object is tested
// twice, the second time with tmax = the first hit's t.
bool bvh_hit(const Scene& S, int node, const ray& r, float tmin, float tmax, hit_record& rec) {
  const bvh_node& n = S.nodes[node];
  if (!aabb_hit(n.box, r, tmin, tmax)) return false;
  hit_record lrec, rrec;
  bool hl, hr;
  if (n.leaf >= 0) {
    const int o = S.top[n.leaf];
    hl = object_hit(S, o, r, tmin, tmax, lrec);
    hr = object_hit(S, o, r, tmin, hl ? lrec.t : tmax, rrec);
    lrec.top = rrec.top = n.leaf;
  } else {
    hl = bvh_hit(S, n.left, r, tmin, tmax, lrec);
    hr = bvh_hit(S, n.right, r, tmin, hl ? lrec.t : tmax, rrec);
  }
  if (hr) rec = rrec;
  if (hl && (!hr || lrec.t < rrec.t)) rec = lrec;
  return hl || hr;
}

bool world_hit(const Scene& S, const ray& r, float tmin, float tmax, hit_record& rec) {
  if (S.nodes.empty()) return false;  // bvh_node over an empty list: left = right = nullptr
  return bvh_hit(S, 0, r, tmin, tmax, rec);
}

// ---- perlin.cuh:6-82 ----
inline uint32_t wanghash(uint32_t x) { x = (x ^ 61u) ^ (x >> 16); x *= 9u; x = x ^ (x >> 4); x *= 0x27d4eb2du; x = x ^ (x >> 15); return x; }
inline uint32_t mix3(int x, int y, int z) { return (uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u; }
inline float u2m11(uint32_t h) { return fmaf((float)((h >> 8) & 0x00FFFFFFu), 1.0f / 8388607.5f, -1.0f); }
inline vec3 perlin_grad(int xi, int yi, int zi) {
  uint32_t h = wanghash(mix3(xi, yi, zi));
  return unit_vector(V(u2m11(h), u2m11(wanghash(h)), u2m11(wanghash(h ^ 0x9e3779b9u))));
}
inline float perlin_smooth(float t) { return (t * t) * (3.0f - (t + t)); }
float perlin_noise(vec3 p) {
  const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
  const float u = p.x - fx, v = p.y - fy, w = p.z - fz;
  const int i = (int)fx, j = (int)fy, k = (int)fz;
  const float uu = perlin_smooth(u), vv = perlin_smooth(v), ww = perlin_smooth(w);
  float accum = 0.0f;
  for (int di = 0; di < 2; ++di)
    for (int dj = 0; dj < 2; ++dj)
      for (int dk = 0; dk < 2; ++dk) {
        const vec3 c = perlin_grad(i + di, j + dj, k + dk);
        const vec3 wt = V(u - (float)di, v - (float)dj, w - (float)dk);
        const float s = ((di ? uu : (1.0f - uu)) * (dj ? vv : (1.0f - vv))) * (dk ? ww : (1.0f - ww));
        accum = fmaf(s, dot(c, wt), accum);
      }
  return accum;
}
float perlin_turb(vec3 p, int depth) {
  float accum = 0.0f, weight = 1.0f;
  vec3 temp = p;
  for (int i = 0; i < depth; ++i) {
    accum = fmaf(weight, perlin_noise(temp), accum);
    weight *= 0.5f;
    temp = 2.0f * temp;
  }
  return fabsf(accum);
}

inline float clamp01(float x) { return x < 0 ? 0 : x > 1 ? 1 : x; }
inline float smoothstep(float e0, float e1, float x) { float t = clamp01((x - e0) / (e1 - e0)); return (t * t) * (3.0f - 2.0f * t); }

// texture::value, texture.cuh:16-164
vec3 texture_value(const Scene& S, int tex, float u, float v, vec3 p) {
  const rt_texture_desc& t = S.tex[tex];
  switch (t.kind) {
    case RT_TEX_SOLID: return V3(t.color);
    case RT_TEX_CHECKER: {
      const int xi = (int)floorf(t.scale * p.x), yi = (int)floorf(t.scale * p.y), zi = (int)floorf(t.scale * p.z);
      return texture_value(S, ((xi + yi + zi) & 1) == 0 ? t.even : t.odd, u, v, p);
    }
    case RT_TEX_IMAGE: {
      const rt_image_desc& im = S.img[t.image];
      const unsigned char* px = t.image < (int)S.img_px.size() ? S.img_px[t.image] : nullptr;
      if (!(px && im.width > 0 && im.height > 0 && im.bpp >= 3)) return V(0, 1, 1);
      u = clamp01(u); v = clamp01(v);
      int i = (int)(u * im.width); if (i > im.width - 1) i = im.width - 1;
      int j = (int)((1.f - v) * im.height); if (j > im.height - 1) j = im.height - 1;
      const int idx = (j * im.width + i) * im.bpp;
      const float inv255 = 1.f / 255.f;
      return V(inv255 * px[idx + 0], inv255 * px[idx + 1], inv255 * px[idx + 2]);
    }
    case RT_TEX_NOISE: {
      const float s = sinf(fmaf(t.scale, p.z, 10.0f * perlin_turb(p, 7)));  // __sinf on the device
      const float q = 0.5f * (1.0f + s);
      return V(q, q, q);
    }
    case RT_TEX_NOODLE: {
      const vec3 d = V(t.p[4], t.p[5], t.p[6]);
      const float uu = dot(p, d);
      const float wig = perlin_turb(t.p[2] * p, (int)t.p[3]);
      const float stripes = fabsf(sinf(fmaf(t.p[0], uu, t.p[1] * wig)));
      const float q = smoothstep(0.75f, 0.98f, stripes);
      const vec3 cN = V(t.p[7], t.p[8], t.p[9]), cG = V(t.p[10], t.p[11], t.p[12]);
      const float omq = 1.f - q;
      return V(fmaf(q, cN.x, omq * cG.x), fmaf(q, cN.y, omq * cG.y), fmaf(q, cN.z, omq * cG.z));
    }
    case RT_TEX_FELT: {
      const float m = perlin_noise(t.p[0] * p);
      // p.x*f_scale + 2*turb: the reference SASS rounds the LEFT product and fuses the right one, FFMA(|turb|, 2, p.x*f_scale)
      const float phase = fmaf(perlin_turb(0.5f * p, 2), 2.0f, p.x * t.p[2]);
      const float fibers = 0.5f * (1.0f + sinf(phase));
      float gain = fmaf(t.p[3], fibers - 0.5f, fmaf(t.p[1], m - 0.5f, 1.0f));
      gain = fminf(fmaxf(gain, 0.7f), 1.2f);
      return gain * V3(t.color);
    }
    case RT_TEX_UV_OFFSET: {
      float uu = u + t.p[0]; uu -= floorf(uu);
      float vv = v + t.p[1]; vv = fminf(fmaxf(vv, 0.f), 1.f);
      return texture_value(S, t.even, uu, vv, p);
    }
  }
  return V(0, 0, 0);
}

// material.cuh:12-18
vec3 random_in_unit_sphere(xorwow& g) {
  while (true) {
    const float a = g.uniform(), b = g.uniform(), c = g.uniform();  // evaluated left to right by the reference build
    const vec3 p = V(fmaf(2.0f, a, -1.0f), fmaf(2.0f, b, -1.0f), fmaf(2.0f, c, -1.0f));
    if (dot(p, p) < 1.0f) return p;
  }
}

// material::emitted, material.cuh:168-172
vec3 emitted(const Scene& S, const rt_material_desc& m, const hit_record& rec) {
  if (m.kind != RT_MAT_DIFFUSE_LIGHT) return V(0, 0, 0);
  return m.tex >= 0 ? texture_value(S, m.tex, rec.u, rec.v, rec.p) : V3(m.albedo);
}

// material::scatter, material.cuh:75-86 (lambertian), 99-109 (metal), 119-159 (dielectric), 174-178 (light), 193-199 (isotropic)
bool scatter(const Scene& S, const rt_material_desc& m, const ray& r_in, const hit_record& rec, vec3& attenuation,
             ray& scattered, xorwow& g) {
  scattered.tm = r_in.tm;
  scattered.A = rec.p;
  switch (m.kind) {
    case RT_MAT_LAMBERTIAN: {
      const vec3 target = (rec.p + rec.normal) + random_in_unit_sphere(g);
      scattered.B = target - rec.p;
      attenuation = m.tex >= 0 ? texture_value(S, m.tex, rec.u, rec.v, rec.p) : V(1, 1, 1);
      return true;
    }
    case RT_MAT_METAL: {
      const vec3 ud = unit_vector(r_in.B);
      float two = dot(ud, rec.normal); two = two + two;
      const vec3 reflected = madd(-two, rec.normal, ud);
      scattered.B = madd(m.param, random_in_unit_sphere(g), reflected);
      attenuation = V3(m.albedo);
      return dot(scattered.B, rec.normal) > 0.0f;
    }
    case RT_MAT_DIELECTRIC: {
      const float ref_idx = m.param;
      const float dn = dot(r_in.B, rec.normal);
      const vec3 reflected = madd(-(dn + dn), rec.normal, r_in.B);
      attenuation = V(1.0f, 1.0f, 1.0f);
      vec3 outward; float ni_over_nt, cosine;
      if (dn > 0.0f) {
        outward = -rec.normal;
        ni_over_nt = ref_idx;
        cosine = dn / length(r_in.B);
        cosine = sqrtf(fmaxf(0.0f, fmaf(-(ref_idx * ref_idx), fmaf(-cosine, cosine, 1.0f), 1.0f)));
      } else {
        outward = rec.normal;
        ni_over_nt = 1.0f / ref_idx;
        cosine = -dn / length(r_in.B);
      }
      const vec3 uv = unit_vector(r_in.B);  // refract(), material.cuh:26-36
      const float dt = dot(uv, outward);
      const float disc = fmaf(-(ni_over_nt * ni_over_nt), fmaf(-dt, dt, 1.0f), 1.0f);
      vec3 refracted = V(0, 0, 0);
      float reflect_prob;
      if (disc > 0.0f) {
        refracted = madd(-sqrtf(disc), outward, ni_over_nt * madd(-dt, outward, uv));
        float r0 = (1.0f - ref_idx) / (1.0f + ref_idx);  // schlick(), material.cuh:38-43
        r0 = r0 * r0;
        reflect_prob = fmaf(1.0f - r0, powf(1.0f - cosine, 5.0f), r0);
      } else {
        reflect_prob = 1.0f;
      }
      scattered.B = (g.uniform() < reflect_prob) ? reflected : refracted;
      return true;
    }
    case RT_MAT_ISOTROPIC:
      scattered.B = random_in_unit_sphere(g);
      attenuation = texture_value(S, m.tex, rec.u, rec.v, rec.p);
      return true;
    default:
      return false;
  }
}

// camera::get_ray, camera.cuh:35-47 (random_in_unit_disk 8-16)
ray get_ray(const rt_camera_desc& c, float s, float t, xorwow& g) {
  float px, py;
  do {
    const float a = g.uniform(), b = g.uniform();
    px = fmaf(2.0f, a, -1.0f); py = fmaf(2.0f, b, -1.0f);
  } while (fmaf(px, px, py * py) >= 1.0f);
  const float rdx = c.lens_radius * px, rdy = c.lens_radius * py;
  const vec3 cu = V3(c.u), cv = V3(c.v);
  const vec3 offset = V(fmaf(rdx, cu.x, rdy * cv.x), fmaf(rdx, cu.y, rdy * cv.y), fmaf(rdx, cu.z, rdy * cv.z));
  const double tm = fma((double)g.uniform(), c.time1 - c.time0, c.time0);
  ray r;
  r.A = V3(c.origin) + offset;
  const vec3 d = madd(t, V3(c.vertical), madd(s, V3(c.horizontal), V3(c.lower_left_corner)));
  r.B = (d - V3(c.origin)) - offset;
  r.tm = (float)tm;
  return r;
}

// color(), main.cu:44-87
vec3 color(const Scene& S, const ray& r0, vec3 background, bool gradient_bg, int max_depth, xorwow& g, unsigned long long& rays) {
  ray cur = r0;
  vec3 throughput = V(1, 1, 1), radiance = V(0, 0, 0);
  for (int bounce = 0; bounce < max_depth; ++bounce) {
    hit_record rec;
    ++rays;
    if (!world_hit(S, cur, 0.001f, FLT_MAX, rec)) {
      vec3 bg = background;
      if (gradient_bg) {
        const float uy = cur.B.y / length(cur.B);
        const float t = 0.5f * (uy + 1.0f);
        const float omt = 1.0f - t;
        bg = V(fmaf(t, 0.5f, omt), fmaf(t, 0.7f, omt), t + omt);
      }
      radiance = V(fmaf(throughput.x, bg.x, radiance.x), fmaf(throughput.y, bg.y, radiance.y), fmaf(throughput.z, bg.z, radiance.z));
      break;
    }
    const rt_material_desc& m = S.mat[rec.mat];
    const vec3 e = emitted(S, m, rec);
    radiance = V(fmaf(throughput.x, e.x, radiance.x), fmaf(throughput.y, e.y, radiance.y), fmaf(throughput.z, e.z, radiance.z));
    ray scattered; vec3 attenuation = V(0, 0, 0);
    if (!scatter(S, m, cur, rec, attenuation, scattered, g)) break;
    throughput = throughput * attenuation;
    cur = scattered;
  }
  return radiance;
}

inline float apply_gamma(float c, float gamma) {  // main.cu:37-42
  if (gamma == 1.0f) return c;
  return powf(fmaxf(c, 0.0f), 1.0f / gamma);
}

}  // namespace

extern "C" {

// images: n_img pointers to decoded 8-bit pixels (may be NULL when no texture is an image)
void* oracle_load2(const void* sd, size_t len, const unsigned char* const* images, int n_images, int with_bvh);
void* oracle_load(const void* sd, size_t len, const unsigned char* const* images, int n_images) {
  return oracle_load2(sd, len, images, n_images, 1);
}
// with_bvh = 0: no reference BVH (its constructor is O(n^2) per level, bvh.cuh:66-77 - hours at 10^6 objects); only
// oracle_primary_ids_brute may be called on such a scene.
void* oracle_load2(const void* sd, size_t len, const unsigned char* const* images, int n_images, int with_bvh) {
  if (!sd || len < sizeof(rt_sd_header)) return nullptr;
  Scene* S = new Scene();
  const unsigned char* p = (const unsigned char*)sd;
  memcpy(&S->h, p, sizeof(rt_sd_header));
  if (S->h.magic != RT_SD_MAGIC) { delete S; return nullptr; }
  size_t off = sizeof(rt_sd_header);
  const size_t need = off + S->h.n_tex * sizeof(rt_texture_desc) + S->h.n_mat * sizeof(rt_material_desc) +
                      S->h.n_obj * sizeof(rt_object_desc) + S->h.n_top * sizeof(int) + S->h.n_img * sizeof(rt_image_desc);
  if (len < need) { delete S; return nullptr; }
  S->tex.resize(S->h.n_tex); memcpy(S->tex.data(), p + off, S->h.n_tex * sizeof(rt_texture_desc)); off += S->h.n_tex * sizeof(rt_texture_desc);
  S->mat.resize(S->h.n_mat); memcpy(S->mat.data(), p + off, S->h.n_mat * sizeof(rt_material_desc)); off += S->h.n_mat * sizeof(rt_material_desc);
  S->obj.resize(S->h.n_obj); memcpy(S->obj.data(), p + off, S->h.n_obj * sizeof(rt_object_desc)); off += S->h.n_obj * sizeof(rt_object_desc);
  S->top.resize(S->h.n_top); memcpy(S->top.data(), p + off, S->h.n_top * sizeof(int)); off += S->h.n_top * sizeof(int);
  S->img.resize(S->h.n_img); memcpy(S->img.data(), p + off, S->h.n_img * sizeof(rt_image_desc));
  for (int i = 0; i < S->h.n_img; ++i) S->img_px.push_back(images && i < n_images ? images[i] : nullptr);
  const int n = (int)S->top.size();
  S->order.resize(n);
  for (int i = 0; i < n; ++i) S->order[i] = i;
  if (n > 0 && with_bvh) { S->nodes.reserve(2 * n); build_bvh(*S, S->order, 0, n); }
  return S;
}
void oracle_free(void* h) { delete (Scene*)h; }
int oracle_n_top(void* h) { return (int)((Scene*)h)->top.size(); }
int oracle_n_nodes(void* h) { return (int)((Scene*)h)->nodes.size(); }

// rank[k] = position of d_list entry k in the reference BVH's leaf order
void oracle_leaf_order(void* h, int* rank) {
  const Scene& S = *(Scene*)h;
  for (int pos = 0; pos < (int)S.order.size(); ++pos) rank[S.order[pos]] = pos;
}

// Centre ray of every pixel (no jitter, no lens offset, time0): closest hit through the reference BVH.
// obj = index into top[] (creation order), -1 = miss; mat = material id of the SD; t = 0 on a miss.
void oracle_primary_ids(void* h, int nx, int ny, int* obj, int* mat, float* t) {
  const Scene& S = *(Scene*)h;
  const rt_camera_desc& c = S.h.cam;
#pragma omp parallel for schedule(dynamic, 4)
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const float s = ((float)i + 0.5f) / (float)nx, tt = ((float)j + 0.5f) / (float)ny;
      ray r;
      r.A = V3(c.origin);
      r.B = madd(tt, V3(c.vertical), madd(s, V3(c.horizontal), V3(c.lower_left_corner))) - V3(c.origin);
      r.tm = (float)c.time0;
      hit_record rec;
      const bool hit = world_hit(S, r, 0.001f, FLT_MAX, rec);
      const int pix = j * nx + i;
      obj[pix] = hit ? rec.top : -1;
      mat[pix] = hit ? rec.mat : -1;
      t[pix] = hit ? rec.t : 0.f;
    }
}

// The same query WITHOUT any hierarchy: every top-level object in creation order, each behind its own box test
// (aabb::hit on the object's bounding box = the leaf node's box, bvh.cuh:38-43) and hit twice like a leaf node
// (left == right, bvh.cuh:100-105), t_max shrinking as hits are found. Equals oracle_primary_ids except for exact-t ties
// between different objects (decided by visiting order). For the C5 scale-up scenes, whose reference BVH cannot be built.
void oracle_primary_ids_brute(void* h, int nx, int ny, int* obj, int* mat, float* t) {
  const Scene& S = *(Scene*)h;
  const rt_camera_desc& c = S.h.cam;
  const int n = (int)S.top.size();
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const float s = ((float)i + 0.5f) / (float)nx, tt = ((float)j + 0.5f) / (float)ny;
      ray r;
      r.A = V3(c.origin);
      r.B = madd(tt, V3(c.vertical), madd(s, V3(c.horizontal), V3(c.lower_left_corner))) - V3(c.origin);
      r.tm = (float)c.time0;
      float closest = FLT_MAX;
      hit_record best;
      int best_top = -1;
      for (int k = 0; k < n; ++k) {
        const int o = S.top[k];
        if (!aabb_hit(obj_box(S, o), r, 0.001f, closest)) continue;
        hit_record lrec, rrec;
        const bool hl = object_hit(S, o, r, 0.001f, closest, lrec);
        const bool hr = object_hit(S, o, r, 0.001f, hl ? lrec.t : closest, rrec);
        if (!hl && !hr) continue;
        hit_record rec;
        if (hr) rec = rrec;
        if (hl && (!hr || lrec.t < rrec.t)) rec = lrec;
        best = rec; best_top = k; closest = rec.t;
      }
      const int pix = j * nx + i;
      obj[pix] = best_top;
      mat[pix] = best_top >= 0 ? best.mat : -1;
      t[pix] = best_top >= 0 ? best.t : 0.f;
    }
}

// render_init + render (main.cu:96-133) for every pixel, reference-RNG mode: seed 1984 + pixel_index.
// fb: ny*nx*3 floats, gamma applied, row 0 = bottom scanline. Returns the number of closest-hit queries.
unsigned long long oracle_render(void* h, int nx, int ny, int ns, int max_depth, const float* background, int gradient,
                                 float gamma, float* fb) {
  const Scene& S = *(Scene*)h;
  unsigned long long rays = 0;
  const vec3 bg = V(background[0], background[1], background[2]);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rays)
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const int pix = j * nx + i;
      xorwow g;
      g.init((unsigned long long)(long long)(1984 + pix));
      vec3 col = V(0, 0, 0);
      for (int s = 0; s < ns; ++s) {
        const float u = ((float)i + g.uniform()) / (float)nx;
        const float v = ((float)j + g.uniform()) / (float)ny;
        const ray r = get_ray(S.h.cam, u, v, g);
        col = col + color(S, r, bg, gradient != 0, max_depth, g, rays);
      }
      const float k = (float)(1.0 / (double)(float)ns);  // vec3::operator/=(float): `float k = 1.0/t`, vec3.cuh:145-153
      fb[3 * pix + 0] = apply_gamma(col.x * k, gamma);
      fb[3 * pix + 1] = apply_gamma(col.y * k, gamma);
      fb[3 * pix + 2] = apply_gamma(col.z * k, gamma);
    }
  return rays;
}

void oracle_xorwow(unsigned long long seed, int n, unsigned int* raw, float* uni) {
  xorwow a, b;
  a.init(seed); b.init(seed);
  for (int i = 0; i < n; ++i) { raw[i] = a.next(); uni[i] = b.uniform(); }
}

}  // extern "C"
