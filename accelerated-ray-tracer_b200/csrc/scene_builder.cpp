// scene_builder.cpp — host restatement of the reference's constructors (see scene_builder.h).
//
// Two arithmetic regimes, both taken from the reference's sm_100 SASS:
//   * "runtime" values (anything that depends on nx/ny, the scene RNG or a loop counter) are computed
//     by the constructors on the GPU with nvcc's default contraction -> fused forms of rt_math.h;
//   * values that depend only on literals (the cameras' look-from/look-at/up, literal quads and
//     boxes) are constant-folded by NVVM BEFORE contraction -> every operation rounded separately.
//     Example: create_world_bouncing's camera v.y is 0.98894989490509 (folded) and not
//     0.98894983530045 (fused). The `folded` flag selects this regime.
#include "scene_builder.h"
#if defined(__linux__)
#include <sys/mman.h>
#endif
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>

namespace rt {

void advise_huge(void* p, size_t bytes) {
#if defined(__linux__)
  const uintptr_t a = ((uintptr_t)p + 4095) & ~(uintptr_t)4095, e = ((uintptr_t)p + bytes) & ~(uintptr_t)4095;
  if (p && bytes >= ((size_t)8 << 20) && e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
#else
  (void)p; (void)bytes;
#endif
}

static void put3(float* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
static V3 get3(const float* d) { return v3(d[0], d[1], d[2]); }

// ---- separately-rounded (constant-folded) forms ----
static float dot_unfused(V3 a, V3 b) { return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z)); }
static V3 cross_unfused(V3 a, V3 b) {
  return v3(fsub(fmul(a.y, b.z), fmul(a.z, b.y)), -fsub(fmul(a.x, b.z), fmul(a.z, b.x)),
            fsub(fmul(a.x, b.y), fmul(a.y, b.x)));
}
static float len_unfused(V3 a) { return fsqrt(dot_unfused(a, a)); }
static V3 unit_unfused(V3 a) { return vdivs(a, len_unfused(a)); }

static rt_texture_desc tex_clear() {
  rt_texture_desc d;
  memset(&d, 0, sizeof(d));
  d.even = d.odd = d.image = -1;
  return d;
}
static rt_object_desc obj_clear() {
  rt_object_desc d;
  memset(&d, 0, sizeof(d));
  d.kind = -1; d.mat = -1; d.child = -1;
  return d;
}

// ---- textures (texture.cuh) ----
int SceneBuilder::solid_color(V3 a) {
  rt_texture_desc d = tex_clear();
  d.kind = RT_TEX_SOLID; put3(d.color, a);
  S.tex.push_back(d);
  return (int)S.tex.size() - 1;
}
int SceneBuilder::checker_texture(float scale, int even, int odd) {
  rt_texture_desc d = tex_clear();
  d.kind = RT_TEX_CHECKER; d.scale = fdiv(1.f, scale);  // inv_scale(1.f/scale), texture.cuh:33
  d.even = even; d.odd = odd;
  S.tex.push_back(d);
  return (int)S.tex.size() - 1;
}
int SceneBuilder::image_texture(int image) {
  rt_texture_desc d = tex_clear();
  d.kind = RT_TEX_IMAGE; d.image = image;
  S.tex.push_back(d);
  return (int)S.tex.size() - 1;
}
int SceneBuilder::noise_texture(float scale) {
  rt_texture_desc d = tex_clear();
  d.kind = RT_TEX_NOISE; d.scale = scale;
  S.tex.push_back(d);
  return (int)S.tex.size() - 1;
}
int SceneBuilder::noodle_texture(float k, float A, float f, int oct, V3 dir, V3 noodle, V3 gap) {
  rt_texture_desc d = tex_clear();
  d.kind = RT_TEX_NOODLE;
  d.p[0] = k; d.p[1] = A; d.p[2] = f; d.p[3] = (float)oct;
  put3(d.p + 4, unit_unfused(dir));  // d(unit_vector(dir)), texture.cuh:92 (literal default -> folded)
  put3(d.p + 7, noodle); put3(d.p + 10, gap);
  S.tex.push_back(d);
  return (int)S.tex.size() - 1;
}
int SceneBuilder::felt_texture(V3 base, float m_scale, float m_amt, float f_scale, float f_amt) {
  rt_texture_desc d = tex_clear();
  d.kind = RT_TEX_FELT; put3(d.color, base);
  d.p[0] = m_scale; d.p[1] = m_amt; d.p[2] = f_scale; d.p[3] = f_amt;
  S.tex.push_back(d);
  return (int)S.tex.size() - 1;
}
int SceneBuilder::uv_offset_texture(int base, float du, float dv) {
  rt_texture_desc d = tex_clear();
  d.kind = RT_TEX_UV_OFFSET; d.even = base; d.p[0] = du; d.p[1] = dv;
  S.tex.push_back(d);
  return (int)S.tex.size() - 1;
}
int SceneBuilder::add_image(const HostImage& im) {
  rt_image_desc d; d.width = im.width; d.height = im.height; d.bpp = im.bpp; d.pad_ = 0;
  S.img.push_back(d);
  S.img_data.push_back(im);
  return (int)S.img.size() - 1;
}

// ---- materials (material.cuh) ----
static rt_material_desc mat_clear() {
  rt_material_desc d;
  memset(&d, 0, sizeof(d));
  d.tex = -1;
  return d;
}
int SceneBuilder::lambertian_tex(int tex) {
  rt_material_desc d = mat_clear();
  d.kind = RT_MAT_LAMBERTIAN; d.tex = tex;
  S.mat.push_back(d);
  return (int)S.mat.size() - 1;
}
int SceneBuilder::metal(V3 a, float f) {
  rt_material_desc d = mat_clear();
  d.kind = RT_MAT_METAL; put3(d.albedo, a);
  d.param = f < 1.0f ? f : 1.0f;  // material.cuh:97
  S.mat.push_back(d);
  return (int)S.mat.size() - 1;
}
int SceneBuilder::dielectric(float ri) {
  rt_material_desc d = mat_clear();
  d.kind = RT_MAT_DIELECTRIC; d.param = ri;
  S.mat.push_back(d);
  return (int)S.mat.size() - 1;
}
int SceneBuilder::diffuse_light(V3 c) {
  rt_material_desc d = mat_clear();
  d.kind = RT_MAT_DIFFUSE_LIGHT; d.tex = -1; put3(d.albedo, c);
  S.mat.push_back(d);
  return (int)S.mat.size() - 1;
}
int SceneBuilder::diffuse_light_tex(int tex) {
  rt_material_desc d = mat_clear();
  d.kind = RT_MAT_DIFFUSE_LIGHT; d.tex = tex;
  S.mat.push_back(d);
  return (int)S.mat.size() - 1;
}
int SceneBuilder::isotropic_tex(int tex) {
  rt_material_desc d = mat_clear();
  d.kind = RT_MAT_ISOTROPIC; d.tex = tex;
  S.mat.push_back(d);
  return (int)S.mat.size() - 1;
}

// ---- hittables ----
static void set_box(rt_object_desc& o, V3 a, V3 b) {  // aabb(a, b), aabb.cuh:17-21
  put3(o.box_min, v3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)));
  put3(o.box_max, v3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)));
}
static void union_box(rt_object_desc& o, V3 mn0, V3 mx0, V3 mn1, V3 mx1) {  // surrounding_box, aabb.cuh:34-43
  V3 sm = v3(fminf(mn0.x, mn1.x), fminf(mn0.y, mn1.y), fminf(mn0.z, mn1.z));
  V3 bg = v3(fmaxf(mx0.x, mx1.x), fmaxf(mx0.y, mx1.y), fmaxf(mx0.z, mx1.z));
  set_box(o, sm, bg);
}

int SceneBuilder::sphere(V3 cen, float r, int mat) {  // sphere.cuh:21-26
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_SPHERE; o.mat = mat; put3(o.c0, cen); put3(o.dc, v3(0, 0, 0)); o.radius = r;
  V3 rv = v3(r, r, r);
  set_box(o, vsub(cen, rv), vadd(cen, rv));
  return push_obj(o);
}
int SceneBuilder::sphere(V3 cen1, V3 cen2, float r, int mat) {  // sphere.cuh:29-38
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_SPHERE; o.mat = mat; o.radius = r;
  V3 B = vsub(cen2, cen1);
  put3(o.c0, cen1); put3(o.dc, B);
  V3 rv = v3(r, r, r);
  V3 c0 = vmad(0.0f, B, cen1);  // center.point_at_parameter(0.0)
  V3 c1 = vmad(1.0f, B, cen1);  // center.point_at_parameter(1.0)
  rt_object_desc b0 = obj_clear(), b1 = obj_clear();
  set_box(b0, vsub(c0, rv), vadd(c0, rv));
  set_box(b1, vsub(c1, rv), vadd(c1, rv));
  union_box(o, get3(b0.box_min), get3(b0.box_max), get3(b1.box_min), get3(b1.box_max));
  return push_obj(o);
}
int SceneBuilder::quad(V3 Q, V3 u, V3 v, int mat, bool inward) {  // quad.cuh:29-41, 49-54
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_QUAD; o.mat = mat; o.inward = inward ? 1 : 0;
  V3 n = folded ? cross_unfused(u, v) : vcross(u, v);
  V3 normal = folded ? unit_unfused(n) : vunit(n);
  if (inward) normal = vneg(normal);
  float D = folded ? dot_unfused(normal, Q) : vdot(normal, Q);
  V3 w = vdivs(n, folded ? dot_unfused(n, n) : vdot(n, n));
  put3(o.Q, Q); put3(o.u, u); put3(o.v, v); put3(o.w, w); put3(o.n, normal); o.D = D;
  // set_bounding_box: d1(Q, Q+u+v), d2(Q+u, Q+v), union, pad(1e-3f)
  rt_object_desc d1 = obj_clear(), d2 = obj_clear(), un = obj_clear();
  set_box(d1, Q, vadd(vadd(Q, u), v));
  set_box(d2, vadd(Q, u), vadd(Q, v));
  union_box(un, get3(d1.box_min), get3(d1.box_max), get3(d2.box_min), get3(d2.box_max));
  V3 d = v3(1e-3f, 1e-3f, 1e-3f);
  set_box(o, vsub(get3(un.box_min), d), vadd(get3(un.box_max), d));
  return push_obj(o);
}
int SceneBuilder::make_box(V3 a, V3 b, int mat) {  // quad.cuh:145-162
  V3 minp = v3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z));
  V3 maxp = v3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z));
  V3 dx = v3(fsub(maxp.x, minp.x), 0.f, 0.f);
  V3 dy = v3(0.f, fsub(maxp.y, minp.y), 0.f);
  V3 dz = v3(0.f, 0.f, fsub(maxp.z, minp.z));
  int f0 = quad(v3(minp.x, minp.y, maxp.z), dx, dy, mat);        // front  +Z
  quad(v3(maxp.x, minp.y, maxp.z), vneg(dz), dy, mat);           // right  +X
  quad(v3(maxp.x, minp.y, minp.z), vneg(dx), dy, mat);           // back   -Z
  quad(v3(minp.x, minp.y, minp.z), dz, dy, mat);                 // left   -X
  quad(v3(minp.x, maxp.y, maxp.z), dx, vneg(dz), mat);           // top    +Y
  quad(v3(minp.x, minp.y, minp.z), dx, dz, mat);                 // bottom -Y
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_BOX; o.mat = mat; o.child = f0;
  // compound6 ctor: union of the six face boxes (quad.cuh:108-121)
  V3 mn = get3(S.obj[f0].box_min), mx = get3(S.obj[f0].box_max);
  for (int i = 1; i < 6; ++i) {
    V3 bmn = get3(S.obj[f0 + i].box_min), bmx = get3(S.obj[f0 + i].box_max);
    mn = v3(fminf(mn.x, bmn.x), fminf(mn.y, bmn.y), fminf(mn.z, bmn.z));
    mx = v3(fmaxf(mx.x, bmx.x), fmaxf(mx.y, bmx.y), fmaxf(mx.z, bmx.z));
  }
  set_box(o, mn, mx);
  return push_obj(o);
}
int SceneBuilder::translate(int obj, V3 d) {  // hittable.cuh:52-54, aabb.cuh:76-79
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_TRANSLATE; o.child = obj; put3(o.offset, d);
  set_box(o, vadd(get3(S.obj[obj].box_min), d), vadd(get3(S.obj[obj].box_max), d));
  return push_obj(o);
}
int SceneBuilder::rotate_y(int obj, float angle_degrees) {  // hittable.cuh:89-116
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_ROTATE_Y; o.child = obj;
  const float rad = fmul(angle_degrees, 0.017453292519943295769f);
  o.sin_t = M.sinf_(rad);
  o.cos_t = M.cosf_(rad);
  V3 bmn = get3(S.obj[obj].box_min), bmx = get3(S.obj[obj].box_max);
  V3 minp = v3(FLT_MAX, FLT_MAX, FLT_MAX), maxp = v3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j)
      for (int k = 0; k < 2; ++k) {
        float x = i ? bmx.x : bmn.x, y = j ? bmx.y : bmn.y, z = k ? bmx.z : bmn.z;
        float nx = ffma(o.cos_t, x, fmul(o.sin_t, z));    // cos*x + sin*z
        float nz = ffma(o.cos_t, z, -fmul(o.sin_t, x));   // -sin*x + cos*z
        minp = v3(fminf(minp.x, nx), fminf(minp.y, y), fminf(minp.z, nz));
        maxp = v3(fmaxf(maxp.x, nx), fmaxf(maxp.y, y), fmaxf(maxp.z, nz));
      }
  set_box(o, minp, maxp);
  return push_obj(o);
}
int SceneBuilder::with_material(int obj, int mat) {  // hittable.cuh:166-168: box = obj->bounding_box()
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_WITH_MATERIAL; o.child = obj; o.mat = mat;
  memcpy(o.box_min, S.obj[obj].box_min, 12);
  memcpy(o.box_max, S.obj[obj].box_max, 12);
  return push_obj(o);
}
int SceneBuilder::bvh_node(const std::vector<int>& members) {  // bvh.cuh:29-84: box = union of the members' boxes
  int head = -1;
  for (int m : members) {
    rt_object_desc o = obj_clear();
    o.kind = RT_OBJ_BVH; o.child = m; o.inward = head;
    V3 mn = get3(S.obj[m].box_min), mx = get3(S.obj[m].box_max);
    if (head >= 0) {
      const V3 bmn = get3(S.obj[head].box_min), bmx = get3(S.obj[head].box_max);
      mn = v3(fminf(mn.x, bmn.x), fminf(mn.y, bmn.y), fminf(mn.z, bmn.z));
      mx = v3(fmaxf(mx.x, bmx.x), fmaxf(mx.y, bmx.y), fmaxf(mx.z, bmx.z));
    }
    set_box(o, mn, mx);
    head = push_obj(o);
  }
  return head;
}
int SceneBuilder::constant_medium_tex(int boundary, float density, int tex) {  // constant_medium.cuh:24-25
  rt_object_desc o = obj_clear();
  o.kind = RT_OBJ_MEDIUM; o.child = boundary;
  o.neg_inv_density = fdiv(-1.0f, density);
  o.mat = isotropic_tex(tex);
  memcpy(o.box_min, S.obj[boundary].box_min, 12);
  memcpy(o.box_max, S.obj[boundary].box_max, 12);
  return push_obj(o);
}
int SceneBuilder::constant_medium(int boundary, float density, V3 albedo) {  // constant_medium.cuh:27-28
  return constant_medium_tex(boundary, density, solid_color(albedo));
}

// camera::init (camera.cuh:59-78). lookfrom/lookat/vup/vfov/aperture/focus_dist are literals in every
// generator -> the basis is constant-folded (separately rounded); `aspect` comes from nx/ny at run
// time -> everything multiplied by half_width/half_height is fused as in the SASS of main.cu:242.
void SceneBuilder::camera(V3 lookfrom, V3 lookat, V3 vup, float vfov, float aspect, float aperture,
                          float focus_dist, double t0, double t1) {
  rt_camera_desc& c = S.cam;
  memset(&c, 0, sizeof(c));
  c.time0 = t0; c.time1 = t1;
  c.lens_radius = fmul(aperture, 0.5f);
  float theta = fdiv(fmul(vfov, 3.141592654f), 180.0f);
  float half_height = M.tanf_(fmul(theta, 0.5f));
  float half_width = fmul(aspect, half_height);
  V3 w = unit_unfused(vsub(lookfrom, lookat));
  V3 u = unit_unfused(cross_unfused(vup, w));
  V3 v = cross_unfused(w, u);
  float hwf = fmul(half_width, focus_dist), hhf = fmul(half_height, focus_dist);
  // lower_left_corner = origin - hwf*u - hhf*v - focus_dist*w ; the last product is a folded constant
  V3 fw = vscale(focus_dist, w);
  V3 llc = vnmad(hwf, u, lookfrom);
  llc = vnmad(hhf, v, llc);
  llc = vsub(llc, fw);
  V3 horizontal = vscale(fmul(fmul(2.0f, half_width), focus_dist), u);
  V3 vertical = vscale(fmul(fmul(2.0f, half_height), focus_dist), v);
  put3(c.origin, lookfrom); put3(c.lower_left_corner, llc);
  put3(c.horizontal, horizontal); put3(c.vertical, vertical);
  put3(c.u, u); put3(c.v, v); put3(c.w, w);
}

// ---- reference leaf order: replay of bvh_node's constructor (bvh.cuh:29-84) on indices ----
// The reference sorts every range with a selection sort (bvh.cuh:46-81): position i receives the FIRST smallest key of
// [i, end) (strict <) and the element that was at i moves to where that one came from. The permutation it leaves
// matters only among equal keys (a grid of boxes has many), and it is not the stable one. Long ranges replay exactly
// these swaps through a tournament tree over the positions (leftmost minimum, two leaf updates per step):
// O(n log n) instead of O(n^2) per range - 10 004 spheres: 0.75 s -> a few ms of rt_build_scene.
static void selection_sort_replay(std::vector<int>& objs, const SceneDesc& sd, int start, int end, int axis) {
  const int n = end - start;
  if (n < 48) {  // the reference's loop as it stands
    for (int i = start; i < end - 1; ++i) {
      int best = i;
      for (int j = i + 1; j < end; ++j)
        if (sd.obj[objs[j]].box_min[axis] < sd.obj[objs[best]].box_min[axis]) best = j;
      if (best != i) std::swap(objs[i], objs[best]);
    }
    return;
  }
  std::vector<float> key((size_t)n);
  for (int p = 0; p < n; ++p) key[p] = sd.obj[objs[start + p]].box_min[axis];
  int sz = 1;
  while (sz < n) sz <<= 1;
  std::vector<int> tr((size_t)2 * sz, -1);  // per node: position of the leftmost minimum of its span, -1: nothing left
  auto better = [&](int a, int b) {         // a lies left of b: b wins only with a strictly smaller key
    if (a < 0) return b;
    if (b < 0) return a;
    return key[b] < key[a] ? b : a;
  };
  for (int p = 0; p < n; ++p) tr[sz + p] = p;
  for (int k = sz - 1; k >= 1; --k) tr[k] = better(tr[2 * k], tr[2 * k + 1]);
  auto fix = [&](int p) { for (int k = (sz + p) >> 1; k >= 1; k >>= 1) tr[k] = better(tr[2 * k], tr[2 * k + 1]); };
  for (int i = 0; i < n - 1; ++i) {
    const int best = tr[1];  // positions < i have been removed: the leftmost minimum of [i, n)
    if (best != i) {
      std::swap(objs[start + i], objs[start + best]);
      std::swap(key[i], key[best]);
      fix(best);
    }
    tr[sz + i] = -1;
    fix(i);
  }
}

static void bvh_order_rec(std::vector<int>& objs, const SceneDesc& sd, int start, int end) {
  const int n = end - start;
  if (n <= 1) return;
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (int i = start; i < end; ++i) {
    const float* b = sd.obj[objs[i]].box_min;
    for (int a = 0; a < 3; ++a) { if (b[a] < mn[a]) mn[a] = b[a]; if (b[a] > mx[a]) mx[a] = b[a]; }
  }
  const float sx = mx[0] - mn[0], sy = mx[1] - mn[1], sz = mx[2] - mn[2];
  int axis = 0;
  if (sy > sx && sy >= sz) axis = 1;
  else if (sz > sx && sz >= sy) axis = 2;
  selection_sort_replay(objs, sd, start, end, axis);
  const int mid = start + (n >> 1);
  bvh_order_rec(objs, sd, start, mid);
  bvh_order_rec(objs, sd, mid, end);
}

std::vector<int> reference_leaf_order(const SceneDesc& sd, int limit) {
  const int n = (int)sd.top.size();
  std::vector<int> rank(n);
  std::vector<int> objs(sd.top);
  if (n <= limit) bvh_order_rec(objs, sd, 0, n);
  // objs[pos] = object id at leaf position pos; map back to the index in `top`
  std::vector<int> where(sd.obj.size(), -1);
  for (int k = 0; k < n; ++k) where[sd.top[k]] = k;
  for (int pos = 0; pos < n; ++pos) rank[where[objs[pos]]] = pos;
  return rank;
}

bool load_ppm(const std::string& path, HostImage& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  int w = 0, h = 0, mx = 0;
  if (fscanf(f, "P6 %d %d %d", &w, &h, &mx) != 3 || w <= 0 || h <= 0 || mx != 255) { fclose(f); return false; }
  fgetc(f);
  out.width = w; out.height = h; out.bpp = 3;
  out.px.resize((size_t)w * h * 3);
  size_t got = fread(out.px.data(), 1, out.px.size(), f);
  fclose(f);
  return got == out.px.size();
}

std::string load_texture_file(const std::string& path, HostImage& out) {
  const size_t dot = path.rfind('.');
  std::string ext = dot == std::string::npos ? "" : path.substr(dot + 1);
  for (auto& c : ext) c = (char)tolower((unsigned char)c);
  if (ext == "jpg" || ext == "jpeg") {
    std::string err;
    return load_jpeg(path, out, err) ? "" : err;
  }
  if (ext == "ppm") return load_ppm(path, out) ? "" : "cannot read " + path + " (binary PPM, P6, maxval 255)";
  return path + ": unknown texture format (expected .jpg or .ppm)";
}

// ---- groups (RT_OBJ_BVH) -> instanced members ----
namespace {
bool is_wrapper(int kind) { return kind == RT_OBJ_TRANSLATE || kind == RT_OBJ_ROTATE_Y || kind == RT_OBJ_WITH_MATERIAL; }
// box of wrapper `w` around a child with box (mn, mx): the wrappers' constructors (aabb.cuh:76-79, hittable.cuh:89-116, 166-168)
void rebox(rt_object_desc& w, const rt_object_desc& child) {
  const V3 bmn = get3(child.box_min), bmx = get3(child.box_max);
  if (w.kind == RT_OBJ_TRANSLATE) {
    const V3 d = get3(w.offset);
    set_box(w, vadd(bmn, d), vadd(bmx, d));
  } else if (w.kind == RT_OBJ_ROTATE_Y) {
    V3 minp = v3(FLT_MAX, FLT_MAX, FLT_MAX), maxp = v3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) {
          const float x = i ? bmx.x : bmn.x, y = j ? bmx.y : bmn.y, z = k ? bmx.z : bmn.z;
          const float nx = ffma(w.cos_t, x, fmul(w.sin_t, z)), nz = ffma(w.cos_t, z, -fmul(w.sin_t, x));
          minp = v3(fminf(minp.x, nx), fminf(minp.y, y), fminf(minp.z, nz));
          maxp = v3(fmaxf(maxp.x, nx), fmaxf(maxp.y, y), fmaxf(maxp.z, nz));
        }
    set_box(w, minp, maxp);
  } else {
    set_box(w, bmn, bmx);
  }
}
struct GroupExpander {
  SceneDesc& out;
  std::vector<int>& origin;
  std::vector<int> chain;  // wrappers above the current object, outermost first
  std::string err;
  size_t budget = (size_t)1 << 24;  // expanded entries (a group instanced many times multiplies)
  GroupExpander(SceneDesc& o, std::vector<int>& og) : out(o), origin(og) {}
  bool ends_in_group(int id) const {
    while (is_wrapper(out.obj[id].kind)) id = out.obj[id].child;
    return out.obj[id].kind == RT_OBJ_BVH;
  }
  void emit(int id, int k) {
    if (!err.empty()) return;
    if (out.top.size() >= budget) { err = "groups expand to more than 2^24 top-level objects"; return; }
    int cur = id;
    if (!chain.empty()) {
      if (out.obj[id].kind == RT_OBJ_MEDIUM) {
        for (int w : chain) if (out.obj[w].kind != RT_OBJ_WITH_MATERIAL) { err = "a medium cannot be wrapped in an instance"; return; }
      }
      int depth = 0;
      for (int w : chain) depth += out.obj[w].kind != RT_OBJ_WITH_MATERIAL;
      for (int c = id; is_wrapper(out.obj[c].kind); c = out.obj[c].child) depth += out.obj[c].kind != RT_OBJ_WITH_MATERIAL;
      if (depth > 4) { err = "more than 4 nested instance wrappers"; return; }
      for (size_t i = chain.size(); i-- > 0;) {
        rt_object_desc w = out.obj[chain[i]];
        w.child = cur;
        rebox(w, out.obj[cur]);
        out.obj.push_back(w);
        cur = (int)out.obj.size() - 1;
      }
    }
    out.top.push_back(cur);
    origin.push_back(k);
  }
  void walk(int id, int k) {
    if (!err.empty()) return;
    const rt_object_desc o = out.obj[id];
    if (is_wrapper(o.kind)) {
      if (!ends_in_group(id)) { emit(id, k); return; }
      chain.push_back(id);
      walk(o.child, k);
      chain.pop_back();
    } else if (o.kind == RT_OBJ_BVH) {
      std::vector<int> members;
      for (int c = id; c >= 0; c = out.obj[c].inward) members.push_back(out.obj[c].child);
      for (size_t i = members.size(); i-- > 0;) walk(members[i], k);  // creation order
    } else {
      emit(id, k);
    }
  }
};
}  // namespace

std::string expand_groups(const SceneDesc& sd, SceneDesc& out, std::vector<int>& origin) {
  out = sd;
  origin.clear();
  bool any = false;
  for (const auto& o : sd.obj) any = any || o.kind == RT_OBJ_BVH;
  if (!any) { origin.resize(sd.top.size()); for (size_t k = 0; k < sd.top.size(); ++k) origin[k] = (int)k; return ""; }
  out.top.clear();
  GroupExpander E(out, origin);
  for (size_t k = 0; k < sd.top.size() && E.err.empty(); ++k) E.walk(sd.top[k], (int)k);
  return E.err;
}

std::string sd_serialize(const SceneDesc& sd) {
  rt_sd_header h;
  memset(&h, 0, sizeof(h));
  h.magic = RT_SD_MAGIC; h.scene_id = sd.scene_id; h.nx = sd.nx; h.ny = sd.ny;
  h.n_tex = (int)sd.tex.size(); h.n_mat = (int)sd.mat.size(); h.n_obj = (int)sd.obj.size();
  h.n_top = (int)sd.top.size(); h.n_img = (int)sd.img.size();
  h.cam = sd.cam;
  std::string s;
  s.append((const char*)&h, sizeof(h));
  s.append((const char*)sd.tex.data(), sd.tex.size() * sizeof(rt_texture_desc));
  s.append((const char*)sd.mat.data(), sd.mat.size() * sizeof(rt_material_desc));
  s.append((const char*)sd.obj.data(), sd.obj.size() * sizeof(rt_object_desc));
  s.append((const char*)sd.top.data(), sd.top.size() * sizeof(int));
  s.append((const char*)sd.img.data(), sd.img.size() * sizeof(rt_image_desc));
  return s;
}

// Inverse of sd_serialize, with validation of every index (the buffer comes from the caller). Image pixels are not
// part of the SD: `images[i]` points to width*height*bpp bytes for image i (or is null: the texture renders cyan
// like the reference's invalid DeviceImage, texture.cuh:52).
std::string sd_deserialize(const void* buf, size_t bytes, const unsigned char* const* images, int n_images, SceneDesc& sd) {
  if (!buf || bytes < sizeof(rt_sd_header)) return "scene description too short";
  rt_sd_header h;
  memcpy(&h, buf, sizeof(h));
  if (h.magic != RT_SD_MAGIC) return "bad scene description magic";
  if (h.n_tex < 0 || h.n_mat < 0 || h.n_obj < 0 || h.n_top < 0 || h.n_img < 0 || h.nx <= 0 || h.ny <= 0) return "bad scene description header";
  const size_t need = sizeof(h) + (size_t)h.n_tex * sizeof(rt_texture_desc) + (size_t)h.n_mat * sizeof(rt_material_desc) +
                      (size_t)h.n_obj * sizeof(rt_object_desc) + (size_t)h.n_top * sizeof(int) + (size_t)h.n_img * sizeof(rt_image_desc);
  if (bytes < need) return "scene description truncated";
  const char* p = (const char*)buf + sizeof(h);
  sd = SceneDesc();
  sd.scene_id = h.scene_id; sd.nx = h.nx; sd.ny = h.ny; sd.cam = h.cam;
  sd.tex.resize(h.n_tex); memcpy(sd.tex.data(), p, sd.tex.size() * sizeof(rt_texture_desc)); p += sd.tex.size() * sizeof(rt_texture_desc);
  sd.mat.resize(h.n_mat); memcpy(sd.mat.data(), p, sd.mat.size() * sizeof(rt_material_desc)); p += sd.mat.size() * sizeof(rt_material_desc);
  sd.obj.resize(h.n_obj); memcpy(sd.obj.data(), p, sd.obj.size() * sizeof(rt_object_desc)); p += sd.obj.size() * sizeof(rt_object_desc);
  sd.top.resize(h.n_top); memcpy(sd.top.data(), p, sd.top.size() * sizeof(int)); p += sd.top.size() * sizeof(int);
  sd.img.resize(h.n_img); memcpy(sd.img.data(), p, sd.img.size() * sizeof(rt_image_desc));
  sd.default_nx = h.nx; sd.default_ny = h.ny; sd.default_spp = 10;
  for (const auto& t : sd.tex) {
    if (t.kind < RT_TEX_SOLID || t.kind > RT_TEX_UV_OFFSET) return "texture kind out of range";
    if (t.kind == RT_TEX_CHECKER && (t.even < 0 || t.even >= h.n_tex || t.odd < 0 || t.odd >= h.n_tex)) return "checker child out of range";
    if (t.kind == RT_TEX_UV_OFFSET && (t.even < 0 || t.even >= h.n_tex)) return "uv_offset base out of range";
    if (t.kind == RT_TEX_IMAGE && (t.image < 0 || t.image >= h.n_img)) return "image id out of range";
    // the octave count is a loop bound on the device (texture.cuh:94-100 -> perlin turb)
    if (t.kind == RT_TEX_NOODLE && !(t.p[3] >= 0.0f && t.p[3] <= 64.0f)) return "noodle texture: octave count outside 0..64";
  }
  for (const auto& d : sd.img) {
    if (d.width < 0 || d.height < 0 || d.width > 65536 || d.height > 65536 || (d.width > 0 && d.height > 0 && (d.bpp < 3 || d.bpp > 4)))
      return "image size / channel count out of range";
  }
  for (const auto& m : sd.mat) {
    if (m.kind < RT_MAT_LAMBERTIAN || m.kind > RT_MAT_ISOTROPIC) return "material kind out of range";
    if (m.tex >= h.n_tex) return "material texture out of range";
    if ((m.kind == RT_MAT_ISOTROPIC) && m.tex < 0) return "isotropic needs a texture";
  }
  for (int i = 0; i < h.n_obj; ++i) {
    const rt_object_desc& o = sd.obj[i];
    if (o.kind < RT_OBJ_SPHERE || o.kind > RT_OBJ_BVH) return "object kind out of range";
    const bool leaf = o.kind == RT_OBJ_SPHERE || o.kind == RT_OBJ_QUAD;
    if ((leaf || o.kind == RT_OBJ_MEDIUM || o.kind == RT_OBJ_WITH_MATERIAL) && (o.mat < 0 || o.mat >= h.n_mat)) return "object material out of range";
    for (int a = 0; a < 3; ++a)
      if (!std::isfinite(o.box_min[a]) || !std::isfinite(o.box_max[a])) return "object bounding box is not finite";
    if (o.kind == RT_OBJ_SPHERE && !(std::isfinite(o.radius) && std::isfinite(o.c0[0]) && std::isfinite(o.c0[1]) && std::isfinite(o.c0[2]) &&
                                     std::isfinite(o.dc[0]) && std::isfinite(o.dc[1]) && std::isfinite(o.dc[2]))) return "sphere is not finite";
    if (o.kind == RT_OBJ_WITH_MATERIAL) {
      if (o.child < 0 || o.child >= i) return "wrapper child must precede the wrapper";
    }
    if (o.kind == RT_OBJ_BVH) {  // a cell of a group: member and next cell were created before it
      if (o.child < 0 || o.child >= i) return "group member must precede the group";
      if (o.inward < -1 || o.inward >= i || (o.inward >= 0 && sd.obj[o.inward].kind != RT_OBJ_BVH)) return "group cell chain is broken";
    }
    if (o.kind == RT_OBJ_MEDIUM && o.child >= 0 && o.child < i) {
      int c = o.child;
      while (sd.obj[c].kind == RT_OBJ_TRANSLATE || sd.obj[c].kind == RT_OBJ_ROTATE_Y || sd.obj[c].kind == RT_OBJ_WITH_MATERIAL) c = sd.obj[c].child;
      if (sd.obj[c].kind == RT_OBJ_BVH) return "a group (bvh_node) as the boundary of a medium is not supported";
    }
    if (o.kind == RT_OBJ_BOX) {
      if (o.child < 0 || o.child + 6 > h.n_obj) return "box faces out of range";
      for (int f = 0; f < 6; ++f) if (sd.obj[o.child + f].kind != RT_OBJ_QUAD) return "box face is not a quad";
    }
    if (o.kind == RT_OBJ_TRANSLATE || o.kind == RT_OBJ_ROTATE_Y || o.kind == RT_OBJ_MEDIUM) {
      // children are created before their wrappers (like the reference's `new` order): no cycles, bounded depth
      if (o.child < 0 || o.child >= i) return "wrapper child must precede the wrapper";
      int depth = 0;  // material overrides vanish when the scene is flattened: they do not count
      for (int c = i; sd.obj[c].kind == RT_OBJ_TRANSLATE || sd.obj[c].kind == RT_OBJ_ROTATE_Y || sd.obj[c].kind == RT_OBJ_MEDIUM ||
                      sd.obj[c].kind == RT_OBJ_WITH_MATERIAL; c = sd.obj[c].child) {
        if ((sd.obj[c].kind == RT_OBJ_TRANSLATE || sd.obj[c].kind == RT_OBJ_ROTATE_Y) && ++depth > 4) return "more than 4 nested instance wrappers";
        if (c != i && sd.obj[c].kind == RT_OBJ_MEDIUM) return "a medium cannot be wrapped in an instance";
      }
    }
  }
  {
    std::vector<char> seen((size_t)h.n_obj, 0);
    for (int t : sd.top) {
      if (t < 0 || t >= h.n_obj) return "top-level object out of range";
      if (seen[t]) return "top-level list names an object twice";
      seen[t] = 1;
    }
  }
  sd.img_data.resize(h.n_img);
  for (int i = 0; i < h.n_img; ++i) {
    const rt_image_desc& d = sd.img[i];
    HostImage& im = sd.img_data[i];
    im.width = d.width; im.height = d.height; im.bpp = d.bpp;
    if (images && i < n_images && images[i] && d.width > 0 && d.height > 0 && d.bpp >= 3)
      im.px.assign(images[i], images[i] + (size_t)d.width * d.height * d.bpp);
  }
  return "";
}

}  // namespace rt
