// scene_builder.h — host-side scene vocabulary of the reference, producing the flat SD
// (include/rt_scene_desc.h).
//
// The reference builds its scenes inside <<<1,1>>> kernels with device-side `new`
// (main.cu:160-635). Here the same vocabulary is a host C++ builder whose methods take the SAME
// arguments with the SAME meaning as the reference constructors and store the SAME floats the
// constructors compute on the GPU (explicit fusion per rt_math.h; sinf/cosf/tanf evaluated by CUDA's
// libdevice through DevMath so they match the reference's device-side calls bit for bit):
//
//   sphere(c, r, mat) / sphere(c0, c1, r, mat)           sphere.cuh:21-38
//   quad(Q, u, v, mat, inward)                           quad.cuh:29-41, 49-54
//   make_box(a, b, mat)                                  quad.cuh:145-162 (+ compound6 94-122)
//   translate(obj, offset) / rotate_y(obj, deg)          hittable.cuh:52-54, 89-116
//   constant_medium(obj, density, albedo | tex)          constant_medium.cuh:24-28
//   lambertian / metal / dielectric / diffuse_light / isotropic      material.cuh:62-201
//   solid_color / checker_texture / image_texture / noise_texture    texture.cuh:16-76
//   noodle_texture / felt_texture / uv_offset_texture                texture.cuh:84-164
//   camera(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, t0, t1)   camera.cuh:21-78
//
// Handles are small ints (ids into the SD arrays). `add(obj)` is `d_list[i++] = obj`.
#pragma once
#include <string>
#include <vector>
#include "rt_math.h"
#include "rt_scene_desc.h"

namespace rt {

// sinf/cosf/tanf as CUDA's libdevice computes them (one tiny kernel launch per call; scene setup only).
struct DevMath {
  virtual ~DevMath() {}
  virtual float sinf_(float x) = 0;
  virtual float cosf_(float x) = 0;
  virtual float tanf_(float x) = 0;
};

struct HostImage {
  int width = 0, height = 0, bpp = 3;
  std::vector<unsigned char> px;
  bool valid() const { return !px.empty() && width > 0 && height > 0 && bpp >= 3; }
};

// Large host arrays (the C5 scale-up scenes: 10^6 objects, ~0.5 GB of descriptions and device tables) spend most of their
// time in the first touch of fresh memory: ask for huge pages before that touch where the kernel hands them out on request.
void advise_huge(void* p, size_t bytes);
template <class T> void big_reserve(std::vector<T>& v, size_t n) { v.reserve(n); advise_huge(v.data(), n * sizeof(T)); }
template <class T> void big_resize(std::vector<T>& v, size_t n) { big_reserve(v, n); v.resize(n); }

struct SceneDesc {
  int scene_id = 0, nx = 0, ny = 0;
  std::vector<rt_texture_desc> tex;
  std::vector<rt_material_desc> mat;
  std::vector<rt_object_desc> obj;
  std::vector<int> top;  // d_list in creation order
  std::vector<rt_image_desc> img;
  std::vector<HostImage> img_data;
  rt_camera_desc cam;
  // host-driver parameters of the reference's scene function (main.cu:654-1305)
  int default_nx = 0, default_ny = 0, default_spp = 0;
  float background[3] = {0, 0, 0};
  int gradient_bg = 0;
};

class SceneBuilder {
 public:
  SceneBuilder(SceneDesc& sd, DevMath& dm) : S(sd), M(dm) {}

  // textures
  int solid_color(V3 albedo);
  int checker_texture(float scale, int even, int odd);
  int image_texture(int image);
  int noise_texture(float scale);
  int noodle_texture(float stripes_k = 3.0f, float wiggle_amp = 3.0f, float wiggle_freq = 0.6f, int oct = 3,
                     V3 dir = v3(0, 0, 1), V3 noodle = v3(0.92f, 0.85f, 0.65f), V3 gap = v3(0.35f, 0.20f, 0.10f));
  int felt_texture(V3 base = v3(0.06f, 0.36f, 0.18f), float mottling_scale = 16.0f, float mottling_amt = 0.08f,
                   float fiber_scale = 4.0f, float fiber_amt = 0.03f);
  int uv_offset_texture(int base, float u_offset_turns, float v_offset = 0.f);
  int add_image(const HostImage& im);

  // materials
  int lambertian(V3 albedo) { return lambertian_tex(solid_color(albedo)); }
  int lambertian_tex(int tex);
  int metal(V3 albedo, float fuzz);
  int dielectric(float ref_idx);
  int diffuse_light(V3 c);
  int diffuse_light_tex(int tex);
  int isotropic_tex(int tex);

  // hittables
  int sphere(V3 cen, float r, int mat);
  int sphere(V3 cen1, V3 cen2, float r, int mat);
  int quad(V3 Q, V3 u, V3 v, int mat, bool inward = false);
  int make_box(V3 a, V3 b, int mat);
  int translate(int obj, V3 offset);
  int rotate_y(int obj, float angle_degrees);
  int with_material(int obj, int mat);  // hittable.cuh:154-178
  int bvh_node(const std::vector<int>& members);  // bvh.cuh:29-84 as an object: a group (RT_OBJ_BVH cells); -1 for an empty list
  int constant_medium(int boundary, float density, V3 albedo);
  int constant_medium_tex(int boundary, float density, int tex);

  void add(int obj) { S.top.push_back(obj); }  // d_list[i++] = obj
  void camera(V3 lookfrom, V3 lookat, V3 vup, float vfov, float aspect, float aperture, float focus_dist,
              double t0, double t1);

  SceneDesc& S;
  DevMath& M;
  // true while building objects whose arguments are literals in the reference's generator, i.e. whose
  // constructor arithmetic NVVM constant-folds (every operation rounded separately, no FMA).
  bool folded = false;

 private:
  int push_obj(const rt_object_desc& o) { S.obj.push_back(o); return (int)S.obj.size() - 1; }
};

// Generators: restatements of create_world_* (main.cu:160-635). scene_id follows main()'s switch
// (main.cu:1311-1320): 1 bouncing, 2 checker, 3 earth, 4 perlin, 5 quads, 6 simple_light, 7 cornell,
// 8 cornell_smoke, 9 final, 10 original. grid_half generalises GRID_MIN/MAX (main.cu:140-141) for
// the C5 scale-up (11 = the reference's 488 spheres). Returns "" or an error message.
std::string generate_scene(SceneDesc& sd, DevMath& dm, int scene_id, int nx, int ny, int grid_half,
                           const std::string& texture_dir);

// The order in which the reference's BVH visits leaves (= d_list after the in-place selection sorts of
// bvh.cuh:46-81). rank[k] = position of top[k] in that order. Used only to break exact-t ties the way
// bvh_node::hit does (bvh.cuh:95-106). O(n^2) like the reference; for n > limit returns identity.
std::vector<int> reference_leaf_order(const SceneDesc& sd, int limit = 20000);

bool load_ppm(const std::string& path, HostImage& out);
bool load_jpeg(const std::string& path, HostImage& out, std::string& err);  // jpeg_baseline.cpp
// .jpg / .jpeg (baseline JPEG, decoded like the reference's stbi_load(path, .., 3)) or .ppm (P6). "" or an error message.
std::string load_texture_file(const std::string& path, HostImage& out);
// Groups (RT_OBJ_BVH) are resolved before the scene is flattened: every top-level entry whose wrapper chain ends in a group
// becomes one entry per member under a copy of that chain (boxes recomputed like the wrappers' constructors do), which is
// what the reference's own final scene does by hand for its sphere cluster (main.cu:545-551). origin[k] = index in sd.top
// of the entry that expanded entry k came from. Returns "" or an error message; out == sd when the scene has no group.
std::string expand_groups(const SceneDesc& sd, SceneDesc& out, std::vector<int>& origin);
std::string sd_serialize(const SceneDesc& sd);  // binary SD file image
std::string sd_deserialize(const void* buf, size_t bytes, const unsigned char* const* images, int n_images, SceneDesc& sd);  // "" or an error

}  // namespace rt
