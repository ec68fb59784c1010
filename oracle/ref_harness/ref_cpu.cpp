// ref_cpu.cpp — TEST/BENCH INFRASTRUCTURE (not product code).
//
// Runs the reference's OWN render code (all headers + main.cu up to the first host function, compiled
// unmodified from /root/reference by oracle/build_ref.sh) on host cores behind oracle/shim. The
// reference has no CPU path as shipped; this is the "shimmed" CPU build: plain
// g++ -O2 -ffp-contract=off, glibc libm, no FMA contraction, __sinf -> sinf. One OpenMP thread plays
// one CUDA thread (one pixel) at a time.
//
// Same argv and outputs as ref_gpu.cu. Extra: --decode in.jpg out.ppm converts a texture with the
// reference's vendored stb_image (external/stb_image.h) so that decoded bytes match the reference's.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include <iostream>
#include <omp.h>
#include "cuda_runtime.h"
#include "curand_kernel.h"
#include "math_constants.h"

thread_local rh_dim3 threadIdx, blockIdx, blockDim;
static int g_grid_half = 11;

#define private public
#include REF_DEVICE_PART
#undef private

#include "ref_common.h"

static DeviceImage load_ppm(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); return DeviceImage{}; }
  int w = 0, h = 0, mx = 0;
  if (fscanf(f, "P6 %d %d %d", &w, &h, &mx) != 3) { fclose(f); return DeviceImage{}; }
  fgetc(f);
  unsigned char* px = (unsigned char*)malloc((size_t)w * h * 3);
  size_t got = fread(px, 1, (size_t)w * h * 3, f);
  fclose(f);
  if (got != (size_t)w * h * 3) return DeviceImage{};
  return DeviceImage{px, w, h, 3};
}

static void set_thread(int i, int j) {
  threadIdx.x = i; threadIdx.y = j; threadIdx.z = 0;
  blockIdx = rh_dim3(); blockDim = rh_dim3();
}

int main(int argc, char** argv) {
  if (argc == 4 && std::string(argv[1]) == "--decode") {
    int w = 0, h = 0, n = 0;
    unsigned char* p = stbi_load(argv[2], &w, &h, &n, 3);  // image_io.h:26
    if (!p) { fprintf(stderr, "stbi_load failed: %s\n", argv[2]); return 1; }
    FILE* f = fopen(argv[3], "wb");
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    fwrite(p, 1, (size_t)w * h * 3, f);
    fclose(f);
    return 0;
  }
  RhArgs a;
  int threads = omp_get_max_threads();
  for (int i = 1; i < argc; ++i) {
    std::string s = argv[i];
    auto nxt = [&]() { return i + 1 < argc ? argv[++i] : (char*)"0"; };
    if (s == "--scene") a.scene = atoi(nxt());
    else if (s == "--nx") a.nx = atoi(nxt());
    else if (s == "--ny") a.ny = atoi(nxt());
    else if (s == "--ns") a.ns = atoi(nxt());
    else if (s == "--reps") a.reps = atoi(nxt());
    else if (s == "--grid") a.grid_half = atoi(nxt());
    else if (s == "--ids") a.ids = atoi(nxt());
    else if (s == "--count") a.count = atoi(nxt());
    else if (s == "--textures") a.tex_dir = nxt();
    else if (s == "--out") a.out = nxt();
    else if (s == "--threads") threads = atoi(nxt());
    else { fprintf(stderr, "unknown arg %s\n", s.c_str()); return 2; }
  }
  omp_set_num_threads(threads);
  g_grid_half = a.grid_half;
  RhSceneParams sp = rh_scene_params(a.scene, a.grid_half);
  if (sp.nx == 0) { fprintf(stderr, "bad scene\n"); return 2; }
  const int nx = a.nx > 0 ? a.nx : sp.nx, ny = a.ny > 0 ? a.ny : sp.ny, ns = a.ns;
  const float gamma = 2.2f;

  std::string td = a.tex_dir;
  DeviceImage earth{}, ball{};
  if (a.scene == 3 || a.scene == 9) earth = load_ppm(td + "/earthmap.ppm");
  if (a.scene == 6) ball = load_ppm(td + "/poolball.ppm");
  if (a.scene == 10) { earth = load_ppm(td + "/porcelain.ppm"); ball = load_ppm(td + "/8ball.ppm"); }
  if ((a.scene == 3 || a.scene == 9 || a.scene == 10) && !earth.valid()) { fprintf(stderr, "texture missing\n"); return 3; }

  const int num_pixels = nx * ny;
  std::vector<vec3> fb(num_pixels);
  std::vector<curandState> rand_state(num_pixels);
  curandState rand_state2;
  set_thread(0, 0);
  rand_init(&rand_state2);

  camera* cam = nullptr; hittable* world = nullptr;
  const int n_list = sp.n_list;
  std::vector<hittable*> list(n_list + 8, nullptr);
  hittable** d_list = list.data();
  auto t0 = std::chrono::steady_clock::now();
  switch (a.scene) {
    case 1: create_world_bouncing(d_list, &world, &cam, nx, ny, &rand_state2); break;
    case 2: create_world_checker(d_list, &world, &cam, nx, ny, &rand_state2); break;
    case 3: create_world_earth(d_list, &world, &cam, nx, ny, earth); break;
    case 4: create_world_perlin(d_list, &world, &cam, nx, ny, 4.0f); break;
    case 5: create_world_quads(d_list, &world, &cam, nx, ny); break;
    case 6: create_world_simple_light(d_list, &world, &cam, nx, ny, ball); break;
    case 7: create_world_cornell(d_list, &world, &cam, nx, ny); break;
    case 8: create_world_cornell_smoke(d_list, &world, &cam, nx, ny); break;
    case 9: create_world_final(d_list, &world, &cam, nx, ny, earth); break;
    case 10: create_world_original(d_list, &world, &cam, nx, ny, earth, ball); break;
  }
  auto t1 = std::chrono::steady_clock::now();
  double build_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();

  // ---- export ----
  RefVptrs vp; rh_probe_vptrs(&vp);
  const int cap_obj = n_list * 8 + 64, cap_mat = n_list + 64, cap_tex = n_list * 2 + 64, cap_img = 8;
  std::vector<rt_texture_desc> tex(cap_tex); std::vector<const void*> tex_ptr(cap_tex);
  std::vector<rt_material_desc> mat(cap_mat); std::vector<const void*> mat_ptr(cap_mat);
  std::vector<rt_object_desc> obj(cap_obj);
  std::vector<rt_image_desc> img(cap_img); std::vector<const void*> img_ptr(cap_img);
  std::vector<int> top(n_list + 8);
  RefExport E{};
  E.vp = &vp;
  E.tex = tex.data(); E.tex_ptr = tex_ptr.data(); E.cap_tex = cap_tex;
  E.mat = mat.data(); E.mat_ptr = mat_ptr.data(); E.cap_mat = cap_mat;
  E.obj = obj.data(); E.cap_obj = cap_obj;
  E.img = img.data(); E.img_ptr = img_ptr.data(); E.cap_img = cap_img;
  E.top = top.data();
  rh_export_scene(E, d_list, n_list);
  rt_camera_desc camd; memset(&camd, 0, sizeof(camd));
  rh_export_camera(&camd, cam);
  int leaf_bad = 0;
  {
    std::vector<const hittable*> leaves(n_list + 1);
    int n = 0;
    rh_leaf_order(&vp, world, leaves.data(), &n);
    leaf_bad = (n != n_list);
    for (int k = 0; k < n_list && k < n; ++k) leaf_bad += (leaves[k] != d_list[k]);
  }
  if (a.out) {
    rt_sd_header h; memset(&h, 0, sizeof(h));
    h.magic = RT_SD_MAGIC; h.scene_id = a.scene; h.nx = nx; h.ny = ny;
    h.n_tex = E.n_tex; h.n_mat = E.n_mat; h.n_obj = E.n_obj; h.n_top = E.n_top; h.n_img = E.n_img;
    h.cam = camd;
    std::string p = std::string(a.out) + ".sd";
    FILE* f = fopen(p.c_str(), "wb");
    fwrite(&h, sizeof(h), 1, f);
    fwrite(tex.data(), sizeof(rt_texture_desc), E.n_tex, f);
    fwrite(mat.data(), sizeof(rt_material_desc), E.n_mat, f);
    fwrite(obj.data(), sizeof(rt_object_desc), E.n_obj, f);
    fwrite(top.data(), sizeof(int), E.n_top, f);
    fwrite(img.data(), sizeof(rt_image_desc), E.n_img, f);
    fclose(f);
  }

  const vec3 bg(sp.bg[0], sp.bg[1], sp.bg[2]);

  // ---- primary-hit AOV ----
  int id_mismatch = -1;
  if (a.ids && a.out) {
    std::vector<int> o(num_pixels), m(num_pixels); std::vector<float> t(num_pixels), bt(num_pixels);
    RefIds R{o.data(), t.data(), m.data(), bt.data()};
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) rh_primary_ids(E, d_list, n_list, &world, cam, i, j, nx, ny, R);
    id_mismatch = 0;
    for (int k = 0; k < num_pixels; ++k) id_mismatch += (memcmp(&t[k], &bt[k], 4) != 0);
    std::string p = std::string(a.out) + ".ids";
    FILE* f = fopen(p.c_str(), "wb");
    int hdr[2] = {nx, ny}; fwrite(hdr, sizeof(int), 2, f);
    fwrite(o.data(), 4, num_pixels, f); fwrite(t.data(), 4, num_pixels, f);
    fwrite(m.data(), 4, num_pixels, f); fwrite(bt.data(), 4, num_pixels, f);
    fclose(f);
  }

  // ---- stock render: render_init + render called once per pixel (main.cu:96-133) ----
  double best_ms = 1e30, sum_ms = 0;
  for (int rep = 0; rep < a.reps; ++rep) {
    auto r0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        set_thread(i, j);
        render_init(nx, ny, rand_state.data());
        render(fb.data(), nx, ny, ns, gamma, &cam, &world, rand_state.data(), bg, sp.gradient);
      }
    auto r1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(r1 - r0).count();
    sum_ms += ms; if (ms < best_ms) best_ms = ms;
  }
  if (a.out && a.reps > 0) {
    std::string p = std::string(a.out) + ".fb";
    FILE* f = fopen(p.c_str(), "wb");
    int hdr[3] = {nx, ny, ns}; fwrite(hdr, sizeof(int), 3, f);
    fwrite(fb.data(), sizeof(vec3), num_pixels, f);
    fclose(f);
  }

  // ---- ray count ----
  unsigned long long rays = 0; long long fb_mismatch = -1;
  if (a.count) {
    std::vector<vec3> fb2(num_pixels);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rays)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        unsigned long long r = 0;
        fb2[j * nx + i] = rh_render_pixel_counted(i, j, nx, ny, ns, gamma, &cam, &world, bg, sp.gradient, r);
        rays += r;
      }
    if (a.reps > 0) {
      fb_mismatch = 0;
      for (int k = 0; k < num_pixels; ++k) fb_mismatch += (memcmp(&fb2[k], &fb[k], sizeof(vec3)) != 0);
    }
  }

  const double samples = (double)num_pixels * ns;
  const double mean_ms = a.reps > 0 ? sum_ms / a.reps : 0.0;
  printf("{\"impl\": \"reference-cpu-shim\", \"scene\": %d, \"name\": \"%s\", \"nx\": %d, \"ny\": %d, \"ns\": %d, "
         "\"grid_half\": %d, \"threads\": %d, \"n_top\": %d, \"n_obj\": %d, \"n_mat\": %d, \"n_tex\": %d, "
         "\"export_error\": %d, \"leaf_order_mismatch\": %d, \"id_t_mismatch\": %d, \"build_ms\": %.3f, "
         "\"render_ms_best\": %.4f, \"render_ms_mean\": %.4f, \"reps\": %d, \"rays\": %llu, \"samples\": %.0f, "
         "\"rays_per_sample\": %.4f, \"counted_fb_mismatch\": %lld, \"mrays_per_s\": %.3f, \"msamples_per_s\": %.3f}\n",
         a.scene, sp.name, nx, ny, ns, a.grid_half, threads, E.n_top, E.n_obj, E.n_mat, E.n_tex, E.error, leaf_bad,
         id_mismatch, build_ms, a.reps > 0 ? best_ms : 0.0, mean_ms, a.reps, rays, samples, rays / samples,
         fb_mismatch, mean_ms > 0 ? rays / (mean_ms * 1e3) : 0.0, mean_ms > 0 ? samples / (mean_ms * 1e3) : 0.0);
  return 0;
}
