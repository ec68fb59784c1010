// generators.cpp — host restatements of the reference's ten scene generators
// (create_world_*, main.cu:160-635) on top of SceneBuilder, plus the per-scene host parameters of the
// reference's scene functions (main.cu:654-1305: resolution, spp, background, gradient flag).
//
// Same constants, same object order (d_list[i++]), same scene-RNG draw order (one XORWOW stream,
// curand_init(1984,0,0), main.cu:92,164). `folded` marks literal-only constructor calls that NVVM
// constant-folds in the reference build (see scene_builder.cpp).
#include "scene_builder.h"
#include "rng.h"

namespace rt {

namespace {

struct SceneRng {
  Xorwow st;
  SceneRng() { st.init(1984); }       // rand_init, main.cu:89-94
  float rnd() { return st.uniform(); }  // RND, main.cu:137
};

V3 pick_ut_color(float r) {  // main.cu:149-158
  if (r < 0.25f) return v3(1.0f, 1.0f, 1.0f);
  else if (r < 0.50f) return v3(1.0f, 0.51f, 0.0f);
  else if (r < 0.75f) return v3(0.60f, 0.60f, 0.60f);
  else return v3(0.0f, 0.0f, 0.0f);
}

V3 random_in_unit_cube(int seed) {  // util.cuh:3-11
  uint32_t s = 1103515245u * (uint32_t)(seed + 1) + 12345u;
  auto next01 = [&]() {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    return fmul((float)(s & 0xFFFFFF), 1.0f / 16777216.0f);
  };
  float a = next01(), b = next01(), c = next01();
  return v3(a, b, c);
}

float aspect_of(int nx, int ny) { return fdiv((float)nx, (float)ny); }

// main.cu:160-244 (grid_half = 11 in the reference; larger for the C5 scale-up)
void world_bouncing(SceneBuilder& B, int nx, int ny, int grid_half) {
  SceneRng R;
  {  // one object, material and (mostly) texture per grid cell: no vector regrowth at the C5 scale-up sizes, huge pages
    const size_t cells = (size_t)4 * grid_half * grid_half + 8;
    big_reserve(B.S.obj, cells); big_reserve(B.S.mat, cells); big_reserve(B.S.tex, cells); big_reserve(B.S.top, cells);
  }
  const V3 UT_ORANGE = v3(1.0f, 0.51f, 0.0f);
  int checker = B.checker_texture(0.64f, B.solid_color(v3(1.0f, 1.0f, 1.0f)), B.solid_color(UT_ORANGE));
  B.add(B.sphere(v3(0.0f, -1000.0f, -1.0f), 1000.0f, B.lambertian_tex(checker)));
  const float P_EMISSIVE = 0.10f, EMIT_POWER = 4.0f;
  for (int a = -grid_half; a < grid_half; a++) {
    for (int b = -grid_half; b < grid_half; b++) {
      float choose_mat = R.rnd();
      float rx = R.rnd(), rz = R.rnd();
      V3 center = v3(ffma(0.9f, rx, (float)a), 0.2f, ffma(0.9f, rz, (float)b));
      if (choose_mat < 0.8f) {
        float ry = R.rnd(), rv = R.rnd();
        // vel(0, 0.5*RND, 0.25*(RND-0.5)); center2 = center + vel, fused per the SASS of main.cu:193-194
        V3 center2 = v3(fadd(center.x, 0.0f), ffma(0.5f, ry, 0.2f), ffma(fsub(rv, 0.5f), 0.25f, center.z));
        if (R.rnd() < P_EMISSIVE) {
          B.add(B.sphere(center, center2, 0.2f, B.diffuse_light(vscale(EMIT_POWER, UT_ORANGE))));
        } else {
          V3 albedo = pick_ut_color(R.rnd());
          B.add(B.sphere(center, center2, 0.2f, B.lambertian(albedo)));
        }
      } else if (choose_mat < 0.95f) {
        V3 albedo = pick_ut_color(R.rnd());
        if (fadd(fadd(albedo.x, albedo.y), albedo.z) < 1e-5f) albedo = v3(0.15f, 0.15f, 0.15f);
        float fuzz = fmul(0.5f, R.rnd());
        B.add(B.sphere(center, 0.2f, B.metal(albedo, fuzz)));
      } else {
        B.add(B.sphere(center, 0.2f, B.dielectric(1.5f)));
      }
    }
  }
  B.add(B.sphere(v3(0.0f, 1.0f, 0.0f), 1.0f, B.dielectric(1.5f)));
  B.add(B.sphere(v3(-4.0f, 1.0f, 0.0f), 1.0f, B.lambertian(v3(0.4f, 0.2f, 0.1f))));
  B.add(B.sphere(v3(4.0f, 1.0f, 0.0f), 1.0f, B.metal(v3(0.7f, 0.6f, 0.5f), 0.0f)));
  V3 lookfrom = v3(13.0f, 2.0f, 3.0f), lookat = v3(0, 0, 0);
  V3 d = vsub(lookfrom, lookat);
  float dist_to_focus = fsqrt(fadd(fadd(fmul(d.x, d.x), fmul(d.y, d.y)), fmul(d.z, d.z)));  // folded
  B.camera(lookfrom, lookat, v3(0, 1, 0), 30.0f, aspect_of(nx, ny), 0.1f, dist_to_focus, 0.0, 1.0);
}

// main.cu:246-280
void world_checker(SceneBuilder& B, int nx, int ny) {
  int checker = B.checker_texture(0.32f, B.solid_color(v3(0.2f, 0.3f, 0.1f)), B.solid_color(v3(0.9f, 0.9f, 0.9f)));
  int lam = B.lambertian_tex(checker);
  B.add(B.sphere(v3(0, -10, 0), 10.0f, lam));
  B.add(B.sphere(v3(0, 10, 0), 10.0f, lam));
  B.camera(v3(13.0f, 2.0f, 3.0f), v3(0, 0, 0), v3(0, 1, 0), 20.0f, aspect_of(nx, ny), 0.0f, 10.0f, 0.0, 1.0);
}

// main.cu:282-308
void world_earth(SceneBuilder& B, int nx, int ny, int earth_img) {
  int lam = B.lambertian_tex(B.image_texture(earth_img));
  B.add(B.sphere(v3(0, 0, 0), 2.0f, lam));
  B.camera(v3(0, 0, 12.0f), v3(0, 0, 0), v3(0, 1, 0), 20.0f, aspect_of(nx, ny), 0.0f, 12.0f, 0.0, 1.0);
}

// main.cu:310-329 (scale = 4, main.cu:903)
void world_perlin(SceneBuilder& B, int nx, int ny, float scale) {
  int lam = B.lambertian_tex(B.noise_texture(scale));
  B.add(B.sphere(v3(0, -1000, 0), 1000.f, lam));
  B.add(B.sphere(v3(0, 2, 0), 2.f, lam));
  B.camera(v3(13, 2, 3), v3(0, 0, 0), v3(0, 1, 0), 20.0f, aspect_of(nx, ny), 0.0f, 10.0f, 0.0, 1.0);
}

// main.cu:331-358
void world_quads(SceneBuilder& B, int nx, int ny) {
  int left_red = B.lambertian(v3(1.0f, 0.2f, 0.2f));
  int back_green = B.lambertian(v3(0.2f, 1.0f, 0.2f));
  int right_blue = B.lambertian(v3(0.2f, 0.2f, 1.0f));
  int upper_orange = B.lambertian(v3(1.0f, 0.5f, 0.0f));
  int lower_teal = B.lambertian(v3(0.2f, 0.8f, 0.8f));
  B.folded = true;
  B.add(B.quad(v3(-3, -2, 5), v3(0, 0, -4), v3(0, 4, 0), left_red));
  B.add(B.quad(v3(-2, -2, 0), v3(4, 0, 0), v3(0, 4, 0), back_green));
  B.add(B.quad(v3(3, -2, 1), v3(0, 0, 4), v3(0, 4, 0), right_blue));
  B.add(B.quad(v3(-2, 3, 1), v3(4, 0, 0), v3(0, 0, 4), upper_orange));
  B.add(B.quad(v3(-2, -3, 5), v3(4, 0, 0), v3(0, 0, -4), lower_teal));
  B.folded = false;
  B.camera(v3(0, 0, 9), v3(0, 0, 0), v3(0, 1, 0), 80.0f, aspect_of(nx, ny), 0.0f, 10.0f, 0.0, 1.0);
}

// main.cu:360-400
void world_simple_light(SceneBuilder& B, int nx, int ny, int ball_img) {
  int feltlam = B.lambertian_tex(B.felt_texture(v3(0.06f, 0.36f, 0.18f), 16.0f, 0.08f, 4.0f, 0.03f));
  B.add(B.sphere(v3(0, -1000, 0), 1000.f, feltlam));
  int base_img = B.image_texture(ball_img);
  float u_rot_turns = fdiv(60.0f, 360.0f);
  int ball_diffuse = B.lambertian_tex(B.uv_offset_texture(base_img, u_rot_turns));
  const V3 C = v3(0, 2, 0);
  const float Rr = 2.0f;
  B.add(B.sphere(C, Rr, ball_diffuse));
  B.add(B.sphere(C, fadd(Rr, 0.02f), B.dielectric(1.5f)));
  int light1 = B.diffuse_light(v3(4, 4, 4));
  int light2 = B.diffuse_light(v3(4, 4, 4));
  B.add(B.sphere(v3(0, 7, 0), 2.f, light1));
  B.folded = true;
  B.add(B.quad(v3(3, 1, -2), v3(2, 0, 0), v3(0, 2, 0), light2));
  B.folded = false;
  V3 lookfrom = v3(26, 3, 6), lookat = v3(0, 2, 0);
  V3 d = vsub(lookfrom, lookat);
  float dist_to_focus = fsqrt(fadd(fadd(fmul(d.x, d.x), fmul(d.y, d.y)), fmul(d.z, d.z)));
  B.camera(lookfrom, lookat, v3(0, 1, 0), 20.0f, aspect_of(nx, ny), 0.0f, dist_to_focus, 0.0, 1.0);
}

float len_folded(V3 d) { return fsqrt(fadd(fadd(fmul(d.x, d.x), fmul(d.y, d.y)), fmul(d.z, d.z))); }

// main.cu:402-450
void world_cornell(SceneBuilder& B, int nx, int ny) {
  int red = B.lambertian(v3(.65f, .05f, .05f));
  int blue = B.lambertian(v3(.15f, .15f, .75f));
  int white = B.lambertian(v3(.73f, .73f, .73f));
  int light = B.diffuse_light(v3(15.f, 15.f, 15.f));
  B.folded = true;
  B.add(B.quad(v3(0, 0, 0), v3(0, 555, 0), v3(0, 0, 555), blue, true));
  B.add(B.quad(v3(555, 0, 555), v3(0, 555, 0), v3(0, 0, -555), red, true));
  B.add(B.quad(v3(0, 0, 0), v3(555, 0, 0), v3(0, 0, 555), white, true));
  B.add(B.quad(v3(0, 555, 555), v3(555, 0, 0), v3(0, 0, -555), white, true));
  B.add(B.quad(v3(555, 0, 555), v3(-555, 0, 0), v3(0, 555, 0), white, true));
  B.add(B.quad(v3(213, 554, 227), v3(130, 0, 0), v3(0, 0, 105), light, true));
  int proto_short = B.make_box(v3(0, 0, 0), v3(165, 165, 165), white);
  int proto_tall = B.make_box(v3(0, 0, 0), v3(165, 330, 165), white);
  B.folded = false;
  B.add(B.translate(B.rotate_y(proto_short, -18.f), v3(130.f, 0.f, 65.f)));
  B.add(B.translate(B.rotate_y(proto_tall, 15.f), v3(265.f, 0.f, 295.f)));
  int glass = B.dielectric(1.5f);
  B.add(B.sphere(v3(278.f, 335.f, 150.f), 60.f, glass));
  B.add(B.sphere(v3(278.f, 335.f, 150.f), -59.0f, glass));
  V3 lookfrom = v3(278, 278, -800), lookat = v3(278, 278, 0);
  B.camera(lookfrom, lookat, v3(0, 1, 0), 40.0f, aspect_of(nx, ny), 0.0f, len_folded(vsub(lookfrom, lookat)), 0.0, 1.0);
}

// main.cu:452-486
void world_cornell_smoke(SceneBuilder& B, int nx, int ny) {
  int red = B.lambertian(v3(.65f, .05f, .05f));
  int white = B.lambertian(v3(.73f, .73f, .73f));
  int green = B.lambertian(v3(.12f, .45f, .15f));
  int light = B.diffuse_light(v3(7.f, 7.f, 7.f));
  B.folded = true;
  B.add(B.quad(v3(555, 0, 0), v3(0, 555, 0), v3(0, 0, 555), green, true));
  B.add(B.quad(v3(0, 0, 0), v3(0, 555, 0), v3(0, 0, 555), red, true));
  B.add(B.quad(v3(0, 555, 0), v3(555, 0, 0), v3(0, 0, 555), white, true));
  B.add(B.quad(v3(0, 0, 0), v3(555, 0, 0), v3(0, 0, 555), white, true));
  B.add(B.quad(v3(0, 0, 555), v3(555, 0, 0), v3(0, 555, 0), white, true));
  B.add(B.quad(v3(113, 554, 127), v3(330, 0, 0), v3(0, 0, 305), light, true));
  int b1 = B.make_box(v3(0, 0, 0), v3(165, 330, 165), white);
  B.folded = false;
  b1 = B.translate(B.rotate_y(b1, 15.f), v3(265.f, 0.f, 295.f));
  B.folded = true;
  int b2 = B.make_box(v3(0, 0, 0), v3(165, 165, 165), white);
  B.folded = false;
  b2 = B.translate(B.rotate_y(b2, -18.f), v3(130.f, 0.f, 65.f));
  B.add(B.constant_medium(b1, 0.01f, v3(0.5f, 0.5f, 0.5f)));
  B.add(B.constant_medium(b2, 0.01f, v3(1, 1, 1)));
  V3 lookfrom = v3(278, 278, -800), lookat = v3(278, 278, 0);
  B.camera(lookfrom, lookat, v3(0, 1, 0), 40.0f, aspect_of(nx, ny), 0.0f, len_folded(vsub(lookfrom, lookat)), 0.0, 1.0);
}

V3 rotate_y_deg(DevMath& M, V3 p, float deg) {  // main.cu:489-496
  float r = fmul(deg, 0.017453292519943295f);
  float c = M.cosf_(r), s = M.sinf_(r);
  return v3(ffma(c, p.x, fmul(s, p.z)), p.y, ffma(c, p.z, -fmul(s, p.x)));
}

void ground_boxes(SceneBuilder& B, int ground) {  // main.cu:505-515 / 571-581
  const int S = 20;
  for (int ix = 0; ix < S; ++ix)
    for (int iz = 0; iz < S; ++iz) {
      float w = 100.0f;
      float x0 = ffma((float)ix, w, -1000.0f);
      float z0 = ffma((float)iz, w, -1000.0f);
      float y1 = fadd(1.0f, fdiv(fmul(100.0f, (float)((ix * 13 + iz * 37) % 100)), 100.0f));
      B.add(B.make_box(v3(x0, 0, z0), v3(fadd(x0, w), y1, fadd(z0, w)), ground));
    }
}

void cluster(SceneBuilder& B, int white) {  // main.cu:546-552 / 625-631
  const int ns = 1000;
  for (int j = 0; j < ns; ++j) {
    const V3 r = random_in_unit_cube(j);
    V3 p = vscale(165.0f, r);
    p = vadd(rotate_y_deg(B.M, p, 15.0f), v3(-100, 270, 395));
    p.y = ffma(r.y, 165.0f, 270.0f);  // y passes through the rotation: r.y*165 + 270 fuses (SASS of main.cu:550)
    B.add(B.sphere(p, 10.0f, white));
  }
}

// main.cu:498-562
void world_final(SceneBuilder& B, int nx, int ny, int earth_img) {
  int white = B.lambertian(v3(.73f, .73f, .73f));
  int ground = B.lambertian(v3(0.48f, 0.83f, 0.53f));
  int light = B.diffuse_light(v3(7, 7, 7));
  ground_boxes(B, ground);
  B.folded = true;
  B.add(B.quad(v3(123, 554, 147), v3(300, 0, 0), v3(0, 0, 265), light, true));
  B.folded = false;
  V3 c1 = v3(400, 400, 200), c2 = vadd(c1, v3(30, 0, 0));
  B.add(B.sphere(c1, c2, 50.f, B.lambertian(v3(0.7f, 0.3f, 0.1f))));
  B.add(B.sphere(v3(260, 150, 45), 50.f, B.dielectric(1.5f)));
  B.add(B.sphere(v3(0, 150, 145), 50.f, B.metal(v3(0.8f, 0.8f, 0.9f), 1.0f)));
  B.add(B.sphere(v3(360, 150, 145), 70.f, B.dielectric(1.5f)));
  B.add(B.constant_medium(B.sphere(v3(360, 150, 145), 70.f, B.dielectric(1.5f)), 0.2f, v3(0.2f, 0.4f, 0.9f)));
  B.add(B.constant_medium(B.sphere(v3(0, 0, 0), 5000.f, B.dielectric(1.5f)), 0.0001f, v3(1, 1, 1)));
  B.add(B.sphere(v3(400, 200, 400), 100.f, B.lambertian_tex(B.image_texture(earth_img))));
  B.add(B.sphere(v3(220, 280, 300), 80.f, B.lambertian_tex(B.noise_texture(0.2f))));
  cluster(B, white);
  V3 lookfrom = v3(478, 278, -600), lookat = v3(278, 278, 0);
  B.camera(lookfrom, lookat, v3(0, 1, 0), 40.0f, aspect_of(nx, ny), 0.0f, len_folded(vsub(lookfrom, lookat)), 0.0, 1.0);
}

// main.cu:564-635
void world_original(SceneBuilder& B, int nx, int ny, int ball_img) {
  int white = B.lambertian(v3(.73f, .73f, .73f));
  int ground = B.lambertian(v3(0.88f, 0.50f, 0.76f));
  int light = B.diffuse_light(v3(7, 7, 7));
  ground_boxes(B, ground);
  B.folded = true;
  B.add(B.quad(v3(123, 554, 147), v3(300, 0, 0), v3(0, 0, 265), light, true));
  B.folded = false;
  V3 c1 = v3(400, 400, 200), c2 = vadd(c1, v3(30, 0, 0));
  B.add(B.sphere(c1, c2, 50.f, B.lambertian(v3(0.0488f, 0.0148f, 0.0171f))));
  B.add(B.sphere(v3(260, 150, 45), 50.f, B.dielectric(1.5f)));
  B.add(B.sphere(v3(0, 150, 145), 50.f, B.metal(v3(0.6387f, 0.3605f, 0.8826f), 1.0f)));
  int eightball = B.lambertian_tex(B.image_texture(ball_img));
  B.add(B.sphere(v3(360.f, 150.f, 145.f), 70.f, eightball));
  B.add(B.sphere(v3(360, 150, 145), fadd(70.f, 0.5f), B.dielectric(1.5f)));
  B.add(B.constant_medium(B.sphere(v3(0, 0, 0), 5000.f, B.dielectric(1.5f)), 0.0001f, v3(1, 1, 1)));
  B.add(B.sphere(v3(400, 200, 400), 100.f, B.metal(v3(0.23f, 0.24f, 0.85f), 0.02f)));
  B.add(B.sphere(v3(220, 280, 300), 80.f, B.lambertian_tex(B.noodle_texture(0.2f))));
  cluster(B, white);
  V3 lookfrom = v3(478, 278, -600), lookat = v3(278, 278, 0);
  B.camera(lookfrom, lookat, v3(0, 1, 0), 40.0f, aspect_of(nx, ny), 0.0f, len_folded(vsub(lookfrom, lookat)), 0.0, 1.0);
}

struct HostParams { int nx, ny, ns; float bg[3]; int gradient; };
HostParams host_params(int scene) {
  switch (scene) {
    case 1: return {1200, 600, 10000, {0, 0, 0}, 0};                    // main.cu:654-707
    case 2: return {1200, 600, 500, {0, 0, 0}, 1};                      // main.cu:746-774
    case 3: return {1200, 600, 500, {0, 0, 0}, 1};                      // main.cu:802-847
    case 4: return {1200, 600, 500, {0, 0, 0}, 1};                      // main.cu:882-911
    case 5: return {1200, 600, 500, {0, 0, 0}, 1};                      // main.cu:939-967
    case 6: return {1200, 600, 10000, {0, 0, 0}, 0};                    // main.cu:995-1041
    case 7: return {600, 600, 10000, {0, 0, 0}, 0};                     // main.cu:1072-1101
    case 8: return {600, 600, 1000, {0, 0, 0}, 0};                      // main.cu:1129-1156
    case 9: return {800, 800, 10000, {0, 0, 0}, 0};                     // main.cu:1178-1208
    case 10: return {800, 800, 10000, {0.043f, 0.030f, 0.094f}, 0};     // main.cu:1239-1276
  }
  return {0, 0, 0, {0, 0, 0}, 0};
}

}  // namespace

std::string generate_scene(SceneDesc& sd, DevMath& dm, int scene_id, int nx, int ny, int grid_half,
                           const std::string& texture_dir) {
  HostParams hp = host_params(scene_id);
  if (hp.nx == 0) return "unknown scene id (expected 1..10, main.cu:1311-1320)";
  sd = SceneDesc();
  sd.scene_id = scene_id;
  sd.default_nx = hp.nx; sd.default_ny = hp.ny; sd.default_spp = hp.ns;
  sd.background[0] = hp.bg[0]; sd.background[1] = hp.bg[1]; sd.background[2] = hp.bg[2];
  sd.gradient_bg = hp.gradient;
  if (nx <= 0) nx = hp.nx;
  if (ny <= 0) ny = hp.ny;
  sd.nx = nx; sd.ny = ny;
  if (grid_half <= 0) grid_half = 11;
  SceneBuilder B(sd, dm);
  auto load = [&](const char* name, int& id) -> std::string {
    HostImage im;
    // The reference opens textures/<name>.jpg relative to the CWD (main.cu:1186, image_io.h:24-41): <name>.jpg in
    // texture_dir is decoded here (jpeg_baseline.cpp: the same bytes as the reference's decoder); a pre-decoded binary
    // PPM (P6) <name>.ppm is taken when there is no .jpg.
    const std::string dir = texture_dir.empty() ? std::string("textures") : texture_dir;
    std::string e = load_texture_file(dir + "/" + name + ".jpg", im);
    if (!e.empty()) {
      const std::string e2 = load_texture_file(dir + "/" + name + ".ppm", im);
      if (!e2.empty()) return "cannot load texture: " + e + "; " + e2;
    }
    id = B.add_image(im);
    return "";
  };
  int img = -1, img2 = -1;
  std::string err;
  switch (scene_id) {
    case 1: world_bouncing(B, nx, ny, grid_half); break;
    case 2: world_checker(B, nx, ny); break;
    case 3: if (!(err = load("earthmap", img)).empty()) return err; world_earth(B, nx, ny, img); break;
    case 4: world_perlin(B, nx, ny, 4.0f); break;
    case 5: world_quads(B, nx, ny); break;
    case 6: if (!(err = load("poolball", img)).empty()) return err; world_simple_light(B, nx, ny, img); break;
    case 7: world_cornell(B, nx, ny); break;
    case 8: world_cornell_smoke(B, nx, ny); break;
    case 9: if (!(err = load("earthmap", img)).empty()) return err; world_final(B, nx, ny, img); break;
    case 10: if (!(err = load("8ball", img2)).empty()) return err; world_original(B, nx, ny, img2); break;
  }
  return "";
}

}  // namespace rt
