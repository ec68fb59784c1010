// rt_cli.cpp — command-line front end of librt_b200.so: replaces the reference's main() and its ten
// hard-wired scene functions (main.cu:654-1322). Like the reference it writes an ASCII P3 PPM to
// stdout and diagnostics to stderr, and exits 99 on a CUDA/library error (main.cu:23-35). Unlike the
// reference the scene, resolution, spp and GPU count are arguments (the reference picks the scene with
// a hard-coded `switch (10)`, main.cu:1309).
//
//   rt_cli --scene 9 [--nx 800 --ny 800] [--spp 10000] [--rng philox|reference] [--gpus N] [--split tile|spp]
//          [--textures DIR] [--out file] [--format p3|p6|png] [--clamp] [--depth 50] [--seed 1984] [--grid-half G]
//
// --gpus N: the job is split over N devices of this box from ONE process, one host thread and one scene replica per GPU:
//   --split tile (default)  interleaved scanlines, the shares are assembled on the host; bit-identical to N = 1;
//   --split spp             every GPU renders all pixels for 1/N of the sample numbers; the linear-radiance buffers are
//                           summed on GPU 0 over peer memory (rt_accum_reduce) and resolved there (Philox mode only);
//   --split dynamic         tile shares (--chunks C of them, default 8 per GPU) handed out at run time to whichever GPU is
//                           free (rt_render_queue); bit-identical to N = 1.
// --format: p3 = the reference's text PPM (default), p6 = binary PPM, png; p6/png clamp to 0..255, --clamp does it for p3.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "rt_api.h"

struct Share { int rc = 0; std::string err; std::vector<float> fb; rt_render_stats st{}; rt_scene_info info{}; rt_scene* sc = nullptr; };

int main(int argc, char** argv) {
  rt_scene_desc d; memset(&d, 0, sizeof(d));
  rt_render_params p; memset(&p, 0, sizeof(p));
  d.scene_id = 10;  // what the reference's main() runs
  d.device = -1;
  int gpus = 1, format = 0, clamp = 0, spp_split = 0, dynamic = 0, chunks = 0;
  std::string tex = "textures", out;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
    if (a == "--scene") d.scene_id = atoi(val());
    else if (a == "--nx") d.nx = atoi(val());
    else if (a == "--ny") d.ny = atoi(val());
    else if (a == "--grid-half") d.grid_half = atoi(val());
    else if (a == "--spp") p.spp = atoi(val());
    else if (a == "--depth") p.max_depth = atoi(val());
    else if (a == "--seed") p.seed = strtoull(val(), nullptr, 10);
    else if (a == "--rng") { std::string v = val(); p.rng_mode = (v == "reference" || v == "ref" || v == "xorwow") ? 1 : 0; }
    else if (a == "--gpus") gpus = atoi(val());
    else if (a == "--split") {
      std::string v = val();
      if (v != "tile" && v != "spp" && v != "dynamic") { fprintf(stderr, "--split tile|spp|dynamic\n"); return 2; }
      spp_split = v == "spp"; dynamic = v == "dynamic";
    } else if (a == "--chunks") chunks = atoi(val());
    else if (a == "--textures") tex = val();
    else if (a == "--out") out = val();
    else if (a == "--format") {
      std::string v = val();
      if (v == "p3") format = 0; else if (v == "p6") format = 1; else if (v == "png") format = 2;
      else { fprintf(stderr, "--format p3|p6|png\n"); return 2; }
    } else if (a == "--clamp") clamp = 1;
    else if (a == "--help" || a == "-h") {
      fprintf(stderr, "usage: rt_cli --scene 1..10 [--nx N --ny N] [--spp N] [--rng philox|reference] [--gpus N] [--split tile|spp]\n"
                      "              [--textures DIR] [--out FILE] [--format p3|p6|png] [--clamp] [--depth N] [--seed N] [--grid-half N]\n");
      return 0;
    } else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
  }
  if (gpus < 1) gpus = 1;
  if (gpus == 1) spp_split = 0;
  d.texture_dir = tex.c_str();
  if (dynamic) {
    // one scene replica per GPU, then the queue: every replica's host thread pulls tile shares until none is left
    std::vector<rt_scene*> sc(gpus, nullptr);
    int rc = 0;
    for (int r = 0; r < gpus && !rc; ++r) {
      rt_scene_desc dd = d; dd.device = gpus > 1 ? r : d.device;
      if (rt_build_scene(&dd, &sc[r])) { fprintf(stderr, "rt_cli: GPU %d: %s\n", r, rt_last_error()); rc = 99; }
    }
    rt_scene_info info{}; rt_queue_stats qs{};
    std::vector<float> img;
    if (!rc) {
      rt_scene_info_get(sc[0], &info);
      img.assign((size_t)info.nx * info.ny * 3, 0.f);
      if (rt_render_queue(sc.data(), gpus, &p, chunks, img.data(), &qs)) { fprintf(stderr, "rt_cli: %s\n", rt_last_error()); rc = 99; }
    }
    for (auto* s : sc) if (s) rt_destroy(s);
    if (rc) return rc;
    double ms = 0; for (int r = 0; r < gpus; ++r) if (qs.device_ms[r] > ms) ms = qs.device_ms[r];
    fprintf(stderr, "Rendering a %dx%d image, scene %d, %d GPU(s) (dynamic queue, %d shares:", info.nx, info.ny, d.scene_id, gpus, qs.n_chunks);
    for (int r = 0; r < gpus; ++r) fprintf(stderr, " %d", qs.chunks_per_device[r]);
    fprintf(stderr, "): %.3f ms on the busiest device, %llu rays, %.1f Mrays/s\n", ms, (unsigned long long)qs.rays, ms > 0 ? qs.rays / ms / 1e3 : 0.0);
    if (rt_write_image(out.empty() ? nullptr : out.c_str(), img.data(), info.nx, info.ny, format, clamp, d.scene_id == 1 ? 1 : 0) < 0) {
      fprintf(stderr, "rt_cli: %s\n", rt_last_error());
      return 99;
    }
    return 0;
  }
  std::vector<Share> sh(gpus);
  std::vector<std::thread> th;
  for (int r = 0; r < gpus; ++r)
    th.emplace_back([&, r]() {
      Share& S = sh[r];
      rt_scene_desc dd = d; dd.device = gpus > 1 ? r : d.device;
      rt_render_params pp = p; pp.rank = r; pp.world = gpus; pp.split_mode = spp_split;
      if ((S.rc = rt_build_scene(&dd, &S.sc)) != 0) { S.err = rt_last_error(); return; }
      rt_scene_info_get(S.sc, &S.info);
      if ((S.rc = rt_render(S.sc, &pp, nullptr, nullptr)) != 0) { S.err = rt_last_error(); return; }
      rt_render_stats_get(S.sc, &S.st);
      if (!spp_split) {
        S.fb.resize((size_t)S.st.rows_local * S.st.nx * 3);
        if ((S.rc = rt_readback(S.sc, S.fb.data(), nullptr, nullptr)) != 0) S.err = rt_last_error();
      }
    });
  for (auto& t : th) t.join();
  int rc = 0;
  for (int r = 0; r < gpus; ++r)
    if (sh[r].rc) { fprintf(stderr, "rt_cli: GPU %d: %s\n", r, sh[r].err.c_str()); rc = 99; }
  const int nx = sh[0].info.nx, ny = sh[0].info.ny;
  std::vector<float> img(rc ? 0 : (size_t)nx * ny * 3);
  unsigned long long rays = 0; double ms = 0;
  if (!rc) {
    for (int r = 0; r < gpus; ++r) { rays += sh[r].st.rays; if (sh[r].st.device_ms > ms) ms = sh[r].st.device_ms; }
    if (spp_split) {
      // sum of the per-GPU linear-radiance buffers on GPU 0, then / total spp and gamma there
      std::vector<rt_scene*> others;
      for (int r = 1; r < gpus; ++r) others.push_back(sh[r].sc);
      const int total_spp = p.spp > 0 ? p.spp : sh[0].info.default_spp;
      if (rt_accum_reduce(sh[0].sc, others.data(), (int)others.size()) || rt_resolve(sh[0].sc, total_spp, 0.f) ||
          rt_readback(sh[0].sc, img.data(), nullptr, nullptr)) {
        fprintf(stderr, "rt_cli: %s\n", rt_last_error());
        rc = 99;
      }
    } else {
      for (int r = 0; r < gpus; ++r)
        for (int lr = 0; lr < sh[r].st.rows_local; ++lr)
          memcpy(&img[(size_t)(lr * gpus + r) * nx * 3], &sh[r].fb[(size_t)lr * nx * 3], (size_t)nx * 3 * sizeof(float));
    }
  }
  for (int r = 0; r < gpus; ++r) if (sh[r].sc) rt_destroy(sh[r].sc);
  if (rc) return rc;
  fprintf(stderr, "Rendering a %dx%d image, scene %d, %d GPU(s)%s: %.3f ms on the device, %llu rays, %.1f Mrays/s\n", nx, ny,
          d.scene_id, gpus, gpus > 1 ? (spp_split ? " (spp split)" : " (tile split)") : "", ms, rays, ms > 0 ? rays / ms / 1e3 : 0.0);
  // bouncing_spheres alone scales with a double 255.99 (main.cu:722-724)
  if (rt_write_image(out.empty() ? nullptr : out.c_str(), img.data(), nx, ny, format, clamp, d.scene_id == 1 ? 1 : 0) < 0) {
    fprintf(stderr, "rt_cli: %s\n", rt_last_error());
    return 99;
  }
  return 0;
}
