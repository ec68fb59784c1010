/* rt_api.h — C ABI of the B200-native render path (librt_b200.so).
 *
 * The reference (slbouknight/accelerated-ray-tracer) has no FFI: its "API" is one host function per
 * scene that allocates, launches create_world_* / render_init / render and prints a PPM
 * (main.cu:654-1322). Each entry point below replaces one stage of those host functions; the
 * citation says which. Plain pointers and sizes only; the library owns all device memory, the
 * caller owns every host buffer. Functions return 0 on success, non-zero on error with the message
 * in rt_last_error() (the reference prints and exit(99)s instead, main.cu:23-35).
 *
 * Threading: one host thread drives one rt_scene; scenes are independent. The library never falls
 * back to a CPU renderer: without a CUDA device rt_build_scene fails.
 */
#ifndef RT_API_H
#define RT_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rt_scene rt_scene; /* opaque */

/* Which generator to run and at what resolution. Replaces the hard-coded `switch (10)` and the
 * per-scene `int nx=.., ny=..` (main.cu:1307-1322, 654-1305). */
typedef struct rt_scene_desc {
  int32_t scene_id;        /* 1..10 = main()'s cases: 1 bouncing, 2 checker, 3 earth, 4 perlin, 5 quads,
                              6 simple_light, 7 cornell, 8 cornell_smoke, 9 final, 10 original */
  int32_t nx, ny;          /* <= 0: the scene function's own resolution; nx/ny feeds the camera aspect */
  int32_t grid_half;       /* scene 1 only: GRID_MIN/MAX (main.cu:140-141); <= 0 means 11 (488 spheres) */
  int32_t device;          /* CUDA device ordinal; < 0: current device */
  const char* texture_dir; /* directory with the reference's <name>.jpg textures (or pre-decoded <name>.ppm, P6);
                              NULL = "textures" relative to the CWD like the reference (main.cu:1186) */
} rt_scene_desc;

/* Replaces the arguments of render<<<>>> and the per-scene constants (main.cu:107-109, 1179, 1208). */
typedef struct rt_render_params {
  int32_t spp;             /* ns; <= 0: the scene function's own value */
  int32_t max_depth;       /* <= 0: 50 (main.cu:54) */
  float gamma;             /* <= 0: 2.2 */
  float t_min;             /* <= 0: 0.001 (main.cu:57) */
  float background[3];     /* used when override_background != 0, else the scene function's */
  int32_t gradient_bg;
  int32_t override_background;
  uint64_t seed;           /* 0: 1984 (main.cu:92, 104) */
  int32_t rng_mode;        /* 0 = Philox4x32-10 (counter based), 1 = reference XORWOW streams, curand_init(1984+pixel,0,0) */
  int32_t split_mode;      /* 0 = tile split (rank owns scanlines j = rank mod world), 1 = spp split */
  int32_t rank, world;     /* this process's share; world <= 0 means 1 */
  int32_t slots;           /* Philox mode: path slots in flight; <= 0: automatic (32 Mi). Reference-RNG mode: one per pixel */
  int32_t aov;             /* != 0: also produce primary-hit object id / material id / t buffers */
  int32_t accumulate;      /* != 0: progressive pass: the linear sums of this call are ADDED to the accumulation buffer of the
                              previous call (same resolution and split); resolve with rt_resolve(total spp so far) */
  int32_t profile;         /* != 0: CUDA events around every k_trace / k_shade launch (rt_render_stats.trace_ms / shade_ms) */
} rt_render_params;

typedef struct rt_scene_info {
  int32_t scene_id, nx, ny;
  int32_t n_top, n_obj, n_mat, n_tex, n_img, n_bvh_nodes;
  int32_t default_nx, default_ny, default_spp, gradient_bg;
  float background[3];
  float bvh_build_ms;      /* device time of the BVH build */
  uint64_t h2d_bytes;      /* bytes uploaded by rt_build_scene */
} rt_scene_info;

typedef struct rt_render_stats {
  double device_ms;        /* cudaEvent span over all render kernels (k_start .. k_resolve) */
  uint64_t rays;           /* closest-hit queries = bounce-loop iterations of color() (main.cu:54-57) */
  uint64_t samples;        /* camera samples rendered by this rank */
  int32_t waves, kernel_launches;
  int32_t rows_local, nx;  /* shape of this rank's framebuffer share */
  int32_t nonfinite_samples; /* Philox mode: samples dropped because their radiance was inf/NaN (normally 0) */
  int32_t n_slots;
  uint32_t stack_overflow; /* must be 0 */
  int32_t profiled_waves;  /* waves covered by trace_ms / shade_ms (profile != 0) */
  double trace_ms, shade_ms; /* summed device time of the k_trace / k_shade launches */
} rt_render_stats;

/* Adaptive per-tile sampling (the step after the reference's fixed `for s < ns` loop, main.cu:119): passes of pass_spp
 * samples; after each pass every still-active tile's error is estimated from two half-buffers and tiles below `threshold`
 * stop. threshold <= 0: nothing ever stops (the result then equals rt_render(spp = max_spp) bit for bit). */
typedef struct rt_adaptive_params {
  int32_t min_spp;         /* no tile stops before this many samples per pixel (at least one pass) */
  int32_t max_spp;         /* <= 0: rt_render_params.spp, else the scene function's own value */
  int32_t pass_spp;        /* samples per pixel per pass (even; <= 0: 16) */
  int32_t tile;            /* tile edge in pixels (<= 0: 16) */
  float threshold;         /* mean over the tile of |I_even - I_odd|_1 / sqrt(|I|_1 + 1e-3), linear radiance */
  int32_t reserved[3];
} rt_adaptive_params;
typedef struct rt_adaptive_stats {
  uint64_t samples, rays;
  double device_ms;
  int32_t passes, tiles, tiles_converged;
  int32_t min_spp_used, max_spp_used;
  float mean_spp;
  float err_min, err_max;  /* smallest / largest tile error of the last estimate (over the tiles active at that time) */
  int32_t err_spp;         /* samples per pixel those tiles had when it was taken */
} rt_adaptive_stats;

/* create_world_*<<<1,1>>> + texture upload (main.cu:1186-1204): host generator -> H2D -> device BVH build. */
int rt_build_scene(const rt_scene_desc* desc, rt_scene** out);
/* The same from a caller-made scene: `sd` is a scene description in the flat format of rt_scene_desc.h (what the
 * reference's device-side `new sphere(...) / new quad(...) / ...; new bvh_node(d_list, 0, n)` would have built, e.g. the
 * output of rt_scene_export or of the SceneBuilder in csrc/scene_builder.h); image_pixels[i] = decoded 8-bit pixels of
 * image i (width*height*bpp bytes, may be NULL). Every index in the buffer is validated. The scene function's host
 * parameters (spp, background) are not part of an SD: pass them in rt_render_params (defaults: 10 spp, black). */
int rt_build_scene_sd(const void* sd, size_t sd_bytes, const unsigned char* const* image_pixels, int32_t n_images,
                      int32_t device, rt_scene** out);
/* Host-only check of a scene description (no GPU): the validation rt_build_scene_sd applies, then the resolution of
 * bvh_node groups (RT_OBJ_BVH) into instanced members. Writes the flattened scene (an SD without groups, every member
 * under a copy of the group's wrapper chain; *needed = its size, copied if cap suffices) and origin[k] = index in the
 * input's top-level list that flattened top-level entry k came from (up to origin_cap ints; *n_top_out = their number).
 * Returns 0 or 1 with the reason in rt_last_error(). */
int rt_sd_flatten(const void* sd, size_t sd_bytes, void* buf, size_t cap, size_t* needed, int32_t* origin, int32_t origin_cap,
                  int32_t* n_top_out);
/* render_init + render (main.cu:1207-1209). device_ms / rays may be NULL. */
int rt_render(rt_scene* s, const rt_render_params* p, double* device_ms, uint64_t* rays);
int rt_render_stats_get(rt_scene* s, rt_render_stats* out);
/* Counter-based RNG, tile split only. Leaves the framebuffer resolved (per-pixel mean, gamma) and the accumulation buffer
 * holding per-pixel MEANS; rt_readback_spp returns the samples each pixel of this rank's share received. */
int rt_render_adaptive(rt_scene* s, const rt_render_params* p, const rt_adaptive_params* a, rt_adaptive_stats* out);
int rt_readback_spp(rt_scene* s, int32_t* spp);
/* The managed-memory framebuffer read (main.cu:1212-1221). rgb: rows_local*nx*3 floats, gamma applied,
 * local row lr is image row j = lr*world + rank, j = 0 is the BOTTOM scanline like the reference's fb.
 * obj_id / mat_id / t (each rows_local*nx, nullable) need aov != 0 in the last rt_render. */
int rt_readback(rt_scene* s, float* rgb, int32_t* obj_id, int32_t* mat_id);
int rt_readback_t(rt_scene* s, float* t);
/* free_world + cudaFree (main.cu:1223-1236). */
void rt_destroy(rt_scene* s);
const char* rt_last_error(void);

int rt_scene_info_get(rt_scene* s, rt_scene_info* out);
/* The scene in the flat SD format of rt_scene_desc.h (for parity tests). *needed = size; copies if cap suffices.
 * rank[] (nullable, n_top ints): position of each top-level object in the reference BVH's leaf order. */
int rt_scene_export(rt_scene* s, void* buf, size_t cap, size_t* needed, int32_t* rank);
/* Same, generated without a GPU (sinf/cosf/tanf from the host libm instead of libdevice: fields derived
 * from them may differ from the GPU build in the last place). Tooling and CPU-side tests only. */
int rt_scene_export_host(const rt_scene_desc* desc, void* buf, size_t cap, size_t* needed, int32_t* rank, int32_t rank_cap);

/* Multi-GPU plumbing (spp split): the linear per-pixel radiance sums of this rank live in a device
 * buffer of rows_local*nx*3 floats that the caller may reduce across ranks (NCCL) before resolving. */
int rt_accum_device_ptr(rt_scene* s, void** dptr, size_t* n_floats);
int rt_resolve(rt_scene* s, int32_t total_spp, float gamma); /* accum -> framebuffer: /ns, gamma (main.cu:128-132) */
/* Spp split inside ONE process (one rt_scene per device, each rendered with split_mode = 1 and its own rank): adds the
 * accumulation buffers of src[0..n_src) into dst's on dst's device - the add kernel reads the peers' buffers directly over
 * NVLink when peer access is available, else through one cudaMemcpyPeerAsync each. Then rt_resolve(dst, total spp).
 * (Across processes the same sum is an NCCL reduce on rt_accum_device_ptr, see pyrt/dist.py.) */
int rt_accum_reduce(rt_scene* dst, rt_scene* const* src, int32_t n_src);
/* Dynamic tile queue inside ONE process: scenes[] = replicas of the same scene (normally one per device; several on one
 * device also work), each driven by its own host thread that pulls the next of n_chunks tile shares (interleaved
 * scanlines, n_chunks <= 0: 8 per replica) from a shared counter, renders it and copies its rows into rgb_full
 * (ny*nx*3 floats, host). Load-balances unequal GPUs / unequal rows; the image does not depend on who rendered what. */
#define RT_QUEUE_MAX_DEVICES 16
typedef struct rt_queue_stats {
  int32_t n_chunks;
  int32_t chunks_per_device[RT_QUEUE_MAX_DEVICES];
  double device_ms[RT_QUEUE_MAX_DEVICES];
  uint64_t rays_per_device[RT_QUEUE_MAX_DEVICES];
  uint64_t rays;
} rt_queue_stats;
int rt_render_queue(rt_scene* const* scenes, int32_t n_scenes, const rt_render_params* p, int32_t n_chunks, float* rgb_full,
                    rt_queue_stats* out);
/* Device address of this rank's framebuffer share (rows_local*nx*3 floats), e.g. for an NCCL gather of tiles. */
int rt_fb_device_ptr(rt_scene* s, void** dptr, size_t* n_floats);

/* rt_destroy keeps device blocks >= 1 MiB in a per-process free list for the next rt_build_scene / rt_render on the
 * same device (cudaMalloc/cudaFree of the ~0.4 GB path state cost more than a scene build); this releases them. */
void rt_trim_device_cache(void);

/* Host-only: decodes an image texture file the way rt_build_scene does - <name>.jpg (baseline JPEG; the bytes equal
 * what the reference's stbi_load(path, &w, &h, &n, 3) returns, image_io.h:24-41) or binary <name>.ppm - to 8-bit RGB.
 * *w / *h receive the size; pixels are copied if cap >= w*h*3 (call with rgb = NULL to query the size). */
int rt_load_texture(const char* path, unsigned char* rgb, size_t cap, int32_t* w, int32_t* h);

/* PPM writer of the scene functions (main.cu:1212-1221): "P3\n{nx} {ny}\n255\n", rows j = ny-1..0,
 * int(255.99f*c) per channel, no clamp. rgb is a FULL image (ny*nx*3, row 0 = bottom). double_scale != 0
 * reproduces bouncing_spheres' `int(255.99*c)` in double (main.cu:722-724). Returns bytes written or < 0. */
long rt_write_ppm(const char* path /* NULL = stdout */, const float* rgb, int32_t nx, int32_t ny, int32_t double_scale);
/* The optional outputs next to it: format 0 = that P3 text (clamp != 0 limits the integers to 0..255, which the reference
 * does not do), 1 = binary P6, 2 = PNG (8-bit RGB, stored deflate blocks; no external library). P6 / PNG always clamp. */
long rt_write_image(const char* path /* NULL = stdout */, const float* rgb, int32_t nx, int32_t ny, int32_t format, int32_t clamp,
                    int32_t double_scale);

#ifdef __cplusplus
}
#endif
#endif /* RT_API_H */
