/* rt_scene_desc.h — flat, POD "scene description" (SD) shared by the product, the oracle and the
 * reference harnesses. It is a wire format, not logic: plain arrays of fixed-size records that say
 * what the reference's device-side object graph contains after a create_world_* kernel ran
 * (/root/reference/src/main.cu:160-635).
 *
 * Vocabulary follows the reference (sphere.cuh, quad.cuh, hittable.cuh, constant_medium.cuh,
 * material.cuh, texture.cuh, camera.cuh). Every float is the value the reference's constructor
 * stores, so two SDs can be compared bit for bit.
 *
 * Binary file layout (little endian), used by rt_scene_export / the harness dumps:
 *   rt_sd_header | rt_texture_desc[n_tex] | rt_material_desc[n_mat] | rt_object_desc[n_obj]
 *   | int32 top[n_top] | rt_image_desc[n_img]     (image pixels are NOT in the file)
 */
#ifndef RT_SCENE_DESC_H
#define RT_SCENE_DESC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_SD_MAGIC 0x31445352u /* "RSD1" */

/* texture.cuh:16-164 */
enum rt_tex_kind {
  RT_TEX_SOLID = 0,    /* solid_color(albedo)                     texture.cuh:16-23  */
  RT_TEX_CHECKER = 1,  /* checker_texture(scale, even, odd)       texture.cuh:25-43  */
  RT_TEX_IMAGE = 2,    /* image_texture(DeviceImage)              texture.cuh:45-60  */
  RT_TEX_NOISE = 3,    /* noise_texture(scale) (marble)           texture.cuh:62-76  */
  RT_TEX_NOODLE = 4,   /* noodle_texture(...)                     texture.cuh:84-103 */
  RT_TEX_FELT = 5,     /* felt_texture(...)                       texture.cuh:109-148 */
  RT_TEX_UV_OFFSET = 6 /* uv_offset_texture(base, du, dv)         texture.cuh:151-164 */
};

typedef struct rt_texture_desc {
  int32_t kind;
  int32_t even, odd;   /* checker children; uv_offset: even = base texture id */
  int32_t image;       /* image id (RT_TEX_IMAGE), else -1 */
  float color[3];      /* solid: albedo; felt: base_col */
  float scale;         /* checker: inv_scale (=1/scale as stored); noise: scale */
  /* noodle: k, A, f, octaves(as float), d[3], cN[3], cG[3]  -> p[0..12]
   * felt  : m_scale, m_amt, f_scale, f_amt                  -> p[0..3]
   * uv_offset: du, dv                                       -> p[0..1] */
  float p[13];
  int32_t pad_;
} rt_texture_desc;

/* material.cuh:62-201 */
enum rt_mat_kind {
  RT_MAT_LAMBERTIAN = 0,
  RT_MAT_METAL = 1,
  RT_MAT_DIELECTRIC = 2,
  RT_MAT_DIFFUSE_LIGHT = 3,
  RT_MAT_ISOTROPIC = 4
};

typedef struct rt_material_desc {
  int32_t kind;
  int32_t tex;      /* lambertian/isotropic: texture id; diffuse_light: texture id or -1 (solid) */
  float albedo[3];  /* metal: albedo; diffuse_light(tex==-1): solid emission */
  float param;      /* metal: fuzz (already clamped to <=1); dielectric: ref_idx */
  int32_t pad_[2];
} rt_material_desc;

/* hittable vocabulary */
enum rt_obj_kind {
  RT_OBJ_SPHERE = 0,    /* sphere.cuh:21-38 (static: dc == 0)                       */
  RT_OBJ_QUAD = 1,      /* quad.cuh:29-41                                          */
  RT_OBJ_BOX = 2,       /* compound6 from make_box, quad.cuh:94-162; child = first of 6 quads */
  RT_OBJ_TRANSLATE = 3, /* hittable.cuh:40-69                                      */
  RT_OBJ_ROTATE_Y = 4,  /* hittable.cuh:77-149                                     */
  RT_OBJ_MEDIUM = 5,    /* constant_medium.cuh:17-79; child = boundary, mat = phase function */
  RT_OBJ_WITH_MATERIAL = 6, /* with_material(obj, mat), hittable.cuh:154-178: child's geometry and box, every hit reports
                              `mat` (the OUTERMOST override wins: each wrapper sets rec.mat_ptr after its child's hit) */
  RT_OBJ_BVH = 7        /* a bvh_node used as an OBJECT (bvh.cuh:29: it is a hittable, so it can sit under translate /
                           rotate_y / with_material or in d_list): the closest hit over its members. The member list is a
                           chain of cells: child = one member, inward = id of the next cell of the same group (-1: last);
                           the group is named by its head cell, whose box is the union of all member boxes. */
};

typedef struct rt_object_desc {
  int32_t kind;
  int32_t mat;        /* sphere/quad: material; box: material of face 0; medium: isotropic phase material;
                         with_material: the override */
  int32_t child;      /* wrapper: wrapped object; box: id of face 0 (faces are child..child+5) */
  int32_t inward;     /* quad: inward flag; bvh cell: next cell of the group or -1 */
  float c0[3];        /* sphere: center.A */
  float dc[3];        /* sphere: center.B (= c1 - c0; 0 for static) */
  float radius;       /* sphere (may be negative) */
  float Q[3], u[3], v[3], w[3], n[3]; /* quad */
  float D;            /* quad plane constant */
  float offset[3];    /* translate */
  float sin_t, cos_t; /* rotate_y */
  float neg_inv_density; /* medium */
  float box_min[3], box_max[3]; /* bounding_box() */
} rt_object_desc;

typedef struct rt_image_desc {
  int32_t width, height, bpp;
  int32_t pad_;
} rt_image_desc;

/* camera.cuh:49-57 public members after init() */
typedef struct rt_camera_desc {
  float origin[3];
  float lower_left_corner[3];
  float horizontal[3];
  float vertical[3];
  float u[3], v[3], w[3];
  float lens_radius;
  double time0, time1;
} rt_camera_desc;

typedef struct rt_sd_header {
  uint32_t magic;
  int32_t scene_id;
  int32_t nx, ny;
  int32_t n_tex, n_mat, n_obj, n_top, n_img;
  int32_t pad_;
  rt_camera_desc cam;
} rt_sd_header;

#ifdef __cplusplus
}
#endif
#endif /* RT_SCENE_DESC_H */
