// rt_scene_dev.h — the flattened scene as it lives in HBM (host + device view).
//
// The reference keeps a heap-scattered graph of polymorphic objects (hittable.cuh, material.cuh,
// texture.cuh) and a recursive-pointer binary BVH (bvh.cuh). Here everything is tagged-union POD in
// 16-byte-aligned arrays:
//
//   spheres[]  32 B   quads[]  80 B (box = 6 consecutive quads)   xforms[] 32 B   media[] 16 B
//   mats[]     32 B   texs[]   96 B                               images[] 24 B
//   tlp[]      16 B   one per top-level object (d_list entry): geometry ref, material, kind, rank
//   nodes[]   128 B   4-wide BVH over the top-level objects, child boxes in SoA, built on the device
//
// A geometry "ref" is type<<28 | index.
#pragma once
#include <stdint.h>
#include "rt_math.h"

namespace rt {

enum GeomType : uint32_t { G_SPHERE = 0, G_QUAD = 1, G_BOX = 2, G_XFORM = 3, G_MEDIUM = 4 };
RT_HD uint32_t make_ref(uint32_t type, uint32_t idx) { return (type << 28) | idx; }
RT_HD uint32_t ref_type(uint32_t r) { return r >> 28; }
RT_HD uint32_t ref_index(uint32_t r) { return r & 0x0FFFFFFFu; }

struct alignas(16) DSphere {  // sphere.cuh:95-100
  float cx, cy, cz, radius;   // center.A, radius
  float dx, dy, dz;           // center.B (= c1 - c0; 0 for a static sphere)
  int mat;
};
struct alignas(16) DQuad {    // quad.cuh:14-21; plane (normal, D) first: a box test reads the six planes before anything else
  float nx, ny, nz, D;
  float Qx, Qy, Qz; int mat;
  float ux, uy, uz, pad0;
  float vx, vy, vz, pad1;
  float wx, wy, wz, pad2;
};
enum XformKind : int { X_TRANSLATE = 0, X_ROTATE_Y = 1 };
struct alignas(16) DXform {   // hittable.cuh:40-149
  int kind; uint32_t child;
  float a, b, c;              // translate: offset.xyz ; rotate_y: a = sin_t, b = cos_t
  float pad[3];
};
struct alignas(16) DMedium {  // constant_medium.cuh:17-22
  uint32_t boundary; float neg_inv_density; int mat; int pad;
};
enum MatKind : int { M_LAMBERTIAN = 0, M_METAL = 1, M_DIELECTRIC = 2, M_LIGHT = 3, M_ISOTROPIC = 4 };
struct alignas(16) DMat {
  int kind; int tex;
  float ax, ay, az;           // metal albedo / light solid colour
  float param;                // fuzz / ref_idx
  int needs_uv;               // texture chain reads the sphere (u, v): image / uv_offset (acosf/atan2f skipped otherwise)
  int pad;
};
enum TexKind : int { T_SOLID = 0, T_CHECKER = 1, T_IMAGE = 2, T_NOISE = 3, T_NOODLE = 4, T_FELT = 5, T_UV_OFFSET = 6 };
struct alignas(16) DTex {
  int kind; int even, odd, image;
  float cx, cy, cz, scale;
  float p[13]; float pad[3];
};
struct DImage {
  const unsigned char* data; int width, height, bpp, pad;
};

// Shade-queue classes: what the trace kernel sorts paths by.
// Q_PROCEDURAL: lambertian / isotropic whose texture chain evaluates Perlin noise (marble, noodle, felt): ~20x the
// instructions of any other shade, so those paths get warps of their own.
enum QueueId : int { Q_MISS = 0, Q_LIGHT = 1, Q_LAMBERTIAN = 2, Q_METAL = 3, Q_DIELECTRIC = 4, Q_ISOTROPIC = 5, Q_PROCEDURAL = 6, Q_COUNT = 7 };

struct alignas(16) DTlp {     // one per top-level object
  uint32_t ref;               // geometry ref
  int mat;                    // material id (medium: phase function)
  int queue;                  // QueueId of that material
  int rank;                   // position in the reference BVH's leaf order (tie-breaks only)
};

// 4-wide BVH node, 128 B = one L1 line. Child i's box is (lox[i], loy[i], loz[i])-(hix[i], ...).
// child[i]: 0x80000000|node index (internal), geometry ref (leaf), 0xFFFFFFFF (empty, box inverted).
// tlp[i]: top-level object index for leaf children.
#define RT_NODE_FLAG 0x80000000u
#define RT_NODE_EMPTY 0xFFFFFFFFu
struct alignas(16) BVH4Node {
  float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
  uint32_t child[4];
  uint32_t tlp[4];
};

struct DCamera {              // camera.cuh:49-57
  V3 origin, llc, horizontal, vertical, u, v;
  float lens_radius;
  double time0, time1;
};

struct DScene {
  const DSphere* spheres; const DQuad* quads; const DXform* xforms; const DMedium* media;
  const DMat* mats; const DTex* texs; const DImage* images;
  const DTlp* tlp; const BVH4Node* nodes;
  const float4* qplanes;      // (normal, D) of quads[i], contiguous: a box test reads its six planes from 96 consecutive bytes
  int n_tlp, n_nodes;
  DCamera cam;
};

}  // namespace rt
