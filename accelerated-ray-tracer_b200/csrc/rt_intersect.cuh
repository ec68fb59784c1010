// rt_intersect.cuh — primitive intersection on the flattened scene (device only).
//
// Each routine restates one reference hit function with the float expression trees and fusion the
// reference's sm_100 SASS shows (see rt_math.h), so that t, p and the normal are bit-identical:
//   sphere_hit   sphere.cuh:51-89      quad_hit  quad.cuh:60-90     box_hit   quad.cuh:124-139
//   xform chain  hittable.cuh:56-65, 118-145                        medium_hit constant_medium.cuh:36-76
// The polymorphic call chain of the reference (virtual hit -> virtual hit -> ...) becomes a switch on
// the 4-bit geometry tag; nesting (medium -> translate -> rotate_y -> box) is unrolled by a bounded loop.
#pragma once
#include "rt_scene_dev.h"
#include "rng.h"

namespace rt {

struct Ray { V3 o, d; float tm; };
struct Rec { float t; V3 p, n; float u, v; int face; };  // face: which of a box's six quads was hit

// ---- sphere (sphere.cuh:51-89) ----
// T-only form: returns the accepted root in (tmin, tmax) exclusive, like the reference. TIE = true (traversal only)
// also lets t == tmax through, so that leaf_accept can settle the exact tie by reference leaf order instead of by
// whichever object this traversal happened to meet first.
template <bool TIE = false>
RT_D bool sphere_t(const DSphere& s, const Ray& r, float tmin, float tmax, float& t_out, V3& cc) {
  cc = vmad(r.tm, v3(s.dx, s.dy, s.dz), v3(s.cx, s.cy, s.cz));  // center.point_at_parameter(time)
  V3 oc = vsub(r.o, cc);
  float a = vdot(r.d, r.d);
  float b = vdot(oc, r.d);
  float c = ffma(-s.radius, s.radius, vdot(oc, oc));  // dot(oc,oc) - r*r
  float disc = ffma(b, b, -fmul(a, c));               // b*b - a*c
  if (!(disc > 0.0f)) return false;
  float sq = fsqrt(disc);
  float t = fdiv(fsub(-b, sq), a);
  if (t > tmin && (TIE ? t <= tmax : t < tmax)) { t_out = t; return true; }
  t = fdiv(fsub(sq, b), a);
  if (t > tmin && (TIE ? t <= tmax : t < tmax)) { t_out = t; return true; }
  return false;
}
RT_D void sphere_uv(V3 n, float& u, float& v) {  // get_sphere_uv, sphere.cuh:42-49
  float theta = acosf(-n.y);
  float phi = fadd(atan2f(-n.z, n.x), 3.141592654f);
  u = fdiv(phi, 6.283185308f);  // 2 * CUDART_PI_F
  v = fdiv(theta, 3.141592654f);
}
RT_D void sphere_fill(const DSphere& s, const Ray& r, float t, V3 cc, bool want_uv, Rec& rec) {
  rec.t = t;
  rec.p = vmad(t, r.d, r.o);
  rec.n = vdivs(vsub(rec.p, cc), s.radius);
  rec.u = rec.v = 0.f;
  if (want_uv) sphere_uv(rec.n, rec.u, rec.v);
}

// ---- quad (quad.cuh:60-90); bounds inclusive: rejects t < tmin || t > tmax ----
// Split in two so that a box can find its candidate face from the six plane distances before it pays for an
// interior test: quad_plane = lines 61-64, quad_interior = lines 66-70 of the reference's quad::hit.
RT_D bool quad_plane(const float4 nD, const Ray& r, float tmin, float tmax, float& t_out) {
  const V3 n = v3(nD.x, nD.y, nD.z);
  const float denom = vdot(n, r.d);
  if (fabsf(denom) < 1e-8f) return false;
  const float t = fdiv(fsub(nD.w, vdot(n, r.o)), denom);
  if (t < tmin || t > tmax) return false;
  t_out = t;
  return true;
}
RT_D bool quad_interior(const DQuad& q, const Ray& r, float t, float& alpha, float& beta) {
  const V3 P = vmad(t, r.d, r.o);
  const V3 pl = vsub(P, v3(q.Qx, q.Qy, q.Qz));
  const V3 w = v3(q.wx, q.wy, q.wz);
  const float a = vdot(w, vcross(pl, v3(q.vx, q.vy, q.vz)));
  const float b = vdot(w, vcross(v3(q.ux, q.uy, q.uz), pl));
  if (a < 0.f || a > 1.f || b < 0.f || b > 1.f) return false;
  alpha = a; beta = b;
  return true;
}
RT_D bool quad_hit(const DQuad& q, const Ray& r, float tmin, float tmax, float& t_out, float& alpha, float& beta) {
  float t;
  if (!quad_plane(make_float4(q.nx, q.ny, q.nz, q.D), r, tmin, tmax, t)) return false;
  if (!quad_interior(q, r, t, alpha, beta)) return false;
  t_out = t;
  return true;
}
RT_D void quad_fill(const DQuad& q, const Ray& r, float t, float alpha, float beta, Rec& rec) {
  rec.t = t;
  rec.p = vmad(t, r.d, r.o);
  rec.u = alpha; rec.v = beta;
  V3 n = v3(q.nx, q.ny, q.nz);
  if (vdot(n, r.d) > 0.f) n = vneg(n);
  rec.n = n;
}

// ---- compound6 (quad.cuh:124-139): closest of six faces, later face wins ties ----
// The reference walks the faces in order with a shrinking t_max; since a face's interior test does not depend on
// t_max, its answer is the face with the smallest plane distance among those inside [tmin, tmax] whose interior
// test passes, the LATER face on an exact tie (the interval test is inclusive). Computed here as: six plane
// distances first (independent loads), then interior tests in ascending-t order until one passes - usually one
// interior test instead of six.
RT_D bool box_hit(const DQuad* faces, const float4* planes, const Ray& r, float tmin, float tmax, float& t_out, int& face, float& alpha,
                  float& beta) {
  float tf[6];
  unsigned valid = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 nD = __ldg(planes + i);  // the six (normal, D) of a box are contiguous: 96 B instead of six 80-B strides
    tf[i] = FLT_MAX;
    if (quad_plane(nD, r, tmin, tmax, tf[i])) valid |= 1u << i;
  }
  while (valid) {
    int bi = -1; float bt = FLT_MAX;
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if ((valid >> i & 1u) && tf[i] <= bt) { bt = tf[i]; bi = i; }  // <=: the later face wins a tie
    if (bi < 0) return false;  // only NaN distances left (a ray with a NaN direction)
    float a, b;
    if (quad_interior(faces[bi], r, bt, a, b)) { t_out = bt; face = bi; alpha = a; beta = b; return true; }
    valid &= ~(1u << bi);
  }
  return false;
}

// ---- instance wrappers + leaves: generic hit of a non-medium geometry ref ----
#define RT_MAX_XFORM 4
// known_face >= 0 (shading): the box face k_trace found; only that quad is intersected again.
// SC: anything with the geometry tables (spheres, quads, qplanes, xforms): DScene, or the GeomTab of pointers that the
// out-of-line copy used by the traversal takes (below).
template <bool FULL, bool TIE = false, class SC = DScene>
RT_D bool geom_hit(const SC& S, uint32_t ref, Ray r, float tmin, float tmax, bool want_uv, Rec& rec, int known_face = -1) {
  uint32_t chain[RT_MAX_XFORM];
  V3 dir_in[RT_MAX_XFORM];  // ray direction as each rotate_y saw it (for its normal flip)
  int n = 0;
  while (ref_type(ref) == G_XFORM && n < RT_MAX_XFORM) {
    const DXform x = S.xforms[ref_index(ref)];
    chain[n] = ref_index(ref);
    dir_in[n] = r.d;
    ++n;
    if (x.kind == X_TRANSLATE) {  // hittable.cuh:58
      r.o = vsub(r.o, v3(x.a, x.b, x.c));
    } else {  // rotate_y, hittable.cuh:120-127: a = sin, b = cos
      float ox = ffma(x.b, r.o.x, -fmul(x.a, r.o.z));
      float oz = ffma(x.a, r.o.x, fmul(x.b, r.o.z));
      float dx = ffma(x.b, r.d.x, -fmul(x.a, r.d.z));
      float dz = ffma(x.a, r.d.x, fmul(x.b, r.d.z));
      r.o = v3(ox, r.o.y, oz);
      r.d = v3(dx, r.d.y, dz);
    }
    ref = x.child;
  }
  const uint32_t ty = ref_type(ref), ix = ref_index(ref);
  if (ty == G_SPHERE) {
    const DSphere s = S.spheres[ix];
    float t; V3 cc;
    if (!sphere_t<TIE>(s, r, tmin, tmax, t, cc)) return false;
    rec.t = t;
    if (FULL) sphere_fill(s, r, t, cc, want_uv, rec);
  } else if (ty == G_QUAD) {
    const DQuad q = S.quads[ix];
    float t, a, b;
    if (!quad_hit(q, r, tmin, tmax, t, a, b)) return false;
    rec.t = t;
    if (FULL) quad_fill(q, r, t, a, b, rec);
  } else if (ty == G_BOX) {
    float t, a, b; int face = 0;
    if (FULL && known_face >= 0) {
      face = known_face;
      if (!quad_hit(S.quads[ix + face], r, tmin, tmax, t, a, b)) return false;
    } else if (!box_hit(S.quads + ix, S.qplanes + ix, r, tmin, tmax, t, face, a, b)) return false;
    rec.t = t; rec.face = face;
    if (FULL) quad_fill(S.quads[ix + face], r, t, a, b, rec);
  } else {
    return false;
  }
  if (FULL) {
    for (int k = n - 1; k >= 0; --k) {
      const DXform x = S.xforms[chain[k]];
      if (x.kind == X_TRANSLATE) {  // hittable.cuh:62
        rec.p = vadd(rec.p, v3(x.a, x.b, x.c));
      } else {  // hittable.cuh:130-142
        float px = ffma(x.b, rec.p.x, fmul(x.a, rec.p.z));
        float pz = ffma(x.b, rec.p.z, -fmul(x.a, rec.p.x));
        float nx = ffma(x.b, rec.n.x, fmul(x.a, rec.n.z));
        float nz = ffma(x.b, rec.n.z, -fmul(x.a, rec.n.x));
        rec.p = v3(px, rec.p.y, pz);
        rec.n = vunit(v3(nx, rec.n.y, nz));
        if (vdot(rec.n, dir_in[k]) > 0.f) rec.n = vneg(rec.n);
      }
    }
  }
  return true;
}

// ---- constant_medium (constant_medium.cuh:36-76), always entered through the 4-argument overload:
// the free-flight sample comes from a throw-away XORWOW seeded by a hash of the ray (:69-74), never
// from the pixel's stream.
// BP: the boundary's two queries of lines 43-47, (float& t1, float& t2) -> bool: t1 = first hit in (-FLT_MAX, FLT_MAX),
// t2 = first hit in (t1 + 1e-4, FLT_MAX).
// The reference asks the boundary first and draws the free-flight distance after; every exit before the draw returns
// false and the draw depends on the ray alone, so the order is free. Here the distance comes first, because it allows an
// EXACT early exit: t2' <= tmax and t1' >= tmin, rounding is monotone, so distance_inside <= (tmax - tmin) * |d| as
// computed below; a free-flight distance beyond that is rejected by line 66 whatever the boundary says. The thin global
// mist of the Book-2 final scene (density 1e-4 inside a sphere of radius 5000 that every ray is in) leaves here
// ~97% of the time once a surface has been found, without its two sphere queries.
template <class BP>
RT_D bool medium_hit_with(const DMedium& m, const Ray& r, float tmin, float tmax, float& t_out, BP boundary_pair) {
  const float ray_len = vlen(r.d);
  if (ray_len <= 0.0f || !isfinite(ray_len)) return false;
  uint32_t seed = 1337u ^ f2u(r.o.x) ^ f2u(fmul(r.o.y, 3.1f)) ^ f2u(fmul(r.d.z, 5.7f));
  Xorwow fake;
  fake.init(seed);
  float U = fmaxf(1e-6f, fake.uniform());
  const float hit_distance = fmul(m.neg_inv_density, logf(U));
#ifndef RT_NO_MEDIUM_EARLY_OUT
  if (hit_distance > fmul(fsub(tmax, tmin), ray_len)) return false;
#endif
  float t1, t2;
  if (!boundary_pair(t1, t2)) return false;
  if (t1 < tmin) t1 = tmin;
  if (t2 > tmax) t2 = tmax;
  if (t1 >= t2) return false;
  if (t1 < 0) t1 = 0;
  const float distance_inside = fmul(fsub(t2, t1), ray_len);
  if (hit_distance > distance_inside) return false;
  t_out = fadd(t1, fdiv(hit_distance, ray_len));
  return true;
}

// ONE out-of-line copy of the quad / box / instance-chain code for the traversal kernels: it is needed by the geometry
// pass of the leaf phase and twice by every medium test, and inlined three times it pushed k_trace to 50 KB of SASS
// against a 32 KB instruction cache. A DScene reference would have to be spilled to local memory for a real call; four
// pointers travel in registers. Returns (bit pattern of t, face) or (0xFFFFFFFF, 0). TIE semantics (t == tmax passes):
// the medium's boundary queries run with tmax = FLT_MAX, where that makes no difference.
struct GeomTab { const DSphere* spheres; const DQuad* quads; const float4* qplanes; const DXform* xforms; };
RT_D GeomTab geom_tab(const DScene& S) { GeomTab G; G.spheres = S.spheres; G.quads = S.quads; G.qplanes = S.qplanes; G.xforms = S.xforms; return G; }
#ifndef RT_GEOM_INLINE
__device__ __noinline__
#else
RT_D
#endif
uint2 geom_hit_call(GeomTab G, uint32_t ref, float ox, float oy, float oz, float dx, float dy, float dz, float tm, float tmin, float tmax) {
  Ray r; r.o = v3(ox, oy, oz); r.d = v3(dx, dy, dz); r.tm = tm;
  Rec rec; rec.face = 0;
  if (!geom_hit<false, true>(G, ref, r, tmin, tmax, false, rec)) return make_uint2(0xFFFFFFFFu, 0u);
  return make_uint2(f2u(rec.t), (uint32_t)rec.face);
}
RT_D bool medium_hit(const DScene& S, const DMedium& m, const Ray& r, float tmin, float tmax, float& t_out) {
  const GeomTab G = geom_tab(S);
  return medium_hit_with(m, r, tmin, tmax, t_out, [&](float& t1, float& t2) {
    if (ref_type(m.boundary) == G_SPHERE) {
      // A plain sphere boundary (both media of the Book-2 final scene): the two queries share everything up to the two
      // roots (sphere.cuh:51-89 evaluated twice on the same ray gives the same roots), so they are computed once.
      const DSphere s = S.spheres[ref_index(m.boundary)];
      const V3 cc = vmad(r.tm, v3(s.dx, s.dy, s.dz), v3(s.cx, s.cy, s.cz));
      const V3 oc = vsub(r.o, cc);
      const float a = vdot(r.d, r.d), b = vdot(oc, r.d), c = ffma(-s.radius, s.radius, vdot(oc, oc));
      const float disc = ffma(b, b, -fmul(a, c));
      if (!(disc > 0.0f)) return false;
      const float sq = fsqrt(disc);
      const float r0 = fdiv(fsub(-b, sq), a), r1 = fdiv(fsub(sq, b), a);
      if (r0 > -FLT_MAX && r0 < FLT_MAX) t1 = r0; else if (r1 > -FLT_MAX && r1 < FLT_MAX) t1 = r1; else return false;
      const float lo = fadd(t1, 1e-4f);
      if (r0 > lo && r0 < FLT_MAX) t2 = r0; else if (r1 > lo && r1 < FLT_MAX) t2 = r1; else return false;
      return true;
    }
    const uint2 h1 = geom_hit_call(G, m.boundary, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.tm, -FLT_MAX, FLT_MAX);
    if (h1.x == 0xFFFFFFFFu) return false;
    t1 = u2f(h1.x);
    const uint2 h2 = geom_hit_call(G, m.boundary, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.tm, fadd(t1, 1e-4f), FLT_MAX);
    if (h2.x == 0xFFFFFFFFu) return false;
    t2 = u2f(h2.x);
    return true;
  });
}

// Is the object's own interval test inclusive at tmax (quad: t > tmax rejects) or exclusive (sphere: t < tmax)?
RT_D bool ref_inclusive(const DScene& S, uint32_t ref) {
  while (ref_type(ref) == G_XFORM) ref = S.xforms[ref_index(ref)].child;
  const uint32_t ty = ref_type(ref);
  return ty == G_QUAD || ty == G_BOX;
}

// ---- closest hit over the 4-wide BVH ----------------------------------------------------------
// Semantics reproduced from bvh_node::hit (bvh.cuh:95-106): the winner is the object with the
// smallest t among those whose OWN box passes the reference slab test (aabb.cuh:45-61, here applied
// to the leaf child's box, which is the object's box bit for bit); interior boxes are exact unions,
// and the slab test is monotone in the box, so they never reject a ray a leaf box would accept.
// Exact-t ties are resolved as the reference's in-order leaf walk would (later leaf wins iff its own
// test is inclusive), using the precomputed leaf rank — so the result does not depend on the order
// in which this traversal meets the leaves.
//
// Execution model: the WHOLE WARP calls closest_hit (lanes without a ray pass active = false) and
// alternates between two warp-uniform phases, chosen by ballot:
//   node phase   lanes that hold a node expand it: 4 slab tests, sorting network, nearest interior child
//                becomes the next node, the others go to the lane's stack, leaf children go to the
//                lane's pending-leaf queue;
//   leaf phase   all lanes drain their queues together, one primitive class at a time (spheres, then
//                quads/boxes/instances, then media), so that a warp runs ONE intersection routine at a
//                time instead of interleaving node steps, sphere, box and medium code lane by lane.
// (First version: leaf tests inline in the node loop ran at 2-6 active lanes per instruction on the
// Book-2 final scene, profiles/r01.)
#ifndef RT_STACK
#define RT_STACK 48
#endif
#ifndef RT_LEAFQ
#define RT_LEAFQ 8         // pending leaves per lane
#endif
#ifndef RT_LEAF_WAIT_MAX
#define RT_LEAF_WAIT_MAX 6   // ... or when this many lanes have finished their traversal and only wait for their leaves (A/B: 4..8 +5%)
#endif
// Media wait for the END of the ray (deferred list, RT_MEDQ entries per lane): a constant_medium is the most expensive
// leaf by far (free-flight draw + two boundary queries), its result depends on t_max only through the clamp of the exit
// distance, and with the closest surface already known most of them leave through the exact early exit of
// medium_hit_with. A lane that collects more than RT_MEDQ - RT_LEAFQ of them makes the warp run a media phase early.
#define RT_MEDQ (RT_LEAFQ + 2)
#ifndef RT_NODE_MIN
#define RT_NODE_MIN 3      // leaf phase starts when fewer lanes than this can still expand a node (A/B on C4: 3 beats 1, 8, 12, 16)
#endif
struct Hit { float t; int tlp; };  // tlp: -1 (miss) or the PACKED hit word: object index | box face << 25 | shade class << 28
#ifdef RT_STATS  // diagnostics build only (tools/): per-ray work counters
__device__ unsigned long long g_stats[8];
__device__ unsigned long long g_stats2[4];  // per node phase: lanes finished, lanes blocked on a full leaf queue, lanes expanding  // 0 node expansions, 1 sphere tests, 2 geom tests, 3 medium tests, 4 node phases, 5 leaf phases, 6 leaves queued, 7 leaves culled
#define RT_COUNT(i, n) atomicAdd(&g_stats[i], (unsigned long long)(n))
#else
#define RT_COUNT(i, n) do { } while (0)
#endif

// The tlp word of a leaf child (BVH4Node::tlp, Hit::tlp) carries the shade-queue class of the object's material in
// its top bits, so that k_trace can bin a finished ray without another lookup: word = class << 28 | object index.
#define RT_TLP_MASK 0x01FFFFFFu   // 2^25 top-level objects (a packed hit is index | face << 25 | class << 28)
RT_D int tlp_index(int word) { return word & (int)RT_TLP_MASK; }
RT_D int tlp_class(int word) { return (word >> 28) & 7; }

// A pending leaf is remembered as its node and child slot (`entry` = node index << 2 | slot), and so is the best hit
// while a ray is traced (Best::e = entry | box face << 29); the child's tlp word is fetched from the node ONCE per ray,
// when the hit is final (Best::resolve), not with every node that is expanded or every hit that is accepted.
#define RT_ENTRY_MASK 0x1FFFFFFFu
#define RT_NO_ENTRY 0xFFFFFFFFu
RT_D uint32_t entry_tlp(const DScene& S, uint32_t entry) { return __ldg(&S.nodes[entry >> 2].tlp[entry & 3u]); }
RT_D uint32_t entry_ref(const DScene& S, uint32_t entry) { return __ldg(&S.nodes[entry >> 2].child[entry & 3u]); }
struct Best {
  float t; uint32_t e;
  RT_D bool none() const { return e == RT_NO_ENTRY; }
  RT_D Hit resolve(const DScene& S) const {
    Hit h; h.t = t;
    h.tlp = none() ? -1 : (int)(entry_tlp(S, e & RT_ENTRY_MASK) | ((e >> 29) << 25));
    return h;
  }
};
RT_D void leaf_accept(const DScene& S, uint32_t ref, uint32_t entry, float t, int face, Best& best) {
  const uint32_t e = entry | ((uint32_t)face << 29);
  if (t < best.t) { best.t = t; best.e = e; return; }
  if (best.none()) {  // t == the caller's t_max: only an inclusive test accepts that (quad.cuh:64 vs sphere.cuh:63)
    if (ref_inclusive(S, ref)) { best.t = t; best.e = e; }
    return;
  }
  // exact tie (t == best.t; only inclusive tests get here): order semantics of bvh_node::hit
  const uint32_t tn = entry_tlp(S, entry), tb = entry_tlp(S, best.e & RT_ENTRY_MASK);
  const int rn = S.tlp[tlp_index((int)tn)].rank, rb = S.tlp[tlp_index((int)tb)].rank;
  const bool take = (rn > rb) ? ref_inclusive(S, ref) : !ref_inclusive(S, S.tlp[tlp_index((int)tb)].ref);
  if (take) { best.t = t; best.e = e; }
}

// The per-lane arrays live OUTSIDE Trav (a struct with dynamically indexed arrays is kept in local memory as a whole:
// 84 LDL / 58 STL in k_trace instead of 20 / 15) and are handed to the phases by pointer.
// (lq_tlp / mq_tlp hold `entry` words: node index << 2 | child slot)
// A stack entry carries the sort key of its node (entry distance bits, low two bits = child slot) and is dropped at pop
// time when the closest hit found since the push is nearer than the node; the node a lane holds is looked at again after
// every leaf / media phase. The expansion of such a node enters no child (a child's box lies inside its parent's, so its
// lo is >= the parent's lo >= best.t >= its hi): skipping it changes no result and no ray count, only the number of node
// phases a warp runs. A/B against -DRT_NO_STACK_CULL (tools/gpu_cull.sh, same box): C4 4526 -> 4661 Mrays/s (+3.0 %),
// C5 at 10^6 spheres 3210 -> 3264 (+1.7 %), Cornell box 6295 -> 6226 (-1 %: three nodes, nothing to drop).
// Scenes whose BVH is a handful of nodes (both Cornell boxes: 3) have nothing to drop and pay for the check (-1 % / -3 %):
// they run the instantiation without it (TravT<false>, k_trace_small; RT_CULL_MIN_NODES in rt_host.cu).
#ifdef RT_NO_STACK_CULL
#define RT_CULL_DEFAULT false
#else
#define RT_CULL_DEFAULT true
#endif
template <bool CULL> struct StackEntT { typedef uint32_t type; };  // node index
template <> struct StackEntT<true> { typedef uint2 type; };        // (node index, key)
#define RT_TRAV_ARRAYS(name, CULL) typename StackEntT<CULL>::type name##_stack[RT_STACK]; uint32_t name##_lq_ref[RT_LEAFQ], name##_lq_tlp[RT_LEAFQ], name##_mq_tlp[RT_MEDQ]; \
  float name##_lq_tn[RT_LEAFQ], name##_mq_tn[RT_MEDQ]
#define RT_TRAV_ARGS(name) name##_stack, name##_lq_ref, name##_lq_tlp, name##_lq_tn

// (p - o) * inv for the four children of one plane vector: two FADD2 + two FMUL2 (add.rn.f32x2 / mul.rn.f32x2)
RT_D void slab2(const float4 p, float o, float inv, float* out) {
#if defined(__CUDA_ARCH__)
  unsigned long long p01, p23, no2, i2, d01, d23;
  const float no = -o;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p01) : "f"(p.x), "f"(p.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(p23) : "f"(p.z), "f"(p.w));
  asm("mov.b64 %0, {%1, %1};" : "=l"(no2) : "f"(no));
  asm("mov.b64 %0, {%1, %1};" : "=l"(i2) : "f"(inv));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(p01), "l"(no2));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(p23), "l"(no2));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(d01), "l"(i2));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(d23), "l"(i2));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(out[0]), "=f"(out[1]) : "l"(d01));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(out[2]), "=f"(out[3]) : "l"(d23));
#endif
}

// Per-lane traversal state. The warp-level driver (closest_hit below for one ray per lane; k_trace's loop, which also
// refills finished lanes with new rays) decides by ballot which phase runs next.
template <bool CULL>
struct TravT {
  typedef typename StackEntT<CULL>::type StackEnt;
  Ray r;
  float ix, iy, iz;      // 1.0f / direction, aabb.cuh:48
  float tmin;
  Best best;
  int sp, nl, nm;        // stack entries, pending leaves, deferred media
  uint32_t cur;          // node to expand next, RT_NODE_EMPTY: none
  uint32_t cur_key;      // CULL: its sort key (0 for the root)
  // lo with its low two bits cleared is <= lo: a node is dropped only when lo >= best.t for certain (positive floats order like their bits)
  RT_D bool live(uint32_t key) const { return (key & ~3u) < __float_as_uint(best.t); }
  RT_D void pop_live(const uint2* stack) {
    cur = RT_NODE_EMPTY;
    while (sp > 0) {
      const uint2 e = stack[--sp];
      if (live(e.y)) { cur = e.x; cur_key = e.y; break; }
    }
  }
  // after a phase that may have brought best.t down: the node chosen before it may be behind the hit now
  RT_D void revalidate(const StackEnt* stack) {
    if constexpr (CULL) { if (have() && !live(cur_key)) pop_live(stack); }
  }
  RT_D bool have() const { return cur != RT_NODE_EMPTY; }

  RT_D void reset() { nl = 0; nm = 0; sp = 0; cur = RT_NODE_EMPTY; best.t = FLT_MAX; best.e = RT_NO_ENTRY; }
  RT_D void begin(const Ray& ray, float tmin_, float tmax0) {
    r = ray; tmin = tmin_;
    best.t = tmax0; best.e = RT_NO_ENTRY;
    ix = frcp(r.d.x); iy = frcp(r.d.y); iz = frcp(r.d.z);
    sp = 0; nl = 0; nm = 0; cur = 0;
    if constexpr (CULL) cur_key = 0u;
  }
  RT_D bool can_expand() const { return have() && nl <= RT_LEAFQ - 4; }
  RT_D bool finished() const { return !have() && nl == 0; }

  // ---------------- node phase (lanes with can_expand()) ----------------
  RT_D void node_step(const DScene& S, unsigned int* overflow, StackEnt* stack, uint32_t* lq_ref, uint32_t* lq_tlp, float* lq_tn) {
    RT_COUNT(0, 1);
    const float4* np = reinterpret_cast<const float4*>(S.nodes + cur);
    // aabb::hit (aabb.cuh:45-61): t0 = (min - o) * invD, t1 = (max - o) * invD, swapped when invD < 0. The swap is
    // done by the LOAD: per-ray offsets pick the near / far plane vectors of the node, no per-child selects.
    // float4 index of the NEAR plane vector of each axis inside a node (lox 0, loy 1, loz 2, hix 3, hiy 4, hiz 5)
    const int onx = ix < 0.0f ? 3 : 0, ony = iy < 0.0f ? 4 : 1, onz = iz < 0.0f ? 5 : 2;
    const float4 nxp = __ldg(np + onx), fxp = __ldg(np + (3 - onx));
    const float4 nyp = __ldg(np + ony), fyp = __ldg(np + (5 - ony));
    const float4 nzp = __ldg(np + onz), fzp = __ldg(np + (7 - onz));
    const uint4 ch = __ldg(reinterpret_cast<const uint4*>(np + 6));
#ifdef RT_NO_PACKED_SLAB
    const float ax[4] = {nxp.x, nxp.y, nxp.z, nxp.w}, bx[4] = {fxp.x, fxp.y, fxp.z, fxp.w};
    const float ay[4] = {nyp.x, nyp.y, nyp.z, nyp.w}, by[4] = {fyp.x, fyp.y, fyp.z, fyp.w};
    const float az[4] = {nzp.x, nzp.y, nzp.z, nzp.w}, bz[4] = {fzp.x, fzp.y, fzp.z, fzp.w};
#endif
    const uint32_t cr[4] = {ch.x, ch.y, ch.z, ch.w};
    uint32_t key[4];  // interior children that are entered: (entry distance bits, child slot); else 0xFFFFFFFF
#ifndef RT_NO_PACKED_SLAB
    // The 48 subtractions and multiplications of the four slab tests as 24 packed f32x2 instructions (sm_100 FADD2 / FMUL2:
    // two IEEE round-to-nearest operations per issue slot, same bits as the scalar forms; the kernel is issue-bound).
    float tn_[3][4], tf_[3][4];
    slab2(nxp, r.o.x, ix, tn_[0]); slab2(fxp, r.o.x, ix, tf_[0]);
    slab2(nyp, r.o.y, iy, tn_[1]); slab2(fyp, r.o.y, iy, tf_[1]);
    slab2(nzp, r.o.z, iz, tn_[2]); slab2(fzp, r.o.z, iz, tf_[2]);
#endif
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      // tmin = t0 > tmin ? t0 : tmin (== fmaxf, NaN keeps tmin); reject when tmax <= tmin.
#ifndef RT_NO_PACKED_SLAB
      float lo = fmaxf(tn_[0][c], tmin), hi = fminf(tf_[0][c], best.t);
      lo = fmaxf(tn_[1][c], lo); hi = fminf(tf_[1][c], hi);
      lo = fmaxf(tn_[2][c], lo); hi = fminf(tf_[2][c], hi);
#else
      float lo = fmaxf(fmul(fsub(ax[c], r.o.x), ix), tmin), hi = fminf(fmul(fsub(bx[c], r.o.x), ix), best.t);
      lo = fmaxf(fmul(fsub(ay[c], r.o.y), iy), lo); hi = fminf(fmul(fsub(by[c], r.o.y), iy), hi);
      lo = fmaxf(fmul(fsub(az[c], r.o.z), iz), lo); hi = fminf(fmul(fsub(bz[c], r.o.z), iz), hi);
#endif
      const bool entered = hi > lo && cr[c] != RT_NODE_EMPTY;
      const bool interior = (cr[c] & RT_NODE_FLAG) != 0;
      // lo >= tmin > 0, so its bit pattern orders like the float; the low two mantissa bits carry the child slot
      key[c] = (entered && interior) ? ((__float_as_uint(lo) & ~3u) | (uint32_t)c) : 0xFFFFFFFFu;
      if (entered && !interior) { lq_ref[nl] = cr[c]; lq_tlp[nl] = (cur << 2) | (uint32_t)c; lq_tn[nl] = lo; ++nl; }  // leaves need no order
    }
    // interior children: nearest is the next node, the others are stacked far-to-near (5-comparator network on keys)
#define RT_KSWAP(a, b) do { const uint32_t lo_ = min(key[a], key[b]), hi_ = max(key[a], key[b]); key[a] = lo_; key[b] = hi_; } while (0)
    RT_KSWAP(0, 1); RT_KSWAP(2, 3); RT_KSWAP(0, 2); RT_KSWAP(1, 3); RT_KSWAP(1, 2);
#undef RT_KSWAP
    if (sp > RT_STACK - 3) { atomicOr(overflow, 1u); sp = RT_STACK - 3; }  // rt_render fails when the flag is set
#pragma unroll
    for (int k = 3; k >= 1; --k) {
      if (key[k] != 0xFFFFFFFFu) {
        const uint32_t i = key[k] & 3u;
        const uint32_t c = (i & 2u) ? ((i & 1u) ? cr[3] : cr[2]) : ((i & 1u) ? cr[1] : cr[0]);
        if constexpr (CULL) stack[sp++] = make_uint2(c & 0x7FFFFFFFu, key[k]);
        else stack[sp++] = c & 0x7FFFFFFFu;
      }
    }
    uint32_t next = RT_NODE_EMPTY;
    if (key[0] != 0xFFFFFFFFu) {
      const uint32_t i = key[0] & 3u;
      next = ((i & 2u) ? ((i & 1u) ? cr[3] : cr[2]) : ((i & 1u) ? cr[1] : cr[0])) & 0x7FFFFFFFu;
    }
    cur = next;
    if constexpr (CULL) {
      cur_key = key[0];
      if (next == RT_NODE_EMPTY) pop_live(stack);
    } else {
      if (next == RT_NODE_EMPTY && sp > 0) cur = stack[--sp];
    }
  }

  // ---------------- leaf phase (the whole warp) ----------------
  // Returns (warp-uniform) whether a media phase has to run before the next leaf phase (a deferred list is nearly full).
  RT_D bool leaf_phase(const DScene& S, const uint32_t* lq_ref, const uint32_t* lq_tlp, const float* lq_tn, uint32_t* mq_tlp, float* mq_tn) {
    RT_COUNT(6, nl);
    const int nmax = __reduce_max_sync(0xFFFFFFFFu, nl);
    // spheres
    for (int k = 0; k < nmax; ++k) {
      if (k < nl && ref_type(lq_ref[k]) == G_SPHERE && lq_tn[k] < best.t) {
        float t; V3 cc;
        RT_COUNT(1, 1);
        if (sphere_t<true>(S.spheres[ref_index(lq_ref[k])], r, tmin, best.t, t, cc)) leaf_accept(S, lq_ref[k], lq_tlp[k], t, 0, best);
      }
    }
    // quads, boxes, instances: every lane walks ITS geometry leaves with its own cursor, so that the j-th box test of
    // all lanes runs in the same iteration (a shared position index left 6 of 32 lanes active in the box code).
    // The same walk moves the media it passes to the lane's deferred list.
    {
      int kk = 0;
      while (true) {
        while (kk < nl) {
          const uint32_t ty = ref_type(lq_ref[kk]);
          if (ty == G_MEDIUM) { mq_tlp[nm] = lq_tlp[kk]; mq_tn[nm] = lq_tn[kk]; ++nm; }
          else if (ty != G_SPHERE) break;
          ++kk;
        }
        const bool has = kk < nl;
        if (__ballot_sync(0xFFFFFFFFu, has) == 0u) break;
        if (has && lq_tn[kk] < best.t) {
          RT_COUNT(2, 1);
          const uint2 h = geom_hit_call(geom_tab(S), lq_ref[kk], r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z, r.tm, tmin, best.t);
          if (h.x != 0xFFFFFFFFu) leaf_accept(S, lq_ref[kk], lq_tlp[kk], u2f(h.x), (int)h.y, best);
        }
        ++kk;
      }
    }
    nl = 0;
    return __any_sync(0xFFFFFFFFu, nm > RT_MEDQ - RT_LEAFQ);
  }

  // ---------------- media phase (the whole warp; lanes with active = true evaluate and clear their deferred list) ----------------
  RT_D void media_phase(const DScene& S, const uint32_t* mq_tlp, const float* mq_tn, bool active) {
    const int n = active ? nm : 0;
    const int nmax = __reduce_max_sync(0xFFFFFFFFu, n);
    for (int k = 0; k < nmax; ++k) {
      if (k < n && mq_tn[k] < best.t) {
        const uint32_t entry = mq_tlp[k];
        const uint32_t ref = entry_ref(S, entry);
        float t;
        RT_COUNT(3, 1);
        if (medium_hit(S, S.media[ref_index(ref)], r, tmin, best.t, t)) leaf_accept(S, ref, entry, t, 0, best);
      }
    }
    if (active) nm = 0;
  }

  // Which phase next? Node phase unless too few lanes can expand a node or too many only wait for their leaves.
  RT_D static bool pick_node_phase(unsigned mexp, unsigned mleaf, unsigned mwait) {
    return mleaf == 0u || (__popc(mexp) >= RT_NODE_MIN && __popc(mwait) < RT_LEAF_WAIT_MAX);
  }
};
typedef TravT<RT_CULL_DEFAULT> Trav;

// One ray per lane, no refill (k_aov, k_finish): the whole warp calls it, lanes without a ray pass active = false.
RT_D Hit closest_hit(const DScene& S, const Ray& r, bool active, float tmin, float tmax0, unsigned int* overflow) {
  Trav T;
  RT_TRAV_ARRAYS(m, RT_CULL_DEFAULT);
  T.reset();
  if (active) T.begin(r, tmin, tmax0);
  bool flush = false;  // warp-uniform: a deferred media list is nearly full
  while (true) {
    const bool can = T.can_expand();
    const unsigned mexp = __ballot_sync(0xFFFFFFFFu, can);
    const unsigned mleaf = __ballot_sync(0xFFFFFFFFu, T.nl > 0);
    const bool done = (mexp | mleaf) == 0u;
    if (done || flush) {
      if (__any_sync(0xFFFFFFFFu, T.nm > 0)) T.media_phase(S, m_mq_tlp, m_mq_tn, true);
      flush = false;
      if (done) break;
      T.revalidate(m_stack);
      continue;
    }
    const unsigned mwait = __ballot_sync(0xFFFFFFFFu, !T.have() && T.nl > 0);  // traversal done, leaves pending
    if (Trav::pick_node_phase(mexp, mleaf, mwait)) {
      if ((threadIdx.x & 31) == 0) RT_COUNT(4, 1);
      if (can) T.node_step(S, overflow, RT_TRAV_ARGS(m));
    } else {
      if ((threadIdx.x & 31) == 0) RT_COUNT(5, 1);
      flush = T.leaf_phase(S, m_lq_ref, m_lq_tlp, m_lq_tn, m_mq_tlp, m_mq_tn);
      T.revalidate(m_stack);
    }
  }
  return T.best.resolve(S);
}

}  // namespace rt
