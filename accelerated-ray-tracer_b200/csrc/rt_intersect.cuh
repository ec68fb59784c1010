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
// T-only form used by traversal: returns the accepted root in (tmin, tmax) exclusive, like the reference.
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
  if (t > tmin && t < tmax) { t_out = t; return true; }
  t = fdiv(fsub(sq, b), a);
  if (t > tmin && t < tmax) { t_out = t; return true; }
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
RT_D bool quad_hit(const DQuad& q, const Ray& r, float tmin, float tmax, float& t_out, float& alpha, float& beta) {
  V3 n = v3(q.nx, q.ny, q.nz);
  float denom = vdot(n, r.d);
  if (fabsf(denom) < 1e-8f) return false;
  float t = fdiv(fsub(q.D, vdot(n, r.o)), denom);
  if (t < tmin || t > tmax) return false;
  V3 P = vmad(t, r.d, r.o);
  V3 pl = vsub(P, v3(q.Qx, q.Qy, q.Qz));
  V3 w = v3(q.wx, q.wy, q.wz);
  float a = vdot(w, vcross(pl, v3(q.vx, q.vy, q.vz)));
  float b = vdot(w, vcross(v3(q.ux, q.uy, q.uz), pl));
  if (a < 0.f || a > 1.f || b < 0.f || b > 1.f) return false;
  t_out = t; alpha = a; beta = b;
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
RT_D bool box_hit(const DQuad* faces, const Ray& r, float tmin, float tmax, float& t_out, int& face, float& alpha,
                  float& beta) {
  bool any = false;
  float closest = tmax;
#pragma unroll 1
  for (int i = 0; i < 6; ++i) {
    float t, a, b;
    if (quad_hit(faces[i], r, tmin, closest, t, a, b)) { any = true; closest = t; face = i; alpha = a; beta = b; }
  }
  t_out = closest;
  return any;
}

// ---- instance wrappers + leaves: generic hit of a non-medium geometry ref ----
#define RT_MAX_XFORM 4
// known_face >= 0 (shading): the box face k_trace found; only that quad is intersected again.
template <bool FULL>
RT_D bool geom_hit(const DScene& S, uint32_t ref, Ray r, float tmin, float tmax, bool want_uv, Rec& rec, int known_face = -1) {
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
    if (!sphere_t(s, r, tmin, tmax, t, cc)) return false;
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
    } else if (!box_hit(S.quads + ix, r, tmin, tmax, t, face, a, b)) return false;
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
RT_D bool medium_hit(const DScene& S, const DMedium& m, const Ray& r, float tmin, float tmax, float& t_out) {
  Rec r1, r2;
  if (!geom_hit<false>(S, m.boundary, r, -FLT_MAX, FLT_MAX, false, r1)) return false;
  if (!geom_hit<false>(S, m.boundary, r, fadd(r1.t, 1e-4f), FLT_MAX, false, r2)) return false;
  float t1 = r1.t, t2 = r2.t;
  if (t1 < tmin) t1 = tmin;
  if (t2 > tmax) t2 = tmax;
  if (t1 >= t2) return false;
  if (t1 < 0) t1 = 0;
  const float ray_len = vlen(r.d);
  if (ray_len <= 0.0f || !isfinite(ray_len)) return false;
  const float distance_inside = fmul(fsub(t2, t1), ray_len);
  uint32_t seed = 1337u ^ f2u(r.o.x) ^ f2u(fmul(r.o.y, 3.1f)) ^ f2u(fmul(r.d.z, 5.7f));
  Xorwow fake;
  fake.init(seed);
  float U = fmaxf(1e-6f, fake.uniform());
  const float hit_distance = fmul(m.neg_inv_density, logf(U));
  if (hit_distance > distance_inside) return false;
  t_out = fadd(t1, fdiv(hit_distance, ray_len));
  return true;
}

// Top-level object test used by traversal: t only.
RT_D bool tlp_hit_t(const DScene& S, uint32_t ref, const Ray& r, float tmin, float tmax, float& t_out) {
  const uint32_t ty = ref_type(ref), ix = ref_index(ref);
  if (ty == G_SPHERE) {
    V3 cc;
    return sphere_t(S.spheres[ix], r, tmin, tmax, t_out, cc);
  } else if (ty == G_MEDIUM) {
    return medium_hit(S, S.media[ix], r, tmin, tmax, t_out);
  } else {
    Rec rec;
    if (!geom_hit<false>(S, ref, r, tmin, tmax, false, rec)) return false;
    t_out = rec.t;
    return true;
  }
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
#define RT_STACK 48
#define RT_LEAFQ 8         // pending leaves per lane
#ifndef RT_NODE_MIN
#define RT_NODE_MIN 12     // leaf phase starts when fewer lanes than this can still expand a node
#endif
struct Hit { float t; int tlp; int face; };
#ifdef RT_STATS  // diagnostics build only (tools/): per-ray work counters
__device__ unsigned long long g_stats[8];  // 0 node expansions, 1 sphere tests, 2 geom tests, 3 medium tests, 4 node phases, 5 leaf phases, 6 leaves queued, 7 leaves culled
#define RT_COUNT(i, n) atomicAdd(&g_stats[i], (unsigned long long)(n))
#else
#define RT_COUNT(i, n) do { } while (0)
#endif

RT_D void leaf_accept(const DScene& S, uint32_t ref, uint32_t tlp, float t, int face, Hit& best) {
  if (t < best.t || best.tlp < 0) { best.t = t; best.tlp = (int)tlp; best.face = face; return; }
  // exact tie (t == best.t; only inclusive tests get here): order semantics of bvh_node::hit
  const int rn = S.tlp[tlp].rank, rb = S.tlp[best.tlp].rank;
  const bool take = (rn > rb) ? ref_inclusive(S, ref) : !ref_inclusive(S, S.tlp[best.tlp].ref);
  if (take) { best.t = t; best.tlp = (int)tlp; best.face = face; }
}

#define RT_CSWAP(a, b) do { if (tn[b] < tn[a]) { float tf_ = tn[a]; tn[a] = tn[b]; tn[b] = tf_; \
  uint32_t tu_ = cr[a]; cr[a] = cr[b]; cr[b] = tu_; tu_ = ct[a]; ct[a] = ct[b]; ct[b] = tu_; } } while (0)

RT_D Hit closest_hit(const DScene& S, const Ray& r, bool active, float tmin, float tmax0, unsigned int* overflow) {
  Hit best; best.t = tmax0; best.tlp = -1; best.face = 0;
  const float ix = frcp(r.d.x), iy = frcp(r.d.y), iz = frcp(r.d.z);  // 1.0f / direction, aabb.cuh:48
  const bool nx = ix < 0.0f, ny = iy < 0.0f, nz = iz < 0.0f;
  uint32_t stack[RT_STACK];  // interior nodes only
  uint32_t lq_ref[RT_LEAFQ], lq_tlp[RT_LEAFQ];
  float lq_tn[RT_LEAFQ];
  int sp = 0, nl = 0;
  uint32_t cur = 0;
  bool have = active;
  while (true) {
    const bool can = have && nl <= RT_LEAFQ - 4;
    const unsigned mexp = __ballot_sync(0xFFFFFFFFu, can);
    const unsigned mleaf = __ballot_sync(0xFFFFFFFFu, nl > 0);
    if ((mexp | mleaf) == 0u) break;
    if (mleaf == 0u || __popc(mexp) >= RT_NODE_MIN) {
      // ---------------- node phase ----------------
      if ((threadIdx.x & 31) == 0) RT_COUNT(4, 1);
      if (can) {
        RT_COUNT(0, 1);
        const float4* np = reinterpret_cast<const float4*>(S.nodes + cur);
        const float4 lox = __ldg(np + 0), loy = __ldg(np + 1), loz = __ldg(np + 2);
        const float4 hix = __ldg(np + 3), hiy = __ldg(np + 4), hiz = __ldg(np + 5);
        const uint4 ch = __ldg(reinterpret_cast<const uint4*>(np + 6));
        const uint4 tl = __ldg(reinterpret_cast<const uint4*>(np + 7));
        const float lx[4] = {lox.x, lox.y, lox.z, lox.w}, ly[4] = {loy.x, loy.y, loy.z, loy.w}, lz[4] = {loz.x, loz.y, loz.z, loz.w};
        const float hx[4] = {hix.x, hix.y, hix.z, hix.w}, hy[4] = {hiy.x, hiy.y, hiy.z, hiy.w}, hz[4] = {hiz.x, hiz.y, hiz.z, hiz.w};
        float tn[4]; uint32_t cr[4] = {ch.x, ch.y, ch.z, ch.w}, ct[4] = {tl.x, tl.y, tl.z, tl.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // aabb::hit (aabb.cuh:45-61): t0 = (min - o) * invD, t1 = (max - o) * invD, swapped when invD < 0;
          // tmin = t0 > tmin ? t0 : tmin (== fmaxf, NaN keeps tmin); reject when tmax <= tmin.
          float t0 = fmul(fsub(nx ? hx[c] : lx[c], r.o.x), ix), t1 = fmul(fsub(nx ? lx[c] : hx[c], r.o.x), ix);
          float lo = fmaxf(t0, tmin), hi = fminf(t1, best.t);
          t0 = fmul(fsub(ny ? hy[c] : ly[c], r.o.y), iy); t1 = fmul(fsub(ny ? ly[c] : hy[c], r.o.y), iy);
          lo = fmaxf(t0, lo); hi = fminf(t1, hi);
          t0 = fmul(fsub(nz ? hz[c] : lz[c], r.o.z), iz); t1 = fmul(fsub(nz ? lz[c] : hz[c], r.o.z), iz);
          lo = fmaxf(t0, lo); hi = fminf(t1, hi);
          tn[c] = (hi > lo && cr[c] != RT_NODE_EMPTY) ? lo : FLT_MAX;  // FLT_MAX = not entered (a real entry is < best.t <= FLT_MAX)
          if (!(hi > lo)) cr[c] = RT_NODE_EMPTY;
        }
        RT_CSWAP(0, 1); RT_CSWAP(2, 3); RT_CSWAP(0, 2); RT_CSWAP(1, 3); RT_CSWAP(1, 2);  // nearest first
        // leaves -> pending queue; interior children: nearest is next, the others are stacked far-to-near
        uint32_t next = RT_NODE_EMPTY;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
          const uint32_t c = cr[k];
          if (c != RT_NODE_EMPTY && (c & RT_NODE_FLAG)) {
            if (next != RT_NODE_EMPTY) { if (sp < RT_STACK) stack[sp++] = next; else atomicOr(overflow, 1u); }
            next = c & 0x7FFFFFFFu;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t c = cr[k];
          if (c != RT_NODE_EMPTY && !(c & RT_NODE_FLAG)) { lq_ref[nl] = c; lq_tlp[nl] = ct[k]; lq_tn[nl] = tn[k]; ++nl; }
        }
        if (next != RT_NODE_EMPTY) cur = next;
        else if (sp > 0) cur = stack[--sp];
        else have = false;
      }
    } else {
      // ---------------- leaf phase ----------------
      if ((threadIdx.x & 31) == 0) RT_COUNT(5, 1);
      RT_COUNT(6, nl);
      const int nmax = __reduce_max_sync(0xFFFFFFFFu, nl);
      // spheres
      for (int k = 0; k < nmax; ++k) {
        if (k < nl && ref_type(lq_ref[k]) == G_SPHERE && lq_tn[k] < best.t) {
          float t; V3 cc;
          RT_COUNT(1, 1);
          if (sphere_t(S.spheres[ref_index(lq_ref[k])], r, tmin, best.t, t, cc)) leaf_accept(S, lq_ref[k], lq_tlp[k], t, 0, best);
        }
      }
      // quads, boxes, instances
      const unsigned mgeo = __ballot_sync(0xFFFFFFFFu, [&] { bool any = false; for (int k = 0; k < nl; ++k) { const uint32_t ty = ref_type(lq_ref[k]); any |= (ty != G_SPHERE && ty != G_MEDIUM); } return any; }());
      if (mgeo) {
        for (int k = 0; k < nmax; ++k) {
          if (k < nl) {
            const uint32_t ty = ref_type(lq_ref[k]);
            if (ty != G_SPHERE && ty != G_MEDIUM && lq_tn[k] < best.t) {
              Rec rec; rec.face = 0;
              RT_COUNT(2, 1);
              if (geom_hit<false>(S, lq_ref[k], r, tmin, best.t, false, rec)) leaf_accept(S, lq_ref[k], lq_tlp[k], rec.t, rec.face, best);
            }
          }
        }
      }
      // media
      const unsigned mmed = __ballot_sync(0xFFFFFFFFu, [&] { bool any = false; for (int k = 0; k < nl; ++k) any |= ref_type(lq_ref[k]) == G_MEDIUM; return any; }());
      if (mmed) {
        for (int k = 0; k < nmax; ++k) {
          if (k < nl && ref_type(lq_ref[k]) == G_MEDIUM && lq_tn[k] < best.t) {
            float t;
            RT_COUNT(3, 1);
            if (medium_hit(S, S.media[ref_index(lq_ref[k])], r, tmin, best.t, t)) leaf_accept(S, lq_ref[k], lq_tlp[k], t, 0, best);
          }
        }
      }
      nl = 0;
    }
  }
  return best;
}

}  // namespace rt
