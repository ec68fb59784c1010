// rt_shade.cuh — camera, textures and materials on the flattened scene (device only).
//
// Restated from camera.cuh:8-47, texture.cuh:16-164, perlin.cuh:6-82 and material.cuh:12-200 with
// the fusion pattern of the reference's sm_100 SASS (rt_math.h). All random draws go through the
// template parameter RNG (Xorwow in reference-RNG mode, Philox otherwise) in the reference's order.
#pragma once
#include "rt_intersect.cuh"

namespace rt {

// ---- camera ----
template <class RNG>
RT_D void random_in_unit_disk(RNG& g, float& x, float& y) {  // camera.cuh:8-16
  do {
    float a = g.uniform();  // evaluated left to right in the reference build
    float b = g.uniform();
    x = ffma(2.0f, a, -1.0f);
    y = ffma(2.0f, b, -1.0f);
  } while (ffma(x, x, fmul(y, y)) >= 1.0f);  // dot(p,p) with p.z = 0
}

// Philox mode draws the same DISTRIBUTION (uniform in the unit disk) without the rejection loop: one Philox block,
// no divergence. Only reference-RNG mode has to reproduce the reference's draw sequence.
#ifndef RT_PHILOX_REJECTION
template <>
RT_D void random_in_unit_disk<Philox>(Philox& g, float& x, float& y) {
  uint32_t o[4];
  g.next4(o);
  const float rad = sqrtf(u32_to_uniform(o[0]));
  float s, c;
  __sincosf(6.283185307f * u32_to_uniform(o[1]), &s, &c);
  x = rad * c; y = rad * s;
}
#endif

template <class RNG>
RT_D Ray camera_get_ray(const DCamera& c, float s, float t, RNG& g) {  // camera.cuh:35-47
  float px, py;
  random_in_unit_disk(g, px, py);
  const float rdx = fmul(c.lens_radius, px), rdy = fmul(c.lens_radius, py);
  // offset = u*rd.x + v*rd.y -> fma(rd.x, u, rd.y*v)
  V3 offset = v3(ffma(rdx, c.u.x, fmul(rdy, c.v.x)), ffma(rdx, c.u.y, fmul(rdy, c.v.y)), ffma(rdx, c.u.z, fmul(rdy, c.v.z)));
  const double tm = fma((double)g.uniform(), c.time1 - c.time0, c.time0);
  Ray r;
  r.o = vadd(c.origin, offset);
  V3 d = vmad(t, c.vertical, vmad(s, c.horizontal, c.llc));  // llc + s*horizontal + t*vertical
  r.d = vsub(vsub(d, c.origin), offset);
  r.tm = (float)tm;  // every consumer converts ray.time() to float (sphere.cuh:54)
  return r;
}

// ---- perlin (perlin.cuh) ----
RT_D uint32_t wanghash(uint32_t x) {
  x = (x ^ 61u) ^ (x >> 16); x *= 9u; x = x ^ (x >> 4); x *= 0x27d4eb2du; x = x ^ (x >> 15);
  return x;
}
RT_D uint32_t mix3(int x, int y, int z) { return (uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u; }
RT_D float u2m11(uint32_t h) { return ffma((float)((h >> 8) & 0x00FFFFFFu), 1.0f / 8388607.5f, -1.0f); }
RT_D V3 perlin_grad(int xi, int yi, int zi) {
  uint32_t h = wanghash(mix3(xi, yi, zi));
  V3 v = v3(u2m11(h), u2m11(wanghash(h)), u2m11(wanghash(h ^ 0x9e3779b9u)));
  return vunit(v);
}
RT_D float perlin_smooth(float t) { return fmul(fmul(t, t), fsub(3.0f, fadd(t, t))); }
RT_D float perlin_noise(V3 p) {  // perlin.cuh:52-70 + perlin_interp 34-50
  const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
  const float u = fsub(p.x, fx), v = fsub(p.y, fy), w = fsub(p.z, fz);
  const int i = (int)fx, j = (int)fy, k = (int)fz;
  const float uu = perlin_smooth(u), vv = perlin_smooth(v), ww = perlin_smooth(w);
  float accum = 0.0f;
#pragma unroll
  for (int di = 0; di < 2; ++di)
#pragma unroll
    for (int dj = 0; dj < 2; ++dj)
#pragma unroll
      for (int dk = 0; dk < 2; ++dk) {
        V3 c = perlin_grad(i + di, j + dj, k + dk);
        V3 wt = v3(fsub(u, (float)di), fsub(v, (float)dj), fsub(w, (float)dk));
        float s = fmul(fmul(di ? uu : fsub(1.0f, uu), dj ? vv : fsub(1.0f, vv)), dk ? ww : fsub(1.0f, ww));
        accum = ffma(s, vdot(c, wt), accum);
      }
  return accum;
}
RT_D float perlin_turb(V3 p, int depth) {  // perlin.cuh:72-82
  float accum = 0.0f, weight = 1.0f;
  V3 temp = p;
  for (int i = 0; i < depth; ++i) {
    accum = ffma(weight, perlin_noise(temp), accum);
    weight = fmul(weight, 0.5f);
    temp = vscale(2.0f, temp);
  }
  return fabsf(accum);
}

RT_D float clamp01(float x) { return x < 0 ? 0 : x > 1 ? 1 : x; }
RT_D float smoothstep_(float e0, float e1, float x) {
  float t = clamp01(fdiv(fsub(x, e0), fsub(e1, e0)));
  return fmul(fmul(t, t), fsub(3.0f, fmul(2.0f, t)));
}

// ---- textures (texture.cuh) ----
RT_D V3 texture_value(const DScene& S, int tex, float u, float v, V3 p) {
  // checker / uv_offset recurse into child textures: unrolled as a bounded loop
  for (int depth = 0; depth < 8; ++depth) {
    const DTex t = S.texs[tex];
    switch (t.kind) {
      case T_SOLID: return v3(t.cx, t.cy, t.cz);
      case T_CHECKER: {  // texture.cuh:35-42
        int xi = (int)floorf(fmul(t.scale, p.x));
        int yi = (int)floorf(fmul(t.scale, p.y));
        int zi = (int)floorf(fmul(t.scale, p.z));
        tex = (((xi + yi + zi) & 1) == 0) ? t.even : t.odd;
        break;
      }
      case T_UV_OFFSET: {  // texture.cuh:156-160
        float uu = fadd(u, t.p[0]); uu = fsub(uu, floorf(uu));
        float vv = fadd(v, t.p[1]); vv = fminf(fmaxf(vv, 0.f), 1.f);
        u = uu; v = vv; tex = t.even;
        break;
      }
      case T_IMAGE: {  // texture.cuh:51-59
        const DImage im = S.images[t.image];
        if (!(im.data && im.width > 0 && im.height > 0 && im.bpp >= 3)) return v3(0, 1, 1);
        u = clamp01(u); v = clamp01(v);
        int i = min((int)fmul(u, (float)im.width), im.width - 1);
        int j = min((int)fmul(fsub(1.f, v), (float)im.height), im.height - 1);
        int idx = (j * im.width + i) * im.bpp;
        const float inv255 = 1.f / 255.f;
        return v3(fmul(inv255, (float)im.data[idx + 0]), fmul(inv255, (float)im.data[idx + 1]), fmul(inv255, (float)im.data[idx + 2]));
      }
      case T_NOISE: {  // texture.cuh:67-72
        float s = __sinf(ffma(t.scale, p.z, fmul(10.0f, perlin_turb(p, 7))));
        float q = fmul(0.5f, fadd(1.0f, s));
        return v3(q, q, q);
      }
      case T_NOODLE: {  // texture.cuh:94-100
        V3 d = v3(t.p[4], t.p[5], t.p[6]);
        float uu = vdot(p, d);
        float wig = perlin_turb(vscale(t.p[2], p), (int)t.p[3]);
        float stripes = fabsf(__sinf(ffma(t.p[0], uu, fmul(t.p[1], wig))));
        float q = smoothstep_(0.75f, 0.98f, stripes);
        V3 cG = v3(t.p[10], t.p[11], t.p[12]), cN = v3(t.p[7], t.p[8], t.p[9]);
        float omq = fsub(1.f, q);
        // (1-t)*cG + t*cN: the reference SASS fuses the RIGHT product here: fma(t, cN, (1-t)*cG)
        return v3(ffma(q, cN.x, fmul(omq, cG.x)), ffma(q, cN.y, fmul(omq, cG.y)), ffma(q, cN.z, fmul(omq, cG.z)));
      }
      case T_FELT: {  // texture.cuh:125-147
        float m = perlin_noise(vscale(t.p[0], p));
        // p.x*f_scale + 2*turb: the reference SASS rounds the LEFT product and fuses the right one - FFMA(|turb|, 2,
        // p.x*f_scale) - unlike every other a*b + c*d site (rt_math.h); fusing the left one was the 1-ulp difference scene
        // 10 showed in 0.5 % of its pixels in round 1
        float phase = ffma(perlin_turb(vscale(0.5f, p), 2), 2.0f, fmul(p.x, t.p[2]));
        float fibers = fmul(0.5f, fadd(1.0f, __sinf(phase)));
        float gain = ffma(t.p[3], fsub(fibers, 0.5f), ffma(t.p[1], fsub(m, 0.5f), 1.0f));
        gain = fminf(fmaxf(gain, 0.7f), 1.2f);
        return vscale(gain, v3(t.cx, t.cy, t.cz));
      }
      default: return v3(0, 0, 0);
    }
  }
  return v3(0, 0, 0);
}

// ---- materials (material.cuh) ----
template <class RNG>
RT_D V3 random_in_unit_sphere(RNG& g) {  // material.cuh:12-18
  while (true) {
    float a = g.uniform(), b = g.uniform(), c = g.uniform();
    V3 p = v3(ffma(2.0f, a, -1.0f), ffma(2.0f, b, -1.0f), ffma(2.0f, c, -1.0f));
    if (vsqlen(p) < 1.0f) return p;
  }
}

// Philox mode: uniform in the unit ball from ONE Philox block, no rejection loop (the reference's loop runs 1.9
// iterations on average and was 39% of k_shade's instructions at 11.6 active lanes, profiles/r01d): radius = cbrt(u),
// direction uniform on the sphere. Same distribution, different draw sequence - which only reference-RNG mode must match.
#ifndef RT_PHILOX_REJECTION
template <>
RT_D V3 random_in_unit_sphere<Philox>(Philox& g) {
  uint32_t o[4];
  g.next4(o);
  const float z = ffma(-2.0f, u32_to_uniform(o[1]), 1.0f);
  const float rxy = sqrtf(fmaxf(0.0f, ffma(-z, z, 1.0f)));
  float s, c;
  __sincosf(6.283185307f * u32_to_uniform(o[2]), &s, &c);
  const float rad = cbrtf(u32_to_uniform(o[0]));
  return v3(rad * rxy * c, rad * rxy * s, rad * z);
}
#endif

// Returns false when the path ends (light, or metal scattering below the surface).
template <class RNG>
RT_D bool material_scatter(const DScene& S, const DMat& m, const Ray& rin, const Rec& rec, RNG& g, V3& attenuation, Ray& out) {
  out.tm = rin.tm;
  out.o = rec.p;
  switch (m.kind) {
    case M_LAMBERTIAN: {  // material.cuh:75-86: target = p + n + r ; dir = target - p
      V3 r = random_in_unit_sphere(g);
      V3 target = vadd(vadd(rec.p, rec.n), r);
      out.d = vsub(target, rec.p);
      attenuation = m.tex >= 0 ? texture_value(S, m.tex, rec.u, rec.v, rec.p) : v3(1, 1, 1);
      return true;
    }
    case M_METAL: {  // material.cuh:99-109
      V3 ud = vunit(rin.d);
      float two = vdot(ud, rec.n); two = fadd(two, two);
      V3 reflected = vnmad(two, rec.n, ud);  // v - 2*dot(v,n)*n
      V3 r = random_in_unit_sphere(g);
      out.d = vmad(m.param, r, reflected);
      attenuation = v3(m.ax, m.ay, m.az);
      return vdot(out.d, rec.n) > 0.0f;
    }
    case M_DIELECTRIC: {  // material.cuh:119-159
      const float ref_idx = m.param;
      const float dn = vdot(rin.d, rec.n);
      const float two = fadd(dn, dn);
      V3 reflected = vnmad(two, rec.n, rin.d);
      attenuation = v3(1.0f, 1.0f, 1.0f);
      V3 outward; float ni_over_nt, cosine;
      if (dn > 0.0f) {
        outward = vneg(rec.n);
        ni_over_nt = ref_idx;
        cosine = fdiv(dn, vlen(rin.d));
        cosine = fsqrt(fmaxf(0.0f, ffma(-fmul(ref_idx, ref_idx), ffma(-cosine, cosine, 1.0f), 1.0f)));
      } else {
        outward = rec.n;
        ni_over_nt = frcp(ref_idx);
        cosine = fdiv(-dn, vlen(rin.d));
      }
      // refract(), material.cuh:26-36
      V3 uv = vunit(rin.d);
      float dt = vdot(uv, outward);
      float disc = ffma(-fmul(ni_over_nt, ni_over_nt), ffma(-dt, dt, 1.0f), 1.0f);
      V3 refracted = v3(0, 0, 0);
      float reflect_prob;
      if (disc > 0.0f) {
        refracted = vnmad(fsqrt(disc), outward, vscale(ni_over_nt, vnmad(dt, outward, uv)));
        // schlick(), material.cuh:38-43
        float r0 = fdiv(fsub(1.0f, ref_idx), fadd(1.0f, ref_idx));
        r0 = fmul(r0, r0);
        reflect_prob = ffma(fsub(1.0f, r0), powf(fsub(1.0f, cosine), 5.0f), r0);
      } else {
        reflect_prob = 1.0f;
      }
      out.d = (g.uniform() < reflect_prob) ? reflected : refracted;
      return true;
    }
    case M_ISOTROPIC: {  // material.cuh:193-199
      out.d = random_in_unit_sphere(g);
      attenuation = texture_value(S, m.tex, rec.u, rec.v, rec.p);
      return true;
    }
    default:  // M_LIGHT: lights don't scatter (material.cuh:174-178)
      return false;
  }
}

RT_D V3 material_emitted(const DScene& S, const DMat& m, const Rec& rec) {  // material.cuh:168-172
  if (m.kind != M_LIGHT) return v3(0, 0, 0);
  return m.tex >= 0 ? texture_value(S, m.tex, rec.u, rec.v, rec.p) : v3(m.ax, m.ay, m.az);
}

}  // namespace rt
