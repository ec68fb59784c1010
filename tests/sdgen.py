"""tests/sdgen.py — random scenes in the flat SD format (include/rt_scene_desc.h), built in numpy from the reference's
scene vocabulary: sphere / moving sphere / quad / make_box / translate / rotate_y / constant_medium, the five materials,
solid / checker / noise textures, thin-lens camera. Derived fields follow the reference constructors (sphere.cuh:21-38,
quad.cuh:29-54, 94-162, hittable.cuh:52-54, 89-116, camera.cuh:59-78) in float32; the product and the oracle consume the
same bytes, so only self-consistency matters here, not bit-equality with what the reference's GPU constructors compute."""
import numpy as np
import pyrt

f32 = np.float32


class SDBuilder:
    def __init__(self, nx, ny):
        self.nx, self.ny = nx, ny
        self.tex, self.mat, self.obj, self.top = [], [], [], []
        self.cam = np.zeros((), dtype=pyrt.CAM_DT)

    # ---- textures / materials ----
    def solid(self, c):
        t = np.zeros((), dtype=pyrt.TEX_DT); t["kind"] = 0; t["even"] = t["odd"] = t["image"] = -1; t["color"] = c
        self.tex.append(t); return len(self.tex) - 1

    def checker(self, scale, even, odd):
        t = np.zeros((), dtype=pyrt.TEX_DT); t["kind"] = 1; t["even"] = even; t["odd"] = odd; t["image"] = -1; t["scale"] = f32(1.0) / f32(scale)
        self.tex.append(t); return len(self.tex) - 1

    def noise(self, scale):
        t = np.zeros((), dtype=pyrt.TEX_DT); t["kind"] = 3; t["even"] = t["odd"] = t["image"] = -1; t["scale"] = scale
        self.tex.append(t); return len(self.tex) - 1

    def _mat(self, kind, tex=-1, albedo=(0, 0, 0), param=0.0):
        m = np.zeros((), dtype=pyrt.MAT_DT); m["kind"] = kind; m["tex"] = tex; m["albedo"] = albedo; m["param"] = param
        self.mat.append(m); return len(self.mat) - 1

    def lambertian(self, tex): return self._mat(0, tex)
    def metal(self, albedo, fuzz): return self._mat(1, -1, albedo, min(fuzz, 1.0))
    def dielectric(self, ior): return self._mat(2, -1, (0, 0, 0), ior)
    def light(self, c): return self._mat(3, -1, c)
    def isotropic(self, tex): return self._mat(4, tex)

    # ---- hittables ----
    def _obj(self, kind):
        o = np.zeros((), dtype=pyrt.OBJ_DT); o["kind"] = kind; o["mat"] = -1; o["child"] = -1
        return o

    def _push(self, o):
        self.obj.append(o); return len(self.obj) - 1

    def sphere(self, c0, r, mat, c1=None):
        o = self._obj(0); c0 = np.asarray(c0, f32); o["c0"] = c0; o["radius"] = r; o["mat"] = mat
        rv = np.full(3, f32(r), f32)
        lo, hi = np.minimum(c0 - rv, c0 + rv), np.maximum(c0 - rv, c0 + rv)
        if c1 is not None:
            dc = np.asarray(c1, f32) - c0; o["dc"] = dc
            e = c0 + dc
            lo, hi = np.minimum(lo, np.minimum(e - rv, e + rv)), np.maximum(hi, np.maximum(e - rv, e + rv))
        o["box_min"], o["box_max"] = lo, hi
        return self._push(o)

    def quad(self, Q, u, v, mat, inward=False):
        o = self._obj(1); Q, u, v = (np.asarray(a, f32) for a in (Q, u, v))
        n = np.cross(u, v).astype(f32)
        normal = (n / f32(np.sqrt(f32(np.dot(n, n))))).astype(f32)
        if inward:
            normal = -normal
        o["Q"], o["u"], o["v"], o["n"], o["mat"], o["inward"] = Q, u, v, normal, mat, int(inward)
        o["D"] = f32(np.dot(normal, Q)); o["w"] = (n / f32(np.dot(n, n))).astype(f32)
        pts = np.stack([Q, Q + u + v, Q + u, Q + v])
        o["box_min"], o["box_max"] = pts.min(0) - f32(1e-3), pts.max(0) + f32(1e-3)
        return self._push(o)

    def box(self, a, b, mat):
        a, b = np.asarray(a, f32), np.asarray(b, f32)
        mn, mx = np.minimum(a, b), np.maximum(a, b)
        dx, dy, dz = np.array([mx[0] - mn[0], 0, 0], f32), np.array([0, mx[1] - mn[1], 0], f32), np.array([0, 0, mx[2] - mn[2]], f32)
        first = len(self.obj)
        faces = [((mn[0], mn[1], mx[2]), dx, dy), ((mx[0], mn[1], mx[2]), -dz, dy), ((mx[0], mn[1], mn[2]), -dx, dy),
                 ((mn[0], mn[1], mn[2]), dz, dy), ((mn[0], mx[1], mx[2]), dx, -dz), ((mn[0], mn[1], mn[2]), dx, dz)]
        for Q, u, v in faces:
            self.quad(Q, u, v, mat)
        o = self._obj(2); o["child"] = first; o["mat"] = mat
        o["box_min"] = np.min([self.obj[first + i]["box_min"] for i in range(6)], axis=0)
        o["box_max"] = np.max([self.obj[first + i]["box_max"] for i in range(6)], axis=0)
        return self._push(o)

    def translate(self, child, off):
        o = self._obj(3); o["child"] = child; o["offset"] = off
        o["box_min"] = self.obj[child]["box_min"] + np.asarray(off, f32); o["box_max"] = self.obj[child]["box_max"] + np.asarray(off, f32)
        return self._push(o)

    def rotate_y(self, child, deg):
        o = self._obj(4); o["child"] = child
        rad = f32(deg) * f32(0.017453292519943295769)
        s, c = f32(np.sin(rad)), f32(np.cos(rad)); o["sin_t"], o["cos_t"] = s, c
        b0, b1 = self.obj[child]["box_min"], self.obj[child]["box_max"]
        pts = []
        for x in (b0[0], b1[0]):
            for y in (b0[1], b1[1]):
                for z in (b0[2], b1[2]):
                    pts.append((c * x + s * z, y, -s * x + c * z))
        pts = np.asarray(pts, f32)
        o["box_min"], o["box_max"] = pts.min(0), pts.max(0)
        return self._push(o)

    def medium(self, boundary, density, tex):
        o = self._obj(5); o["child"] = boundary; o["mat"] = self.isotropic(tex); o["neg_inv_density"] = f32(-1.0) / f32(density)
        o["box_min"], o["box_max"] = self.obj[boundary]["box_min"], self.obj[boundary]["box_max"]
        return self._push(o)

    def with_material(self, child, mat):
        o = self._obj(6); o["child"] = child; o["mat"] = mat
        o["box_min"], o["box_max"] = self.obj[child]["box_min"], self.obj[child]["box_max"]
        return self._push(o)

    def bvh(self, members):
        """bvh_node(list, n) used as an object (bvh.cuh:29): a chain of cells, one per member; returns the head cell."""
        head = -1
        for m in members:
            o = self._obj(7); o["child"] = m; o["inward"] = head
            lo, hi = self.obj[m]["box_min"].copy(), self.obj[m]["box_max"].copy()
            if head >= 0:
                lo, hi = np.minimum(lo, self.obj[head]["box_min"]), np.maximum(hi, self.obj[head]["box_max"])
            o["box_min"], o["box_max"] = lo, hi
            head = self._push(o)
        return head

    def add(self, o): self.top.append(o)

    def camera(self, lookfrom, lookat, vup, vfov, aperture, focus, t0=0.0, t1=1.0):
        lf, la, vup = (np.asarray(a, f32) for a in (lookfrom, lookat, vup))
        aspect = f32(self.nx) / f32(self.ny)
        hh = f32(np.tan(f32(vfov) * f32(np.pi) / f32(180.0) * f32(0.5))); hw = aspect * hh
        unit = lambda a: (a / f32(np.sqrt(f32(np.dot(a, a))))).astype(f32)
        w = unit(lf - la); u = unit(np.cross(vup, w).astype(f32)); v = np.cross(w, u).astype(f32)
        c = self.cam
        c["origin"], c["u"], c["v"], c["w"] = lf, u, v, w
        c["lower_left_corner"] = lf - hw * f32(focus) * u - hh * f32(focus) * v - f32(focus) * w
        c["horizontal"] = f32(2.0) * hw * f32(focus) * u; c["vertical"] = f32(2.0) * hh * f32(focus) * v
        c["lens_radius"] = f32(aperture) * f32(0.5); c["time0"], c["time1"] = t0, t1

    def to_bytes(self):
        h = np.zeros((), dtype=pyrt.HDR_DT)
        h["magic"] = pyrt.SD_MAGIC; h["scene_id"] = 0; h["nx"], h["ny"] = self.nx, self.ny
        h["n_tex"], h["n_mat"], h["n_obj"], h["n_top"], h["n_img"] = len(self.tex), len(self.mat), len(self.obj), len(self.top), 0
        h["cam"] = self.cam
        return (h.tobytes() + b"".join(t.tobytes() for t in self.tex) + b"".join(m.tobytes() for m in self.mat) +
                b"".join(o.tobytes() for o in self.obj) + np.asarray(self.top, "<i4").tobytes())


def random_scene(seed, nx=160, ny=120, n_spheres=60, n_boxes=12, media=True, overrides=False, groups=False):
    """A room of random primitives that exercises every hittable / material branch, including the reference's odd
    corners: negative-radius spheres (hollow glass), moving spheres, nested translate(rotate_y(box)), media whose
    boundary is a sphere or an instanced box, objects that touch and overlap."""
    rng = np.random.default_rng(seed)
    B = SDBuilder(nx, ny)
    white, red = B.solid((0.73, 0.73, 0.73)), B.solid((0.65, 0.05, 0.05))
    chk = B.checker(0.8, B.solid((0.2, 0.3, 0.1)), B.solid((0.9, 0.9, 0.9)))
    mats = [B.lambertian(white), B.lambertian(red), B.lambertian(chk), B.metal((0.8, 0.8, 0.9), 0.0), B.metal((0.7, 0.6, 0.5), 0.4),
            B.dielectric(1.5), B.light((6, 6, 6)), B.lambertian(B.noise(1.3))]
    B.add(B.quad((-12, 0, -12), (24, 0, 0), (0, 0, 24), mats[2]))                 # floor
    B.add(B.quad((-12, 10, -12), (24, 0, 0), (0, 0, 24), mats[0], inward=True))   # ceiling
    B.add(B.quad((-4, 9.99, -4), (8, 0, 0), (0, 0, 8), mats[6]))                  # light
    B.add(B.quad((-12, 0, -12), (0, 10, 0), (0, 0, 24), mats[1]))                 # wall
    for _ in range(n_spheres):
        c = rng.uniform((-9, 0.3, -9), (9, 7, 9)); r = rng.uniform(0.2, 1.1); m = mats[rng.integers(len(mats))]
        kind = rng.integers(5)
        if kind == 0:
            B.add(B.sphere(c, r, m, c1=c + rng.uniform(-0.5, 0.5, 3)))            # moving
        elif kind == 1:
            B.add(B.sphere(c, r, mats[5])); B.add(B.sphere(c, -0.9 * r, mats[5]))  # hollow glass: negative radius
        else:
            B.add(B.sphere(c, r, m))
    for _ in range(n_boxes):
        a = rng.uniform(0.4, 2.0, 3); m = mats[rng.integers(5)]
        bx = B.box((0, 0, 0), a, m)
        kind = rng.integers(3)
        if kind >= 1:
            bx = B.rotate_y(bx, float(rng.uniform(-60, 60)))
        if kind == 2 or rng.random() < 0.8:
            bx = B.translate(bx, rng.uniform((-8, 0, -8), (8, 4, 8)))
        B.add(bx)
    if overrides:
        # with_material (hittable.cuh:154-178): shared geometry, per-instance look. One sphere and one box reused under
        # different materials; override of an instanced box, instance of an overridden box, override on top of an
        # override (the outermost wins), and an overridden medium (the phase function is replaced)
        ball = B.sphere((0, 0, 0), 0.8, mats[0])
        crate = B.box((0, 0, 0), (1.2, 1.8, 1.2), mats[1])
        for k in range(6):
            off = rng.uniform((-8, 0.8, -8), (8, 5, 8))
            B.add(B.translate(B.with_material(ball, mats[(3 + k) % len(mats)]), off))
        B.add(B.with_material(B.translate(B.rotate_y(crate, 30.0), (-3, 0, 5)), mats[3]))
        B.add(B.translate(B.rotate_y(B.with_material(crate, mats[7]), -20.0), (4, 0, 6)))
        B.add(B.with_material(B.with_material(B.translate(crate, (0, 0, 7)), mats[4]), mats[6]))
        if media:
            B.add(B.with_material(B.medium(B.sphere((-4, 2, -3), 1.5, mats[5]), 0.8, white), B.isotropic(red)))
    if groups:
        # bvh_node used as an object (bvh.cuh:29, RT_OBJ_BVH): a cluster in its own frame under translate(rotate_y(.)) - the
        # Book-2 final scene's sphere cluster as the book writes it -, the same group instanced twice, a group under a
        # material override, a group inside a group, members that are instances themselves, and a group straight in d_list
        cluster = B.bvh([B.sphere(rng.uniform(-1.5, 1.5, 3), 0.25, mats[rng.integers(len(mats))]) for _ in range(40)])
        B.add(B.translate(B.rotate_y(cluster, 15.0), (-5, 3, -4)))
        B.add(B.translate(cluster, (6, 5, -6)))
        pile = B.bvh([B.box((0, 0, 0), (0.6, 0.6, 0.6), mats[1]), B.translate(B.box((0, 0, 0), (0.5, 0.5, 0.5), mats[3]), (0.2, 0.6, 0.1)),
                      B.translate(B.rotate_y(B.box((0, 0, 0), (0.4, 0.4, 0.4), mats[4]), 40.0), (0.3, 1.1, 0.3))])
        B.add(B.with_material(B.translate(pile, (2, 0, 6)), mats[2]))
        B.add(B.translate(B.rotate_y(pile, -25.0), (-2, 0, 7)))
        nest = B.bvh([B.translate(cluster, (0, 2, 0)), B.sphere((0, 0, 0), 0.9, mats[5]), B.quad((-1, -1, 1.2), (2, 0, 0), (0, 2, 0), mats[0])])
        B.add(B.translate(nest, (-7, 2, 5)))
        B.add(B.bvh([B.sphere(rng.uniform((-9, 6, -9), (9, 9, 9)), 0.4, mats[6]) for _ in range(5)]))
    if media:
        B.add(B.medium(B.sphere((3, 3, 0), 2.0, mats[5]), 0.4, B.solid((0.2, 0.4, 0.9))))
        B.add(B.medium(B.translate(B.rotate_y(B.box((0, 0, 0), (2.5, 2.5, 2.5), mats[0]), 25.0), (-6, 0.5, 2)), 0.3, white))
        B.add(B.medium(B.sphere((0, 0, 0), 60.0, mats[5]), 0.004, white))       # fog around everything
    B.camera((0.5, 4.5, 15.5), (0, 3.5, 0), (0, 1, 0), 42.0, 0.15, 15.0)
    return B.to_bytes()
