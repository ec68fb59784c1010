"""pyrt — thin ctypes mirror of include/rt_api.h (librt_b200.so).

Host-side plumbing only: every pixel is produced by the CUDA kernels behind the C ABI. There is no
Python or CPU rendering path; if the shared library or a CUDA device is missing, calls fail loudly.

The object model mirrors the reference's host scene functions (main.cu:654-1305):
    Scene(scene_id, nx, ny)   ~ create_world_*<<<1,1>>> + texture upload
    Scene.render(spp=...)     ~ render_init + render
    Scene.framebuffer()       ~ reading `fb` back;  write_ppm() ~ the P3 loop (main.cu:1212-1221)
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "librt_b200.so")  # RT_LIB: diagnostics builds
REPO_ROOT = os.path.dirname(os.path.dirname(_HERE))

SCENE_NAMES = {1: "bouncing", 2: "checker", 3: "earth", 4: "perlin", 5: "quads", 6: "simple_light",
               7: "cornell", 8: "cornell_smoke", 9: "final", 10: "original"}


class RtError(RuntimeError):
    pass


class SceneDescC(C.Structure):
    _fields_ = [("scene_id", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("grid_half", C.c_int32),
                ("device", C.c_int32), ("texture_dir", C.c_char_p)]


class RenderParamsC(C.Structure):
    _fields_ = [("spp", C.c_int32), ("max_depth", C.c_int32), ("gamma", C.c_float), ("t_min", C.c_float),
                ("background", C.c_float * 3), ("gradient_bg", C.c_int32), ("override_background", C.c_int32),
                ("seed", C.c_uint64), ("rng_mode", C.c_int32), ("split_mode", C.c_int32), ("rank", C.c_int32),
                ("world", C.c_int32), ("slots", C.c_int32), ("aov", C.c_int32), ("accumulate", C.c_int32),
                ("profile", C.c_int32)]


class SceneInfoC(C.Structure):
    _fields_ = [("scene_id", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("n_top", C.c_int32),
                ("n_obj", C.c_int32), ("n_mat", C.c_int32), ("n_tex", C.c_int32), ("n_img", C.c_int32),
                ("n_bvh_nodes", C.c_int32), ("default_nx", C.c_int32), ("default_ny", C.c_int32),
                ("default_spp", C.c_int32), ("gradient_bg", C.c_int32), ("background", C.c_float * 3),
                ("bvh_build_ms", C.c_float), ("h2d_bytes", C.c_uint64)]


class RenderStatsC(C.Structure):
    _fields_ = [("device_ms", C.c_double), ("rays", C.c_uint64), ("samples", C.c_uint64), ("waves", C.c_int32),
                ("kernel_launches", C.c_int32), ("rows_local", C.c_int32), ("nx", C.c_int32),
                ("nonfinite_samples", C.c_int32), ("n_slots", C.c_int32), ("stack_overflow", C.c_uint32),
                ("profiled_waves", C.c_int32), ("trace_ms", C.c_double), ("shade_ms", C.c_double)]


class AdaptiveParamsC(C.Structure):
    _fields_ = [("min_spp", C.c_int32), ("max_spp", C.c_int32), ("pass_spp", C.c_int32), ("tile", C.c_int32),
                ("threshold", C.c_float), ("reserved", C.c_int32 * 3)]


class AdaptiveStatsC(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("device_ms", C.c_double), ("passes", C.c_int32),
                ("tiles", C.c_int32), ("tiles_converged", C.c_int32), ("min_spp_used", C.c_int32), ("max_spp_used", C.c_int32),
                ("mean_spp", C.c_float), ("err_min", C.c_float), ("err_max", C.c_float), ("err_spp", C.c_int32)]


class QueueStatsC(C.Structure):
    _fields_ = [("n_chunks", C.c_int32), ("chunks_per_device", C.c_int32 * 16), ("device_ms", C.c_double * 16),
                ("rays_per_device", C.c_uint64 * 16), ("rays", C.c_uint64)]


EXPORTS = ["rt_build_scene", "rt_build_scene_sd", "rt_render", "rt_render_stats_get", "rt_readback", "rt_readback_t", "rt_destroy",
           "rt_last_error", "rt_scene_info_get", "rt_scene_export", "rt_scene_export_host", "rt_accum_device_ptr",
           "rt_resolve", "rt_fb_device_ptr", "rt_trim_device_cache", "rt_write_ppm", "rt_load_texture", "rt_write_image", "rt_accum_reduce", "rt_render_adaptive", "rt_readback_spp", "rt_render_queue",
           "rt_sd_flatten"]

_lib = None


def lib():
    """Load librt_b200.so (built by __graft_entry__.build()). Fails loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtError("librt_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.rt_last_error.restype = C.c_char_p
        L.rt_build_scene.argtypes = [C.POINTER(SceneDescC), C.POINTER(C.c_void_p)]
        L.rt_build_scene_sd.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.rt_render.argtypes = [C.c_void_p, C.POINTER(RenderParamsC), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.rt_render_stats_get.argtypes = [C.c_void_p, C.POINTER(RenderStatsC)]
        L.rt_readback.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rt_readback_t.argtypes = [C.c_void_p, C.c_void_p]
        L.rt_destroy.argtypes = [C.c_void_p]
        L.rt_destroy.restype = None
        L.rt_scene_info_get.argtypes = [C.c_void_p, C.POINTER(SceneInfoC)]
        L.rt_scene_export.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]
        L.rt_scene_export_host.argtypes = [C.POINTER(SceneDescC), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                           C.c_void_p, C.c_int32]
        L.rt_accum_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.rt_resolve.argtypes = [C.c_void_p, C.c_int32, C.c_float]
        L.rt_fb_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.rt_write_ppm.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.rt_write_ppm.restype = C.c_long
        L.rt_write_image.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.rt_write_image.restype = C.c_long
        L.rt_accum_reduce.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32]
        L.rt_sd_flatten.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p, C.c_int32,
                                    C.POINTER(C.c_int32)]
        L.rt_render_adaptive.argtypes = [C.c_void_p, C.POINTER(RenderParamsC), C.POINTER(AdaptiveParamsC), C.POINTER(AdaptiveStatsC)]
        L.rt_readback_spp.argtypes = [C.c_void_p, C.c_void_p]
        L.rt_render_queue.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(RenderParamsC), C.c_int32, C.c_void_p, C.POINTER(QueueStatsC)]
        L.rt_load_texture.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RtError(lib().rt_last_error().decode())


def load_texture(path):
    """Decode a texture file (.jpg baseline JPEG / .ppm) exactly as rt_build_scene does: (h, w, 3) uint8."""
    w, h = C.c_int32(0), C.c_int32(0)
    _check(lib().rt_load_texture(path.encode(), None, 0, C.byref(w), C.byref(h)))
    out = np.empty((h.value, w.value, 3), dtype=np.uint8)
    _check(lib().rt_load_texture(path.encode(), out.ctypes.data, out.nbytes, C.byref(w), C.byref(h)))
    return out


def default_texture_dir():
    for d in (os.environ.get("RT_TEXTURE_DIR"), os.path.join(REPO_ROOT, "textures"),
              os.path.join(REPO_ROOT, "tests", "golden", "textures")):
        if d and os.path.isdir(d):
            return d
    return "textures"


# ---- SD (include/rt_scene_desc.h) as numpy structured arrays ----
TEX_DT = np.dtype([("kind", "<i4"), ("even", "<i4"), ("odd", "<i4"), ("image", "<i4"), ("color", "<f4", 3),
                   ("scale", "<f4"), ("p", "<f4", 13), ("pad_", "<i4")])
MAT_DT = np.dtype([("kind", "<i4"), ("tex", "<i4"), ("albedo", "<f4", 3), ("param", "<f4"), ("pad_", "<i4", 2)])
OBJ_DT = np.dtype([("kind", "<i4"), ("mat", "<i4"), ("child", "<i4"), ("inward", "<i4"), ("c0", "<f4", 3),
                   ("dc", "<f4", 3), ("radius", "<f4"), ("Q", "<f4", 3), ("u", "<f4", 3), ("v", "<f4", 3),
                   ("w", "<f4", 3), ("n", "<f4", 3), ("D", "<f4"), ("offset", "<f4", 3), ("sin_t", "<f4"),
                   ("cos_t", "<f4"), ("neg_inv_density", "<f4"), ("box_min", "<f4", 3), ("box_max", "<f4", 3)])
IMG_DT = np.dtype([("width", "<i4"), ("height", "<i4"), ("bpp", "<i4"), ("pad_", "<i4")])
CAM_DT = np.dtype([("origin", "<f4", 3), ("lower_left_corner", "<f4", 3), ("horizontal", "<f4", 3),
                   ("vertical", "<f4", 3), ("u", "<f4", 3), ("v", "<f4", 3), ("w", "<f4", 3), ("lens_radius", "<f4"),
                   ("time0", "<f8"), ("time1", "<f8")])
HDR_DT = np.dtype([("magic", "<u4"), ("scene_id", "<i4"), ("nx", "<i4"), ("ny", "<i4"), ("n_tex", "<i4"),
                   ("n_mat", "<i4"), ("n_obj", "<i4"), ("n_top", "<i4"), ("n_img", "<i4"), ("pad_", "<i4"),
                   ("cam", CAM_DT)])
SD_MAGIC = 0x31445352
assert CAM_DT.itemsize == 104 and HDR_DT.itemsize == 144 and OBJ_DT.itemsize == 156 and TEX_DT.itemsize == 88


class SD:
    """Parsed scene description."""

    def __init__(self, raw):
        raw = np.frombuffer(bytes(raw), dtype=np.uint8)
        self.raw = raw
        h = raw[:HDR_DT.itemsize].view(HDR_DT)[0]
        if int(h["magic"]) != SD_MAGIC:
            raise RtError("bad SD magic")
        self.hdr = h
        off = HDR_DT.itemsize

        def take(dt, n):
            nonlocal off
            a = raw[off:off + dt.itemsize * n].view(dt)
            off += dt.itemsize * n
            return a

        self.tex = take(TEX_DT, int(h["n_tex"]))
        self.mat = take(MAT_DT, int(h["n_mat"]))
        self.obj = take(OBJ_DT, int(h["n_obj"]))
        self.top = take(np.dtype("<i4"), int(h["n_top"]))
        self.img = take(IMG_DT, int(h["n_img"]))
        self.cam = h["cam"]

    # canonical, id-free description of object `i` (bit patterns of the floats), for deep comparison
    def _f(self, a):
        return tuple(int(x) for x in np.atleast_1d(np.asarray(a, dtype=np.float32)).view(np.uint32))

    def tex_key(self, t):
        if t < 0:
            return None
        d = self.tex[t]
        k = int(d["kind"])
        if k == 0:
            return ("solid", self._f(d["color"]))
        if k == 1:
            return ("checker", self._f(d["scale"]), self.tex_key(int(d["even"])), self.tex_key(int(d["odd"])))
        if k == 2:
            im = self.img[int(d["image"])]
            return ("image", int(im["width"]), int(im["height"]), int(im["bpp"]))
        if k == 3:
            return ("noise", self._f(d["scale"]))
        if k == 4:
            return ("noodle", self._f(d["p"]))
        if k == 5:
            return ("felt", self._f(d["color"]), self._f(d["p"][:4]))
        if k == 6:
            return ("uv_offset", self._f(d["p"][:2]), self.tex_key(int(d["even"])))
        return ("?", k)

    def mat_key(self, m):
        if m < 0:
            return None
        d = self.mat[m]
        k = int(d["kind"])
        if k in (0, 4):
            return (("lambertian", "", "", "", "isotropic")[k], self.tex_key(int(d["tex"])))
        if k == 1:
            return ("metal", self._f(d["albedo"]), self._f(d["param"]))
        if k == 2:
            return ("dielectric", self._f(d["param"]))
        if k == 3:
            return ("light", self.tex_key(int(d["tex"])), self._f(d["albedo"]) if int(d["tex"]) < 0 else None)
        return ("?", k)

    def obj_key(self, i, with_box=True):
        d = self.obj[i]
        k = int(d["kind"])
        box = (self._f(d["box_min"]), self._f(d["box_max"])) if with_box else None
        if k == 0:
            return ("sphere", self._f(d["c0"]), self._f(d["dc"]), self._f(d["radius"]), self.mat_key(int(d["mat"])), box)
        if k == 1:
            return ("quad", self._f(d["Q"]), self._f(d["u"]), self._f(d["v"]), self._f(d["w"]), self._f(d["n"]),
                    self._f(d["D"]), int(d["inward"]), self.mat_key(int(d["mat"])), box)
        if k == 2:
            c = int(d["child"])
            return ("box", tuple(self.obj_key(c + f, with_box) for f in range(6)), box)
        if k == 3:
            return ("translate", self._f(d["offset"]), self.obj_key(int(d["child"]), with_box), box)
        if k == 4:
            return ("rotate_y", self._f(d["sin_t"]), self._f(d["cos_t"]), self.obj_key(int(d["child"]), with_box), box)
        if k == 5:
            return ("medium", self._f(d["neg_inv_density"]), self.obj_key(int(d["child"]), with_box),
                    self.mat_key(int(d["mat"])), box)
        if k == 6:
            return ("with_material", self.obj_key(int(d["child"]), with_box), self.mat_key(int(d["mat"])), box)
        if k == 7:  # a bvh_node used as an object: the members of its cell chain, in creation order
            members, c = [], i
            while c >= 0:
                members.append(int(self.obj[c]["child"])); c = int(self.obj[c]["inward"])
            return ("bvh", tuple(self.obj_key(m, with_box) for m in reversed(members)), box)
        return ("?", k)

    def top_keys(self, with_box=True):
        return [self.obj_key(int(t), with_box) for t in self.top]


def sd_flatten(raw):
    """Host-only (no GPU): validate a scene description and resolve its bvh_node groups into instanced members.
    Returns (flattened SD, origin array); raises RtError with the validator's reason."""
    L = lib()
    need, ntop = C.c_size_t(0), C.c_int32(0)
    _check(L.rt_sd_flatten(raw, len(raw), None, 0, C.byref(need), None, 0, C.byref(ntop)))
    buf = (C.c_uint8 * need.value)()
    origin = np.zeros(ntop.value, dtype=np.int32)
    _check(L.rt_sd_flatten(raw, len(raw), buf, need.value, C.byref(need), origin.ctypes.data, ntop.value, C.byref(ntop)))
    return SD(bytes(buf)), origin


def export_host(scene_id, nx=0, ny=0, grid_half=0, texture_dir=None):
    """Scene description generated on the host only (no GPU). Returns (SD, rank array)."""
    L = lib()
    td = (texture_dir or default_texture_dir()).encode()
    d = SceneDescC(scene_id, nx, ny, grid_half, -1, td)
    need = C.c_size_t(0)
    _check(L.rt_scene_export_host(C.byref(d), None, 0, C.byref(need), None, 0))
    buf = (C.c_uint8 * need.value)()
    hdr_only = SD_peek_ntop(L, d)
    rank = np.zeros(hdr_only, dtype=np.int32)
    _check(L.rt_scene_export_host(C.byref(d), buf, need.value, C.byref(need), rank.ctypes.data, hdr_only))
    return SD(bytes(buf)), rank


def SD_peek_ntop(L, d):
    need = C.c_size_t(0)
    _check(L.rt_scene_export_host(C.byref(d), None, 0, C.byref(need), None, 0))
    buf = (C.c_uint8 * need.value)()
    _check(L.rt_scene_export_host(C.byref(d), buf, need.value, C.byref(need), None, 0))
    return int(np.frombuffer(bytes(buf[:HDR_DT.itemsize]), dtype=HDR_DT)[0]["n_top"])


class Scene:
    """A scene resident on one GPU: generator -> H2D -> device BVH build (rt_build_scene)."""

    def __init__(self, scene_id=0, nx=0, ny=0, grid_half=0, texture_dir=None, device=-1, sd=None, images=()):
        """scene_id 1..10: one of the reference's generators. sd = bytes of a scene description (rt_scene_desc.h)
        instead: the generic path (rt_build_scene_sd); images = decoded uint8 arrays for its image textures."""
        L = lib()
        self._h = C.c_void_p()
        if sd is not None:
            raw = bytes(sd)
            imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
            arr = (C.c_void_p * max(len(imgs), 1))(*[im.ctypes.data for im in imgs])
            _check(L.rt_build_scene_sd(raw, len(raw), arr, len(imgs), device, C.byref(self._h)))
        else:
            td = (texture_dir or default_texture_dir()).encode()
            self._td = td
            d = SceneDescC(scene_id, nx, ny, grid_half, device, td)
            _check(L.rt_build_scene(C.byref(d), C.byref(self._h)))
        info = SceneInfoC()
        _check(L.rt_scene_info_get(self._h, C.byref(info)))
        self.info = info
        self.nx, self.ny = info.nx, info.ny
        self.stats = None
        self._world, self._rank, self._split = 1, 0, 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().rt_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def render(self, spp=0, rng_mode=0, rank=0, world=1, split_mode=0, slots=0, aov=False, seed=0,
               max_depth=0, gamma=0.0, background=None, gradient_bg=None, profile=False, accumulate=False):
        p = RenderParamsC()
        p.spp, p.max_depth, p.gamma, p.t_min = spp, max_depth, gamma, 0.0
        if background is not None:
            p.override_background = 1
            p.background[0], p.background[1], p.background[2] = background
            p.gradient_bg = int(bool(gradient_bg))
        p.seed, p.rng_mode, p.split_mode, p.rank, p.world = seed, rng_mode, split_mode, rank, world
        p.slots, p.aov, p.profile = slots, int(bool(aov)), int(bool(profile))
        p.accumulate = int(bool(accumulate))
        ms, rays = C.c_double(0), C.c_uint64(0)
        _check(lib().rt_render(self._h, C.byref(p), C.byref(ms), C.byref(rays)))
        st = RenderStatsC()
        _check(lib().rt_render_stats_get(self._h, C.byref(st)))
        self.stats = st
        self._world, self._rank, self._split = world, rank, split_mode
        return st

    def render_adaptive(self, max_spp, threshold, min_spp=0, pass_spp=16, tile=16, rank=0, world=1, seed=0, slots=0,
                        max_depth=0, gamma=0.0, background=None, gradient_bg=None):
        """Adaptive per-tile sampling (rt_render_adaptive): returns AdaptiveStatsC; framebuffer() / spp_map() afterwards."""
        p = RenderParamsC()
        p.spp, p.max_depth, p.gamma = max_spp, max_depth, gamma
        if background is not None:
            p.override_background = 1
            p.background[0], p.background[1], p.background[2] = background
            p.gradient_bg = int(bool(gradient_bg))
        p.seed, p.rng_mode, p.split_mode, p.rank, p.world, p.slots = seed, 0, 0, rank, world, slots
        a = AdaptiveParamsC(min_spp, max_spp, pass_spp, tile, threshold)
        out = AdaptiveStatsC()
        _check(lib().rt_render_adaptive(self._h, C.byref(p), C.byref(a), C.byref(out)))
        st = RenderStatsC()
        _check(lib().rt_render_stats_get(self._h, C.byref(st)))
        self.stats = st
        self._world, self._rank, self._split = world, rank, 0
        return out

    def spp_map(self):
        """Samples every pixel of this rank's share received in the last adaptive render: int32 (rows_local, nx)."""
        st = self.stats
        m = np.empty((st.rows_local, st.nx), dtype=np.int32)
        _check(lib().rt_readback_spp(self._h, m.ctypes.data))
        return m

    def framebuffer(self):
        """This rank's share: float32 (rows_local, nx, 3), gamma applied, local row 0 = lowest owned scanline."""
        st = self.stats
        fb = np.empty((st.rows_local, st.nx, 3), dtype=np.float32)
        _check(lib().rt_readback(self._h, fb.ctypes.data, None, None))
        return fb

    def aov(self):
        st = self.stats
        obj = np.empty((st.rows_local, st.nx), dtype=np.int32)
        mat = np.empty((st.rows_local, st.nx), dtype=np.int32)
        t = np.empty((st.rows_local, st.nx), dtype=np.float32)
        _check(lib().rt_readback(self._h, None, obj.ctypes.data, mat.ctypes.data))
        _check(lib().rt_readback_t(self._h, t.ctypes.data))
        return obj, mat, t

    def accum_ptr(self):
        p, n = C.c_void_p(), C.c_size_t(0)
        _check(lib().rt_accum_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def fb_ptr(self):
        p, n = C.c_void_p(), C.c_size_t(0)
        _check(lib().rt_fb_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def resolve(self, total_spp=0, gamma=0.0):
        _check(lib().rt_resolve(self._h, total_spp, gamma))

    def export(self):
        L = lib()
        need = C.c_size_t(0)
        _check(L.rt_scene_export(self._h, None, 0, C.byref(need), None))
        buf = (C.c_uint8 * need.value)()
        rank = np.zeros(self.info.n_top, dtype=np.int32)
        _check(L.rt_scene_export(self._h, buf, need.value, C.byref(need), rank.ctypes.data))
        return SD(bytes(buf)), rank


def render_queue(scenes, spp, n_chunks=0, rng_mode=0, seed=0, max_depth=0, background=None, gradient_bg=None):
    """Dynamic tile queue over scene replicas of this process (rt_render_queue): returns (full (ny, nx, 3) image, QueueStatsC)."""
    p = RenderParamsC()
    p.spp, p.rng_mode, p.seed, p.max_depth = spp, rng_mode, seed, max_depth
    if background is not None:
        p.override_background = 1
        p.background[0], p.background[1], p.background[2] = background
        p.gradient_bg = int(bool(gradient_bg))
    arr = (C.c_void_p * len(scenes))(*[sc._h for sc in scenes])
    out = np.zeros((scenes[0].ny, scenes[0].nx, 3), dtype=np.float32)
    st = QueueStatsC()
    _check(lib().rt_render_queue(arr, len(scenes), C.byref(p), n_chunks, out.ctypes.data, C.byref(st)))
    return out, st


def assemble_rows(parts, ny):
    """Interleave tile-split shares (rank r owns scanlines r, r+world, ...) into a full (ny, nx, ...) array."""
    world = len(parts)
    out = np.empty((ny,) + parts[0].shape[1:], dtype=parts[0].dtype)
    for r, p in enumerate(parts):
        out[r::world] = p
    return out


def to_8bit(fb, double_scale=False):
    """The reference's PPM quantisation: int(255.99f * c), no clamp (main.cu:1216-1218)."""
    if double_scale:
        return (np.float64(255.99) * fb.astype(np.float64)).astype(np.int32)
    return (np.float32(255.99) * fb.astype(np.float32)).astype(np.int32)


def write_ppm(path, fb, double_scale=False):
    """P3 PPM exactly as the reference prints it (rows top to bottom). fb: (ny, nx, 3), row 0 = bottom."""
    fb = np.ascontiguousarray(fb, dtype=np.float32)
    ny, nx = fb.shape[:2]
    n = lib().rt_write_ppm(path.encode() if path else None, fb.ctypes.data, nx, ny, int(double_scale))
    if n < 0:
        raise RtError(lib().rt_last_error().decode())
    return n


class DevicePtrView:
    """Zero-copy view of library-owned device memory for torch (torch.as_tensor(view, device="cuda")).
    Plumbing for the NCCL reduce / gather of torch.distributed; no compute happens in torch."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}
