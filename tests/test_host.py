"""CPU tests (no GPU): the C ABI loads and exports what include/rt_api.h declares, the host scene generators
reproduce the scenes the reference built on the GPU (tests/golden), the PPM writer is byte-exact, and the
multi-rank host logic works under torch.distributed (gloo, world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, PKG, texture_dir

SCENES = [("c1_400x225_10", 1), ("c2_300x300_16", 7), ("c3_300x300_16", 8), ("c4_400x400_16", 9), ("s2_300x150_8", 2),
          ("s3_300x150_8", 3), ("s4_300x150_8", 4), ("s5_300x150_8", 5), ("s6_300x150_8", 6), ("s10_300x150_8", 10)]
NEEDS_TEX = {3: ["earthmap"], 9: ["earthmap"], 6: ["poolball"], 10: ["porcelain", "8ball"]}


def test_library_exports_every_declared_symbol(pyrt):
    hdr = open(os.path.join(ROOT, "include", "rt_api.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", hdr))
    assert {"rt_build_scene", "rt_render", "rt_readback", "rt_destroy", "rt_last_error"} <= declared
    lib = pyrt.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "librt_b200.so does not export %s" % name
    assert declared == set(pyrt.EXPORTS), "pyrt.EXPORTS out of sync with include/rt_api.h"


def test_struct_layouts_match_header(pyrt):
    # sizes the C compiler gives the ABI structs (checked against a tiny C program built with gcc)
    src = r'''
#include <stdio.h>
#include "rt_api.h"
#include "rt_scene_desc.h"
int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rt_scene_desc), sizeof(rt_render_params),
  sizeof(rt_scene_info), sizeof(rt_render_stats), sizeof(rt_texture_desc), sizeof(rt_material_desc), sizeof(rt_object_desc),
  sizeof(rt_camera_desc), sizeof(rt_sd_header)); printf("%zu %zu %zu\n", sizeof(rt_adaptive_params), sizeof(rt_adaptive_stats), sizeof(rt_queue_stats)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    sizes = [int(x) for x in out]
    assert sizes[:4] == [C.sizeof(pyrt.SceneDescC), C.sizeof(pyrt.RenderParamsC), C.sizeof(pyrt.SceneInfoC), C.sizeof(pyrt.RenderStatsC)]
    assert sizes[4:9] == [pyrt.TEX_DT.itemsize, pyrt.MAT_DT.itemsize, pyrt.OBJ_DT.itemsize, pyrt.CAM_DT.itemsize, pyrt.HDR_DT.itemsize]
    assert sizes[9:] == [C.sizeof(pyrt.AdaptiveParamsC), C.sizeof(pyrt.AdaptiveStatsC), C.sizeof(pyrt.QueueStatsC)]


JPEG_DIR = os.path.join(ROOT, "oracle", "_ref", "textures_jpg")
PPM_DIR = os.path.join(ROOT, "oracle", "_ref", "textures")


@pytest.mark.parametrize("name", ["earthmap", "poolball", "porcelain", "8ball", "hardwood"])
def test_jpeg_decoder_gives_the_reference_decoders_bytes(pyrt, name):
    """The product decodes textures/<name>.jpg itself (csrc/jpeg_baseline.cpp) where the reference calls stbi_load
    (image_io.h:24-41). Texels enter the render bit for bit, so the decode must equal the reference decoder's output byte
    for byte: oracle/_ref/textures/<name>.ppm is what the reference's own decoder (compiled into oracle/_ref/ref_cpu)
    made of the same file. Covers 4:4:4 (earthmap) and 4:2:0 (the others) baseline files, 1024x512 and 824x360."""
    jpg, ppm = os.path.join(JPEG_DIR, name + ".jpg"), os.path.join(PPM_DIR, name + ".ppm")
    if not (os.path.exists(jpg) and os.path.exists(ppm)):
        pytest.skip("oracle/_ref textures not built (needs /root/reference once: oracle/build_ref.sh cpu)")
    mine = pyrt.load_texture(jpg)
    want = pyrt.load_texture(ppm)
    assert mine.shape == want.shape and mine.shape[2] == 3
    assert np.array_equal(mine, want), "%d bytes differ" % int((mine != want).sum())


def test_jpeg_decoder_rejects_what_it_cannot_decode_exactly(pyrt, tmp_path):
    with pytest.raises(pyrt.RtError, match="cannot open"):
        pyrt.load_texture(str(tmp_path / "missing.jpg"))
    (tmp_path / "junk.jpg").write_bytes(b"not a jpeg at all")
    with pytest.raises(pyrt.RtError, match="not a JPEG"):
        pyrt.load_texture(str(tmp_path / "junk.jpg"))
    with pytest.raises(pyrt.RtError, match="unknown texture format"):
        pyrt.load_texture(str(tmp_path / "x.png"))
    jpg = os.path.join(JPEG_DIR, "poolball.jpg")
    if os.path.exists(jpg):
        raw = bytearray(open(jpg, "rb").read())
        sof = raw.index(b"\xff\xc0")
        prog = bytearray(raw); prog[sof + 1] = 0xC2  # a progressive frame header: refused, never decoded differently
        (tmp_path / "prog.jpg").write_bytes(prog)
        with pytest.raises(pyrt.RtError, match="progressive"):
            pyrt.load_texture(str(tmp_path / "prog.jpg"))
        (tmp_path / "cut.jpg").write_bytes(raw[: len(raw) // 3])  # truncated entropy data: decodes what is there or fails, no crash
        try:
            img = pyrt.load_texture(str(tmp_path / "cut.jpg"))
            assert img.shape == (512, 1024, 3)
        except pyrt.RtError:
            pass
        (tmp_path / "hdr.jpg").write_bytes(raw[: sof + 6])  # cut inside the frame header
        with pytest.raises(pyrt.RtError):
            pyrt.load_texture(str(tmp_path / "hdr.jpg"))


def test_image_writers_p3_p6_png(pyrt, tmp_path):
    """rt_write_image: P3 is the reference's text (unclamped unless asked), P6 and PNG carry the same clamped bytes;
    the PNG is decoded back here with zlib and checked chunk by chunk (CRC, Adler, IHDR)."""
    import struct, zlib
    rng = np.random.default_rng(3)
    nx, ny = 37, 23
    img = rng.random((ny, nx, 3), dtype=np.float32) * 1.3 - 0.1   # values below 0 and above 1
    L = pyrt.lib()
    want = np.clip((np.float32(255.99) * img).astype(np.int64), 0, 255).astype(np.uint8)[::-1]   # top row first
    unclamped = (np.float32(255.99) * img).astype(np.int64)[::-1]
    paths = {k: str(tmp_path / ("o." + k)).encode() for k in ("p3", "p3c", "p6", "png")}
    assert L.rt_write_image(paths["p3"], img.ctypes.data, nx, ny, 0, 0, 0) > 0
    assert L.rt_write_image(paths["p3c"], img.ctypes.data, nx, ny, 0, 1, 0) > 0
    assert L.rt_write_image(paths["p6"], img.ctypes.data, nx, ny, 1, 0, 0) > 0
    assert L.rt_write_image(paths["png"], img.ctypes.data, nx, ny, 2, 0, 0) > 0
    tok = open(paths["p3"]).read().split()
    assert tok[:4] == ["P3", str(nx), str(ny), "255"]
    assert np.array_equal(np.array(tok[4:], dtype=np.int64).reshape(ny, nx, 3), unclamped)
    assert unclamped.max() > 255 and unclamped.min() <= 0
    tokc = open(paths["p3c"]).read().split()
    assert np.array_equal(np.array(tokc[4:], dtype=np.int64).reshape(ny, nx, 3), want)
    raw = open(paths["p6"], "rb").read()
    hdr = b"P6\n%d %d\n255\n" % (nx, ny)
    assert raw.startswith(hdr) and np.array_equal(np.frombuffer(raw[len(hdr):], dtype=np.uint8).reshape(ny, nx, 3), want)
    png = open(paths["png"], "rb").read()
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    off, chunks = 8, []
    while off < len(png):
        n, typ = struct.unpack(">I4s", png[off:off + 8])
        data = png[off + 8:off + 8 + n]
        assert struct.unpack(">I", png[off + 8 + n:off + 12 + n])[0] == zlib.crc32(typ + data)
        chunks.append((typ, data))
        off += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    assert struct.unpack(">IIBBBBB", chunks[0][1]) == (nx, ny, 8, 2, 0, 0, 0)
    rows = np.frombuffer(zlib.decompress(chunks[1][1]), dtype=np.uint8).reshape(ny, 1 + 3 * nx)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(ny, nx, 3), want)
    big = np.zeros((300, 300, 3), dtype=np.float32)   # > 65535 bytes: several stored deflate blocks
    assert L.rt_write_image(paths["png"], big.ctypes.data, 300, 300, 2, 0, 0) > 0
    png = open(paths["png"], "rb").read()
    i = png.index(b"IDAT")
    n = struct.unpack(">I", png[i - 4:i])[0]
    assert len(zlib.decompress(png[i + 4:i + 4 + n])) == 300 * 901
    assert L.rt_write_image(paths["png"], img.ctypes.data, nx, ny, 7, 0, 0) < 0  # unknown format


def test_no_cpu_fallback(pyrt):
    """Without a CUDA device the render path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pyrt.RtError, match="no CUDA device"):
        pyrt.Scene(7, 16, 16)
    # and the product never loads, links or calls the oracle
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(root, f), errors="ignore").read()
                for needle in ("oracle_py", "liboracle", "rt_oracle", "oracle/", "ref_cpu"):
                    assert needle not in txt, "%s mentions %s" % (f, needle)


def test_library_carries_native_sm100a_kernels(pyrt):
    """The shipped library holds sm_100a machine code (no PTX-only build, no other arch) for every kernel of the path, both
    trace kernels among them, and the node step's slab arithmetic is the packed f32x2 form (FADD2 / FMUL2, sm_100 only)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not installed")
    elf = subprocess.run([cuobjdump, "-lelf", pyrt.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in elf and not [l for l in elf.splitlines() if "sm_" in l and "sm_100a" not in l], elf
    sass = subprocess.run([cuobjdump, "-sass", pyrt.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    funcs = [l.split(":", 1)[1].strip() for l in sass.splitlines() if l.strip().startswith("Function :")]
    for k in ("k_init", "k_traceE", "k_trace_small", "k_shade", "k_finish", "k_accumulate", "k_resolve", "k_aov", "k_ploc", "k_bvh_collapse"):
        assert any(k in f for f in funcs), "kernel %s missing from the library" % k
    body = {}
    cur = None
    for l in sass.splitlines():
        if l.strip().startswith("Function :"):
            cur = l.split(":", 1)[1].strip()
            body[cur] = [0, 0]
        elif cur and ("FADD2" in l or "FMUL2" in l):
            body[cur][0] += 1
        elif cur and ("STL.64" in l or "LDL.64" in l):
            body[cur][1] += 1
    kt = [f for f in funcs if "7k_traceE" in f][0]
    ks = [f for f in funcs if "k_trace_small" in f][0]
    assert body[kt][0] == 24 and body[ks][0] == 24   # 48 subtractions and multiplications of four slab tests, two per instruction
    assert body[kt][1] > 0 and body[ks][1] == 0       # 64-bit (node, distance) stack entries only where they are dropped by distance


def _fkey_close(a, b, tol):
    """Deep compare of SD keys: ints/strings exact, float bit patterns within `tol` ulps."""
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b) and all(_fkey_close(x, y, tol) for x, y in zip(a, b))
    if isinstance(a, int) and isinstance(b, int) and (a > 1 << 16 or b > 1 << 16):
        return abs(a - b) <= tol
    return a == b


@pytest.mark.parametrize("name,sid", SCENES, ids=[s[0] for s in SCENES])
def test_host_generators_reproduce_reference_scene(pyrt, golden, name, sid):
    """rt_scene_export_host (generators + builder + reference leaf order on the host) vs the scene dump of the
    reference's CUDA build. Exact for everything that does not come out of sinf/cosf/tanf (host libm here, CUDA
    libdevice on the GPU: the `-m gpu` test is bit-exact); those fields within 4 ulps."""
    g = golden(name)
    td = texture_dir()
    for t in NEEDS_TEX.get(sid, []):
        if td is None or not os.path.exists(os.path.join(td, t + ".ppm")):
            pytest.skip("decoded texture %s.ppm not present" % t)
    ref = pyrt.SD(g["sd"].tobytes())
    mine, rank = pyrt.export_host(sid, int(g["nx"]), int(g["ny"]), texture_dir=td)
    assert len(mine.top) == len(ref.top) and len(mine.obj) == len(ref.obj)
    inv = np.argsort(rank)
    trig = sid in (7, 8)  # rotate_y instances: sin/cos in the object, boxes derived from them
    for pos in range(len(ref.top)):
        a, b = ref.obj_key(int(ref.top[pos])), mine.obj_key(int(mine.top[inv[pos]]))
        assert a == b or (trig and _fkey_close(a, b, 64)), "leaf %d" % pos
    for n in pyrt.CAM_DT.names:
        a, b = np.atleast_1d(ref.cam[n]), np.atleast_1d(mine.cam[n])
        assert np.allclose(a, b, rtol=1e-6, atol=1e-6), n  # camera goes through tanf


def test_oracle_leaf_order_equals_host_leaf_order(pyrt, built):
    """Two independent restatements of the reference BVH constructor's ordering (bvh.cuh:46-81)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    for sid, gh in [(1, 0), (7, 0), (5, 0), (1, 20)]:
        sd, rank = pyrt.export_host(sid, 64, 64, grid_half=gh, texture_dir=texture_dir())
        o = oracle_py.Oracle(sd.raw.tobytes())
        assert np.array_equal(o.leaf_order(), rank), sid


def test_ppm_writer_matches_reference_format(pyrt, tmp_path):
    """main.cu:1212-1221: 'P3\\n{nx} {ny}\\n255\\n', rows j = ny-1..0, int(255.99f*c) per channel, no clamp."""
    rng = np.random.default_rng(1)
    fb = (rng.random((5, 7, 3), dtype=np.float32) * 1.3).astype(np.float32)
    fb[0, 0] = [15.0 ** (1 / 2.2), 0.0, 1.0]  # an emitter: values above 255 are printed as they are
    p = str(tmp_path / "o.ppm")
    n = pyrt.write_ppm(p, fb)
    txt = open(p).read()
    assert n == len(txt)
    lines = txt.split("\n")
    assert lines[0] == "P3" and lines[1] == "7 5" and lines[2] == "255"
    want = []
    for j in range(4, -1, -1):
        for i in range(7):
            c = fb[j, i]
            want.append("%d %d %d" % tuple(int(np.float32(255.99) * np.float32(x)) for x in c))
    assert lines[3:3 + 35] == want
    assert lines[3 + 28] == "%d 0 255" % int(np.float32(255.99) * np.float32(15.0 ** (1 / 2.2)))  # pixel (0,0) is on the last row
    # bouncing_spheres prints with a DOUBLE 255.99 (main.cu:722-724)
    pyrt.write_ppm(p, fb, double_scale=True)
    l2 = open(p).read().split("\n")
    c = fb[4, 0]
    assert l2[3] == "%d %d %d" % tuple(int(255.99 * float(x)) for x in c)
    assert np.array_equal(pyrt.to_8bit(fb), (np.float32(255.99) * fb).astype(np.int32))


def test_partition_arithmetic():
    from pyrt import dist
    for ny in (1, 7, 225, 800):
        for world in (1, 2, 3, 8):
            rows = [dist.rows_of_rank(ny, r, world) for r in range(world)]
            assert sorted(sum(rows, [])) == list(range(ny))
            assert [len(x) for x in rows] == [dist.rows_local(ny, r, world) for r in range(world)]
    for spp in (1, 10, 1000, 10007):
        for world in (1, 2, 8):
            shares = [dist.sample_share(spp, r, world) for r in range(world)]
            assert shares[0][0] == 0 and shares[-1][1] == spp
            assert all(shares[i][1] == shares[i + 1][0] for i in range(world - 1))
    parts = [np.arange(12).reshape(4, 3)[r::3] for r in range(3)]
    assert np.array_equal(dist.assemble_rows(parts, 4), np.arange(12).reshape(4, 3))


WORKER = r'''
import os, sys
sys.path.insert(0, %(pkg)r)
import numpy as np, torch, torch.distributed as td
from pyrt import dist
world, rank, local = dist.init_process_group("gloo")
assert world == 2 and td.get_backend() == "gloo"
ny, nx = 9, 5
full = torch.arange(ny * nx * 3, dtype=torch.float32).view(ny, nx, 3)
# tile split: each rank holds its scanlines; rank 0 must end up with the whole image
mine = full[rank::world].contiguous()
assert mine.shape[0] == dist.rows_local(ny, rank, world)
got = dist.gather_rows_to_root(mine, ny)
if rank == 0:
    assert torch.equal(got, full)
else:
    assert got is None
# spp split: per-rank linear sums reduced to rank 0
base, end = dist.sample_share(10, rank, world)
acc = torch.full((ny * nx * 3,), float(end - base))
dist.reduce_sum_to_root(acc)
if rank == 0:
    assert torch.all(acc == 10.0)
# dynamic tile queue: 7 chunks pulled from the shared counter by 2 ranks; rank 1 is "slow" and must end up with fewer
import time
def fake_render(c, n):
    if rank == 1:
        time.sleep(0.05)
    return full[c::n].contiguous()
img, mine = dist.render_dynamic(fake_render, 7, ny, nx)
counts = [None, None]
td.all_gather_object(counts, mine)
assert sorted(counts[0] + counts[1]) == list(range(7)), counts
assert len(counts[0]) > len(counts[1]) >= 0, counts
if rank == 0:
    assert torch.equal(img, full)
else:
    assert img is None
img2, mine2 = dist.render_dynamic(fake_render, 3, ny, nx)   # a second queue in the same group gets a fresh counter
td.all_gather_object(counts, mine2)
assert sorted(counts[0] + counts[1]) == [0, 1, 2]
td.barrier()
td.destroy_process_group()
print("rank %%d ok" %% rank)
'''


def test_multi_rank_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"pkg": PKG})
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_bench_reference_arm_contract():
    """bench.py --impl reference: ranks other than 0 exit 0 without work (the driver launches it under torchrun)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], env=dict(os.environ, RANK="1", WORLD_SIZE="2"), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_c5_grid_generator_sizes(pyrt):
    """C5 scale-up: GRID_MIN/MAX (main.cu:140-141) generalised by grid_half; (2G)^2 grid cells + ground + 3 big spheres."""
    for g in (11, 20, 50):
        sd, rank = pyrt.export_host(1, 64, 36, grid_half=g)
        assert len(sd.top) == 4 * g * g + 4
        assert sorted(rank.tolist()) == list(range(len(sd.top)))
        kinds = sd.obj["kind"][sd.top]
        assert (kinds == 0).all()  # spheres only
        moving = (np.abs(sd.obj["dc"][sd.top]).sum(axis=1) > 0).mean()
        assert 0.6 < moving < 0.9  # ~80% of the grid is diffuse (moving), SURVEY §8a7


def test_scene_builder_bvh_node_matches_python_builder_and_flattens(pyrt, tmp_path):
    """The source-level API (csrc/scene_builder.h): a C++ program builds translate(rotate_y(bvh_node(spheres))) + with_material
    of the same group with rt::SceneBuilder and writes the SD; the numpy builder of tests/sdgen.py builds the same scene.
    Same object graph (floats bit for bit, boxes included) and the same flattening through rt_sd_flatten."""
    csrc = os.path.join(ROOT, "accelerated-ray-tracer_b200", "csrc")
    src = tmp_path / "g.cpp"
    src.write_text(r'''
#include <cmath>
#include <cstdio>
#include "scene_builder.h"
using namespace rt;
struct HM : DevMath { float sinf_(float x) override { return sinf(x); } float cosf_(float x) override { return cosf(x); }
                      float tanf_(float x) override { return tanf(x); } };
int main(int, char** argv) {
  HM dm; SceneDesc sd; sd.nx = 64; sd.ny = 48;
  SceneBuilder B(sd, dm);
  const int grey = B.lambertian(v3(0.5f, 0.5f, 0.5f)), gold = B.metal(v3(0.8f, 0.6f, 0.2f), 0.25f);
  std::vector<int> members;
  for (int i = 0; i < 5; ++i) members.push_back(B.sphere(v3(0.75f * i + 0.25f, 0.5f, -0.25f * i - 0.5f), 0.3f, grey));
  members.push_back(B.translate(B.sphere(v3(0.5f, 0.75f, 0.5f), v3(0.5f, 1.0f, 0.5f), 0.25f, gold), v3(1.0f, 0.0f, 1.5f)));  // a moving sphere, instanced
  const int g = B.bvh_node(members);
  B.add(B.translate(B.rotate_y(g, 0.0f), v3(-2.0f, 1.0f, 0.5f)));   // 0 degrees: sin / cos are exact on every libm
  B.add(B.with_material(g, gold));
  B.add(B.sphere(v3(0, -100.5f, 0), 100.0f, grey));
  B.camera(v3(0, 2, 8), v3(0, 1, 0), v3(0, 1, 0), 40.0f, 64.0f / 48.0f, 0.0f, 8.0f, 0.0, 1.0);
  const std::string bin = sd_serialize(sd);
  FILE* f = fopen(argv[1], "wb"); fwrite(bin.data(), 1, bin.size(), f); fclose(f);
  return 0;
}
''')
    exe = tmp_path / "g"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I", csrc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                           os.path.join(csrc, "scene_builder.cpp"), os.path.join(csrc, "generators.cpp"), os.path.join(csrc, "jpeg_baseline.cpp")])
    out = tmp_path / "scene.sd"
    subprocess.check_call([str(exe), str(out)])
    cpp = pyrt.SD(out.read_bytes())
    sys.path.insert(0, os.path.dirname(__file__))
    from sdgen import SDBuilder
    P = SDBuilder(64, 48)
    grey, gold = P.lambertian(P.solid((0.5, 0.5, 0.5))), P.metal((0.8, 0.6, 0.2), 0.25)
    members = [P.sphere((0.75 * i + 0.25, 0.5, -0.25 * i - 0.5), 0.3, grey) for i in range(5)]
    members.append(P.translate(P.sphere((0.5, 0.75, 0.5), 0.25, gold, c1=(0.5, 1.0, 0.5)), (1.0, 0.0, 1.5)))
    g = P.bvh(members)
    P.add(P.translate(P.rotate_y(g, 0.0), (-2.0, 1.0, 0.5)))
    P.add(P.with_material(g, gold))
    P.add(P.sphere((0, -100.5, 0), 100.0, grey))
    P.camera((0, 2, 8), (0, 1, 0), (0, 1, 0), 40.0, 0.0, 8.0)
    py = pyrt.SD(P.to_bytes())
    assert cpp.top_keys() == py.top_keys()
    fc, oc = pyrt.sd_flatten(out.read_bytes())
    fp, op = pyrt.sd_flatten(P.to_bytes())
    assert list(oc) == list(op) == [0] * 6 + [1] * 6 + [2]
    assert fc.top_keys() == fp.top_keys()
    kinds = [int(fc.obj["kind"][t]) for t in fc.top]
    assert kinds == [3] * 6 + [6] * 6 + [0]   # translate(rotate_y(member)) x 6, with_material(member) x 6, the ground sphere


def test_leaf_order_replay_equals_the_reference_selection_sort(tmp_path):
    """reference_leaf_order replays bvh_node's constructor (bvh.cuh:29-84) to know the reference's leaf order (exact-t ties are
    broken by it). Long ranges replay the selection sort's swaps through a tournament tree; tests/cpp/leaf_order_replay.cpp
    checks that against the plain O(n^2) loop on 300 random / tie-heavy key sets."""
    csrc = os.path.join(ROOT, "accelerated-ray-tracer_b200", "csrc")
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I", csrc, "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "cpp", "leaf_order_replay.cpp"), os.path.join(csrc, "jpeg_baseline.cpp")])
    out = subprocess.check_output([str(exe)], text=True)
    assert "sorttest: 0 mismatches" in out


def test_stack_culling_argument_holds_in_float32():
    """k_trace drops a stacked BVH node at pop time when its entry distance, with the two low mantissa bits cleared (they
    carry the child slot of the sort key), is >= the closest hit found since (rt_intersect.cuh, TravT<true>::live). The
    claim that this changes no result: the reference slab test (aabb.cuh:45-61) applied to any box INSIDE that node's box
    - every descendant's, leaf boxes included - with t_max = that hit rejects. Checked here in float32 with the kernel's
    own expression ((plane - o) * (1 / d), near / far plane picked by the sign of 1 / d, fmaxf / fminf that ignore NaN),
    on nested random boxes, rays with zero direction components and origins on box planes included."""
    rng = np.random.default_rng(7)
    n = 400000
    f = np.float32
    lo_p = rng.uniform(-10, 10, (n, 3)).astype(f)
    hi_p = (lo_p + rng.uniform(0, 8, (n, 3)).astype(f)).astype(f)
    a, b = rng.uniform(0, 1, (n, 3)).astype(f), rng.uniform(0, 1, (n, 3)).astype(f)
    a, b = np.minimum(a, b), np.maximum(a, b)
    shared = rng.uniform(0, 1, (n, 3)) < 0.2           # a child often shares planes with its parent (exact unions)
    lo_c = np.where(shared, lo_p, (lo_p + a * (hi_p - lo_p)).astype(f)).astype(f)
    hi_c = np.where(rng.uniform(0, 1, (n, 3)) < 0.2, hi_p, (lo_p + b * (hi_p - lo_p)).astype(f)).astype(f)
    lo_c, hi_c = np.clip(lo_c, lo_p, hi_p), np.clip(hi_c, lo_p, hi_p)
    lo_c, hi_c = np.minimum(lo_c, hi_c), np.maximum(lo_c, hi_c)
    o = rng.uniform(-20, 20, (n, 3)).astype(f)
    on_plane = rng.uniform(0, 1, (n, 3)) < 0.05
    o = np.where(on_plane, lo_p, o).astype(f)          # (plane - o) = 0: 0 * inf = NaN when the direction component is 0
    d = rng.normal(0, 1, (n, 3)).astype(f)
    aim = rng.uniform(0, 1, (n, 1)) < 0.7              # most rays point at the node's box, so that it is entered
    d = np.where(aim, (lo_p + rng.uniform(0, 1, (n, 3)).astype(f) * (hi_p - lo_p) - o), d).astype(f)
    d = np.where(rng.uniform(0, 1, (n, 3)) < 0.1, f(0), d).astype(f)
    d = np.where(rng.uniform(0, 1, (n, 3)) < 0.02, f(-0.0), d).astype(f)
    tmin = f(0.001)

    def slab(lo, hi, tmax):
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            inv = (f(1) / d).astype(f)
            near = np.where(inv < 0, hi, lo)
            far = np.where(inv < 0, lo, hi)
            tn = ((near - o).astype(f) * inv).astype(f)
            tf = ((far - o).astype(f) * inv).astype(f)
        t0 = np.full(n, tmin, f)
        t1 = tmax.astype(f).copy()
        for ax in range(3):                               # fmaxf / fminf: a NaN operand is ignored (np.fmax / np.fmin too)
            t0 = np.fmax(tn[:, ax], t0)
            t1 = np.fmin(tf[:, ax], t1)
        return t0, t1

    lo_parent, hi_parent = slab(lo_p, hi_p, np.full(n, np.finfo(f).max, f))
    entered = hi_parent > lo_parent                       # the node was pushed: its box passed the test at push time
    key = lo_parent.view(np.uint32) & np.uint32(0xFFFFFFFC)
    for frac in (0.25, 0.9, 1.0, "1ulp", 1.1):            # a hit found later, nearer than / at / just beyond the node's entry
        if frac == "1ulp":   # (the case that needs the two key bits cleared: with `key | slot` a node 1 ulp before the hit would go)
            best_t, frac = np.nextafter(lo_parent, f(np.inf)), 1.0000001
        else:
            best_t = np.nextafter((lo_parent * f(frac)).astype(f), f(0)) if frac != 1.0 else lo_parent.copy()
        best_t = np.maximum(best_t, np.nextafter(tmin, f(1))).astype(f)
        dropped = entered & ~(key < best_t.view(np.uint32))
        t0, t1 = slab(lo_c, hi_c, best_t)
        assert not np.any(dropped & (t1 > t0)), frac      # nothing inside a dropped node would have been entered
        if frac < 1.0:
            assert dropped[entered & (lo_parent > tmin)].all()   # and the rule does drop the nodes behind the hit
        elif frac > 1.0:
            assert not dropped[entered & (lo_parent > tmin)].any()
    assert (entered & (lo_parent > tmin)).mean() > 0.2
