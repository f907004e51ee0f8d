"""CPU tests: the host-side mirror of the reference's scene/camera API against the oracle's restatement of the
same producers, and the C-ABI libraries themselves (load, symbols, no-GPU behaviour)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- Camera.init (src/camera.zig:118-154): host mirror == oracle, bit for bit ---------------------------------
@pytest.mark.parametrize("kw", [
    dict(image_width=400, samples_per_pixel=10, max_depth=50),
    dict(image_width=1200, samples_per_pixel=500, max_depth=50),
    dict(image_width=7680, samples_per_pixel=1024, max_depth=50),
    dict(image_width=600, aspect_ratio=1.0, vfov=40.0, lookfrom=(278, 278, -800), lookat=(278, 278, 0),
         defocus_angle=0.0, max_depth=200, samples_per_pixel=200),          # cornellBox camera (main.zig:191-200)
    dict(image_width=333, image_height=77, vfov=63.0, lookfrom=(1, 2, 3), lookat=(-4, 0.5, 9), vup=(0.1, 1, 0.2),
         defocus_angle=2.5, focus_dist=3.3),
])
def test_camera_init_matches_oracle(pkg, orc, kw):
    cam = pkg.Camera(**kw)
    assert bytes(cam.init()) == bytes(orc.camera_init(cam))


def test_camera_defaults_are_the_reference_defaults(pkg):
    o = pkg.RtwCameraOptions()
    pkg._ffi.rtw().rtw_camera_defaults(C.byref(o))     # src/camera.zig:70-91
    assert o.aspect_ratio == pytest.approx(16 / 9) and o.image_width == 800 and o.image_height == 0
    assert o.samples_per_pixel == 100 and o.max_depth == 16 and list(o.background) == [0, 0, 0]
    assert o.vfov == 20 and list(o.lookfrom) == [13, 2, 3] and list(o.lookat) == [0, 0, 0] and list(o.vup) == [0, 1, 0]
    assert o.defocus_angle == pytest.approx(0.6) and o.focus_dist == 10


def test_camera_field_ranges_are_the_reference_types(pkg):
    # image_width/height u16, samples_per_pixel u16, max_depth u8 (src/camera.zig:71-79)
    cam = pkg.RtbCamera()
    for kw in (dict(image_width=70000), dict(samples_per_pixel=70000), dict(max_depth=256), dict(image_width=0)):
        o = pkg.Camera(**kw).options()
        assert pkg._ffi.rtw().rtw_camera_init(C.byref(o), C.byref(cam)) == pkg.RTB_ERR_INVALID_ARGUMENT


# ---- BVHTree.init (src/bvh.zig:22-103): host mirror == oracle on the same RNG stream ---------------------------
def _tree_signature(desc, i, objkey):
    nd = desc.nodes[i]
    box = tuple(nd.bmin[:]) + tuple(nd.bmax[:])
    if nd.leaf >= 0:
        return ("leaf", box, objkey(desc.hittables[nd.leaf]))
    return ("node", box, _tree_signature(desc, nd.left, objkey), _tree_signature(desc, nd.right, objkey))


def test_bvh_build_matches_oracle(pkg, orc):
    ffi = pkg._ffi
    rng = np.random.default_rng(5)
    n = 203
    world = pkg.World.new()
    hs = (ffi.RtbHittable * n)()
    boxes = np.zeros((n, 6), np.float32)
    for i in range(n):
        c1 = rng.uniform(-10, 10, 3).astype(np.float32)
        r = float(np.float32(rng.uniform(0.05, 0.8)))
        moving = i % 3 == 0
        c2 = (c1 + rng.uniform(0, 0.5, 3).astype(np.float32)).astype(np.float32) if moving else None
        world.add_sphere(c1, r, pkg.material_spec(color=(i / n, 0, 0)), center2=c2)
        hs[i].type, hs[i].material, hs[i].radius, hs[i].is_moving = ffi.RTB_HITTABLE_SPHERE, i, r, int(moving)
        hs[i].a = (ffi.f32 * 3)(*c1)
        if moving:
            hs[i].b = (ffi.f32 * 3)(*(c2 - c1))     # center_vec = center2 - center1 (objects.zig:91)
        orc.lib.orc_sphere_bbox(c1.ctypes.data, c2.ctypes.data if moving else None, r, boxes[i, :3].ctypes.data,
                                boxes[i, 3:].ctypes.data)
    world.build(bvh_seed=77)
    nodes = (ffi.RtbBvhNode * (2 * n - 1))()
    seed = C.c_uint64(77)
    root = orc.lib.orc_bvh_build(hs, boxes.ctypes.data, n, C.byref(seed), nodes)
    host = world.desc.contents
    assert host.n_nodes == 2 * n - 1 and host.n_hittables == n

    class OracleDesc:
        pass
    od = OracleDesc()
    od.nodes, od.hittables = nodes, hs
    key = lambda h: (tuple(h.a[:]), tuple(h.b[:]), h.radius, h.is_moving)
    assert _tree_signature(host, host.root, key) == _tree_signature(od, root, key)
    # the object array is permuted identically (world_objects.items after the in-place sorts)
    assert [key(host.hittables[i]) for i in range(n)] == [key(hs[i]) for i in range(n)]
    for i in (0, 1, n // 2, n - 1):
        assert np.array_equal(world.object_box(i), boxes[i])


def test_perlin_init_matches_oracle(pkg, orc):
    w = pkg.World.new()
    w.add_sphere((0, 0, 0), 1.0, pkg.material_spec(texture=pkg.RTB_TEX_NOISE, scale=4.0, perlin_seed=1234))
    w.build()
    host = w.desc.contents.perlins[0]
    ref = pkg.RtbPerlin()
    seed = C.c_uint64(1234)
    orc.lib.orc_perlin_init(C.byref(seed), C.byref(ref))
    assert bytes(host) == bytes(ref)


# ---- lowering + scene builders (src/main.zig:88-125, :253-312) -----------------------------------------------
def _check_desc(pkg, d):
    assert d.abi_version == pkg.RTB_ABI_VERSION and d.n_nodes == 2 * d.n_hittables - 1 and d.root == 0
    seen = np.zeros(d.n_hittables, int)
    stack = [d.root]
    visited = 0
    while stack:
        nd = d.nodes[stack.pop()]
        visited += 1
        if nd.leaf >= 0:
            seen[nd.leaf] += 1
            assert nd.left == -1 and nd.right == -1
        else:
            stack += [nd.right, nd.left]
            for ch in (d.nodes[nd.left], d.nodes[nd.right]):   # Aabb.fromBoxes: the parent contains its children
                assert all(ch.bmin[a] >= nd.bmin[a] and ch.bmax[a] <= nd.bmax[a] for a in range(3))
    assert visited == d.n_nodes and (seen == 1).all()
    for i in range(d.n_hittables):
        m = d.materials[d.hittables[i].material]
        assert m.type <= pkg.RTB_MAT_ISOTROPIC
        if m.type in (pkg.RTB_MAT_LAMBERTIAN, pkg.RTB_MAT_DIFFUSE_LIGHT):
            assert m.texture < d.n_textures


def test_book1_world(pkg):
    w = pkg.World.book1()
    d = w.desc.contents
    _check_desc(pkg, d)
    assert 470 <= d.n_hittables <= 488            # 22 x 22 grid minus rejected + ground + 3 big spheres
    types = np.bincount([d.materials[d.hittables[i].material].type for i in range(d.n_hittables)], minlength=3)
    small = d.n_hittables - 4
    assert 0.72 < (types[0] - 2) / small < 0.88 and 0.08 < (types[1] - 1) / small < 0.22    # 80 / 15 / 5 %
    radii = sorted(d.hittables[i].radius for i in range(d.n_hittables))
    assert radii[-1] == 1000 and radii[-4:-1] == [1, 1, 1]
    moving = sum(d.hittables[i].is_moving for i in range(d.n_hittables))
    assert moving == types[0] - 2                 # every small diffuse sphere is initMoving (main.zig:279-281)
    fuzz = [d.materials[d.hittables[i].material].fuzz for i in range(d.n_hittables)
            if d.materials[d.hittables[i].material].type == pkg.RTB_MAT_METAL]
    assert max(fuzz) <= 0.5
    # same seeds -> same world; different scene seed -> different world
    w2, w3, wc = pkg.World.book1(), pkg.World.book1(scene_seed=99), pkg.World.book1(checker_ground=True, moving=False)
    d2 = w2.desc.contents   # (keep the worlds alive: desc points into them)
    assert bytes(d.hittables[5]) == bytes(d2.hittables[5]) and d2.n_hittables == d.n_hittables
    d3 = w3.desc.contents
    assert any(bytes(d.hittables[i]) != bytes(d3.hittables[i]) for i in range(min(d.n_hittables, d3.n_hittables)))
    # HEAD flags: checker ground (scale 0.32 -> inv_scale 3.125) and static spheres
    dc = wc.desc.contents
    ground = [dc.hittables[i] for i in range(dc.n_hittables) if dc.hittables[i].radius == 1000][0]
    t = dc.textures[dc.materials[ground.material].texture]
    assert t.type == pkg.RTB_TEX_CHECKER and t.scale == pytest.approx(1 / 0.32)
    assert sum(dc.hittables[i].is_moving for i in range(dc.n_hittables)) == 0


def test_other_scene_builders(pkg, earthmap):
    w = pkg.World.create(pkg.RTW_SCENE_TWO_SPHERES)
    d = w.desc.contents
    _check_desc(pkg, d)
    assert d.n_hittables == 2 and d.textures[d.materials[0].texture].scale == pytest.approx(1 / 0.8)
    w = pkg.World.create(pkg.RTW_SCENE_TWO_PERLIN)
    d = w.desc.contents
    _check_desc(pkg, d)
    assert d.n_hittables == 2 and d.n_perlins == 1           # both spheres share ONE NoiseTexture (main.zig:116-117)
    assert d.textures[d.materials[0].texture].scale == 4.0
    w = pkg.World.create(pkg.RTW_SCENE_EARTH, image=earthmap)
    d = w.desc.contents
    assert d.n_hittables == 1 and d.n_nodes == 1 and d.n_images == 1
    assert (d.images[0].width, d.images[0].height, d.images[0].bytes_per_row) == (1024, 512, 4096)
    with pytest.raises(pkg.RtbError):
        pkg.World.create(pkg.RTW_SCENE_EARTH)                # needs the image
    w = pkg.World.create(pkg.RTW_SCENE_TEXTURED, image=earthmap)
    d = w.desc.contents
    _check_desc(pkg, d)
    assert {d.textures[i].type for i in range(d.n_textures)} >= {pkg.RTB_TEX_CHECKER, pkg.RTB_TEX_IMAGE, pkg.RTB_TEX_NOISE}
    w = pkg.World.create(pkg.RTW_SCENE_RANDOM_SPHERES, n_spheres=5000)
    d = w.desc.contents
    _check_desc(pkg, d)
    assert d.n_hittables == 5001


def test_earthmap_fixture_is_the_reference_decode(earthmap):
    # content/earthmap.jpg decoded by the reference's vendored stb_image v2.28 (tools/make_earthmap_fixture.py)
    assert earthmap.shape == (512, 1024, 4) and (earthmap[..., 3] == 255).all()
    import hashlib
    assert hashlib.sha256(earthmap[..., :3].tobytes()).hexdigest()[:16] == EARTHMAP_SHA16


EARTHMAP_SHA16 = "23b97e148b8f4871"


def test_write_ppm(pkg, tmp_path):
    rgba = np.array([[255, 0, 1, 255], [2, 3, 4, 255], [5, 6, 7, 255], [8, 9, 10, 255], [11, 12, 13, 255],
                     [14, 15, 16, 255]], np.uint8)
    path = str(tmp_path / "o.ppm")
    pkg.write_ppm(path, rgba, 3, 2)
    lines = open(path).read().split("\n")
    assert lines[:3] == ["P3", "3 2", "255"] and lines[3] == "255 0 1" and lines[8] == "14 15 16"   # stdout.zig:8


def test_write_png(pkg, tmp_path):
    """Decode the file with nothing but the PNG spec: signature, chunk CRCs, IHDR, zlib-inflate IDAT, filter 0."""
    import struct
    import zlib
    rng = np.random.default_rng(4)
    for w, h in ((3, 2), (257, 130)):      # the larger one needs more than one 65 535-byte stored block
        rgba = rng.integers(0, 256, (h * w, 4)).astype(np.uint8)
        path = str(tmp_path / f"o{w}.png")
        pkg.write_png(path, rgba, w, h)
        blob = open(path, "rb").read()
        assert blob[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(blob):
            n, kind = struct.unpack(">I4s", blob[pos:pos + 8])
            data = blob[pos + 8:pos + 8 + n]
            (crc,) = struct.unpack(">I", blob[pos + 8 + n:pos + 12 + n])
            assert crc == zlib.crc32(kind + data)
            chunks.append((kind, data))
            pos += 12 + n
        assert [k for k, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        assert struct.unpack(">IIBBBBB", chunks[0][1]) == (w, h, 8, 2, 0, 0, 0)
        raw = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, 1 + 3 * w)
        assert (raw[:, 0] == 0).all()
        assert np.array_equal(raw[:, 1:].reshape(h * w, 3), rgba[:, :3])


# ---- the C-ABI libraries ------------------------------------------------------------------------------------------
def _declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_\w+)\s*\(", text)))


def test_libraries_export_every_declared_symbol(pkg):
    ffi = pkg._ffi
    assert _declared("rtb.h", "rtb") == sorted(ffi.RTB_SYMBOLS)
    assert _declared("rtw_host.h", "rtw") == sorted(ffi.RTW_SYMBOLS)
    rtb, rtw = ffi.rtb(), ffi.rtw()
    for s in ffi.RTB_SYMBOLS:
        assert getattr(rtb, s) is not None
    for s in ffi.RTW_SYMBOLS:
        assert getattr(rtw, s) is not None
    assert rtb.rtb_abi_version() == ffi.RTB_ABI_VERSION


def test_pod_sizes_match_the_header(pkg):
    ffi = pkg._ffi
    assert C.sizeof(ffi.RtbHittable) == 64 and C.sizeof(ffi.RtbMaterial) == 32 and C.sizeof(ffi.RtbTexture) == 48
    assert C.sizeof(ffi.RtbBvhNode) == 40 and C.sizeof(ffi.RtbRay) == 36 and C.sizeof(ffi.RtbHit) == 52
    assert C.sizeof(ffi.RtbPerlin) == 256 * 12 + 3 * 512
    assert np.dtype(ffi.RAY_DTYPE).itemsize == 36 and np.dtype(ffi.HIT_DTYPE).itemsize == 52


def test_product_does_not_depend_on_the_oracle(pkg):
    for lib in ("librtb.so", "librtw_host.so"):
        path = os.path.join(pkg._ffi.LIB_DIR, lib)
        out = subprocess.run(["nm", "-D", path], capture_output=True, text=True).stdout
        assert "orc_" not in out
        ldd = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "liboracle" not in ldd
    pkg_dir = os.path.dirname(pkg.__file__)
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_ffi" not in text and "liboracle" not in text and "orc_" not in text, f


def test_sm100a_code_is_in_the_library(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(pkg._ffi.LIB_DIR, "librtb.so")], capture_output=True,
                         text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_fails_loudly_without_fallback(pkg):
    n = C.c_int(-1)
    rc = pkg._ffi.rtb().rtb_device_count(C.byref(n))
    if rc == pkg.RTB_OK:
        pytest.skip("a CUDA device is visible here")
    assert rc == pkg.RTB_ERR_NO_DEVICE and n.value == 0
    assert b"no CPU fallback" in pkg._ffi.rtb().rtb_last_error()
    with pytest.raises(pkg.RtbError) as e:
        pkg.Scene(pkg.World.create(pkg.RTW_SCENE_TWO_SPHERES))
    assert e.value.code == pkg.RTB_ERR_NO_DEVICE
    with pytest.raises(pkg.RtbError):
        pkg.resolve(np.ones((4, 4), np.float32))


def test_scene_desc_validation_runs_before_any_device_work(pkg):
    ffi = pkg._ffi
    w = pkg.World.create(pkg.RTW_SCENE_TWO_SPHERES)
    d = ffi.RtbSceneDesc.from_buffer_copy(bytes(w.desc.contents))
    h = C.c_void_p()
    d.abi_version = 999
    assert ffi.rtb().rtb_scene_create(C.byref(d), 0, C.byref(h)) == ffi.RTB_ERR_INVALID_ARGUMENT
    d.abi_version = ffi.RTB_ABI_VERSION
    d.root = 17
    assert ffi.rtb().rtb_scene_create(C.byref(d), 0, C.byref(h)) == ffi.RTB_ERR_INVALID_ARGUMENT
    assert ffi.rtb().rtb_scene_create(None, 0, C.byref(h)) == ffi.RTB_ERR_INVALID_ARGUMENT
    assert ffi.rtb().rtb_scene_destroy(None) == ffi.RTB_OK


def _build_abi_smoke(pkg, tmp_path):
    """include/rtb.h must be usable from plain C99 (no C++ / ctypes in between): compile tests/abi_smoke.c strictly."""
    exe = str(tmp_path / "abi_smoke")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi_smoke.c"), "-L", pkg._ffi.LIB_DIR, "-lrtb", f"-Wl,-rpath,{pkg._ffi.LIB_DIR}",
           "-lm", "-o", exe]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    return exe


def test_header_is_c99_clean_and_links(pkg, tmp_path):
    exe = _build_abi_smoke(pkg, tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    n = C.c_int(0)
    if pkg._ffi.rtb().rtb_device_count(C.byref(n)) == pkg.RTB_OK:
        assert out.returncode == 0 and "abi_smoke ok" in out.stdout, out.stdout + out.stderr
    else:   # no device here: the C program must see the loud RTB_ERR_NO_DEVICE, not a crash or a silent fallback
        assert out.returncode == 77 and "no CPU fallback" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_abi_smoke_runs_from_c(pkg, tmp_path):
    exe = _build_abi_smoke(pkg, tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "abi_smoke ok" in out.stdout, out.stdout + out.stderr


def test_zig_shim_declares_every_entry_point_of_the_header(pkg):
    """integration/rtb.zig ships as source (no Zig toolchain here): at least keep it in step with include/rtb.h —
    every function a Zig host needs is declared `pub extern fn` with the header's name, and the struct fields appear in
    the header's order."""
    zig = open(os.path.join(ROOT, "integration", "rtb.zig")).read()
    declared = set(re.findall(r"pub extern fn (rtb_\w+)\(", zig))
    test_only = {"rtb_debug_build_layout", "rtb_debug_packed_layout", "rtb_philox_device_selftest", "rtb_measure_fp32_peak"}
    assert declared == set(pkg._ffi.RTB_SYMBOLS) - test_only
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rtb.h")).read(), flags=re.S)
    for name in ("RtbHittable", "RtbMaterial", "RtbTexture", "RtbBvhNode", "RtbSceneDesc", "RtbCamera", "RtbRenderOptions",
                 "RtbRenderStats"):
        body = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", header, flags=re.S).group(1)
        c_fields = [re.sub(r"\[.*", "", f.strip().split()[-1]).lstrip("*") for f in body.split(";") if f.strip()]
        z_body = re.search(r"pub const " + name + r" = extern struct \{(.*?)\n\};", zig, flags=re.S).group(1)
        z_fields = re.findall(r"^\s*(\w+):", z_body, flags=re.M)
        assert z_fields == c_fields, (name, z_fields, c_fields)
    lower = open(os.path.join(ROOT, "integration", "lower.zig")).read()
    for fn in ("lowerTexture", "lowerMaterial", "lowerHittable", "lowerNode", "lowerImages", "lowerWorld", "lowerCamera",
               "startRender", "stopRender", "shouldStopRender", "renderOnAllGpus"):
        assert re.search(r"pub fn " + fn + r"\(", lower), fn
