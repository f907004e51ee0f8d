"""CPU tests of the library's host-side scene re-layout (rtb_debug_build_layout — no device needed): the threaded
per-octant node arrays the kernels walk, for all three traversal modes.  A plain-Python walker applies the kernels'
rules (slab test on pre-swapped planes, skip links, end sentinel) and must find the oracle's hits."""
import ctypes as C

import numpy as np
import pytest

END = 0xFFFFFFFF


def _layout(pkg, world, mode, octant):
    n = C.c_uint32(0)
    assert pkg._ffi.rtb().rtb_debug_build_layout(world.desc, mode, octant, None, C.byref(n)) == 0
    out = np.zeros((n.value + 1, 8), np.float32)
    assert pkg._ffi.rtb().rtb_debug_build_layout(world.desc, mode, octant, out.ctypes.data, C.byref(n)) == 0
    return out, n.value


@pytest.mark.parametrize("scene", ["book1", "quads", "textured"])
def test_layout_invariants(pkg, scene, earthmap):
    world = {"book1": lambda: pkg.World.book1(), "quads": lambda: pkg.World.create(pkg.RTW_SCENE_QUADS),
             "textured": lambda: pkg.World.create(pkg.RTW_SCENE_TEXTURED, image=earthmap)}[scene]()
    d = world.desc.contents
    ref_boxes = None
    for mode in (0, 1, 2):
        for octant in range(8):
            L, n = _layout(pkg, world, mode, octant)
            meta_all = L[:, 3].view(np.uint32)
            # mode 2 (SAH) starts with the "huge" objects (box >= half of the scene's: the ground sphere) as plain leaves
            huge = 0
            while mode == 2 and huge < n and meta_all[huge] >= (1 << 30):
                huge += 1
            rest = d.n_hittables - huge
            # modes 0/1: the host's 2k-1 nodes; mode 2: the huge leaves, then the tree of the rest with a box node in
            # front of each leaf and without the root's own box (the walk starts with the root's two subtrees)
            assert n == (d.n_nodes if mode < 2 else huge + 3 * rest - (2 if rest > 1 else 1))
            assert huge == ({"book1": 1}.get(scene, 0) if mode == 2 else 0)   # the r = 1000 ground sphere
            meta = L[:, 3].view(np.uint32)
            assert meta[n] == END                                       # the sentinel closes every octant
            kind, idx = meta[:n] >> 30, meta[:n] & 0x3FFFFFFF
            leaves = kind != 0
            assert sorted(idx[leaves].tolist()) == list(range(d.n_hittables))   # every object exactly once
            inner = np.nonzero(~leaves)[0]
            if mode < 2:
                assert (idx[inner] > inner + 2).all() and (idx[inner] <= n).all()   # skip jumps over >= 2 children
            else:
                leaf_box = inner[idx[inner] == inner + 2]          # {object box, skip past the leaf}, {leaf}
                assert leaves[leaf_box + 1].all() and len(leaf_box) == rest
                assert (idx[inner] >= inner + 2).all() and (idx[inner] <= n).all()
                # the leaf's box is the host's box for that object, padded outwards by <= 2^-20 of the scene extent
                obj = idx[leaf_box + 1]
                host = np.array([world.object_box(int(k)) for k in obj], np.float32)
                lo = np.minimum(L[leaf_box, :3], L[leaf_box, 4:7])
                hi = np.maximum(L[leaf_box, :3], L[leaf_box, 4:7])
                everything = np.array([world.object_box(k) for k in range(d.n_hittables)], np.float32)
                extent = np.abs(everything).reshape(-1, 2, 3).max(axis=(0, 1))   # of the whole scene, huge objects too
                assert (lo <= host[:, :3]).all() and (hi >= host[:, 3:]).all()
                assert (host[:, :3] - lo <= extent * 2.0 ** -20).all() and (hi - host[:, 3:] <= extent * 2.0 ** -20).all()
            # pre-swapped slabs: entry plane = max where the octant bit is set, min otherwise
            for a in range(3):
                lo, hi = L[inner, a], L[inner, 4 + a]
                assert (hi <= lo).all() if (octant >> a) & 1 else (lo <= hi).all()
            # skip(i) = end of i's subtree: the subtree of i is exactly [i, skip(i)) and nests properly
            for i in inner[:200]:
                sub = np.arange(i + 1, idx[i])
                inner_sub = sub[kind[sub] == 0]
                assert (idx[inner_sub] <= idx[i]).all()
            # every octant of a mode holds the same set of boxes (only swapped / reordered)
            boxes = sorted(map(tuple, np.concatenate([np.minimum(L[inner, :3], L[inner, 4:7]),
                                                      np.maximum(L[inner, :3], L[inner, 4:7])], axis=1).tolist()))
            if octant == 0:
                ref_boxes = boxes
            assert boxes == ref_boxes


def _walk(L, n, hittables, ray):
    """The kernels' traverse_octant, in Python (float32 arithmetic via numpy scalars)."""
    f = np.float32
    o, d, time = ray["origin"].astype(f), ray["direction"].astype(f), f(ray["time"])
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = f(1) / d
        best_t, best_obj = f(np.inf), -1
        tmin = f(0.001)
        a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        i, n_box, n_obj = 0, 0, 0
        meta = L[:, 3].view(np.uint32)
        while True:
            m = int(meta[i])
            if m < (1 << 30):
                n_box += 1
                t0 = (L[i, :3] - o) * inv
                t1 = (L[i, 4:7] - o) * inv
                lo = max(np.fmax.reduce(t0), tmin)      # fmax/fmin ignore NaN like PTX max/min
                hi = min(np.fmin.reduce(t1), best_t)
                i = m if hi <= lo else i + 1
                continue
            if m == END:
                break
            kind, obj = m >> 30, m & 0x3FFFFFFF
            n_obj += 1
            if kind != 3:
                c = L[i, :3] + (time * L[i, 4:7] if kind == 2 else f(0))
                oc = o - c
                hb = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2]
                cc = (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - L[i, 7] * L[i, 7]
                disc = hb * hb - a * cc
                if disc >= 0:
                    s = np.sqrt(disc)
                    root = (-hb - s) / a
                    if not (tmin < root < best_t):
                        root = (-hb + s) / a
                    if tmin < root < best_t:
                        best_t, best_obj = root, obj
            i += 1
    return best_obj, best_t, n_box, n_obj


def test_python_walk_of_every_layout_finds_the_oracle_hits(pkg, orc):
    world = pkg.World.book1()
    d = world.desc.contents
    cam = pkg.book1_camera(400, 10, 50).init()
    rng = np.random.default_rng(3)
    rays = orc.get_rays(cam, 9, rng.integers(0, 400 * 225, 120), 0)
    extra = np.zeros(120, dtype=rays.dtype)
    extra["origin"] = rng.uniform(-8, 8, (120, 3)).astype(np.float32)
    extra["origin"][:, 1] = np.abs(extra["origin"][:, 1]) * 0.2 + 0.05
    extra["direction"] = rng.normal(size=(120, 3)).astype(np.float32)
    extra["time"] = rng.random(120).astype(np.float32)
    extra["t_min"], extra["t_max"] = 0.001, np.inf
    rays = np.concatenate([rays, extra])
    cpu = orc.trace_rays(world.desc, rays)
    layouts = {(m, o): _layout(pkg, world, m, o) for m in (0, 1, 2) for o in range(8)}
    total = {0: 0, 1: 0, 2: 0}
    total_obj = {0: 0, 1: 0, 2: 0}
    for k, ray in enumerate(rays):
        with np.errstate(divide="ignore"):
            inv = np.float32(1) / ray["direction"]
        octant = int(inv[0] < 0) | (int(inv[1] < 0) << 1) | (int(inv[2] < 0) << 2)
        for mode in (0, 1, 2):
            L, n = layouts[(mode, octant)]
            obj, t, n_box, n_obj = _walk(L, n, d.hittables, ray)
            assert obj == cpu["object"][k], (k, mode)
            if obj >= 0:
                assert np.float32(t) == cpu["t"][k]
            if mode == 0:
                assert n_box == cpu["n_box_tests"][k]      # reference order: the very same node visits
            total[mode] += n_box
            total_obj[mode] += n_obj
    # SAH re-partition with box-tested leaves: far fewer slab tests and far fewer primitive tests
    assert total[1] <= total[0] and total[2] < 0.65 * total[0] and total_obj[2] < 0.35 * total_obj[0]


def test_octant_from_sign_bits_equals_octant_from_reciprocal():
    """push_ray bins rays by octant with ray_octant_of_direction (rtb_device.cuh): 0x80000000 <= bits(d) < 0xff800000.
    It must equal Aabb.hit's own predicate `1/d < 0` (src/aabb.zig:97) for EVERY float, or a ray would be walked with
    another octant's pre-swapped slabs."""
    rng = np.random.default_rng(0)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, -np.nan, 1e-45, -1e-45, 3.4028235e38, -3.4028235e38,
                        1.17549435e-38, -1.17549435e-38, 1.0, -1.0], np.float32)
    bits = np.concatenate([special.view(np.uint32), rng.integers(0, 2 ** 32, 2_000_000, dtype=np.uint64).astype(np.uint32)])
    d = bits.view(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = (np.float32(1.0) / d) < 0
    got = (bits - np.uint32(0x80000000)) < np.uint32(0x7F800000)      # uint32 arithmetic wraps, like the device code
    assert np.array_equal(got, want)


def test_sah_leaf_box_of_a_non_axis_aligned_quad_covers_all_four_corners(pkg):
    """Quad.init's box is fromPoints(q, q+u+v).pad() (src/objects.zig:209) — the diagonal only.  The reference never
    box-tests a leaf, the SAH layout does, so its leaf box must cover q+u and q+v too (a diamond quad
    q=(0,0,0), u=(1,1,0), v=(1,-1,0) spans y in [-1,1] but its host box is y in +-5e-5)."""
    w = pkg.World.new()
    w.add_quad((0, 0, 0), (1, 1, 0), (1, -1, 0), pkg.material_spec())
    w.add_sphere((0, 0, -5), 0.5, pkg.material_spec())
    w.add_sphere((3, 0, -5), 0.5, pkg.material_spec())
    w.build()
    d = w.desc.contents                                   # constructTree sorts the object list in place (bvh.zig:64)
    quad = [i for i in range(d.n_hittables) if d.hittables[i].type == pkg.RTB_HITTABLE_QUAD][0]
    host = w.object_box(quad)
    assert host[4] - host[1] < 1e-3                       # the reference's (short) box, kept as is on the host side
    for octant in range(8):
        L, n = _layout(pkg, w, 2, octant)
        meta = L[:n, 3].view(np.uint32)
        leaf = np.nonzero((meta >> 30 == 3) & ((meta & 0x3FFFFFFF) == quad))[0]
        assert len(leaf) == 1 and meta[leaf[0] - 1] == leaf[0] + 1     # its box node sits right in front of it
        box = L[leaf[0] - 1]
        lo, hi = np.minimum(box[:3], box[4:7]), np.maximum(box[:3], box[4:7])
        corners = np.array([(0, 0, 0), (1, 1, 0), (1, -1, 0), (2, 0, 0)], np.float32)
        assert (lo <= corners.min(axis=0)).all() and (hi >= corners.max(axis=0)).all()


# ------------------------------------------------------------------------------------------------
# RTB_TRAVERSAL_SAH16: 16-byte packed box nodes (binary16 planes) + conservative half2 slab test
# ------------------------------------------------------------------------------------------------
def _packed(pkg, world, octant):
    n = C.c_uint32(0)
    center, scale = (C.c_float * 3)(), (C.c_float * 3)()
    lib = pkg._ffi.rtb()
    assert lib.rtb_debug_packed_layout(world.desc, octant, None, C.byref(n), C.byref(center), C.byref(scale)) == 0
    out = np.zeros((n.value, 4), np.uint32)
    assert lib.rtb_debug_packed_layout(world.desc, octant, out.ctypes.data, C.byref(n), C.byref(center), C.byref(scale)) == 0
    return out, np.array(list(center), np.float32), np.array(list(scale), np.float32)


def _halves(words):
    """uint32 -> (lo half, hi half) as float64"""
    w = np.asarray(words, np.uint32)
    lo = (w & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float64)
    hi = (w >> 16).astype(np.uint16).view(np.float16).astype(np.float64)
    return lo, hi


@pytest.mark.parametrize("scene", ["book1", "quads", "cornell"])
def test_packed_layout_contains_the_f32_layout(pkg, scene):
    world = {"book1": lambda: pkg.World.book1(), "quads": lambda: pkg.World.create(pkg.RTW_SCENE_QUADS),
             "cornell": lambda: pkg.World.create(pkg.RTW_SCENE_CORNELL_BOX)}[scene]()
    for octant in range(8):
        L, n = _layout(pkg, world, 2, octant)
        S, center, scale = _packed(pkg, world, octant)
        meta = L[:, 3].view(np.uint32)
        # slot index of every entry of the f32 layout: box 1 slot, leaf 2, sentinel 1
        size = np.where((meta < (1 << 30)) | (meta == END), 1, 2)
        slot = np.concatenate([[0], np.cumsum(size)])[:-1]
        assert len(S) == size.sum() and S[slot[n], 3] == END
        for k in range(n):
            s = S[slot[k]]
            if meta[k] < (1 << 30):
                assert s[3] == slot[meta[k]]                                   # skip link in slot units
                e, x = _halves(s[:3])
                e, x = e * scale + center, x * scale + center                   # back to world space (float64)
                lo, hi = np.minimum(e, x), np.maximum(e, x)
                blo, bhi = np.minimum(L[k, :3], L[k, 4:7]), np.maximum(L[k, :3], L[k, 4:7])
                assert (lo <= blo).all() and (hi >= bhi).all()                   # contains the f32 box ...
                assert (blo - lo <= scale * 2.0 ** -9).all() and (hi - bhi <= scale * 2.0 ** -9).all()   # ... tightly
                for a in range(3):                                               # and keeps the pre-swapped order
                    assert (x[a] <= e[a]) if (octant >> a) & 1 else (e[a] <= x[a])
            else:
                assert s[3] == meta[k] and (s[:3] == L[k, :3].view(np.uint32)).all()
                assert S[slot[k] + 1, 3] & 0x80000000                            # a leaf's second slot is never a link
                if meta[k] >> 30 != 3:
                    assert np.float32(-abs(L[k, 7])).view(np.uint32) == S[slot[k] + 1, 3]
                    assert (S[slot[k] + 1, :3] == L[k, 4:7].view(np.uint32)).all()


def _walk_packed(S, center, scale, ray):
    """traverse_packed (rtb_device.cuh) in Python: binary16 constants, one rounding per packed FMA."""
    f = np.float32
    h = lambda v: np.float64(np.float16(v))                      # round to binary16 (nearest even), back to float64
    o, d, time = ray["origin"].astype(f), ray["direction"].astype(f), f(ray["time"])
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        I, N = np.zeros((3, 2)), np.zeros((3, 2))
        terms = []
        for a in range(3):
            on = (o[a] - center[a]) * (f(1) / scale[a])
            inv = scale[a] * (f(1) / d[a])
            nod = -(on * inv)
            terms.append((inv, nod, abs(inv) + abs(nod)))
        m = max([t[2] for t in terms if t[2] < 3e38] + [f(0)])
        e = (int(np.float32(m).view(np.uint32)) >> 23) - 127
        sigma = f(2.0) ** -(min(e - 12, 120) if e > 12 else 0)          # t' = sigma * t fits binary16
        for a, (inv, nod, mag) in enumerate(terms):
            if not (mag < 3e38):
                I[a], N[a] = (0, 0), (-np.inf, -np.inf)
                continue
            k = f(1.01) / f(2048)
            E = mag * (f(1.01) / f(2048)) + f(1e-7)
            I[a] = h(sigma * (inv * (f(1) - k))), h(sigma * -(inv * (f(1) + k)))
            N[a] = h(sigma * (nod * (f(1) - k) - E)), h(sigma * (-(nod * (f(1) + k)) - E))
        tmin = f(0.001)
        rd = lambda v: np.float64(np.nextafter(np.float16(v), np.float16(-np.inf))) if np.float64(np.float16(v)) > v else np.float64(np.float16(v))
        ru = lambda v: np.float64(np.nextafter(np.float16(v), np.float16(np.inf))) if np.float64(np.float16(v)) < v else np.float64(np.float16(v))
        K = [rd(sigma * tmin), -np.inf]
        best_t, best_obj = f(np.inf), -1
        aa = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        i, n_box, n_obj = 0, 0, 0
        while True:
            s = S[i]
            m = int(s[3])
            if m < (1 << 30):
                n_box += 1
                e, x = _halves(s[:3])
                r = [K[0], K[1]]
                for a in range(3):
                    te, tx = h(e[a] * I[a][0] + N[a][0]), h(x[a] * I[a][1] + N[a][1])
                    r[0] = np.fmax(r[0], te)
                    r[1] = np.fmax(r[1], tx)
                i = m if r[1] >= -r[0] else i + 1
                continue
            if m == END:
                break
            kind, obj = m >> 30, m & 0x3FFFFFFF
            n_obj += 1
            assert kind != 3
            c1 = s[:3].view(f)
            s1 = S[i + 1]
            cv, radius = s1[:3].view(f), s1[3:4].view(f)[0]
            c = c1 + (time * cv if kind == 2 else f(0))
            oc = o - c
            hb = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2]
            cc = (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - radius * radius
            disc = hb * hb - aa * cc
            if disc >= 0:
                sq = np.sqrt(disc)
                root = (-hb - sq) / aa
                if not (tmin < root < best_t):
                    root = (-hb + sq) / aa
                if tmin < root < best_t:
                    best_t, best_obj = root, obj
                    K[1] = -ru(sigma * root)
            i += 2
    return best_obj, best_t, n_box, n_obj


def test_python_walk_of_the_packed_layout_finds_the_oracle_hits(pkg, orc):
    world = pkg.World.book1()
    cam = pkg.book1_camera(400, 10, 50).init()
    rng = np.random.default_rng(5)
    rays = orc.get_rays(cam, 9, rng.integers(0, 400 * 225, 150), 0)
    extra = np.zeros(250, dtype=rays.dtype)
    extra["origin"] = rng.uniform(-9, 9, (250, 3)).astype(np.float32)
    extra["origin"][:, 1] = np.abs(extra["origin"][:, 1]) * 0.1 + 0.01
    extra["direction"] = rng.normal(size=(250, 3)).astype(np.float32)
    extra["direction"][:40, rng.integers(0, 3)] = 0.0                 # axis-parallel rays: that axis is dropped
    extra["direction"][40:60, 1] *= 1e-6                              # nearly parallel: |1/d| overflows binary16
    extra["origin"][60:80] *= 40.0                                    # far origins: the margin E grows with |o|
    extra["time"] = rng.random(250).astype(np.float32)
    extra["t_min"], extra["t_max"] = 0.001, np.inf
    rays = np.concatenate([rays, extra])
    cpu = orc.trace_rays(world.desc, rays)
    layouts = {o: _packed(pkg, world, o) for o in range(8)}
    f32_layouts = {o: _layout(pkg, world, 2, o) for o in range(8)}
    tot16 = tot32 = 0
    for k, ray in enumerate(rays):
        bits = ray["direction"].view(np.uint32)
        octant = sum((1 << a) for a in range(3) if np.uint32(bits[a] - np.uint32(0x80000000)) < 0x7F800000)
        S, center, scale = layouts[octant]
        obj, t, n_box, n_obj = _walk_packed(S, center, scale, ray)
        assert obj == cpu["object"][k], k
        if obj >= 0:
            assert np.float32(t) == cpu["t"][k]
        L, n = f32_layouts[octant]
        _, _, n_box32, n_obj32 = _walk(L, n, None, ray)
        assert n_obj >= n_obj32                                        # a superset of the leaves is tested
        if k < 150 or k >= 150 + 80:                                   # ordinary rays (not the stress cases above)
            tot16 += n_box
            tot32 += n_box32
        if 150 + 60 <= k < 150 + 80:                                   # far origins: scaled to fit binary16, not dropped
            assert n_box <= n_box32 + 8, (k, n_box, n_box32)
    assert tot32 <= tot16 <= 1.08 * tot32                              # ... at a small price in extra visits


def test_packed_layout_of_a_scene_too_large_for_shared_memory(pkg, orc):
    """Since r3h the packed SAH16 layout is also built for scenes whose layout does not fit shared memory (it is then walked
    from global memory).  Its binary16 planes then span a much larger extent relative to the objects: the Python model of
    the device arithmetic must still find the oracle's hit for every ray (conservative), at a moderate price in visits."""
    world = pkg.World.create(pkg.RTW_SCENE_RANDOM_SPHERES, n_spheres=6000)
    S0, _, _ = _packed(pkg, world, 0)
    assert S0.shape[0] * 16 > 96 * 1024                                # one octant does not fit the 96 KB staging limit
    cam = pkg.million_camera(320, 4, 50).init()
    rng = np.random.default_rng(11)
    rays = orc.get_rays(cam, 3, rng.integers(0, cam.image_width * cam.image_height, 120), 0)
    inside = np.zeros(120, dtype=rays.dtype)
    inside["origin"] = rng.uniform(-30, 30, (120, 3)).astype(np.float32)
    inside["origin"][:, 1] = rng.uniform(0.1, 5, 120).astype(np.float32)
    inside["direction"] = rng.normal(size=(120, 3)).astype(np.float32)
    inside["time"] = rng.random(120).astype(np.float32)
    inside["t_min"], inside["t_max"] = 0.001, np.inf
    rays = np.concatenate([rays, inside])
    cpu = orc.trace_rays(world.desc, rays)
    assert (cpu["object"] >= 0).mean() > 0.3
    layouts = {}
    f32_layouts = {}
    tot16 = tot32 = 0
    for k, ray in enumerate(rays):
        bits = ray["direction"].view(np.uint32)
        octant = sum((1 << a) for a in range(3) if int(bits[a]) >= 0x80000000 and int(bits[a]) - 0x80000000 < 0x7F800000)
        if octant not in layouts:
            layouts[octant] = _packed(pkg, world, octant)
            f32_layouts[octant] = _layout(pkg, world, 2, octant)
        S, center, scale = layouts[octant]
        obj, t, n_box, n_obj = _walk_packed(S, center, scale, ray)
        assert obj == cpu["object"][k], k
        if obj >= 0:
            assert np.float32(t) == cpu["t"][k]
        L, n = f32_layouts[octant]
        _, _, n_box32, n_obj32 = _walk(L, n, None, ray)
        assert n_obj >= n_obj32
        tot16 += n_box
        tot32 += n_box32
    assert tot32 <= tot16 <= 1.35 * tot32, (tot16, tot32)
