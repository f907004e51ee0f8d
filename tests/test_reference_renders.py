"""The oracle and the CUDA path against OUTPUTS OF THE REFERENCE PROGRAM: the two 800 x 450 renders committed at the
reference's root (image.png, image2.png; fixture tests/golden/ref_renders_top160.npz, made by
tools/make_ref_render_fixture.py).

The reference placed its small spheres with an unseeded generator (src/main.zig:262-300 + std.crypto.random), so those
frames cannot be reproduced pixel for pixel.  Above the horizon, though, they show only what the reference fixes:
  * the sky: Camera defaults (src/camera.zig:70-91) -> getRay (:169-180) -> the gradient (:204-206) -> toGamma2 and the
    8-bit truncation (src/color.zig) — a deterministic function of the pixel, no noise;
  * the silhouettes of the three big spheres (src/main.zig:303-309; Sphere.hit, src/objects.zig) through the thin-lens
    camera (defocus 0.6, focus 10);
  * the upper cap of the metal sphere (albedo 0.7/0.6/0.5, fuzz 0): one mirror bounce into the sky
    (Metal.scatter + vec3.reflect, src/material.zig);
  * the band of the glass sphere that shows the sky upside down (Dielectric.scatter, refract, Schlick), as block means;
  * the upper part of the brown lambertian sphere, which sees only sky (Lambertian.scatter), as block means.
The same Book-1 scene (our seeded small spheres never rise above y = 0.4, i.e. stay below these rows) rendered by the
oracle — and by the CUDA path — must reproduce those regions to within 8-bit rounding.  This is the independent anchor
of the oracle for the camera / sphere / all three Book-1 materials / colour pipeline; the BVH visiting order and tie
rule, textures, quads and media have no such anchor."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS, W, H = 160, 800, 450


@pytest.fixture(scope="module")
def ref_renders():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_renders_top160.npz"))
    return {k: z[k].astype(np.int32) for k in ("image_png", "image2_png")}


def _dilate(mask, n):
    m = mask.copy()
    for _ in range(n):
        g = m.copy()
        g[1:] |= m[:-1]
        g[:-1] |= m[1:]
        g[:, 1:] |= m[:, :-1]
        g[:, :-1] |= m[:, 1:]
        m = g
    return m


def _objects(img):
    """Pixels that are not sky: they differ from the colour at the left edge of their row (sky in every frame)."""
    return np.abs(img - img[:, :1, :]).max(axis=2) > 6


def _check_against_reference_render(img, ref, name):
    """img, ref: [ROWS, W, 3] int32."""
    d = np.abs(img - ref).max(axis=2)
    m_img, m_ref = _objects(img), _objects(ref)
    above = np.arange(ROWS)[:, None] < 96  # rows that hold nothing but sky and the three big spheres
    # 1. sky, away from any silhouette: the reference's value +- 1 (image2.png sits one count above image.png on ~40 % of
    #    the sky: two code versions of the reference; +-1 covers both)
    #    (rows 0..87: the defocused horizon begins to show in the last rows of the fixture)
    sky = ~_dilate(m_img | m_ref, 2)
    sky[88:] = False
    assert sky.sum() > 50_000
    assert d[sky].max() <= 1, (name, d[sky].max())
    # 2. the silhouettes of the three big spheres: same pixels (the glass sphere's rim refracts the unseeded small
    #    spheres, hence not exactly 1)
    iou = (m_img & m_ref & above).sum() / ((m_img | m_ref) & above).sum()
    assert m_ref.sum() > 15_000 and iou >= 0.97, (name, iou)
    # 3. the upper half of the metal sphere away from its rim and from the glass sphere behind it (columns 500..620,
    #    rows 70..150): sky seen in a mirror
    cap = np.zeros_like(m_ref)
    cap[70:150, 500:620] = True
    cap &= ~_dilate(~(m_img & m_ref), 3)
    assert cap.sum() > 9_000
    assert d[cap].max() <= 2, (name, d[cap].max())
    # 4. inside the glass sphere, the band where it shows the sky upside down (rows 140..155; 8 x 8 blocks whose mean does
    #    not depend on where the small spheres are — found by rendering three differently seeded scenes): two refractions
    #    and the Schlick-weighted choice between reflecting and refracting (Dielectric.scatter, vec3.refract,
    #    src/material.zig:80-117) reproduce the reference's block means to 8-bit rounding + sampling noise
    for (y0, x0) in ((140, 324), (140, 332), (140, 356), (140, 364), (140, 372), (148, 364), (148, 372)):
        got = img[y0:y0 + 8, x0:x0 + 8].reshape(-1, 3).mean(axis=0)
        want = ref[y0:y0 + 8, x0:x0 + 8].reshape(-1, 3).mean(axis=0)
        assert want[2] > 240 and np.abs(got - want).max() <= 2.5, (name, y0, x0, got, want)
    # 5. the upper part of the brown lambertian sphere (albedo 0.4 / 0.2 / 0.1, src/main.zig:306) behind the glass one: its
    #    surface points see nothing but sky, so a pixel is albedo x the sky gradient averaged over the scatter distribution
    #    (Lambertian.scatter: normal + randomUnitVector, src/material.zig:43-54) — block means (seed-independent, as
    #    above) equal the reference's to 8-bit rounding; a uniform-hemisphere scatter would be off by ~4 counts
    for y0, xs in ((44, (304, 312, 320)), (52, (296, 304, 312)), (60, (288, 296, 304)), (68, (280, 288, 296)),
                   (76, (280, 288, 296))):
        for x0 in xs:
            got = img[y0:y0 + 8, x0:x0 + 8].reshape(-1, 3).mean(axis=0)
            want = ref[y0:y0 + 8, x0:x0 + 8].reshape(-1, 3).mean(axis=0)
            assert want[0] > want[2] + 25 and np.abs(got - want).max() <= 2.5, (name, y0, x0, got, want)
    return float(iou), float(d[sky].mean()), int(d[cap].max())


def _top_rows(rgba):
    return rgba.reshape(-1, 4)[: ROWS * W, :3].reshape(ROWS, W, 3).astype(np.int32)


def test_oracle_reproduces_the_reference_renders(pkg, orc, ref_renders):
    world = pkg.World.book1(moving=False)
    cam = orc.camera_init(pkg.book1_camera(W, 32, 50))
    assert (cam.image_width, cam.image_height) == (W, H)
    o = pkg.render_options(seed=7, pixel_begin=0, pixel_count=ROWS * W)
    _, rgba, _ = orc.render(world.desc, cam, o, n_threads=8)
    img = _top_rows(rgba)
    for name, ref in ref_renders.items():
        _check_against_reference_render(img, ref, name)
    # image.png is the code version closest to HEAD's colour pipeline: its sky is reproduced EXACTLY on > 95 % of the pixels
    # (the rest sit on an 8-bit rounding boundary of the gradient and move with the jitter samples)
    sky = ~_dilate(_objects(img) | _objects(ref_renders["image_png"]), 2)
    sky[88:] = False
    assert (np.abs(img - ref_renders["image_png"]).max(axis=2)[sky] == 0).mean() > 0.95


@pytest.mark.gpu
@pytest.mark.parametrize("integrator,traversal", [(0, 0), (1, 0), (1, 3)])
def test_cuda_path_reproduces_the_reference_renders(pkg, ref_renders, integrator, traversal):
    """The product itself (megakernel / wavefront in reference order, and the benchmarked wavefront + SAH16) against the
    reference's own renders, through rtb_render."""
    world = pkg.World.book1(moving=False)
    scene = pkg.Scene(world)
    cam = pkg.book1_camera(W, 32, 50).init()
    o = pkg.render_options(seed=7, pixel_begin=0, pixel_count=ROWS * W, integrator=integrator, traversal=traversal)
    _, rgba, _ = scene.render(cam, o)
    img = _top_rows(rgba)
    for name, ref in ref_renders.items():
        _check_against_reference_render(img, ref, name)
    scene.close()
