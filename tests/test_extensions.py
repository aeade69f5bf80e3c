"""Extension nodes beyond the reference grammar: union / intersection / difference
(include/lolb200.h; BASELINE config C4 words its scene as "unioned/intersected
primitives", which scene.h:27-35 cannot express).

PARITY UNPINNED for these nodes: the reference has no such objects, so the only
checker is the oracle port's own restatement (oracle/lol_oracle.c: csg_dist).  What
IS pinned: a scene without extension nodes is untouched by them (every other test),
and the three definitions below are checked against plain numpy on sampled points.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol

HEAD = """materials { { shininess = 0, diffuse = (0,0,0), specular = (0,0,0), ambient = (0,0,0) },
  { shininess = 8, diffuse = (0.3,0.2,0.1), specular = (0.1,0.1,0.1), ambient = (0.3,0.2,0.1) } }
scene { camera { point = (0, 1, 4), direction = (0, -0.1, -1), fov = 90 },
  point_light { point = (3, 6, 2), diffuse_intensity = (3,3,3), specular_intensity = (3,3,3) },
"""
A = "sphere { point = (-0.4, 1, -3), radius = 1.2 }"
B = "sphere { point = (0.5, 1.2, -2.6), radius = 1 }"


def _scene(kind, lb):
    return lb.Scene.from_string(HEAD + f"  {kind} {{ material = #1, a = {A}, b = {B} }},\n  plane {{ y = -1, material = #1 }} }}")


def _oracle_sdf(scene, pts):
    d = np.zeros(len(pts), np.float32)
    for i, p in enumerate(pts):
        pt = (C.c_float * 3)(*p.tolist())
        dd, ii = C.c_float(), C.c_uint32()
        ol.port().lolo_sdf(C.cast(scene._ptr, C.c_void_p), 0, C.byref(pt), C.byref(dd), C.byref(ii))
        d[i] = dd.value
    return d


@pytest.mark.parametrize("kind", ["union", "intersection", "difference"])
def test_definitions_against_numpy(kind):
    """union = min(a, b), intersection = max(a, b), difference = max(a, -b) of the two
    sphere distances (each computed like sdSphere, sdf.h:8-10, in float32)."""
    import loltracer_b200 as lb

    scene = _scene(kind, lb)
    rng = np.random.default_rng(5)
    pts = rng.uniform(-3, 3, (500, 3)).astype(np.float32) + np.float32([0, 1, -3])

    def sphere(c, r):
        q = pts - np.float32(c)
        s = (q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]
        return np.sqrt(s).astype(np.float32) - np.float32(r)

    a, b = sphere((-0.4, 1, -3), 1.2), sphere((0.5, 1.2, -2.6), 1)
    want = {"union": np.minimum(a, b), "intersection": np.maximum(a, b), "difference": np.maximum(a, -b)}[kind]
    want = np.minimum(want, pts[:, 1] - np.float32(-1))  # the plane, object 2
    assert np.array_equal(_oracle_sdf(scene, pts), want)


@pytest.mark.parametrize("kind", ["union", "intersection", "difference"])
@pytest.mark.parametrize("variant", [1, 3])
def test_lowered_extension_nodes_equal_oracle_on_cpu(kind, variant, tmp_path):
    import loltracer_b200 as lb

    scene = _scene(kind, lb)
    src = lb.lower_cuda(scene, lb.Options.default(variant=variant, guarded_fastpath=2))
    assert f"lol_csg_{ {'union': 'union', 'intersection': 'inter', 'difference': 'diff'}[kind] }(" in src
    L = ol.cpu_sdf(tmp_path, src, f"ext_{kind}{variant}")
    rng = np.random.default_rng(6)
    pts = (rng.uniform(-3, 3, (2000, 3)) + [0, 1, -3]).astype(np.float32)
    d = np.zeros(len(pts), np.float32)
    ids = np.zeros(len(pts), np.uint32)
    fn = L.eval2 if variant == 3 else L.eval
    fn(pts.ctypes.data_as(C.c_void_p), len(pts), d.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
    assert np.array_equal(d, _oracle_sdf(scene, pts))


def test_csg_synthetic_scene_lowers_to_one_table_loop(tmp_path):
    """The 1024-primitive CSG scene: one loop over 128 rows of U(U(I,D),U(I,D)), pruned by
    bounding boxes (intersection: the smaller child box; difference: a's box)."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    scene = lb.Scene.from_string(scenegen.synthetic_scene_text(csg=True))
    assert scene.struct.n_objects == 129 and scene.flops_per_eval() == 128 * (8 * 10 + 4 + 3 * 13 + 1) + 2
    src = lb.lower_cuda(scene)
    assert "128 x U(U(I(S,S),D(S,S)),U(I(S,S),D(S,S)))" in src and "none can win" in src
    for variant in (1, 3):
        L = ol.cpu_sdf(tmp_path, lb.lower_cuda(scene, lb.Options.default(variant=variant)), f"csg{variant}")
        rng = np.random.default_rng(7)
        pts = np.concatenate([rng.uniform(-10, 10, (150, 3)) + [0, 3, -12],
                              rng.normal(0, 1.5, (150, 3)) + [0, 3, -12]]).astype(np.float32)
        d = np.zeros(len(pts), np.float32)
        ids = np.zeros(len(pts), np.uint32)
        fn = L.eval2 if variant == 3 else L.eval
        fn(pts.ctypes.data_as(C.c_void_p), len(pts), d.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
        want = _oracle_sdf(scene, pts)
        assert np.array_equal(d, want)


def test_extension_syntax_errors():
    import loltracer_b200 as lb

    with pytest.raises(lb.LolB200Error):  # a CSG node needs both children
        lb.Scene.from_string(HEAD + f"  intersection {{ a = {A} }} }}")
    with pytest.raises(lb.LolB200Error):  # smoothness belongs to smooth_union only
        lb.Scene.from_string(HEAD + f"  difference {{ smoothness = 1, a = {A}, b = {B} }} }}")
    nested = lb.Scene.from_string(
        HEAD + f"  smooth_union {{ smoothness = 0.5, a = difference {{ a = {A}, b = {B} }}, b = union {{ a = {B}, b = {A} }} }} }}")
    assert nested.struct.n_nodes == 7 and nested.struct.n_objects == 1


def test_reference_side_refuses_extension_nodes():
    """scene.c has no object type for them: the scene_parse shim (used by the headless
    host and oracle/_ref) must say so instead of handing the reference an unknown type."""
    if not ol.have_ref():
        pytest.skip("oracle/_ref not built here")
    with pytest.raises(RuntimeError):
        ol.RefScene(text=HEAD + f"  union {{ a = {A}, b = {B} }} }}")


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 3])
def test_gpu_renders_csg_scenes_like_the_oracle(variant):
    torch = pytest.importorskip("torch")
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen
    from test_gpu_parity import _check, _render

    for scene, (w, h) in [(_scene("difference", lb), (640, 360)), (_scene("intersection", lb), (333, 200)),
                          (lb.Scene.from_string(scenegen.synthetic_scene_text(csg=True)), (192, 108))]:
        got = _render(lb, scene, w, h, options=lb.Options.default(variant=variant, guarded_fastpath=2))
        _check(got, ol.port_render(scene, w, h))
        got["renderer"].close()


# ---- per-child materials (lolb200_options.child_materials) -------------------------------

CHILD_MATERIALS_SCENE = """materials {
  { shininess = 0, diffuse = (0,0,0), specular = (0,0,0), ambient = (0,0,0) },
  { shininess = 8, diffuse = (0.8,0.1,0.1), specular = (0.3,0.3,0.3), ambient = (0.8,0.1,0.1) },
  { shininess = 30, diffuse = (0.1,0.8,0.1), specular = (0.5,0.5,0.5), ambient = (0.1,0.8,0.1) },
  { shininess = 2, diffuse = (0.1,0.1,0.9), specular = (0.1,0.1,0.1), ambient = (0.1,0.1,0.9) },
  { shininess = 4, diffuse = (0.6,0.6,0.2), specular = (0.2,0.2,0.2), ambient = (0.6,0.6,0.2) } }
scene { ambient { color = (0.2, 0.2, 0.2) },
  camera { point = (0, 2, 5), direction = (0, -0.2, -1), fov = 90 },
  point_light { point = (4, 8, 3), diffuse_intensity = (1,1,1), specular_intensity = (1,1,1) },
  point_light { point = (-6, 3, 0), diffuse_intensity = (0.5,0.5,0.8), specular_intensity = (0.5,0.5,0.8) },
  smooth_union { material = #1, smoothness = 0.8,
    a = sphere { material = #2, point = (-1.2, 1, -5), radius = 1.2 },
    b = smooth_union { smoothness = 0.5,
          a = sphere { material = #3, point = (1.0, 1.2, -5.5), radius = 1.0 },
          b = box { point = (0, 2.6, -5), point2 = (0.6, 0.3, 0.6), radius = 0.1 } } },
  %s
  plane { material = #4, y = -0.5 } }"""
CSG_PART = """difference { material = #2,
    a = intersection { a = sphere { material = #3, point = (3.5, 1, -6), radius = 1.4 },
                       b = box { material = #1, point = (3.5, 1, -6), point2 = (1.1, 1.1, 1.1), radius = 0 } },
    b = sphere { material = #4, point = (3.5, 1.6, -4.9), radius = 0.8 } },"""


@pytest.mark.parametrize("csg", [False, True])
def test_child_materials_on_the_cpu_pipeline(csg, tmp_path):
    """The generated program with options.child_materials (lol_child_material: the winning top-level object's
    tree once more at the hit point, IEEE forms, with a material select per node) compiled for the host against
    the oracle's restatement (mode | 0x100): every pixel, RGB exact.  The unset material of the inner smooth
    union and of the box inherit their parent's; distances, ids and step counts do not depend on the option;
    with the option off the frame is the reference's (children ignored)."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_string(CHILD_MATERIALS_SCENE % (CSG_PART if csg else ""))
    w, h = 96, 54
    frames = {}
    for on in (0, 1):
        src = lb.lower_cuda(scene, lb.Options.default(variant=1, child_materials=on))
        assert f"#define LOL_CHILD_MATERIALS {on}" in src
        L = ol.cpu_pipeline(tmp_path, src, f"cm{int(csg)}{on}")
        got = ol.cpu_pipeline_render(L, lb, scene, w, h)
        want = ol.port_render(scene, w, h, mode=0x100 if on else 0, counts=True)
        for k in ("rgba", "id", "nprimary"):
            assert np.array_equal(got[k], want[k]), (on, k)
        assert np.array_equal(got["dist"].view(np.uint32), want["dist"].view(np.uint32))
        frames[on] = got
    assert np.array_equal(frames[0]["id"], frames[1]["id"])
    assert np.array_equal(frames[0]["dist"].view(np.uint32), frames[1]["dist"].view(np.uint32))
    blob = frames[0]["id"] == 1
    differs = frames[0]["rgba"] != frames[1]["rgba"]
    assert differs[blob].mean() > 0.5          # the blob's children show their own colours
    assert not differs[~(blob | (frames[0]["id"] == 2) & csg)].any()   # nothing else changes
    if not ol.have_ref():
        return
    # option off == the reference itself (children's materials ignored); the extension nodes are not its
    if not csg:
        ref = ol.RefScene(text=CHILD_MATERIALS_SCENE % "").probe(w, h)
        assert np.array_equal(frames[0]["rgba"], ref["rgba"])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_child_materials_on_the_gpu(variant):
    """Every kernel variant with options.child_materials against the oracle's restatement at 1280x720."""
    import torch

    import loltracer_b200 as lb
    from test_gpu_parity import _check, _render

    scene = lb.Scene.from_string(CHILD_MATERIALS_SCENE % CSG_PART)
    w, h = 1280, 720
    want = ol.port_render(scene, w, h, mode=0x100)
    got = _render(lb, scene, w, h, options=lb.Options.default(variant=variant, child_materials=1))
    _check(got, want)
    plain = _render(lb, scene, w, h, options=lb.Options.default(variant=variant))
    _check(plain, ol.port_render(scene, w, h))
    assert (plain["rgba"] != got["rgba"]).mean() > 0.02
    got["renderer"].close()
    plain["renderer"].close()
