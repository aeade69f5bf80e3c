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
