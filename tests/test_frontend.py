"""The .lol front-end against the reference's scene.c (through oracle/_ref) and its
error behaviour (scene-lexer.l, scene-parser.y, scene.c:104-292)."""
import ctypes as C
import os

import pytest

import oracle_lib as ol
from conftest import EXAMPLES


def scene_bytes(scene):
    from loltracer_b200 import api

    s = scene.struct
    return b"|".join([
        bytes(C.string_at(s.materials, s.n_materials * C.sizeof(api.Material))),
        bytes(s.ambient_color), bytes(C.string_at(s.lights, s.n_lights * C.sizeof(api.Light))),
        bytes(C.string_at(s.nodes, s.n_nodes * C.sizeof(api.Object))),
        bytes(C.string_at(s.objects, s.n_objects * 4)), bytes(s.camera)])


@pytest.mark.parametrize("name", EXAMPLES + ["synthetic"])
def test_scene_equals_reference_structs(name, scenes_dir):
    """Our parser + semantics == the reference's scene.c, bit for bit (camera
    normalisation and the deg->rad conversion included)."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    if not ol.have_ref():
        pytest.skip("oracle/_ref not built here")
    text = (scenegen.synthetic_scene_text() if name == "synthetic"
            else open(os.path.join(scenes_dir, name + ".lol")).read())
    ours = lb.Scene.from_string(text)
    theirs = ol.RefScene(text=text).flatten()
    assert scene_bytes(ours) == scene_bytes(theirs)


def test_example_shapes(scenes_dir):
    import loltracer_b200 as lb

    sc = lb.Scene.from_file(os.path.join(scenes_dir, "scene.lol"))
    s = sc.struct  # a view into sc: keep sc alive
    assert (s.n_materials, s.n_lights, s.n_objects, s.n_nodes) == (5, 1, 4, 4)
    assert [s.nodes[s.objects[i]].type for i in range(4)] == [3, 3, 4, 5]  # sphere, sphere, box, plane
    s4 = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    assert (s4.struct.n_objects, s4.struct.n_nodes, s4.struct.n_lights) == (2, 10, 2)
    assert s4.flops_per_eval() == 105  # 5 spheres, 4 smooth nodes, 1 plane, 2 top-level
    assert lb.Scene.from_file(os.path.join(scenes_dir, "scene.lol")).flops_per_eval() == 45
    assert lb.Scene.from_file(os.path.join(scenes_dir, "scene2.lol")).flops_per_eval() == 35
    assert lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol")).flops_per_eval() == 36


MINI = """materials { { shininess = 1, diffuse = (0,0,0), specular = (0,0,0), ambient = (0,0,0) } }
scene { %s }"""


def test_lexer_quirks():
    """Dash spellings, unknown characters dropped, longest keyword match."""
    import loltracer_b200 as lb

    a = lb.Scene.from_string(MINI % "point-light { point = (1,2,3), diffuse-intensity = (1,1,1), "
                                    "specular_intensity = (2,2,2) }, plane { y = -.5, material = #0 }")
    b = lb.Scene.from_string(MINI % "point_light { point = (1,2,3) ;, diffuse_intensity = (1,1,1), "
                                    "specular-intensity = (2,2,2) },\n\n plane { y = -.5, material=#0 } ?!")
    assert scene_bytes(a) == scene_bytes(b)
    assert a.struct.nodes[0].point[1] == -0.5 and a.struct.nodes[0].point[0] == 0.0


def test_defaults_without_camera():
    """scene_new() (scene.c:44-58): camera at the origin looking down +z, fov pi/2."""
    import loltracer_b200 as lb

    sc = lb.Scene.from_string(MINI % "plane { y = 0, material = #0 }")
    s = sc.struct
    assert list(s.camera.point) == [0, 0, 0] and list(s.camera.direction) == [0, 0, 1]
    assert abs(s.camera.fov - 1.5707964) < 1e-6


@pytest.mark.parametrize("body", [
    "sphere { point = (0,0,0), radius = 1, material = #0, y = 3 }",  # unknown sphere property
    "sphere { point = (0,0), radius = 1, material = #0 }",            # v3 needs 3 numbers
    "sphere { point = 3, radius = 1, material = #0 }",                # number where a list is due
    "sphere { point = (0,0,0), radius = 1, material = #1 }",          # material out of range
    "smooth_union { material = #0, smoothness = 1, a = camera { fov = 1 }, "
    "b = sphere { radius = 1 } }",                                     # camera is not an object
    "smooth_union { material = #0, smoothness = 1, a = sphere { radius = 1 } }",  # b missing
    "sphere { point = (0,0,0), radius = 1, material = #0 },",         # trailing comma
    "sphere { point = (0,0,0) radius = 1 }",                          # missing comma
])
def test_rejected_inputs(body):
    import loltracer_b200 as lb

    with pytest.raises(lb.LolB200Error) as e:
        lb.Scene.from_string(MINI % body)
    assert e.value.code == -1


def test_missing_file():
    import loltracer_b200 as lb

    with pytest.raises(lb.LolB200Error):
        lb.Scene.from_file("/nonexistent/scene.lol")


def test_camera_basis_matches_reference_ray(scenes_dir):
    """The hoisted camera basis reproduces get_camera_ray() (checked end to end by the
    golden frames); here: orthogonality and the atan (sic) half extent."""
    import math
    import loltracer_b200 as lb

    cam = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol")).camera
    cb = lb.camera_basis(cam, 320, 240)
    dot = sum(cb.right[i] * cb.dir[i] for i in range(3))
    assert abs(dot) < 1e-6 and abs(cb.right[1]) < 1e-7
    assert abs(cb.height - math.atan(cam.fov / 2)) < 1e-6
    assert abs(cb.width - cb.height * 320 / 240) < 1e-6


def test_deep_nesting_is_a_parse_error_not_a_stack_overflow():
    """`a = smooth_union { a = smooth_union { ...` recurses once per level in the parser and in
    every tree walk behind it; the reference's bison parser gives up at YYMAXDEPTH.  Depth is
    capped (lol_ast.h: LOLB200_MAX_NESTING) and the input is refused with LOLB200_EPARSE."""
    import loltracer_b200 as lb

    def nested(depth):
        leaf = "sphere { point = (0,0,-5), radius = 1 }"
        obj = leaf
        for _ in range(depth):
            obj = "smooth_union { smoothness = 0.5, a = %s, b = %s }" % (obj, leaf)
        return MINI % obj

    ok = lb.Scene.from_string(nested(200))
    assert ok.struct.n_nodes == 401
    with pytest.raises(lb.LolB200Error) as e:
        lb.Scene.from_string(nested(50000))
    assert e.value.code == -1 and "nested deeper" in str(e.value)


def test_many_distinct_smoothness_values_are_each_proved_once():
    """The division-by-constant proof (2^23 significands per distinct k) is cached for the life
    of the process, however many distinct k a scene uses: the second lowering costs no proofs."""
    import time

    import loltracer_b200 as lb

    objs = ", ".join("smooth_union { smoothness = %.3f, a = sphere { point = (%d,0,-8), radius = 1 }, "
                     "b = sphere { point = (%d,1,-8), radius = 1 } }" % (0.2 + 0.013 * i, i, i) for i in range(24))
    scene = lb.Scene.from_string(MINI % objs)
    opt = lb.Options.default(variant=1, guarded_fastpath=2)
    t0 = time.perf_counter()
    a = lb.lower_cuda(scene, opt)
    t1 = time.perf_counter()
    b = lb.lower_cuda(scene, opt)
    t2 = time.perf_counter()
    assert a == b and "#define LOL_DIV_CONST 1" in a
    assert (t2 - t1) < 0.5 * (t1 - t0) or (t2 - t1) < 0.05
