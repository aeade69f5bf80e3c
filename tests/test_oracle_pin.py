"""Pins the oracle (oracle/lol_oracle.c) before anything is compared against it.

The reference ships no tests or golden vectors, so the pins are
outputs of the reference itself: tests/golden/ was produced by the reference's own
naive_renderer.c compiled unmodified (tests/golden/make_golden.py), and where
oracle/_ref travelled it is also run live here.
"""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import EXAMPLES, ROOT

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def frames():
    return np.load(os.path.join(GOLD, "frames.npz"))


@pytest.fixture(scope="module")
def hashes():
    return json.load(open(os.path.join(GOLD, "hashes.json")))


def _same(got, frames, key):
    assert np.array_equal(got["dist"].view(np.uint32), frames[key + "_dist"].view(np.uint32)), key
    assert np.array_equal(got["id"], frames[key + "_id"].astype(np.uint32)), key
    assert np.array_equal(got["rgba"], frames[key + "_rgba"]), key


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("size", [(160, 90), (96, 64)])
def test_port_equals_golden_frames(name, size, frames, scenes_dir):
    """distance, object id and packed pixel of every pixel, bit for bit."""
    import loltracer_b200 as lb

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    _same(ol.port_render(scene, w, h), frames, f"{name}_{w}x{h}")


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("size", [(320, 240), (1920, 1080)])
def test_port_equals_reference_frame_hashes(name, size, hashes, scenes_dir):
    """320x240 is main.c's default window (main.c:136-137); the hashes are also in SURVEY.md 8c."""
    import loltracer_b200 as lb

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    assert ol.frame_hash(ol.port_render(scene, w, h)["rgba"]) == hashes[f"{name}_{w}x{h}"]


def test_port_4k_hash_scene(hashes, scenes_dir):
    """BASELINE size, cheapest scene (about 2 s on 8 cores)."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene.lol"))
    assert ol.frame_hash(ol.port_render(scene, 3840, 2160)["rgba"]) == hashes["scene_3840x2160"]


@pytest.mark.parametrize("k", [0, 16, 32, 48])
def test_port_orbit_cameras(k, frames, hashes, scenes_dir):
    """Config C5: moved cameras (main.c mutates scene->camera between frames)."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    cam = scenegen.orbit_camera(scene.camera, k, 64)
    _same(ol.port_render(scene, 160, 90, camera=cam), frames, f"orbit{k}_160x90")
    assert ol.frame_hash(ol.port_render(scene, 640, 360, camera=cam)["rgba"]) == hashes[f"orbit{k}_640x360"]


def test_port_synthetic_scene(frames, hashes):
    """Config C4: 1024 spheres in 128 smooth-union trees; also pins the generator's text."""
    import hashlib
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    text = scenegen.synthetic_scene_text()
    assert hashlib.sha256(text.encode()).hexdigest() == hashes["synthetic_text_sha256"]
    scene = lb.Scene.from_string(text)
    assert scene.struct.n_objects == 129 and scene.struct.n_nodes == 128 * 15 + 1
    assert scene.flops_per_eval() == 22018
    _same(ol.port_render(scene, 96, 54), frames, "synthetic_96x54")


@pytest.mark.parametrize("name", EXAMPLES)
def test_port_equals_live_reference(name, scenes_dir):
    """Where the compiled reference travelled: random window sizes and cameras."""
    import loltracer_b200 as lb

    if not ol.have_ref():
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(hash(name) % 2**32)
    path = os.path.join(scenes_dir, name + ".lol")
    scene = lb.Scene.from_file(path)
    rs = ol.RefScene(path=path)
    for _ in range(3):
        w, h = int(rng.integers(17, 200)), int(rng.integers(13, 120))
        cam = scene.camera
        p = (np.asarray(list(cam.point), np.float32) + rng.normal(0, 1.0, 3).astype(np.float32))
        d = (np.asarray(list(cam.direction), np.float32) + rng.normal(0, 0.2, 3).astype(np.float32))
        cam2 = lb.Camera.make(p.tolist(), d.tolist(), cam.fov)
        rs.set_camera(p.tolist(), d.tolist())
        want = rs.probe(w, h)
        got = ol.port_render(scene, w, h, camera=cam2)
        assert np.array_equal(got["dist"].view(np.uint32), want["dist"].view(np.uint32))
        assert np.array_equal(got["id"], want["id"])
        assert np.array_equal(got["rgba"], want["rgba"])
        px, _ = rs.render_protocol(w, h, threads=3)  # the unmodified render_thread itself
        assert np.array_equal(px, want["rgba"])


def test_port_sdf_equals_reference_sdf(scenes_dir):
    import ctypes as C
    import loltracer_b200 as lb

    if not ol.have_ref():
        pytest.skip("oracle/_ref not present")
    rng = np.random.default_rng(7)
    for name in EXAMPLES:
        path = os.path.join(scenes_dir, name + ".lol")
        scene = lb.Scene.from_file(path)
        rs = ol.RefScene(path=path)
        for p in rng.uniform(-12, 12, (400, 3)).astype(np.float32):
            pt = (C.c_float * 3)(*p.tolist())
            d1, i1, d2, i2 = C.c_float(), C.c_uint32(), C.c_float(), C.c_uint32()
            ol.ref().lolref_sdf(rs.ptr, C.byref(pt), C.byref(d1), C.byref(i1))
            ol.port().lolo_sdf(C.cast(scene._ptr, C.c_void_p), 0, C.byref(pt), C.byref(d2), C.byref(i2))
            assert (np.float32(d1.value).view(np.uint32), i1.value) == (np.float32(d2.value).view(np.uint32), i2.value)


def test_jit_semantics_mode(scenes_dir):
    """Appendix C of SURVEY.md: the JIT renderer cannot draw boxes, otherwise agrees
    with the naive one up to last-ulp differences in sminf."""
    import loltracer_b200 as lb

    s1 = lb.Scene.from_file(os.path.join(scenes_dir, "scene.lol"))
    naive, jit = ol.port_render(s1, 160, 90, mode=0), ol.port_render(s1, 160, 90, mode=1)
    assert (naive["id"] == 3).any() and not (jit["id"] == 3).any()  # the box is object 3
    s2 = lb.Scene.from_file(os.path.join(scenes_dir, "scene2.lol"))  # spheres + plane only
    naive, jit = ol.port_render(s2, 160, 90, mode=0), ol.port_render(s2, 160, 90, mode=1)
    assert np.array_equal(naive["rgba"], jit["rgba"])
    s4 = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    naive, jit = ol.port_render(s4, 160, 90, mode=0), ol.port_render(s4, 160, 90, mode=1)
    cmp = ol.compare_frames(jit["rgba"], jit["id"], naive["rgba"], naive["id"])
    assert cmp["mask_agree"] > 0.999 and cmp["max_rgb_err"] <= 2


@pytest.mark.parametrize("name", EXAMPLES)
def test_specialised_sdf_mode_equals_naive(name, scenes_dir, tmp_path):
    """Mode 2 -- the JIT-equivalent CPU baseline: the oracle's pipeline around the
    lowering's straight-line distance code compiled by g++ -- renders the naive
    renderer's frame bit for bit (same semantics, no interpreter dispatch)."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    keep = ol.specialised_sdf(scene, tmp_path, name)
    naive = ol.port_render(scene, 200, 112, mode=0, counts=True)
    spec = ol.port_render(scene, 200, 112, mode=2, counts=True)
    for key in ("rgba", "id", "nprimary", "nshadow"):
        assert np.array_equal(naive[key], spec[key]), key
    assert np.array_equal(naive["dist"].view(np.uint32), spec["dist"].view(np.uint32))
    assert keep is not None


def test_reference_built_without_optimisation_renders_the_same_frame(scenes_dir):
    """The reference's Makefile passes no -O flag; oracle/_ref/liblolref_O0.so is that build.
    Same frame as the -O2 build the other pins use (IEEE semantics do not depend on -O)."""
    if not (ol.have_ref() and os.path.exists(ol.REF_O0_PATH)):
        pytest.skip("oracle/_ref not built here")
    path = os.path.join(scenes_dir, "scene4.lol")
    a = ol.RefScene(path=path).probe(96, 54)
    b = ol.RefScene(path=path, lib=ol.ref(ol.REF_O0_PATH)).probe(96, 54)
    assert np.array_equal(a["rgba"], b["rgba"]) and np.array_equal(a["id"], b["id"])
    assert np.array_equal(a["dist"].view(np.uint32), b["dist"].view(np.uint32))
