"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances are BASELINE.json's: hit/miss mask agreement >= 99.99 %, RGB within
1/255 per channel on agreeing hits.  In the exact arithmetic mode the march is
expected to be bit-identical (distance and id of every pixel); only powf differs
between glibc and CUDA, hence the 1/255.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import EXAMPLES

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _render(lb, scene, w, h, camera=None, options=None, shard=None, counts=False):
    r = lb.Renderer(scene, options, device=0)
    dev = "cuda:0"
    frame = torch.zeros((h, w), dtype=torch.int32, device=dev)
    dist = torch.zeros((h, w), dtype=torch.float32, device=dev)
    ids = torch.zeros((h, w), dtype=torch.int32, device=dev)
    npr = torch.zeros((h, w), dtype=torch.int16, device=dev)
    nsh = torch.zeros((h, w), dtype=torch.int16, device=dev)
    aux = lb.Aux(dist=dist.data_ptr(), id=ids.data_ptr(), primary_steps=npr.data_ptr(),
                 shadow_steps=nsh.data_ptr())
    r.render_device(frame.data_ptr(), w, h, camera=camera, aux=aux, shard=shard,
                    stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = dict(rgba=frame.cpu().numpy().view(np.uint32), dist=dist.cpu().numpy(),
               id=ids.cpu().numpy().view(np.uint32), nprimary=npr.cpu().numpy().view(np.uint16),
               nshadow=nsh.cpu().numpy().view(np.uint16), renderer=r)
    return out


def _check(got, want, exact=True):
    cmp = ol.compare_frames(got["rgba"], got["id"], want["rgba"], want["id"])
    assert cmp["mask_agree"] >= 0.9999, cmp
    assert cmp["max_rgb_err"] <= 1, cmp          # 1/255 per channel on agreeing hits
    assert cmp["max_miss_rgb_err"] == 0, cmp     # misses are exactly the background
    if exact:
        assert cmp["n_mask_off"] == 0, cmp
        assert np.array_equal(got["id"], want["id"])
        assert np.array_equal(got["dist"].view(np.uint32), want["dist"].view(np.uint32))
    return cmp


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("size", [(320, 240), (1920, 1080)])
def test_examples_match_oracle(name, size, scenes_dir):
    import loltracer_b200 as lb

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    got = _render(lb, scene, w, h)
    want = ol.port_render(scene, w, h, counts=True)
    _check(got, want)
    # the primary march takes exactly the reference's number of steps per pixel
    assert np.array_equal(got["nprimary"], want["nprimary"])


@pytest.fixture(scope="module")
def golden_frames():
    from conftest import ROOT
    return np.load(os.path.join(ROOT, "tests", "golden", "frames.npz"))


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("size", [(160, 90), (96, 64)])
def test_examples_match_reference_goldens(name, size, scenes_dir, golden_frames):
    """Straight against the reference's own output (tests/golden/frames.npz)."""
    import loltracer_b200 as lb

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    got = _render(lb, scene, w, h)
    key = f"{name}_{w}x{h}"
    want = dict(rgba=golden_frames[key + "_rgba"], id=golden_frames[key + "_id"].astype(np.uint32),
                dist=golden_frames[key + "_dist"])
    _check(got, want)


@pytest.mark.parametrize("name", ["scene2", "scene3", "scene4"])
def test_examples_match_the_jit_renderers_semantics(name, scenes_dir):
    """The reference's second renderer (tracing_jit_renderer.dasc; oracle mode 1 restates its
    deltas, SURVEY.md Appendix C: <= select, fminf/fmaxf in the shadow march, its own sminf
    sequence, no boxes -- hence not scene.lol).  North-star tolerance at 1920x1080: the
    hit/miss mask agrees on every pixel and RGB is within 1/255 -- except where the two
    REFERENCE renderers themselves disagree by more (scene4: one pixel of 2 M, by 5/255, from
    a last-ulp difference in sminf at a shadow edge), which no single frame can match."""
    import loltracer_b200 as lb

    w, h = 1920, 1080
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    got = _render(lb, scene, w, h)
    jit = ol.port_render(scene, w, h, mode=1)
    naive = ol.port_render(scene, w, h, mode=0)
    cmp = ol.compare_frames(got["rgba"], got["id"], jit["rgba"], jit["id"])
    assert cmp["mask_agree"] >= 0.9999 and cmp["max_miss_rgb_err"] == 0, cmp

    def err(a, b):
        e = np.zeros(a.shape, np.int32)
        for sh in (16, 8, 0):
            e = np.maximum(e, np.abs(((a >> sh) & 0xFF).astype(np.int32) - ((b >> sh) & 0xFF).astype(np.int32)))
        return e

    both = (got["id"] != 0) & (jit["id"] != 0)
    beyond = (err(got["rgba"], jit["rgba"]) > 1) & both
    reference_disagrees = err(naive["rgba"], jit["rgba"]) > 0
    assert not (beyond & ~reference_disagrees).any()
    assert beyond.sum() <= 1e-5 * w * h, int(beyond.sum())
    got["renderer"].close()


@pytest.mark.parametrize("name", EXAMPLES)
def test_examples_4k_match_oracle(name, scenes_dir):
    """BASELINE size: every examples/*.lol scene at 3840x2160, every pixel."""
    import loltracer_b200 as lb

    w, h = 3840, 2160
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    got = _render(lb, scene, w, h)
    want = ol.port_render(scene, w, h)
    cmp = _check(got, want)
    # powf differs between glibc and CUDA by an ulp or so: few pixels flip a channel by 1
    assert cmp["n_rgb_off"] < 0.02 * w * h, cmp


@pytest.mark.parametrize("k", [0, 16, 32, 48])
def test_orbit_cameras(k, scenes_dir, golden_frames):
    """Config C5: the camera is a per-frame argument, the kernel is not rebuilt."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    cam = scenegen.orbit_camera(scene.camera, k, 64)
    got = _render(lb, scene, 160, 90, camera=cam)
    key = f"orbit{k}_160x90"
    _check(got, dict(rgba=golden_frames[key + "_rgba"], id=golden_frames[key + "_id"].astype(np.uint32),
                     dist=golden_frames[key + "_dist"]))
    got = _render(lb, scene, 1280, 720, camera=cam)
    _check(got, ol.port_render(scene, 1280, 720, camera=cam))


def test_synthetic_1024_primitives(golden_frames):
    """Config C4: 128 smooth-union trees of 8 spheres, lowered to a table loop."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    scene = lb.Scene.from_string(scenegen.synthetic_scene_text())
    got = _render(lb, scene, 96, 54)
    _check(got, dict(rgba=golden_frames["synthetic_96x54_rgba"],
                     id=golden_frames["synthetic_96x54_id"].astype(np.uint32),
                     dist=golden_frames["synthetic_96x54_dist"]))
    got = _render(lb, scene, 256, 144)
    _check(got, ol.port_render(scene, 256, 144))


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (33, 3), (250, 131), (641, 2)])
def test_ragged_sizes(size, scenes_dir):
    """Frames that are not multiples of the 8x4 tile or of the work chunk."""
    import loltracer_b200 as lb

    w, h = size
    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene.lol"))
    _check(_render(lb, scene, w, h), ol.port_render(scene, w, h))


def test_options_do_not_change_the_image(scenes_dir):
    """The three exact shortcuts are exact: switching them off gives the same frame and
    the reference's full shadow-march counts."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    w, h = 480, 270
    want = ol.port_render(scene, w, h, counts=True)
    on = _render(lb, scene, w, h)
    off = _render(lb, scene, w, h, options=lb.Options.default(
        skip_black_miss=0, cull_backfacing=0, shadow_early_out=0))
    assert np.array_equal(on["rgba"], off["rgba"])
    _check(off, want)
    assert np.array_equal(off["nshadow"], want["nshadow"])  # nothing skipped
    assert on["nshadow"].sum() < 0.7 * off["nshadow"].astype(np.int64).sum()


def test_miss_pixels_shaded_when_material0_is_not_black(scenes_dir):
    """naive_renderer.c shades misses with material 0; the shortcut must switch itself off."""
    import loltracer_b200 as lb

    text = open(os.path.join(scenes_dir, "scene2.lol")).read()
    text = text.replace("ambient = (0, 0, 0)", "ambient = (0.3, 0.5, 0.7)", 1).replace(
        "color = (0.01, 0.01, 0.01)", "color = (0.9, 0.9, 0.9)")
    scene = lb.Scene.from_string(text)
    w, h = 320, 180
    got = _render(lb, scene, w, h)
    want = ol.port_render(scene, w, h)
    assert (want["rgba"][want["id"] == 0] & 0xFFFFFF).min() > 0  # misses are not black here
    cmp = ol.compare_frames(got["rgba"], got["id"], want["rgba"], want["id"])
    assert cmp["n_mask_off"] == 0 and cmp["max_rgb_err"] <= 1
    miss = want["id"] == 0
    d = np.abs(((got["rgba"][miss] >> 8) & 0xFF).astype(int) - ((want["rgba"][miss] >> 8) & 0xFF).astype(int))
    assert d.max() <= 1


def test_pixel_formats_and_pitch(scenes_dir):
    """SDL_MapRGB for other 32-bit layouts (renderer.h:17-22) and a padded surface pitch."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene.lol"))
    w, h = 200, 100
    r = lb.Renderer(scene)
    base = np.zeros((h, w), np.uint32)
    r.render_host(base.ctypes.data, w, h)
    # BGRX: r<<8 | g<<16 | b<<24, no alpha
    f = lb.PixFmt(rshift=8, gshift=16, bshift=24, amask=0)
    other = np.zeros((h, w), np.uint32)
    r.render_host(other.ctypes.data, w, h, fmt=f)
    rr, gg, bb = (base >> 16) & 0xFF, (base >> 8) & 0xFF, base & 0xFF
    assert np.array_equal(other, (rr << 8) | (gg << 16) | (bb << 24))
    # RGB565-style loss in a 32-bit word
    f = lb.PixFmt(rshift=11, gshift=5, bshift=0, rloss=3, gloss=2, bloss=3, amask=0)
    r.render_host(other.ctypes.data, w, h, fmt=f)
    assert np.array_equal(other, ((rr >> 3) << 11) | ((gg >> 2) << 5) | (bb >> 3))
    # pitch larger than the row: bytes beyond w stay untouched
    pitch_px = w + 24
    padded = np.full((h, pitch_px), 0xDEADBEEF, np.uint32)
    r.render_host(padded.ctypes.data, w, h, pitch_bytes=pitch_px * 4)
    assert np.array_equal(padded[:, :w], base) and (padded[:, w:] == 0xDEADBEEF).all()
    r.close()


def test_counters_match_oracle_when_nothing_is_skipped(scenes_dir):
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene2.lol"))
    w, h = 640, 360
    opt = lb.Options.default(skip_black_miss=0, cull_backfacing=0, shadow_early_out=0, counters=1)
    r = lb.Renderer(scene, opt)
    frame = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    r.render_device(frame.data_ptr(), w, h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    c = r.read_counters()
    t = ol.port_render(scene, w, h)["totals"]
    assert (c["primary_evals"], c["normal_evals"], c["shadow_evals"], c["hit_pixels"], c["pixels"]) == (
        t["primary"], t["normal"], t["shadow"], t["hits"], w * h)
    r.close()


def test_frames_are_repeatable_and_counter_rearms(scenes_dir):
    """The work counter is reset by the last CTA: frame after frame, size after size."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene2.lol"))
    r = lb.Renderer(scene)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    for (w, h) in [(320, 240), (1920, 1080), (320, 240), (64, 64), (320, 240)]:
        frame = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
        for _ in range(3):
            frame.zero_()
            r.render_device(frame.data_ptr(), w, h, stream=st)
        torch.cuda.synchronize()
        if (w, h) == (320, 240):
            outs.append(frame.cpu().numpy())
    assert all(np.array_equal(outs[0], o) for o in outs[1:])
    assert (outs[0] != 0).all()
    r.close()


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("guarded", [0, 2])
def test_kernel_variants_are_bit_identical(name, variant, guarded, scenes_dir):
    """Phase-sequential vs ray-compaction vs two-rays-per-thread (packed FP32) kernel,
    IEEE forms vs guarded fast path: the same distance, id and primary step count for
    every pixel, and the oracle's pixels.  (Variant 3 is built from the guarded forms;
    with guarded=0 the lowering hands back variant 1.)"""
    import loltracer_b200 as lb

    w, h = 1001, 562  # odd width: a last column without a partner; not a multiple of the chunk
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    want = ol.port_render(scene, w, h, counts=True)
    got = _render(lb, scene, w, h, options=lb.Options.default(variant=variant, guarded_fastpath=guarded))
    _check(got, want)
    assert np.array_equal(got["nprimary"], want["nprimary"])


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_shadow_step_counts_per_variant(variant, scenes_dir):
    """With the shortcuts off every kernel marches exactly the reference's shadow steps."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    w, h = 400, 226
    opt = lb.Options.default(variant=variant, skip_black_miss=0, cull_backfacing=0, shadow_early_out=0,
                             counters=1)
    got = _render(lb, scene, w, h, options=opt)
    want = ol.port_render(scene, w, h, counts=True)
    _check(got, want)
    assert np.array_equal(got["nshadow"], want["nshadow"])
    c = got["renderer"].read_counters()
    t = want["totals"]
    assert (c["primary_evals"], c["normal_evals"], c["shadow_evals"]) == (t["primary"], t["normal"], t["shadow"])


@pytest.mark.parametrize("variant", [2, 3])
@pytest.mark.parametrize("world", [2, 8])
def test_variant2_shards_and_small_chunks(variant, world, scenes_dir):
    """Compaction and two-ray kernels on shards (narrow chunks) and ragged frames."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    w, h = 333, 129
    r = lb.Renderer(scene, lb.Options.default(variant=variant))
    st = torch.cuda.current_stream().cuda_stream
    frame = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    for rank in range(world):
        r.render_device(frame.data_ptr(), w, h, pitch_px=w,
                        shard=lb.Shard(rank=rank, world=world, dst_full_frame=1), stream=st)
    torch.cuda.synchronize()
    want = ol.port_render(scene, w, h)
    got = frame.cpu().numpy().view(np.uint32)
    err = np.zeros(got.shape, np.int32)
    for s in (16, 8, 0):
        err = np.maximum(err, np.abs(((got >> s) & 0xFF).astype(np.int32) - ((want["rgba"] >> s) & 0xFF).astype(np.int32)))
    assert err.max() <= 1 and err[want["id"] == 0].max() == 0
    r.close()


@pytest.mark.parametrize("name", ["scene", "scene2", "synthetic"])
def test_table_loops_and_pruning_are_exact(name, scenes_dir):
    """Forced table loops (threshold 2) with box pruning, groups, hints and hoisted short
    segments vs fully unrolled code: identical distance, id and pixels."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    scene = (lb.Scene.from_string(scenegen.synthetic_scene_text()) if name == "synthetic"
             else lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol")))
    w, h = (192, 108) if name == "synthetic" else (640, 360)
    a = _render(lb, scene, w, h, options=lb.Options.default(loop_threshold=2, prune_bounds=1))
    b = _render(lb, scene, w, h, options=lb.Options.default(loop_threshold=2, prune_bounds=0))
    for k in ("rgba", "id", "nprimary", "nshadow"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["dist"].view(np.uint32), b["dist"].view(np.uint32))
    if name != "synthetic":
        _check(a, ol.port_render(scene, w, h))


@pytest.mark.parametrize("name", EXAMPLES + ["synthetic"])
def test_two_ray_kernel_4k_bit_identical_to_variant1(name, scenes_dir):
    """Variant 3 (two pixels per thread in packed FADD2/FMUL2/FFMA2 registers) against
    variant 1 at 3840x2160: the SAME frame, distances, ids, primary and shadow step
    counts -- packing changes who issues an instruction, not what it computes."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    w, h = (3840, 2160) if name != "synthetic" else (480, 270)
    scene = (lb.Scene.from_string(scenegen.synthetic_scene_text()) if name == "synthetic" else
             lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol")))
    a = _render(lb, scene, w, h, options=lb.Options.default(variant=1, counters=1))
    b = _render(lb, scene, w, h, options=lb.Options.default(variant=3, counters=1))
    assert "#define LOL_VARIANT 3" in lb.lower_cuda(scene, lb.Options.default(variant=3))
    for key in ("rgba", "id", "nprimary", "nshadow"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["dist"].view(np.uint32), b["dist"].view(np.uint32))
    ca, cb = a["renderer"].read_counters(), b["renderer"].read_counters()
    # what the box tests skipped differs by design: a pair skips only when both rays can, and a
    # ray that had to evaluate its partner's objects may skip more afterwards
    assert {k: v for k, v in ca.items() if k != "skipped_flops"} == {k: v for k, v in cb.items() if k != "skipped_flops"}


@pytest.mark.parametrize("name", EXAMPLES + ["synthetic", "synthetic_csg"])
def test_packed_pairs_bit_identical_to_scalar_code(name, scenes_dir):
    """pack_pairs: two same-shaped subtrees of an object in the halves of FADD2/FMUL2/FFMA2
    (scene4: spheres (0,4), (1,5) and their two smooth unions; the synthetic scenes: the two
    halves of every balanced tree, constants read as 64-bit pairs from the row) against the
    scalar code, guard forced on so that every example takes the packed forms where it can:
    the SAME frame, distances, ids and step counts at 3840x2160, and the oracle's."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    w, h = (3840, 2160) if not name.startswith("synthetic") else (960, 540)
    scene = (lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg"))) if name.startswith("synthetic")
             else lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol")))
    a = _render(lb, scene, w, h, options=lb.Options.default(variant=1, guarded_fastpath=2, pack_pairs=0))
    b = _render(lb, scene, w, h, options=lb.Options.default(variant=1, guarded_fastpath=2, pack_pairs=2))
    src = lb.lower_cuda(scene, lb.Options.default(variant=1, guarded_fastpath=2, pack_pairs=2))
    if name not in ("scene", "scene2"):  # their spheres are separate top-level objects: nothing to pair
        head = src.split("//@@SCENE@@")[0]
        assert "lol_sqrt_fast2(" in head[head.index("__forceinline__ float lol_sdf_try("):]
    for key in ("rgba", "id", "nprimary", "nshadow"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["dist"].view(np.uint32), b["dist"].view(np.uint32))
    if not name.startswith("synthetic"):
        _check(b, ol.port_render(scene, w, h))
    a["renderer"].close()
    b["renderer"].close()


@pytest.mark.parametrize("name", EXAMPLES + ["synthetic"])
def test_shared_first_step_and_division_pretest_do_not_change_the_frame(name, scenes_dir):
    """share_first_step: step 1 of every primary ray is sdf(camera position); one thread per
    CTA evaluates it and every ray takes it from shared memory.  shadow_div_pretest: the shadow
    march divides (50 * d) / t only when a multiplication cannot prove that the quotient leaves
    res alone.  Against every ray evaluating and dividing everything itself: the same frame,
    distances, ids and step counts (the shared step still counts as a step), also from a
    camera inside an object, where the march ends on step 1."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    w, h = (1920, 1080) if name != "synthetic" else (480, 270)
    scene = (lb.Scene.from_string(scenegen.synthetic_scene_text()) if name == "synthetic" else
             lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol")))
    cams = [None, lb.Camera.make([0, 1, -6], [0.2, -0.1, -1], scene.struct.camera.fov)]
    for cam in cams:
        a = _render(lb, scene, w, h, camera=cam,
                    options=lb.Options.default(variant=1, share_first_step=0, shadow_div_pretest=0))
        on = lb.Options.default(variant=1, share_first_step=2, shadow_div_pretest=1)
        b = _render(lb, scene, w, h, camera=cam, options=on)
        src = lb.lower_cuda(scene, on)
        assert "#define LOL_SHARE_FIRST 1" in src and "#define LOL_DIV_PRETEST 1" in src
        for key in ("rgba", "id", "nprimary", "nshadow"):
            assert np.array_equal(a[key], b[key]), key
        assert np.array_equal(a["dist"].view(np.uint32), b["dist"].view(np.uint32))
        a["renderer"].close()
        b["renderer"].close()


@pytest.mark.parametrize("name", EXAMPLES)
def test_guard_outside_the_march_loops_does_not_change_the_frame(name, scenes_dir):
    """guard_out (default on): variant 1's march loops run the guarded arithmetic alone (lol_sdf_try) and
    end when the range guard fails; that march is done again from its start with the IEEE forms.
    Against the guard's fall-back inside every evaluation (guard_out=0): the same frame, distances, ids
    and step counts -- from the file camera, from cameras whose every ray fails the guard at step 1
    (a sphere's centre; beyond 2^60), and from one aimed at a sphere's centre from outside (its march
    ends at the surface; a guard can only fail where a march starts)."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    fov = scene.struct.camera.fov
    st = scene.struct
    centres = [list(st.nodes[i].point) for i in range(st.n_nodes) if st.nodes[i].type == 3]
    c = centres[0] if centres else [0.0, 1.0, -6.0]
    cams = [(None, 1920, 1080),
            (lb.Camera.make(c, [0.2, -0.1, -1], fov), 320, 180),
            (lb.Camera.make([3e19, 1, 0], [-1, 0, 0], fov), 160, 90),
            (lb.Camera.make([c[0], c[1], c[2] + 64.0], [0, 0, -1], fov), 321, 181)]
    on = lb.Options.default(variant=1, guarded_fastpath=2)
    off = lb.Options.default(variant=1, guarded_fastpath=2, guard_out=0)
    assert "#define LOL_GUARD_OUT 3" in lb.lower_cuda(scene, on)
    assert "#define LOL_GUARD_OUT 0" in lb.lower_cuda(scene, off)
    for cam, w, h in cams:
        a = _render(lb, scene, w, h, camera=cam, options=off)
        b = _render(lb, scene, w, h, camera=cam, options=on)
        for key in ("rgba", "id", "nprimary", "nshadow"):
            assert np.array_equal(a[key], b[key]), key
        da, db = a["dist"], b["dist"]
        assert ((da.view(np.uint32) == db.view(np.uint32)) | (np.isnan(da) & np.isnan(db))).all()
        a["renderer"].close()
        b["renderer"].close()


def _rgb_err(got, want_rgba):
    err = np.zeros(got.shape, np.int32)
    for s in (16, 8, 0):
        err = np.maximum(err, np.abs(((got >> s) & 0xFF).astype(np.int32) - ((want_rgba >> s) & 0xFF).astype(np.int32)))
    return err


def test_host_surface_follows_a_resizing_window(scenes_dir):
    """main.c re-fetches the surface every frame and the window is resizable
    (main.c:182): the host-surface entry point must follow changes of size, pitch and
    pixel pointer between frames on one renderer.  Every buffer is DROPPED before the next
    frame, as SDL drops a window surface on resize: the next one often lives at the same
    address, which the library must not mistake for memory it has seen before (it keeps no
    registration on memory it does not own: pageable surfaces go through its own staging)."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    r = lb.Renderer(scene, device=0)
    seen = []
    for (w, h, pad) in [(320, 240, 0), (641, 361, 7), (320, 240, 0), (1280, 720, 64), (1280, 720, 64),
                        (33, 17, 1), (1280, 720, 64)]:
        host = np.full((h, w + pad), 0xDEADBEEF, np.uint32)
        seen.append(host.ctypes.data)
        r.render_host(host.ctypes.data, w, h, pitch_bytes=(w + pad) * 4)
        want = ol.port_render(scene, w, h)
        assert _rgb_err(host[:, :w], want["rgba"]).max() <= 1, (w, h)
        assert (host[:, w:] == 0xDEADBEEF).all(), "padding past the row was written"
        del host
    r.close()


def test_host_surface_kinds_give_the_same_frame(scenes_dir):
    """The same frame through every way a surface can reach the copy engine: pageable memory
    (staged through the renderer's pinned frame), memory its owner pinned with
    lolb200_surface_pin, CUDA-allocated pinned memory (torch), and the zero-copy mode
    (LOLB200_HOST_MODE=mapped) -- and a buffer that is unpinned again goes back to staging."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    r = lb.Renderer(scene, device=0)
    w, h, pitch_px = 1000, 563, 1016
    pageable = np.zeros((h, pitch_px), np.uint32)
    r.render_host(pageable.ctypes.data, w, h, pitch_bytes=pitch_px * 4)
    assert (pageable[:, :w] != 0).any() and (pageable[:, w:] == 0).all()
    owned = np.zeros((h, pitch_px), np.uint32)
    lb.surface_pin(owned.ctypes.data, owned.nbytes)
    r.render_host(owned.ctypes.data, w, h, pitch_bytes=pitch_px * 4)
    assert np.array_equal(owned, pageable)
    os.environ["LOLB200_HOST_MODE"] = "mapped"
    try:
        owned[:] = 0
        r.render_host(owned.ctypes.data, w, h, pitch_bytes=pitch_px * 4)
        assert np.array_equal(owned, pageable)
    finally:
        del os.environ["LOLB200_HOST_MODE"]
    lb.surface_unpin(owned.ctypes.data)
    owned[:] = 0
    r.render_host(owned.ctypes.data, w, h, pitch_bytes=pitch_px * 4)  # pageable again
    assert np.array_equal(owned, pageable)
    pinned = torch.zeros((h, pitch_px), dtype=torch.int32).pin_memory()
    r.render_host(pinned.data_ptr(), w, h, pitch_bytes=pitch_px * 4)
    assert np.array_equal(pinned.numpy().view(np.uint32), pageable)
    r.close()


def test_launches_of_one_renderer_on_two_streams_do_not_interfere(scenes_dir):
    """a renderer has ONE work queue (chunk counter, longest-first buffers).  Launches of
    one renderer on different streams, with no host synchronisation in between, are serialised on
    the device by the library -- every frame complete and equal to the single-stream frame."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene2.lol"))
    r = lb.Renderer(scene, device=0)
    w, h = 1920, 1080
    ref = torch.zeros((h, w), dtype=torch.int32, device="cuda:0")
    r.render_device(ref.data_ptr(), w, h, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(3)]
    frames = [torch.zeros((h, w), dtype=torch.int32, device="cuda:0") for _ in range(9)]
    host = np.zeros((h, w), np.uint32)
    for i, f in enumerate(frames):
        r.render_device(f.data_ptr(), w, h, stream=streams[i % 3].cuda_stream)
        if i == 4:
            r.render_host(host.ctypes.data, w, h)  # slab launches share the first work-counter slot
    torch.cuda.synchronize()
    for f in frames:
        assert torch.equal(f, ref)
    assert np.array_equal(host, ref.cpu().numpy().view(np.uint32))
    r.close()


@pytest.mark.parametrize("csg", [False, True])
def test_pruned_hinted_loops_equal_brute_force(csg):
    """The 1024-primitive scenes with everything on (boxes, Morton-sorted groups, last-winner
    hints, two rays per thread) against the plain loop over all 128 objects in file order
    (prune_bounds=0, variant 1): the same frame, distances, ids and step counts at 1280x720 --
    pruning and evaluation order change what is computed, never the result."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    w, h = 1280, 720
    scene = lb.Scene.from_string(scenegen.synthetic_scene_text(csg=csg))
    brute = _render(lb, scene, w, h, options=lb.Options.default(variant=1, prune_bounds=0))
    for variant in (1, 3):
        fast = _render(lb, scene, w, h, options=lb.Options.default(variant=variant))
        for key in ("rgba", "id", "nprimary", "nshadow"):
            assert np.array_equal(brute[key], fast[key]), (variant, key)
        assert np.array_equal(brute["dist"].view(np.uint32), fast["dist"].view(np.uint32)), variant
        fast["renderer"].close()
    brute["renderer"].close()


@pytest.mark.parametrize("name", EXAMPLES)
def test_straight_line_box_tests_do_not_change_the_frame(name, scenes_dir):
    """Every straight-line box test forced on (prune_bounds=2) against none (0): the same
    frame, distances, ids and step counts at 1920x1080, and the oracle's frame."""
    import loltracer_b200 as lb

    w, h = 1920, 1080
    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    off = _render(lb, scene, w, h, options=lb.Options.default(prune_bounds=0))
    for variant in (1, 3):
        on = _render(lb, scene, w, h, options=lb.Options.default(prune_bounds=2, variant=variant, guarded_fastpath=2))
        for key in ("rgba", "id", "nprimary", "nshadow"):
            assert np.array_equal(off[key], on[key]), (variant, key)
        assert np.array_equal(off["dist"].view(np.uint32), on["dist"].view(np.uint32)), variant
        on["renderer"].close()
    # what the estimate chooses (prune_bounds=1, the default: on scene4 a BALL around one of the blob's own sphere
    # centres) and the same with boxes only (3), also from cameras the estimate did not march
    cams = [None, lb.Camera.make([6.0, 3.0, -2.0], [-0.5, -0.3, -1.0], scene.struct.camera.fov),
            lb.Camera.make([2.0, 2.0, -10.0], [0.1, 0.2, 1.0], scene.struct.camera.fov)]  # the ball's centre on scene4
    for cam in cams:
        ref = off if cam is None else _render(lb, scene, w // 2, h // 2, camera=cam, options=lb.Options.default(prune_bounds=0))
        for prune in (1, 3):
            on = _render(lb, scene, *(ref["rgba"].shape[::-1]), camera=cam, options=lb.Options.default(prune_bounds=prune))
            for key in ("rgba", "id", "nprimary", "nshadow"):
                assert np.array_equal(ref[key], on[key]), (prune, key)
            da, db = ref["dist"], on["dist"]
            assert ((da.view(np.uint32) == db.view(np.uint32)) | (np.isnan(da) & np.isnan(db))).all()
            on["renderer"].close()
        if cam is not None:
            ref["renderer"].close()
    if name == "scene4":
        assert "lol_ball_skips(lol_dot(" in lb.lower_cuda(scene, lb.Options.default())
    _check(off, ol.port_render(scene, w, h))
    off["renderer"].close()


@pytest.mark.parametrize("name,point,direction", [
    ("scene4", (0, 1, -6), (0, 0, -1)),          # the camera sits exactly at a sphere's centre: sqrt(0)
    ("scene4", (-1, 0.5, -3), (0.3, 0.2, -1)),   # another centre, inside the blob
    ("scene4", (3e19, 1, 0), (-1, 0, 0)),        # beyond 2^60: the guard's coordinate range
    ("scene2", (0, 5, -6), (0, -1, 0)),          # straight down: cross(dir, up) = 0, a NaN camera basis
    ("scene", (2, 2, -10), (0, 0, -1)),          # inside the round box
    ("scene", (-0.0, 0, -0.0), (0, 0, -1)),      # -0 + rd * 0 is +0 or -0 depending on rd: the shared first
    ("scene4", (-0.0, 6, 3), (0.3, -0.7, -1)),   # step (share_first_step) must stand aside
])
@pytest.mark.parametrize("variant", [1, 3])
def test_cameras_outside_the_fast_forms_ranges(name, point, direction, variant, scenes_dir):
    """Where the guarded fast path must fall back to the IEEE forms (and where everything is NaN)
    the frame is still the oracle's, pixel for pixel."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    cam = lb.Camera.make(list(point), list(direction), scene.struct.camera.fov)
    w, h = 160, 90
    got = _render(lb, scene, w, h, camera=cam, options=lb.Options.default(variant=variant, guarded_fastpath=2))
    want = ol.port_render(scene, w, h, camera=cam)
    cmp = ol.compare_frames(got["rgba"], got["id"], want["rgba"], want["id"])
    assert cmp["n_mask_off"] == 0 and cmp["max_rgb_err"] <= 1, cmp
    assert np.array_equal(got["id"], want["id"])
    same = (got["dist"].view(np.uint32) == want["dist"].view(np.uint32)) | (np.isnan(got["dist"]) & np.isnan(want["dist"]))
    assert same.all()
    got["renderer"].close()


# ---- straight against the compiled reference (oracle/_ref/liblolref.so travels to the GPU box) ----
#
# The tests above compare the CUDA path with the oracle PORT; the port is pinned to the reference by
# the CPU suite (tests/test_oracle_pin.py).  The tests below close the chain inside `pytest -m gpu`
# itself: the reference's own naive_renderer.c, compiled unmodified, is the checker.

def _need_ref():
    if not ol.have_ref():
        pytest.skip("oracle/_ref/liblolref.so did not travel to this box")


def _check_rows(got, want, ystride, exact=True):
    """`got`: full-frame arrays from the GPU; `want`: the reference's rows y = 0, ystride, ..."""
    sub = {k: got[k][::ystride] for k in ("rgba", "id", "dist")}
    assert sub["id"].shape == want["id"].shape
    return _check(sub, want, exact=exact)


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("size", [(320, 240), (1920, 1080)])
def test_examples_match_the_compiled_reference(name, size, scenes_dir):
    """naive_renderer.c:216-236 itself (RefScene.probe) against the CUDA frame: distance and id of
    every pixel bit-identical, RGB within 1/255 (CUDA powf vs glibc powf), misses exactly black."""
    import loltracer_b200 as lb

    _need_ref()
    w, h = size
    path = os.path.join(scenes_dir, name + ".lol")
    got = _render(lb, lb.Scene.from_file(path), w, h)
    want = ol.RefScene(path=path).probe(w, h)
    _check(got, want)
    got["renderer"].close()


def test_config_c4_at_3840x2160_matches_the_compiled_reference():
    """BASELINE config C4 at its stated size: the 1024-sphere scene at 3840x2160 (pruned table loops,
    packed pairs, hints -- the default kernel) against the reference on 16 scanlines spread over the
    frame (the reference needs ~1.5 core-minutes per scanline of this scene)."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    _need_ref()
    w, h, ystride = 3840, 2160, 135
    text = scenegen.synthetic_scene_text()
    got = _render(lb, lb.Scene.from_string(text), w, h)
    want = ol.RefScene(text=text).probe(w, h, ystride=ystride)
    cmp = _check_rows(got, want, ystride)
    assert (want["id"] != 0).mean() > 0.3, "the sampled rows should see the spheres"
    assert cmp["n_mask_off"] == 0
    got["renderer"].close()


@pytest.mark.parametrize("k", [0, 16, 32, 48])
def test_config_c5_orbit_frames_at_7680x4320_match_the_compiled_reference(k, scenes_dir):
    """BASELINE config C5 at its stated size: orbit frames 0/16/32/48 of scene4 at 7680x4320 against the
    reference (camera mutated as main.c:71-112 does) on 32 scanlines per frame."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    _need_ref()
    w, h, ystride = 7680, 4320, 135
    path = os.path.join(scenes_dir, "scene4.lol")
    scene = lb.Scene.from_file(path)
    cam = scenegen.orbit_camera(scene.camera, k, 64)
    got = _render(lb, scene, w, h, camera=cam)
    rs = ol.RefScene(path=path)
    rs.set_camera(list(cam.point), list(cam.direction))
    want = rs.probe(w, h, ystride=ystride)
    _check_rows(got, want, ystride)
    got["renderer"].close()


@pytest.mark.parametrize("name", ["scene2", "scene3"])
def test_config_c2_at_1920x1080_through_the_unmodified_render_thread(name, scenes_dir):
    """BASELINE config C2: scene2/scene3 at 1920x1080 against the frame the reference's UNMODIFIED
    render_thread() writes under main.c's semaphore protocol (lolref_render_protocol): the image the
    reference program itself would show."""
    import loltracer_b200 as lb

    _need_ref()
    w, h = 1920, 1080
    path = os.path.join(scenes_dir, name + ".lol")
    got = _render(lb, lb.Scene.from_file(path), w, h)
    px, _ = ol.RefScene(path=path).render_protocol(w, h)
    err = _rgb_err(got["rgba"], px)
    assert err.max() <= 1
    assert err[got["id"] == 0].max() == 0
    got["renderer"].close()


@pytest.mark.parametrize("name", ["synthetic", "synthetic_csg"])
def test_candidate_memory_4k_bit_identical(name):
    """options.near_cache (the per-ray candidate memory of pruned table loops, lol_kernel.cuh: struct lol_near)
    on and off at 3840x2160: the SAME frame, distances, ids, primary and shadow step counts and evaluation
    counters; and far fewer FLOPs skipped by box tests that are never made."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    w, h = 3840, 2160
    scene = lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg")))
    a = _render(lb, scene, w, h, options=lb.Options.default(variant=1, near_cache=0, counters=1))
    b = _render(lb, scene, w, h, options=lb.Options.default(variant=1, near_cache=1, counters=1))
    assert "#define LOL_NEAR 1" in b["renderer"].source and "#define LOL_NEAR 0" in a["renderer"].source
    for key in ("rgba", "id", "nprimary", "nshadow"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(a["dist"].view(np.uint32), b["dist"].view(np.uint32))
    ca, cb = a["renderer"].read_counters(), b["renderer"].read_counters()
    assert {k: v for k, v in ca.items() if k != "skipped_flops"} == {k: v for k, v in cb.items() if k != "skipped_flops"}
    # near_cache=2: the warp looks at all rows again whenever one of its lanes has to;
    # near_cache=3: and a look reads the point's cell of the candidate grid (built on the device at creation)
    for near in (2, 3):
        c = _render(lb, scene, w, h, options=lb.Options.default(variant=1, near_cache=near, counters=1))
        assert "#define LOL_NEAR 2" in c["renderer"].source
        assert ("#define LOL_NEAR_GRID (1 &&" in c["renderer"].source) == (near == 3)
        for key in ("rgba", "id", "nprimary", "nshadow"):
            assert np.array_equal(a[key], c[key]), (near, key)
        assert np.array_equal(a["dist"].view(np.uint32), c["dist"].view(np.uint32))
        c["renderer"].close()
    # the grid from cameras inside the crowd and far outside the grid's domain
    for point, direction in (((0.3, 2.8, -9.0), (0.2, -0.1, -1.0)), ((60.0, 30.0, 40.0), (-0.6, -0.3, -0.7))):
        cam = lb.Camera.make(list(point), list(direction), scene.struct.camera.fov)
        p = _render(lb, scene, 1280, 720, camera=cam, options=lb.Options.default(variant=1, near_cache=0))
        q = _render(lb, scene, 1280, 720, camera=cam, options=lb.Options.default(variant=1, near_cache=3))
        for key in ("rgba", "id", "nprimary", "nshadow"):
            assert np.array_equal(p[key], q[key]), (point, key)
        assert np.array_equal(p["dist"].view(np.uint32), q["dist"].view(np.uint32))
        p["renderer"].close()
        q["renderer"].close()
    for r in (a, b):
        r["renderer"].close()
