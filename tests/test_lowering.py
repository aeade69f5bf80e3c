"""The code generator (lol_lower.c) without a GPU: structure of the emitted CUDA,
NVRTC cross-compilation for sm_100a, and the generated distance function itself,
compiled for the CPU with a tiny shim and compared with the oracle bit for bit."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from conftest import EXAMPLES, ROOT

from oracle_lib import HOST_SHIM as SHIM, cpu_sdf


def oracle_sdf(scene, pts):
    d = np.zeros(len(pts), np.float32)
    ids = np.zeros(len(pts), np.uint32)
    for i, p in enumerate(pts):
        pt = (C.c_float * 3)(*p.tolist())
        dd, ii = C.c_float(), C.c_uint32()
        ol.port().lolo_sdf(C.cast(scene._ptr, C.c_void_p), 0, C.byref(pt), C.byref(dd), C.byref(ii))
        d[i], ids[i] = dd.value, ii.value
    return d, ids


def _scene(lb, name, scenes_dir):
    from loltracer_b200 import scenegen
    if name == "synthetic":
        return lb.Scene.from_string(scenegen.synthetic_scene_text())
    return lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))


@pytest.mark.parametrize("name", EXAMPLES + ["synthetic"])
@pytest.mark.parametrize("loops", [0, 2])
def test_generated_sdf_equals_oracle_on_cpu(name, loops, scenes_dir, tmp_path):
    """Each primitive and smooth union as the lowering emits it == sdf.h / float.h.
    loops=2 forces table loops (with box pruning, groups, hints and hoisted short
    segments) onto the small scenes too."""
    import loltracer_b200 as lb

    scene = _scene(lb, name, scenes_dir)
    src = lb.lower_cuda(scene, lb.Options.default(guarded_fastpath=2, loop_threshold=loops))
    if loops == 2 and name in ("scene", "scene2"):
        assert "none can win" in src and "lol_box_skips(" in src  # scene: 2 spheres, scene2: 3 spheres in a row
    L = cpu_sdf(tmp_path, src, f"{name}{loops}")
    rng = np.random.default_rng(11)
    n = 3000 if name != "synthetic" else 300
    pts = np.concatenate([rng.uniform(-12, 12, (n, 3)), rng.normal(0, 2, (n, 3)) + [0, 1, -6]]).astype(np.float32)
    d = np.zeros(len(pts), np.float32)
    ids = np.zeros(len(pts), np.uint32)
    L.eval(pts.ctypes.data_as(C.c_void_p), len(pts), d.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
    wd, wi = oracle_sdf(scene, pts)
    assert np.array_equal(d.view(np.uint32), wd.view(np.uint32))
    assert np.array_equal(ids, wi)


@pytest.mark.parametrize("name", EXAMPLES + ["synthetic"])
@pytest.mark.parametrize("loops", [0, 2])
def test_generated_two_ray_sdf_equals_oracle_on_cpu(name, loops, scenes_dir, tmp_path):
    """Variant 3's lol_sdf2 evaluates two rays per call (packed FP32 on the GPU; a
    plain pair under the host shim): each half must equal the oracle's sdf() of its
    own point, whatever the other half holds -- including far-apart partners, which
    exercise the pruning test's "skip only if neither ray can win" rule."""
    import loltracer_b200 as lb

    scene = _scene(lb, name, scenes_dir)
    src = lb.lower_cuda(scene, lb.Options.default(variant=3, loop_threshold=loops))
    assert "#define LOL_VARIANT 3" in src and "lol_sdf2(" in src
    L = cpu_sdf(tmp_path, src, f"{name}{loops}v3")
    rng = np.random.default_rng(12)
    n = 3000 if name != "synthetic" else 300
    pts = np.concatenate([rng.uniform(-12, 12, (n, 3)), rng.normal(0, 2, (n, 3)) + [0, 1, -6]]).astype(np.float32)
    pts = pts[rng.permutation(len(pts))]  # random partners
    pts[1] = pts[0]                        # a pair of identical rays
    d = np.zeros(len(pts), np.float32)
    ids = np.zeros(len(pts), np.uint32)
    L.eval2(pts.ctypes.data_as(C.c_void_p), len(pts), d.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
    wd, wi = oracle_sdf(scene, pts)
    assert np.array_equal(d.view(np.uint32), wd.view(np.uint32))
    assert np.array_equal(ids, wi)


@pytest.mark.parametrize("name", EXAMPLES)
def test_lowered_source_structure(name, scenes_dir):
    import loltracer_b200 as lb

    scene = _scene(lb, name, scenes_dir)
    src = lb.lower_cuda(scene)
    assert "#define LOL_EXACT 1" in src and "#define LOL_SKIP_MISS 1" in src  # examples: black material 0
    assert "#define LOL_CULL 1" in src and "#define LOL_SHADOW_EARLY 1" in src
    assert f"#define LOL_NLIGHTS {scene.struct.n_lights}" in src
    # the guarded fast form plus its out-of-line IEEE fallback (forced for scene.lol,
    # whose two spheres and a box are below the "guard pays" threshold)
    forced = lb.lower_cuda(scene, lb.Options.default(guarded_fastpath=2))
    assert "#define LOL_GUARDED 1" in forced and "lol_sdf_ref" in forced
    assert forced.count("{ // object ") == 2 * scene.struct.n_objects
    assert ("#define LOL_GUARDED 1" in src) == (name != "scene")
    plain = lb.lower_cuda(scene, lb.Options.default(guarded_fastpath=0))
    assert "#define LOL_GUARDED 0" in plain and plain.count("{ // object ") == scene.struct.n_objects
    assert 'extern "C" __global__' in src and "struct lol_params" in src
    assert "switch (obj" not in src  # no per-node dispatch at run time
    off = lb.lower_cuda(scene, lb.Options.default(skip_black_miss=0, cull_backfacing=0, shadow_early_out=0, arith=1))
    for flag in ("LOL_EXACT", "LOL_SKIP_MISS", "LOL_CULL", "LOL_SHADOW_EARLY"):
        assert f"#define {flag} 0" in off


def test_skips_are_licensed_by_the_scene(scenes_dir):
    """The miss shortcut needs an all-zero material 0; the back-face cull needs finite
    materials with shininess >= 0 (powf(0, negative) is inf)."""
    import loltracer_b200 as lb

    text = open(os.path.join(scenes_dir, "scene2.lol")).read()
    lit = text.replace("ambient = (0, 0, 0)", "ambient = (0.1, 0, 0)", 1)
    assert "#define LOL_SKIP_MISS 0" in lb.lower_cuda(lb.Scene.from_string(lit))
    neg = text.replace("shininess = 50", "shininess = -2")
    src = lb.lower_cuda(lb.Scene.from_string(neg))
    assert "#define LOL_CULL 0" in src and "#define LOL_SKIP_MISS 1" in src


def test_synthetic_scene_becomes_a_table_loop():
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    scene = lb.Scene.from_string(scenegen.synthetic_scene_text())
    src = lb.lower_cuda(scene)
    assert "#define lol_run0 (LOL_TAB + " in src and "128 x U(U(U(S,S),U(S,S)),U(U(S,S),U(S,S)))" in src
    # rows, group boxes and the id -> row map live in ONE array that the kernel copies to shared memory
    assert "lol_run0_groups" in src and "lol_run0_rowof" in src and "#define LOL_TAB_IN_SMEM 1" in src
    words = int(re.search(r"#define LOL_TAB_WORDS (\d+)", src).group(1))
    assert 128 * 50 < words < 128 * 70
    assert src.count("lol_len(") + src.count("lol_sqrt_fast(") < 40  # one unrolled tree, not 1024 spheres
    unrolled = lb.lower_cuda(scene, lb.Options.default(loop_threshold=100000))
    assert unrolled.count("lol_len(") > 1024


@pytest.mark.parametrize("name", EXAMPLES + ["synthetic"])
@pytest.mark.parametrize("arith", [0, 1])
def test_nvrtc_compiles_for_sm_100a(name, arith, scenes_dir, tmp_path):
    """Cross-compiles here; checks the image is an sm_100a cubin with no spills."""
    import loltracer_b200 as lb

    scene = _scene(lb, name, scenes_dir)
    opt = lb.Options.default(arith=arith)
    image = lb.compile_cubin(lb.lower_cuda(scene, opt), opt)
    assert image[:4] == b"\x7fELF"
    path = tmp_path / "k.cubin"
    path.write_bytes(image)
    out = subprocess.run(["cuobjdump", "-res-usage", str(path)], capture_output=True, text=True).stdout
    assert "lol_render" in out
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", out)
    assert m, out
    assert int(m.group(4)) == 0, "local memory (spills) in the render kernel"
    sass = subprocess.run(["cuobjdump", "-sass", str(path)], capture_output=True, text=True).stdout
    assert "code for sm_100a" in sass
    if arith == 0:
        assert "MUFU.RSQ" in sass  # IEEE sqrt = RSQ + Newton fix-up
    assert "ATOMG" in sass or "ATOM" in sass  # the work counter


def test_compile_error_is_reported():
    import loltracer_b200 as lb

    with pytest.raises(lb.LolB200Error) as e:
        lb.compile_cubin("this is not CUDA")
    assert e.value.code == -4 and "error" in str(e.value)


def test_cubin_cache(scenes_dir, tmp_path, monkeypatch):
    """LOLB200_CACHE_DIR: the second compilation of the same program is a file read that
    returns the same image; another program (or arithmetic mode) gets its own entry; a
    damaged entry is recompiled and replaced."""
    import time

    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene2.lol"))
    src = lb.lower_cuda(scene)
    monkeypatch.setenv("LOLB200_CACHE_DIR", str(tmp_path))
    t0 = time.perf_counter()
    first = lb.compile_cubin(src)
    t1 = time.perf_counter()
    again = lb.compile_cubin(src)
    t2 = time.perf_counter()
    files = sorted(tmp_path.glob("lol-*.cubin"))
    assert len(files) == 1 and files[0].read_bytes() == first == again
    assert (t2 - t1) < (t1 - t0) / 5, "a cache hit must not run NVRTC"
    fast = lb.Options.default(arith=1)
    lb.compile_cubin(lb.lower_cuda(scene, fast), fast)
    lb.compile_cubin(lb.lower_cuda(lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))))
    assert len(list(tmp_path.glob("lol-*.cubin"))) == 3
    files[0].write_bytes(b"not a cubin")
    assert lb.compile_cubin(src) == first and files[0].read_bytes() == first
    assert not list(tmp_path.glob("*.tmp"))


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("variant", [1, 3])
def test_straight_line_box_tests_are_exact_on_cpu(name, variant, scenes_dir, tmp_path):
    """prune_bounds=2 switches every straight-line box test on (planes first, the bounded
    objects behind the box around them and their own boxes, tie-aware updates): the same
    distance and id as the oracle at every point, inside the boxes and far outside them."""
    import loltracer_b200 as lb

    scene = _scene(lb, name, scenes_dir)
    src = lb.lower_cuda(scene, lb.Options.default(variant=variant, prune_bounds=2, guarded_fastpath=2))
    assert src.split("//@@SCENE@@")[0].count("lol_box_skips(x, y, z") + src.count("lol_box_skips2(x, y, z") >= 2
    L = cpu_sdf(tmp_path, src, f"{name}box{variant}")
    rng = np.random.default_rng(13)
    pts = np.concatenate([rng.uniform(-40, 40, (2000, 3)), rng.normal(0, 3, (2000, 3)) + [0, 1, -6],
                          rng.uniform(-15, 15, (2000, 3)) * [1, 0.02, 1] + [0, -0.9, -5]]).astype(np.float32)
    d = np.zeros(len(pts), np.float32)
    ids = np.zeros(len(pts), np.uint32)
    fn = L.eval2 if variant == 3 else L.eval
    fn(pts.ctypes.data_as(C.c_void_p), len(pts), d.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
    wd, wi = oracle_sdf(scene, pts)
    assert np.array_equal(d.view(np.uint32), wd.view(np.uint32))
    assert np.array_equal(ids, wi)


def test_box_tests_are_switched_on_where_they_pay(scenes_dir):
    """The sampled estimate (lol_lower.c: est_march): scene4's blob is skipped in about a
    third of all evaluations -> its test is on; scene3's blob fills the view -> off, and the
    code is then the plain file-order code (no reordering, no tie rule)."""
    import loltracer_b200 as lb

    def sdf_text(name, **kw):
        src = lb.lower_cuda(lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol")), lb.Options.default(**kw))
        # the guarded form's arithmetic is lol_sdf_try (lol_sdf is the fall-back around it)
        body = src[src.index("lol_sdf_try(const float x" if "lol_sdf_try(const float x" in src else "lol_sdf(const float x"):]
        return body[:body.index("//@@SCENE@@")]

    # scene4's blob sits behind a BALL around one of its own sphere centres (four instructions instead of the
    # box's sixteen: the squared distance is the sphere's own); with balls off (prune_bounds=3), behind its box
    assert sdf_text("scene4").count("lol_ball_skips(") == 1 and sdf_text("scene4").count("lol_box_skips(") == 0
    assert sdf_text("scene4", prune_bounds=3).count("lol_box_skips(") == 1
    assert sdf_text("scene4", prune_bounds=3).count("lol_ball_skips(") == 0
    assert sdf_text("scene3").count("lol_box_skips(") == 0 and "== best" not in sdf_text("scene3")
    assert sdf_text("scene3") == sdf_text("scene3", prune_bounds=0)
    assert sdf_text("scene3", prune_bounds=2).count("lol_box_skips(") == 1


@pytest.mark.parametrize("name", ["scene4", "scene", "synthetic"])
@pytest.mark.parametrize("variant", [1, 3])
def test_points_outside_the_guard_take_the_ieee_path(name, variant, scenes_dir, tmp_path):
    """Exactly at a sphere's centre (squared length 0 < 2^-101), beyond 2^60, and NaN: the one
    range guard of the fast forms must hand the evaluation to lol_sdf_ref, with the oracle's result."""
    import loltracer_b200 as lb

    scene = _scene(lb, name, scenes_dir)
    st = scene.struct
    centres = [list(st.nodes[i].point) for i in range(st.n_nodes) if st.nodes[i].type == 3][:40]
    pts = np.array(centres + [[3e19, 1, 0], [0, -2e30, 5], [1e-30, 1, -6], [np.nan, 0, 0], [0, np.inf, 0]] +
                   [[c[0], c[1], c[2] + 1e-20] for c in centres[:8]], np.float32)
    if len(pts) % 2:
        pts = np.concatenate([pts, pts[:1]])
    src = lb.lower_cuda(scene, lb.Options.default(variant=variant, guarded_fastpath=2))
    assert "lol_sdf_ref(" in src
    L = cpu_sdf(tmp_path, src, f"guard_{name}{variant}")
    d = np.zeros(len(pts), np.float32)
    ids = np.zeros(len(pts), np.uint32)
    fn = L.eval2 if variant == 3 else L.eval
    fn(pts.ctypes.data_as(C.c_void_p), len(pts), d.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
    wd, wi = oracle_sdf(scene, pts)
    same = (d.view(np.uint32) == wd.view(np.uint32)) | (np.isnan(d) & np.isnan(wd))
    assert same.all(), (pts[~same], d[~same], wd[~same])
    assert np.array_equal(ids, wi)


def test_packed_pairs_structure(scenes_dir):
    """pack_pairs pairs subtrees by SHAPE: scene4's blob U(U(S,S),U(S,U(S,S))) evaluates
    U(S0,S1) with U(S4,S5) -- two packed sphere pairs and one packed smooth union -- and
    keeps S3 and the two outer unions scalar; scene3's U(S,S) becomes one sphere pair; the
    1024-sphere scene's balanced trees pack completely (4 + 2 + 1 packed nodes, one scalar
    root) and read their constants as 64-bit pairs from rows of 16-byte multiples -- by
    default, because that is where packing was measured to pay; straight-line scenes pack
    on request (pack_pairs=2).  Off, no packed instruction is emitted."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    def fast_fn(src):
        body = src.split("//@@SCENE@@")[0]
        body = body[body.index("__forceinline__ float lol_sdf_try("):]
        # (programs with one pruned table loop carry a second form behind it: lol_sdf_slow / lol_sdf_nr)
        return body.split("// the candidate memory's way out")[0]

    s4 = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    f = fast_fn(lb.lower_cuda(s4, lb.Options.default(pack_pairs=2)))
    assert f.count("lol_sqrt_fast2(") == 2 and f.count("lol_smin_c2(") == 1
    assert f.count("lol_sqrt_fast(") == 1 and f.count("lol_smin_c(") == 2
    assert "lol_u32 lol_pairc[] = {" in lb.lower_cuda(s4, lb.Options.default(pack_pairs=2))
    # the default packs inside table loops only (measured: straight-line scenes lose a few percent)
    assert "lol_sqrt_fast2(" not in fast_fn(lb.lower_cuda(s4, lb.Options.default()))
    f = fast_fn(lb.lower_cuda(s4, lb.Options.default(pack_pairs=3)))  # leaves only
    assert f.count("lol_sqrt_fast2(") == 2 and f.count("lol_smin_c2(") == 0 and f.count("lol_smin_c(") == 4
    off = lb.lower_cuda(s4, lb.Options.default(pack_pairs=0)).split("//@@SCENE@@")[0]
    assert "lol_u32 lol_pairc[] = {" not in off
    off = fast_fn(lb.lower_cuda(s4, lb.Options.default(pack_pairs=0)))
    assert "lol_sqrt_fast2(" not in off and off.count("lol_sqrt_fast(") == 5

    s3 = lb.Scene.from_file(os.path.join(scenes_dir, "scene3.lol"))
    f = fast_fn(lb.lower_cuda(s3, lb.Options.default(pack_pairs=2)))
    assert f.count("lol_sqrt_fast2(") == 1 and f.count("lol_smin_c(") == 1

    syn = lb.Scene.from_string(scenegen.synthetic_scene_text())
    src = lb.lower_cuda(syn, lb.Options.default(variant=1))
    f = fast_fn(src)
    assert f.count("lol_sqrt_fast2(") == 4 and f.count("lol_smin_c2v(") == 3 and f.count("lol_smin_c(") == 1
    assert "lol_ld2_(c + 8)" in f
    stride = int(re.search(r"#define LOL_RUN0_STRIDE (\d+)", src).group(1))
    assert stride % 4 == 0
    # the IEEE fallback reads the same rows: a.x at slot 8, b.x next to it, a.y two further on
    ref = src.split("//@@SCENE@@")[0]
    ref = ref[ref.index("lol_u64 lol_sdf_ref("):ref.index("__forceinline__ float lol_sdf_try(")]
    assert "LOL_TF(c[8])" in ref and "LOL_TF(c[9])" in ref and "LOL_TF(c[10])" in ref


def test_folded_smin_equals_reference(tmp_path):
    """lol_smin_c's two folded operations (.5f * (b - a) / k as one proved division by 2k,
    b - n * h for b + (a - b) * h) against sminf as float.h:29-33 computes it: bit-identical
    on 30 M operand pairs per run (random bits, nearly equal, equal, +-0, tiny), for ten
    smoothness values.  The helper's main() returns non-zero on the first mismatch."""
    exe = tmp_path / "smin_fold_check"
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(ROOT, "tests", "helpers", "smin_fold_check.c"), "-lm"])
    out = subprocess.run([str(exe), "3000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert " 0 mismatches" in out.stdout


def test_shadow_division_pretest_never_skips_a_needed_division(tmp_path):
    """LOL_DIV_PRETEST (lol_kernel.cuh, shadow march): whenever the multiplication test says
    "the quotient cannot lower res", minf(res, num / t) as float.h:6-13 computes it returns res
    bit for bit -- 20 M cases: random bits, plausible magnitudes, and quotients within +-2^-19
    of res where the 2^-21 margin decides."""
    exe = tmp_path / "div_pretest_check"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe),
                           os.path.join(ROOT, "tests", "helpers", "div_pretest_check.c"), "-lm"])
    out = subprocess.run([str(exe), "20000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert " 0 mismatches" in out.stdout


def test_the_ball_bound_of_scene4s_blob_is_conservative(scenes_dir):
    """lol_lower.c: ball_row -- scene4's blob is skipped when |p - c|^2 > (1.004 best + r1)^2 for the centre c of one of
    its own spheres.  With the floor moved far away the oracle's distance IS the blob's: on 400 000 random points, and on
    points on the ball's surface, the blob's distance is never below what the ball promises (|p - c| - r1 / 1.004), i.e.
    a skipped blob could not have won."""
    import loltracer_b200 as lb

    text = open(os.path.join(scenes_dir, "scene4.lol")).read()
    src = lb.lower_cuda(lb.Scene.from_string(text), lb.Options.default())
    body = src[src.index("lol_sdf_try(const float x"):]
    m = re.search(r"const float bqx = (.*?), bqy = (.*?), bqz = (.*?);\s*if \(!lol_ball_skips\(lol_dot\([^)]*\), LOL_F\(0x([0-9a-f]+)", body)
    assert m, "scene4's blob is expected behind a ball test"

    def centre(expr, axis):
        if expr.strip() == axis:
            return 0.0
        bits = re.search(r"LOL_F\(0x([0-9a-f]+)", expr).group(1)
        return float(np.array([int(bits, 16)], np.uint32).view(np.float32)[0])

    c = np.array([centre(m.group(1), "x"), centre(m.group(2), "y"), centre(m.group(3), "z")], np.float64)
    r1 = float(np.array([int(m.group(4), 16)], np.uint32).view(np.float32)[0])
    far = lb.Scene.from_string(text.replace("plane { y = -1,", "plane { y = -100000,"))
    rng = np.random.default_rng(5)
    pts = np.concatenate([rng.uniform(-60, 60, (300000, 3)), rng.uniform(-400, 400, (100000, 3))]).astype(np.float32)
    d, ids = oracle_sdf(far, pts)
    assert (ids == 1).all()
    promise = np.sqrt(((pts.astype(np.float64) - c) ** 2).sum(1)) - r1 / 1.004
    assert (d.astype(np.float64) >= promise).all(), float((d - promise).min())
    # the test as the kernel makes it, for running minima of every size: a skipped blob is farther than `best`
    best = rng.uniform(-3, 40, len(pts))
    u = best * 1.004 + r1
    skipped = (u > 0) & (((pts.astype(np.float64) - c) ** 2).sum(1) > u * u)
    assert skipped.sum() > 10000 and (d[skipped] > best[skipped]).all()
