"""The per-pixel pipeline of the generated program, WITHOUT a GPU: lol_kernel.cuh's lol_shade_pixel
(variant 1: camera ray, primary march, normal taps, shadow marches, Phong, gamma, pack) is compiled for
the host together with the generated distance code (tests/oracle_lib.py: cpu_pipeline) and compared with
the oracle pixel for pixel.  On the host powf is glibc's -- the reference's -- so even RGB is exact here;
on the GPU the same text runs with CUDA's powf (<= 1/255, tests/test_gpu_parity.py).

What this buys: the kernel-side logic -- the exact skips, the shared first step, the division pre-test,
packed pairs, table loops, the two-ray bookkeeping of variant 3 -- is checked on every CPU run, and a change to it can be developed without
GPU time.  What it cannot see: ptxas (contraction, MUFU sequences); that stays with the GPU suite.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
from conftest import EXAMPLES


def _same(got, want, shadow_counts=False):
    assert np.array_equal(got["id"], want["id"])
    d, w = got["dist"], want["dist"]
    assert ((d.view(np.uint32) == w.view(np.uint32)) | (np.isnan(d) & np.isnan(w))).all()
    assert np.array_equal(got["nprimary"], want["nprimary"])
    assert np.array_equal(got["rgba"], want["rgba"])  # exact: same libm on both sides
    if shadow_counts:
        assert np.array_equal(got["nshadow"], want["nshadow"])


OPTIONS = {
    "default": dict(),
    "no_skips": dict(skip_black_miss=0, cull_backfacing=0, shadow_early_out=0, share_first_step=0),
    "everything_on": dict(guarded_fastpath=2, pack_pairs=2, share_first_step=2, shadow_div_pretest=1, prune_bounds=2),
    "ieee_forms": dict(guarded_fastpath=0, prune_bounds=0),
    "forced_balls": dict(guarded_fastpath=2, prune_bounds=4),
    "guard_inside_the_loops": dict(guard_out=0, guarded_fastpath=2),
    "forced_loops": dict(loop_threshold=2, guarded_fastpath=2, share_first_step=2),
}


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("opts", list(OPTIONS))
def test_host_compiled_pipeline_equals_oracle(name, opts, scenes_dir, tmp_path):
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    src = lb.lower_cuda(scene, lb.Options.default(variant=1, **OPTIONS[opts]))
    L = ol.cpu_pipeline(tmp_path, src, f"{name}_{opts}")
    for w, h in ((96, 54), (37, 23)):
        got = ol.cpu_pipeline_render(L, lb, scene, w, h)
        want = ol.port_render(scene, w, h, counts=True)
        # with the skips off the shadow marches take exactly the reference's steps
        _same(got, want, shadow_counts=(opts == "no_skips"))


@pytest.mark.parametrize("name", EXAMPLES)
def test_host_compiled_pipeline_at_the_reference_window_size(name, scenes_dir, tmp_path):
    """BASELINE config C1's size (main.c:136-137, 320x240) through the default program of every example --
    on scene4 that is the ball test, the guard's fall-back outside the march loops and the unconditional first
    object -- against the oracle: distance, id, step counts and RGB of all 76 800 pixels, exact."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    src = lb.lower_cuda(scene, lb.Options.default(variant=1))
    L = ol.cpu_pipeline(tmp_path, src, f"{name}_window")
    _same(ol.cpu_pipeline_render(L, lb, scene, 320, 240), ol.port_render(scene, 320, 240, counts=True))


@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("loops", [0, 2])
def test_host_compiled_two_ray_pipeline_equals_oracle(name, loops, scenes_dir, tmp_path):
    """Variant 3 (lol_shade_pair: two horizontally adjacent pixels per call, a finished ray waiting for
    its partner) on the shim's plain pairs: every pixel the oracle's, odd widths included (the last
    pixel of a row has no partner)."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    src = lb.lower_cuda(scene, lb.Options.default(variant=3, loop_threshold=loops, prune_bounds=2 if loops else 1))
    assert "#define LOL_VARIANT 3" in src
    L = ol.cpu_pipeline(tmp_path, src, f"{name}_v3_{loops}")
    for w, h in ((96, 54), (37, 23)):
        _same(ol.cpu_pipeline_render(L, lb, scene, w, h), ol.port_render(scene, w, h, counts=True))


def test_host_compiled_pipeline_on_the_1024_sphere_scene(tmp_path):
    """Pruned table loops with hints, Morton groups and packed pairs under the whole pipeline."""
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen

    scene = lb.Scene.from_string(scenegen.synthetic_scene_text())
    src = lb.lower_cuda(scene, lb.Options.default(variant=1))
    assert "lol_sqrt_fast2(" in src.split("//@@SCENE@@")[0]
    L = ol.cpu_pipeline(tmp_path, src, "synthetic")
    w, h = 40, 22
    _same(ol.cpu_pipeline_render(L, lb, scene, w, h), ol.port_render(scene, w, h, counts=True))


@pytest.mark.parametrize("name,point,direction", [
    ("scene4", (0, 1, -6), (0, 0, -1)),          # at a sphere's centre: sqrt(0), the guard's fallback
    ("scene4", (3e19, 1, 0), (-1, 0, 0)),        # beyond 2^60
    ("scene2", (0, 5, -6), (0, -1, 0)),          # cross(dir, up) = 0: a NaN camera basis
    ("scene", (2, 2, -10), (0, 0, -1)),          # inside the round box: the march ends on step 1
    ("scene", (-0.0, 0, -0.0), (0, 0, -1)),      # -0 + rd * 0 depends on rd: the shared first step stands aside
    ("scene4", (-0.0, 6, 3), (0.3, -0.7, -1)),
    ("scene4", (0, 1, 58), (0, 0, -1)),          # aimed at a sphere's centre from outside (the march ends at the surface: no step may fail the guard)
    ("scene3", (np.nan, 2, 3), (0, 0, -1)),      # a NaN coordinate: min.NaN / max.NaN carry it to the guard's test
])
@pytest.mark.parametrize("pretest", [1, 0])     # 0: the guard outside the march loops (LOL_GUARD_OUT)
def test_host_compiled_pipeline_edge_cameras(name, point, direction, pretest, scenes_dir, tmp_path):
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    cam = lb.Camera.make(list(point), list(direction), scene.struct.camera.fov)
    src = lb.lower_cuda(scene, lb.Options.default(variant=1, guarded_fastpath=2, share_first_step=2, shadow_div_pretest=pretest))
    assert f"#define LOL_GUARD_OUT {3 * (1 - pretest)}" in src
    L = ol.cpu_pipeline(tmp_path, src, "edge")
    w, h = 65, 37  # odd: the centre pixel's ray runs along the camera direction
    _same(ol.cpu_pipeline_render(L, lb, scene, w, h, camera=cam), ol.port_render(scene, w, h, camera=cam, counts=True))


@pytest.mark.parametrize("seed", range(4))
def test_host_compiled_pipeline_on_random_scenes(seed, tmp_path):
    """Random trees, materials (non-black misses, negative shininess: the skips must switch themselves
    off), cameras and 0-3 lights through the whole pipeline, straight-line and as forced table loops."""
    import loltracer_b200 as lb
    from test_lowering_fuzz import random_scene

    scene = lb.Scene.from_string(random_scene(seed + 300, extensions=False, fixed_head=False))
    w, h = 48, 27
    want = ol.port_render(scene, w, h, counts=True)
    for tag, kw in (("s", dict(guarded_fastpath=2, pack_pairs=2, share_first_step=2, shadow_div_pretest=1)),
                    ("l", dict(guarded_fastpath=2, loop_threshold=2, prune_bounds=2)),
                    # every object that has one behind its bounding ball, the guard's fall-back outside the loops
                    ("b", dict(guarded_fastpath=2, loop_threshold=99, prune_bounds=4))):
        src = lb.lower_cuda(scene, lb.Options.default(variant=1, **kw))
        L = ol.cpu_pipeline(tmp_path, src, f"fz{seed}{tag}")
        _same(ol.cpu_pipeline_render(L, lb, scene, w, h), want)


# ---- variant 4 (staged): the resumable pipeline behind "deferring long rays" ----------------------

@pytest.mark.parametrize("name", EXAMPLES)
@pytest.mark.parametrize("opts", ["default", "no_skips", "forced_loops"])
def test_resumable_pipeline_equals_oracle_at_every_cap(name, opts, scenes_dir, tmp_path):
    """lol_pixel_run (lol_kernel.cuh, variant 4): every march stops after `cap` evaluations, the pixel goes
    into a 17-word continuation record and is picked up again in a later pass -- as often as it takes.  At
    caps 48/24 (the plan), 5/3 and 1/1 (every single evaluation is a resume point: primary march, normal,
    each light's setup, shadow march, Phong) the frame, distances, ids and step counts are the oracle's,
    RGB included.  With the skips off the shadow step counts are the reference's too."""
    import ctypes as C

    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, name + ".lol"))
    kw = dict(OPTIONS[opts])
    kw.pop("share_first_step", None)  # variant 1's prologue; variant 4 has none yet
    src = lb.lower_cuda(scene, lb.Options.default(variant=4, **kw))
    assert "#define LOL_VARIANT 4" in src
    L = ol.cpu_pipeline(tmp_path, src, f"{name}_v4_{opts}")
    L.lol_host_get_deferrals.restype = C.c_long
    w, h = 64, 36
    want = ol.port_render(scene, w, h, counts=True)
    seen = []
    for caps in ((256, 128), (48, 24), (5, 3), (1, 1)):
        L.lol_host_set_caps(*caps)
        _same(ol.cpu_pipeline_render(L, lb, scene, w, h), want, shadow_counts=(opts == "no_skips"))
        seen.append(L.lol_host_get_deferrals())
    assert seen[0] == 0 and seen[1] > 0 and seen[3] > seen[2] > seen[1]  # the caps did interrupt marches


@pytest.mark.parametrize("seed", range(3))
def test_resumable_pipeline_on_random_scenes_and_edge_cameras(seed, tmp_path):
    import loltracer_b200 as lb
    from test_lowering_fuzz import random_scene

    scene = lb.Scene.from_string(random_scene(seed + 400, extensions=bool(seed % 2), fixed_head=False))
    src = lb.lower_cuda(scene, lb.Options.default(variant=4, guarded_fastpath=2, pack_pairs=2))
    L = ol.cpu_pipeline(tmp_path, src, f"fz4_{seed}")
    w, h = 40, 22
    cams = [None, lb.Camera.make([-0.0, 2, 6], [0.1, -0.2, -1], 90), lb.Camera.make([0, 5, -6], [0, -1, 0], 90)]
    for cam in cams:
        want = ol.port_render(scene, w, h, camera=cam, counts=True)
        for caps in ((7, 4), (1, 1)):
            L.lol_host_set_caps(*caps)
            _same(ol.cpu_pipeline_render(L, lb, scene, w, h, camera=cam), want)


def test_deferred_rays_variant_compiles_for_sm_100a(scenes_dir):
    """Variant 4's program -- lol_render with caps and the continuation queue, lol_resume -- goes through
    NVRTC for sm_100a, and its march loops are no longer than variant 1's (the cap costs one compare)."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_file(os.path.join(scenes_dir, "scene4.lol"))
    opt = lb.Options.default(variant=4)
    image = lb.compile_cubin(lb.lower_cuda(scene, opt), opt)
    sass = lb.disassemble(image)
    assert "lol_render" in sass and "lol_resume" in sass


@pytest.mark.parametrize("seed", range(6))
def test_shadow_early_out_with_shadow_origins_inside_objects(seed, tmp_path):
    """the shadow march's early-out (leave once res <= 0) must not change a pixel even when
    the shadow ray starts inside or on an object -- negative distances, 0/0 on the first step, res
    going NaN and coming back.  Crowded scenes: overlapping spheres and boxes as separate top-level
    objects (a hit point on one is inside its neighbour one unit along the light direction), lights
    inside objects and on surfaces.  Early-out on, off, and the oracle: the same frame, RGB included
    (glibc powf on both sides); with it off, the reference's shadow step counts."""
    import loltracer_b200 as lb

    rng = np.random.default_rng(9100 + seed)
    objs, centres = [], []
    for i in range(int(rng.integers(5, 9))):
        c = rng.uniform(-2.5, 2.5, 3) * [1, 0.4, 1] + [0, 1.2, -6]
        centres.append(c)
        m = 1 + i % 2
        if rng.random() < 0.7:
            objs.append(f"sphere {{ material = #{m}, point = ({c[0]:.3f},{c[1]:.3f},{c[2]:.3f}), radius = {rng.uniform(1.2, 2.4):.3f} }}")
        else:
            e = rng.uniform(0.8, 1.8, 3)
            objs.append(f"box {{ material = #{m}, point = ({c[0]:.3f},{c[1]:.3f},{c[2]:.3f}), "
                        f"point2 = ({e[0]:.3f},{e[1]:.3f},{e[2]:.3f}), radius = {rng.uniform(0, 0.4):.3f} }}")
    objs.append("plane { material = #2, y = -0.5 }")
    lights = []
    for k in range(int(rng.integers(1, 4))):
        if k == 0:      # inside an object
            p = centres[int(rng.integers(len(centres)))] + rng.uniform(-0.3, 0.3, 3)
        elif k == 1:    # exactly one unit above the floor: p + dir lands on it for hits below
            p = np.array([rng.uniform(-3, 3), 0.5, rng.uniform(-8, -4)])
        else:
            p = rng.uniform(-6, 6, 3) + [0, 6, -3]
        lights.append(f"point_light {{ point = ({p[0]:.3f},{p[1]:.3f},{p[2]:.3f}), diffuse_intensity = (2,2,2), "
                      f"specular_intensity = (1,1,1) }}")
    text = """materials {
  { shininess = 0, diffuse = (0,0,0), specular = (0,0,0), ambient = (0,0,0) },
  { shininess = 8, diffuse = (0.3,0.2,0.1), specular = (0.1,0.1,0.1), ambient = (0.3,0.2,0.1) },
  { shininess = 2, diffuse = (0.1,0.2,0.3), specular = (0.2,0.1,0.1), ambient = (0.1,0.2,0.3) } }
scene { ambient { color = (0.05, 0.05, 0.05) }, camera { point = (0, 2.5, 2), direction = (0, -0.25, -1), fov = 90 },
""" + ",\n".join(lights + objs) + " }"
    scene = lb.Scene.from_string(text)
    w, h = 64, 36
    want = ol.port_render(scene, w, h, counts=True)
    assert (want["nshadow"] > 0).any()
    for early in (1, 0):
        src = lb.lower_cuda(scene, lb.Options.default(variant=1, shadow_early_out=early, cull_backfacing=0))
        assert f"#define LOL_SHADOW_EARLY {early}" in src
        L = ol.cpu_pipeline(tmp_path, src, f"se{seed}_{early}")
        got = ol.cpu_pipeline_render(L, lb, scene, w, h)
        _same(got, want, shadow_counts=False)
        if not early:
            hit = want["id"] != 0
            assert np.array_equal(got["nshadow"][hit], want["nshadow"][hit])


def test_shadow_early_out_is_refused_for_non_finite_scene_constants():
    """The early-out's proof needs sdf() finite at finite points: a scene with an infinite constant
    keeps the full march (lolb200_can_shadow_early)."""
    import loltracer_b200 as lb
    from test_lowering_fuzz import HEAD

    ok = lb.Scene.from_string(HEAD + "sphere { material = #1, point = (0,1,-5), radius = 1 } }")
    assert "#define LOL_SHADOW_EARLY 1" in lb.lower_cuda(ok, lb.Options.default(variant=1))
    s = ok.struct
    s.nodes[0].radius = float("inf")
    assert "#define LOL_SHADOW_EARLY 0" in lb.lower_cuda(ok, lb.Options.default(variant=1))


# ---- the per-ray candidate memory of pruned table loops (lol_kernel.cuh: struct lol_near) -------------

def _near_scene(kind):
    """Scenes with ONE pruned table loop: the 1024-sphere scene; a crowd of overlapping equal spheres (more
    than four rows within reach at once: the memory overflows and the evaluation goes the long way); a
    sparse row of spheres (long stretches where nothing but the floor is near)."""
    from loltracer_b200 import scenegen
    from test_lowering_fuzz import HEAD

    if kind == "synthetic":
        return scenegen.synthetic_scene_text()
    rng = np.random.default_rng(77)
    if kind == "crowd":
        objs = [f"sphere {{ material = #{1 + i % 2}, point = ({rng.uniform(-1.5, 1.5):.3f},{rng.uniform(0.2, 2.2):.3f},"
                f"{rng.uniform(-6.5, -4.5):.3f}), radius = {rng.uniform(0.8, 1.4):.3f} }}" for i in range(24)]
    else:
        objs = [f"sphere {{ material = #{1 + i % 2}, point = ({-14 + 1.5 * i:.3f},{0.5 + 0.1 * (i % 3):.3f},"
                f"{-6 - (i % 5):.3f}), radius = 0.45 }}" for i in range(20)]
    return HEAD + ",\n".join(objs + ["plane { material = #2, y = -0.5 }"]) + " }"


@pytest.mark.parametrize("kind", ["synthetic", "crowd", "sparse"])
def test_candidate_memory_of_pruned_loops_is_exact(kind, tmp_path):
    """options.near_cache: a ray remembers the rows it could not skip and how far it may move before the others
    must be looked at again.  With it and without it: the oracle's frame, distances, ids and step counts (RGB
    exact on the host), also where the memory overflows (crowd) and with every shortcut off."""
    import loltracer_b200 as lb

    scene = lb.Scene.from_string(_near_scene(kind))
    w, h = (64, 36) if kind == "synthetic" else (96, 54)
    want = ol.port_render(scene, w, h, counts=True)
    for tag, kw in (("on", dict(near_cache=1)), ("off", dict(near_cache=0)),
                    ("ieee", dict(near_cache=1, guarded_fastpath=0, pack_pairs=0)),
                    ("noskip", dict(near_cache=1, skip_black_miss=0, cull_backfacing=0, shadow_early_out=0)),
                    # 3: looks are answered by the candidate grid (every cell knows its eight nearest rows)
                    ("grid", dict(near_cache=3)),
                    ("grid_ieee_noskip", dict(near_cache=3, guarded_fastpath=0, pack_pairs=0, skip_black_miss=0,
                                              cull_backfacing=0, shadow_early_out=0))):
        src = lb.lower_cuda(scene, lb.Options.default(variant=1, loop_threshold=8, **kw))
        assert f"#define LOL_NEAR {min(int(kw['near_cache']), 2)}" in src
        assert ("#define LOL_NEAR_GRID (1 &&" in src) == (kw["near_cache"] == 3)
        L = ol.cpu_pipeline(tmp_path, src, f"near_{kind}_{tag}")
        _same(ol.cpu_pipeline_render(L, lb, scene, w, h), want, shadow_counts=(tag == "noskip"))
    # cameras inside the crowd, far away (beyond the coordinate range the memory trusts) and NaN
    cams = [lb.Camera.make([0, 1, -5.5], [0.2, -0.1, -1], 1.5), lb.Camera.make([0, 2, 900], [0, 0, -1], 0.3),
            lb.Camera.make([0, 5, -6], [0, -1, 0], 1.5)]
    for near in (1, 3):
        src = lb.lower_cuda(scene, lb.Options.default(variant=1, loop_threshold=8, near_cache=near))
        L = ol.cpu_pipeline(tmp_path, src, f"near_{kind}_cams{near}")
        for cam in cams:
            _same(ol.cpu_pipeline_render(L, lb, scene, 48, 27, camera=cam), ol.port_render(scene, 48, 27, camera=cam, counts=True))


GRID_CHECK = r"""
// The candidate grid's promise, checked point by point: for a point p of a cell, every LISTED row has
// w <= dbox_row(p) - m1_row, and every row that is NOT listed has rest <= dbox_row(p) - m1_row -- so a
// row left out of a look cannot fail to be skipped.  Exact side in double precision.
#include <cstdint>
extern "C" long grid_check(long n, unsigned long long seed, double* worst) {
    const long cells = 2l * LOL_GRID_N * LOL_GRID_N * LOL_GRID_N;
    for (long ci = 0; ci < cells; ++ci) lol_grid_build_cell((int)ci);
    long bad = 0; *worst = 1e30;
    auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (double)(seed >> 11) / 9007199254740992.0; };
    for (long s = 0; s < n; ++s) {
        const int level = s & 1;
        const double sc = level ? (double)LOL_GRID_OUTER : 1.0;
        const double x0 = level ? (double)LOL_GRID_X1 : (double)LOL_GRID_X0, y0 = level ? (double)LOL_GRID_Y1 : (double)LOL_GRID_Y0,
                     z0 = level ? (double)LOL_GRID_Z1 : (double)LOL_GRID_Z0;
        const float px = (float)(x0 + rnd() * LOL_GRID_N * (double)LOL_GRID_SX * sc), py = (float)(y0 + rnd() * LOL_GRID_N * (double)LOL_GRID_SY * sc),
                    pz = (float)(z0 + rnd() * LOL_GRID_N * (double)LOL_GRID_SZ * sc);
        // the cell as the look computes it (same float arithmetic)
        const float fx = (px - (level ? LOL_GRID_X1 : LOL_GRID_X0)) * (level ? LOL_GRID_IX / LOL_GRID_OUTER : LOL_GRID_IX);
        const float fy = (py - (level ? LOL_GRID_Y1 : LOL_GRID_Y0)) * (level ? LOL_GRID_IY / LOL_GRID_OUTER : LOL_GRID_IY);
        const float fz = (pz - (level ? LOL_GRID_Z1 : LOL_GRID_Z0)) * (level ? LOL_GRID_IZ / LOL_GRID_OUTER : LOL_GRID_IZ);
        const float top = (float)LOL_GRID_N;
        if (!(fx >= 0.f && fx < top && fy >= 0.f && fy < top && fz >= 0.f && fz < top)) continue;
        const lol_cell& c = lol_grid[level * LOL_GRID_N * LOL_GRID_N * LOL_GRID_N + ((int)fz * LOL_GRID_N + (int)fy) * LOL_GRID_N + (int)fx];
        for (int i = 0; i < LOL_RUN0_ROWS; ++i) {
            const lol_u32* ct = lol_tables + lol_run0_offset + i * LOL_RUN0_STRIDE;
            const double qx = std::fmax(std::fabs((double)px - LOL_TF(ct[0])) - LOL_TF(ct[3]), 0.0);
            const double qy = std::fmax(std::fabs((double)py - LOL_TF(ct[1])) - LOL_TF(ct[4]), 0.0);
            const double qz = std::fmax(std::fabs((double)pz - LOL_TF(ct[2])) - LOL_TF(ct[5]), 0.0);
            const double gap = std::sqrt(qx * qx + qy * qy + qz * qz) - (double)LOL_TF(ct[6]);
            double bound = c.rest;
            for (int k = 0; k < 8; ++k)
                if (((c.rows[k >> 2] >> (8 * (k & 3))) & 0xffu) == (lol_u32)i) bound = c.w[k];
            if (gap - bound < *worst) *worst = gap - bound;
            bad += !(bound <= gap);
        }
        for (int k = 1; k < 8; ++k) bad += !(c.w[k - 1] <= c.w[k]);   // ascending
        bad += !(c.w[7] <= c.rest);
    }
    return bad;
}
"""


@pytest.mark.parametrize("kind,cells", [("synthetic", 32), ("synthetic", 64), ("crowd", 16), ("sparse", 48)])
def test_candidate_grid_bounds_are_conservative(kind, cells, tmp_path):
    """lol_grid_build_cell (generated; on the GPU a one-off kernel runs it): 200 000 random points in both
    levels of the grid, every row against its cell's bound in double precision -- no row's true
    dbox(p) - m1 may be below what its cell promises, lists ascend, and the slack is small but positive."""
    import ctypes as C
    import pathlib
    import subprocess
    import loltracer_b200 as lb

    scene = lb.Scene.from_string(_near_scene(kind))
    src = lb.lower_cuda(scene, lb.Options.default(variant=1, loop_threshold=8, near_cache=3, grid_cells=cells))
    assert "#define LOL_NEAR_GRID (1 &&" in src and f"#define LOL_GRID_N {cells}" in src
    cu = pathlib.Path(tmp_path) / "grid.cpp"
    cu.write_text(ol.HOST_SHIM + src + GRID_CHECK)
    so = pathlib.Path(tmp_path) / "grid.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(cu)])
    L = C.CDLL(str(so))
    L.grid_check.restype = C.c_long
    L.grid_check.argtypes = [C.c_long, C.c_ulonglong, C.POINTER(C.c_double)]
    worst = C.c_double()
    assert L.grid_check(200000, 12345, C.byref(worst)) == 0, worst.value
    assert 0.0 <= worst.value < 0.05, worst.value
