"""Random scenes through the code generator, on the CPU: the generated distance function
(host shim, tests/oracle_lib.py) must equal the oracle's sdf() bit for bit on every point --
for straight-line code and forced table loops, with every pruning device on (boxes around
objects and groups, Morton-sorted rows, last-winner hints, tie-aware updates), for one ray
and for two rays per call, with same-shaped subtrees evaluated as packed pairs and one by
one, with the reference's node types and the CSG extensions.

Seeded numpy, not hypothesis: each case costs a g++ run, so the set is fixed and small.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol

HEAD = """materials {
  { shininess = 0, diffuse = (0,0,0), specular = (0,0,0), ambient = (0,0,0) },
  { shininess = 8, diffuse = (0.3,0.2,0.1), specular = (0.1,0.1,0.1), ambient = (0.3,0.2,0.1) },
  { shininess = 2, diffuse = (0.1,0.2,0.3), specular = (0.2,0.1,0.1), ambient = (0.1,0.2,0.3) } }
scene { camera { point = (0, 2, 6), direction = (0, -0.2, -1), fov = 90 },
  point_light { point = (3, 6, 2), diffuse_intensity = (3,3,3), specular_intensity = (3,3,3) },
"""


def _num(v):
    return f"{v:.4f}"


def _vec(v):
    return "(" + ", ".join(_num(x) for x in v) + ")"


def _leaf(rng, extensions):
    kind = rng.choice(["sphere", "sphere", "sphere", "box"])
    c = rng.uniform(-5, 5, 3) * [1, 0.5, 1] + [0, 1, -5]
    if rng.random() < 0.1:
        c[rng.integers(3)] = 0.0  # x - 0 is dropped by the lowering
    if kind == "sphere":
        return f"sphere {{ point = {_vec(c)}, radius = {_num(rng.uniform(0.2, 2.0))} }}"
    return (f"box {{ point = {_vec(c)}, point2 = {_vec(rng.uniform(0.2, 1.5, 3))}, "
            f"radius = {_num(rng.uniform(0.0, 0.5))} }}")


def _tree(rng, depth, extensions):
    if depth == 0 or rng.random() < 0.25:
        return _leaf(rng, extensions)
    a, b = _tree(rng, depth - 1, extensions), _tree(rng, depth - 1, extensions)
    kinds = ["smooth_union"] * 3 + (["union", "intersection", "difference"] if extensions else [])
    kind = rng.choice(kinds)
    if kind == "smooth_union":
        k = rng.choice([0.3, 0.5, 1.0, 3.0, float(rng.uniform(0.05, 4.0))])
        return f"smooth_union {{ smoothness = {_num(k)}, a = {a}, b = {b} }}"
    return f"{kind} {{ a = {a}, b = {b} }}"


def _material(text, m):
    head, rest = text.split("{", 1)
    return f"{head}{{ material = #{m},{rest}"


def random_head(rng):
    """Materials, camera and lights: material 0 is usually black (the miss shortcut is licensed)
    but not always; shininess is sometimes negative (the back-face cull is then not licensed);
    zero to three lights."""
    m0 = "(0,0,0)" if rng.random() < 0.7 else "(0.05,0.02,0.03)"
    shin = 8 if rng.random() < 0.8 else -2
    lights = "".join(
        f"  point_light {{ point = {_vec(rng.uniform(-8, 8, 3) + [0, 6, 0])}, diffuse_intensity = {_vec(rng.uniform(0.5, 4, 3))}, "
        f"specular_intensity = {_vec(rng.uniform(0.5, 4, 3))} }},\n" for _ in range(int(rng.integers(0, 4))))
    cam = rng.uniform(-2, 2, 3) + [0, 2, 6]
    return f"""materials {{
  {{ shininess = 0, diffuse = {m0}, specular = (0,0,0), ambient = {m0} }},
  {{ shininess = {shin}, diffuse = (0.3,0.2,0.1), specular = (0.1,0.1,0.1), ambient = (0.3,0.2,0.1) }},
  {{ shininess = 2, diffuse = (0.1,0.2,0.3), specular = (0.2,0.1,0.1), ambient = (0.1,0.2,0.3) }} }}
scene {{ ambient {{ color = (0.05, 0.05, 0.05) }},
  camera {{ point = {_vec(cam)}, direction = {_vec([-cam[0] * 0.1, -0.2, -1])}, fov = {int(rng.choice([60, 90, 120]))} }},
{lights}"""


def random_scene(seed, extensions, fixed_head=True):
    rng = np.random.default_rng(seed)
    head = HEAD if fixed_head else random_head(np.random.default_rng(seed + 7))
    objs = []
    for _ in range(rng.integers(1, 7)):
        objs.append(_material(_tree(rng, int(rng.integers(0, 4)), extensions), int(rng.integers(1, 3))))
    # a run of same-shaped objects: becomes a table loop when the threshold allows
    if rng.random() < 0.7:
        shape_seed = int(rng.integers(1 << 30))
        for i in range(int(rng.integers(2, 12))):
            # same structure (same seed for the structure), different constants
            srng = np.random.default_rng(shape_seed)
            text = _tree(srng, 2, extensions)
            crng = np.random.default_rng(seed * 1000 + i)
            import re
            text = re.sub(r"-?\d+\.\d+", lambda m: _num(float(m.group()) + crng.uniform(-0.4, 0.4))
                          if float(m.group()) > 0.06 else m.group(), text)
            objs.append(_material(text, 1 + i % 2))
    if rng.random() < 0.8:
        objs.insert(int(rng.integers(0, len(objs) + 1)), f"plane {{ y = {_num(rng.uniform(-3, 0))}, material = #2 }}")
    return head + ",\n".join("  " + o for o in objs) + " }\n"


def oracle_sdf(scene, pts):
    d = np.zeros(len(pts), np.float32)
    ids = np.zeros(len(pts), np.uint32)
    for i, p in enumerate(pts):
        pt = (C.c_float * 3)(*p.tolist())
        dd, ii = C.c_float(), C.c_uint32()
        ol.port().lolo_sdf(C.cast(scene._ptr, C.c_void_p), 0, C.byref(pt), C.byref(dd), C.byref(ii))
        d[i], ids[i] = dd.value, ii.value
    return d, ids


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("extensions", [False, True])
def test_random_scenes_lower_exactly(seed, extensions, tmp_path):
    import loltracer_b200 as lb

    text = random_scene(seed + (100 if extensions else 0), extensions)
    scene = lb.Scene.from_string(text)
    rng = np.random.default_rng(1000 + seed)
    pts = np.concatenate([rng.uniform(-30, 30, (600, 3)), rng.normal(0, 2.5, (1200, 3)) + [0, 1, -5],
                          rng.uniform(-8, 8, (400, 3)) * [1, 0.01, 1] + [0, -1, -5]]).astype(np.float32)
    want_d, want_id = oracle_sdf(scene, pts)
    # (variant, loop threshold, pruning, packed pairs): loops forced from 2 same-shaped neighbours on
    # pruning 4: every object that has a bounding ball around one of its own sphere centres is tested with it
    # (guarded and IEEE forms; scalar and with packed pairs) -- on half of the seeds, to keep the suite short
    sets = [(1, 0, 2, 2), (1, 2, 2, 2), (3, 2, 2, 1), (1, 2, 0, 1), (3, 0, 1, 1), (1, 0, 2, 0), (1, 2, 2, 0), (1, 2, 1, 3)]
    # (every seed through half of the option sets, plus the ball sets on even seeds: the suite stays short)
    sets = (sets[::2] + [(1, 0, 4, 0), (1, 99, 4, 2)]) if seed % 2 == 0 else sets[1::2]
    for variant, loops, prune, pack in sets:
        opt = lb.Options.default(variant=variant, loop_threshold=loops, prune_bounds=prune, guarded_fastpath=2,
                                 pack_pairs=pack)
        src = lb.lower_cuda(scene, opt)
        L = ol.cpu_sdf(tmp_path, src, f"fz{seed}{int(extensions)}_{variant}{loops}{prune}{pack}")
        d = np.zeros(len(pts), np.float32)
        ids = np.zeros(len(pts), np.uint32)
        fn = L.eval2 if "lol_sdf2(" in src.split("//@@SCENE@@")[0] and variant == 3 else L.eval
        fn(pts.ctypes.data_as(C.c_void_p), len(pts), d.ctypes.data_as(C.c_void_p), ids.ctypes.data_as(C.c_void_p))
        same = (d.view(np.uint32) == want_d.view(np.uint32)) | (np.isnan(d) & np.isnan(want_d))
        assert same.all(), (variant, loops, prune, text, pts[~same][:3], d[~same][:3], want_d[~same][:3])
        assert np.array_equal(ids, want_id), (variant, loops, prune, text)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("extensions", [False, True])
def test_random_scenes_render_like_the_oracle(seed, extensions):
    """The whole GPU path on random scenes (forced table loops, every pruning device, one and
    two rays per thread): distance and id of every pixel bit-identical to the oracle's."""
    pytest.importorskip("torch")
    import loltracer_b200 as lb
    from test_gpu_parity import _check, _render

    scene = lb.Scene.from_string(random_scene(seed + (100 if extensions else 0), extensions, fixed_head=False))
    w, h = 200, 112
    want = ol.port_render(scene, w, h)
    for variant, loops, prune, pack in [(1, 2, 2, 2), (3, 2, 2, 1), (1, 0, 1, 2), (1, 2, 2, 0), (1, 0, 2, 0), (1, 0, 4, 0)]:
        got = _render(lb, scene, w, h, options=lb.Options.default(variant=variant, loop_threshold=loops,
                                                                  prune_bounds=prune, guarded_fastpath=2,
                                                                  pack_pairs=pack))
        _check(got, want)
        got["renderer"].close()
