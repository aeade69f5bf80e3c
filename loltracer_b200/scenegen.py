"""Workload generators for the BASELINE configs that are not example files.

* synthetic_scene_text(): config C4, a seeded .lol scene with 1024 spheres in 128
  top-level smooth-union trees.  It is emitted as .lol TEXT so the
  CPU oracle and the GPU backend parse the very same input; numbers are printed
  with %.6f because the lexer has no exponent syntax (scene-lexer.l:12).
* orbit_camera(): config C5, frame k of n of a camera orbit about the vertical
  axis through (0, 1, -6), applied the way main.c mutates scene->camera in place
  (main.c:71-112).
"""
from __future__ import annotations

import math
import random
from typing import Sequence

import numpy as np

from .api import Camera

SYNTHETIC_SEED = 20261018


def _v(x: Sequence[float]) -> str:
    return "(" + ", ".join(f"{c:.6f}" for c in x) + ")"


def _tree(spheres, k: float, indent: str) -> str:
    """Balanced smooth_union tree over a list of (centre, radius)."""
    if len(spheres) == 1:
        c, r = spheres[0]
        return f"sphere {{ point = {_v(c)}, radius = {r:.6f} }}"
    half = len(spheres) // 2
    a = _tree(spheres[:half], k, indent + "\t")
    b = _tree(spheres[half:], k, indent + "\t")
    return (f"smooth_union {{\n{indent}\tsmoothness = {k:.6f},\n{indent}\ta = {a},\n"
            f"{indent}\tb = {b}\n{indent}}}")


def _csg_tree(cells, rng, k: float, indent: str) -> str:
    """Balanced smooth_union tree whose leaves are hard CSG pairs (extension nodes): cell
    2i becomes a lens, intersection { sphere, shifted sphere }, cell 2i+1 a bitten sphere,
    difference { sphere, shifted sphere } -- two primitives per cell."""
    if len(cells) == 1:
        (c, r), kind = cells[0]
        r = 1.4 * r
        off = [rng.uniform(-0.6, 0.6) * r for _ in range(3)]
        c2 = (c[0] + off[0], c[1] + off[1], c[2] + off[2])
        a = f"sphere {{ point = {_v(c)}, radius = {r:.6f} }}"
        b = f"sphere {{ point = {_v(c2)}, radius = {0.9 * r:.6f} }}"
        return f"{kind} {{\n{indent}\ta = {a},\n{indent}\tb = {b}\n{indent}}}"
    half = len(cells) // 2
    a = _csg_tree(cells[:half], rng, k, indent + "\t")
    b = _csg_tree(cells[half:], rng, k, indent + "\t")
    return (f"smooth_union {{\n{indent}\tsmoothness = {k:.6f},\n{indent}\ta = {a},\n"
            f"{indent}\tb = {b}\n{indent}}}")


def synthetic_scene_text(n_trees: int = 128, leaves: int = 8, seed: int = SYNTHETIC_SEED,
                         smoothness: float = 0.3, csg: bool = False) -> str:
    """csg=True: BASELINE config C4 as worded ("unioned/intersected primitives"): the same
    1024 primitives, but every two of them form an intersection or a difference node --
    extension nodes our front-end adds (include/lolb200.h); the reference's grammar and
    renderer have none, so only the oracle port can check that scene."""
    rng = random.Random(seed)
    nx, ny, nz = 16, 8, 8
    assert n_trees * leaves <= nx * ny * nz and nx % leaves == 0
    cells = []
    for iz in range(nz):
        for iy in range(ny):
            for ix in range(nx):
                cx = -8.0 + (ix + 0.5) * (16.0 / nx) + rng.uniform(-0.3, 0.3)
                cy = 0.0 + (iy + 0.5) * (6.0 / ny) + rng.uniform(-0.2, 0.2)
                cz = -20.0 + (iz + 0.5) * (16.0 / nz) + rng.uniform(-0.5, 0.5)
                cells.append(((cx, cy, cz), rng.uniform(0.15, 0.45)))
    mats = [(4, (0, 0, 0), (0, 0, 0), (0, 0, 0)),
            (3, (0.2, 0, 0), (0.2, 0.2, 0.2), (0.2, 0, 0)),
            (50, (0, 0.2, 0), (0.2, 0.2, 0.2), (0, 0.2, 0)),
            (2, (0, 0, 0.2), (0.01, 0.01, 0.01), (0, 0, 0.2)),
            (10, (0.2, 0.2, 0), (0.001, 0.001, 0.001), (0.2, 0.2, 0))]
    out = ["materials {"]
    out.append(",\n".join(
        f"\t{{ shininess = {s}, diffuse = {_v(d)}, specular = {_v(sp)}, ambient = {_v(a)} }}"
        for s, d, sp, a in mats))
    out.append("}\n\nscene {")
    comps = ["\tambient { color = (0.03, 0.03, 0.03) }",
             "\tcamera { point = (0, 4, 4), direction = (0, -0.3, -1), fov = 90 }",
             "\tpoint_light { point = (-2, 10, -1), diffuse_intensity = (4, 4, 4), "
             "specular_intensity = (4, 4, 4) }",
             "\tpoint_light { point = (-7, 2, -5), diffuse_intensity = (1, 1.5, 2), "
             "specular_intensity = (1, 1.5, 2) }"]
    for t in range(n_trees):
        group = cells[t * leaves:(t + 1) * leaves]
        if csg:  # two primitives per cell: every other cell of the group
            kinds = ["intersection", "difference"]
            tree = _csg_tree([(cell, kinds[i % 2]) for i, cell in enumerate(group[::2])], rng, smoothness, "\t")
        else:
            tree = _tree(group, smoothness, "\t")
        # material goes on the top-level node only (naive_renderer.c:102-112)
        head, rest = tree.split("{", 1)
        comps.append(f"\t{head}{{ material = #{1 + t % 4},{rest}")
    comps.append("\tplane { y = -1, material = #4 }")
    out.append(",\n".join(comps))
    out.append("}\n")
    return "\n".join(out)


def orbit_camera(cam0: Camera, k: int, n: int = 64,
                 pivot: Sequence[float] = (0.0, 1.0, -6.0)) -> Camera:
    th = 2.0 * math.pi * k / n
    c, s = math.cos(th), math.sin(th)

    def rot(v):
        return (c * v[0] + s * v[2], v[1], -s * v[0] + c * v[2])

    p0 = [float(x) for x in cam0.point]
    d0 = [float(x) for x in cam0.direction]
    rel = rot([p0[i] - pivot[i] for i in range(3)])
    point = np.asarray([pivot[i] + rel[i] for i in range(3)], np.float32)
    direction = np.asarray(rot(d0), np.float32)
    if k % n == 0:  # frame 0 is the file camera, bit for bit
        point, direction = np.asarray(p0, np.float32), np.asarray(d0, np.float32)
    return Camera.make(point.tolist(), direction.tolist(), cam0.fov)
