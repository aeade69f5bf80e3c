#!/usr/bin/env python
"""Regenerates tests/golden/ from the reference tree.  Run in the build container:

    python tests/golden/make_golden.py            (needs /root/reference)

It cannot run on the GPU box (no reference there); its outputs are committed.

1. scenes/<name>.lol -- the reference's example inputs (examples/*.lol), re-emitted
   token by token in one canonical layout by the writer below.  They are INPUT
   DATA of the benchmark ("every examples/*.lol scene"), not reference source;
   the script asserts that the compiled reference parses original and canonical
   file into bit-identical scenes.
2. frames.npz -- per scene, at small sizes, the (dist, id, pixel) of every pixel
   computed by the reference's own naive_renderer.c (oracle/_ref/liblolref.so).
3. hashes.json -- frame hashes of the reference at 320x240 / 1920x1080 /
   3840x2160 (the 320x240 and 1080p/4K values also appear in SURVEY.md 8c),
   plus hashes of orbit-camera frames and of the synthetic scene.
"""
import ctypes as C
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("LOL_REFERENCE", "/root/reference")

KEYWORDS = ["materials", "scene", "ambient", "camera", "point-light", "point_light", "sphere",
            "box", "plane", "smooth_union", "smooth-union", "shininess", "diffuse", "specular",
            "color", "point", "direction", "fov", "diffuse_intensity", "diffuse-intensity",
            "specular_intensity", "specular-intensity", "radius", "material", "point2", "y",
            "smoothness", "a", "b"]
CANON = {"point-light": "point_light", "smooth-union": "smooth_union",
         "diffuse-intensity": "diffuse_intensity", "specular-intensity": "specular_intensity"}
TYPES = {"ambient", "camera", "point_light", "sphere", "box", "plane", "smooth_union"}


def tokenize(text):
    """scene-lexer.l:10-50: longest match, unknown characters dropped."""
    toks, i = [], 0
    kws = sorted(KEYWORDS, key=len, reverse=True)
    while i < len(text):
        c = text[i]
        if c in " \r\t\n":
            i += 1
            continue
        m = re.match(r"[-.0-9]+", text[i:])
        if m:
            toks.append(("num", m.group(0)))
            i += len(m.group(0))
            continue
        m = re.match(r"#[0-9]+", text[i:])
        if m:
            toks.append(("id", m.group(0)))
            i += len(m.group(0))
            continue
        for k in kws:
            if text.startswith(k, i):
                toks.append(("kw", CANON.get(k, k)))
                i += len(k)
                break
        else:
            if c in ",(){}=":
                toks.append(("p", c))
            i += 1
    return toks


def canonical(text):
    """Compact layout: a block without nested blocks goes on one line."""
    toks = tokenize(text)
    pos = 0

    def parse_block():
        """After '{': returns the list of comma-separated entries up to '}'."""
        nonlocal pos
        entries, cur = [], []
        while True:
            kind, v = toks[pos]
            pos += 1
            if kind == "p" and v == "{":
                cur.append(parse_block())
            elif kind == "p" and v == "}":
                if cur:
                    entries.append(cur)
                return entries
            elif kind == "p" and v == ",":
                entries.append(cur)
                cur = []
            elif kind == "p" and v == "(":
                nums = []
                while toks[pos] != ("p", ")"):
                    if toks[pos][0] == "num":
                        nums.append(toks[pos][1])
                    pos += 1
                pos += 1
                cur.append("(" + ", ".join(nums) + ")")
            elif kind == "p" and v == "=":
                cur.append("=")
            else:
                cur.append(v)

    def flat(entries):
        return all(not isinstance(x, list) for e in entries for x in e)

    def fmt_entry(e, depth):
        return " ".join(fmt_block(x, depth) if isinstance(x, list) else x for x in e)

    def fmt_block(entries, depth):
        if flat(entries):
            return "{ " + ", ".join(fmt_entry(e, depth) for e in entries) + " }"
        pad = "\t" * (depth + 1)
        return ("{\n" + ",\n".join(pad + fmt_entry(e, depth + 1) for e in entries) + "\n" +
                "\t" * depth + "}")

    out = []
    while pos < len(toks):
        kind, v = toks[pos]
        pos += 1
        if kind == "p" and v == "{":
            out.append(fmt_block(parse_block(), 0))
        else:
            out.append(v)
    return out[0] + " " + out[1] + "\n\n" + out[2] + " " + out[3] + "\n"


def scene_bytes(flat):
    """Everything the renderer reads, as bytes, for equality checks."""
    from loltracer_b200 import api

    s = flat.struct
    parts = [bytes(C.string_at(s.materials, s.n_materials * C.sizeof(api.Material))),
             bytes(s.ambient_color),
             bytes(C.string_at(s.lights, s.n_lights * C.sizeof(api.Light))),
             bytes(C.string_at(s.nodes, s.n_nodes * C.sizeof(api.Object))),
             bytes(C.string_at(s.objects, s.n_objects * 4)), bytes(s.camera)]
    return b"|".join(parts)


def main():
    import __graft_entry__ as entry

    entry.build()
    import loltracer_b200 as lb
    from loltracer_b200 import scenegen
    import oracle_lib as ol

    assert ol.have_ref(), "oracle/_ref/liblolref.so missing (needs /root/reference)"
    os.makedirs(os.path.join(HERE, "scenes"), exist_ok=True)
    names = ["scene", "scene2", "scene3", "scene4"]

    for n in names:
        src = open(os.path.join(REF, "examples", n + ".lol")).read()
        canon = canonical(src)
        a = ol.RefScene(text=src).flatten()
        b = ol.RefScene(text=canon).flatten()
        assert scene_bytes(a) == scene_bytes(b), n
        with open(os.path.join(HERE, "scenes", n + ".lol"), "w") as f:
            f.write(canon)

    frames, hashes = {}, {}
    small = [(160, 90), (96, 64)]  # 16:9 and 3:2; the window is resizable (main.c:155-157)
    for n in names:
        rs = ol.RefScene(path=os.path.join(HERE, "scenes", n + ".lol"))
        for (w, h) in small:
            p = rs.probe(w, h)
            px, _ = rs.render_protocol(w, h)
            assert np.array_equal(px, p["rgba"]), "probe and render_thread disagree"
            frames[f"{n}_{w}x{h}_dist"] = p["dist"]
            frames[f"{n}_{w}x{h}_id"] = p["id"].astype(np.uint8)
            frames[f"{n}_{w}x{h}_rgba"] = p["rgba"]
        for (w, h) in [(320, 240), (1920, 1080), (3840, 2160)]:
            px, ms = rs.render_protocol(w, h)
            hashes[f"{n}_{w}x{h}"] = ol.frame_hash(px)
            print(n, w, h, hashes[f"{n}_{w}x{h}"], f"{ms[0]:.0f} ms", flush=True)

    # config C5: orbit cameras around scene4 (frames 0, 16, 32, 48 of 64)
    scene4 = lb.Scene.from_file(os.path.join(HERE, "scenes", "scene4.lol"))
    rs = ol.RefScene(path=os.path.join(HERE, "scenes", "scene4.lol"))
    for k in (0, 16, 32, 48):
        cam = scenegen.orbit_camera(scene4.camera, k, 64)
        rs.set_camera(list(cam.point), list(cam.direction))
        p = rs.probe(160, 90)
        frames[f"orbit{k}_160x90_dist"] = p["dist"]
        frames[f"orbit{k}_160x90_id"] = p["id"].astype(np.uint8)
        frames[f"orbit{k}_160x90_rgba"] = p["rgba"]
        px, _ = rs.render_protocol(640, 360)
        hashes[f"orbit{k}_640x360"] = ol.frame_hash(px)

    # config C4: the synthetic 1024-primitive scene, small frame (seconds on CPU)
    text = scenegen.synthetic_scene_text()
    rs = ol.RefScene(text=text)
    p = rs.probe(96, 54)
    frames["synthetic_96x54_dist"] = p["dist"]
    frames["synthetic_96x54_id"] = p["id"].astype(np.uint16)
    frames["synthetic_96x54_rgba"] = p["rgba"]
    hashes["synthetic_96x54"] = ol.frame_hash(p["rgba"])
    import hashlib
    hashes["synthetic_text_sha256"] = hashlib.sha256(text.encode()).hexdigest()

    np.savez_compressed(os.path.join(HERE, "frames.npz"), **frames)
    with open(os.path.join(HERE, "hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1, sort_keys=True)
    print("wrote", len(frames), "arrays and", len(hashes), "hashes")


if __name__ == "__main__":
    main()
