"""Debug helper: render one fuzz scene on the GPU and list pixels whose colour differs from the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import loltracer_b200 as lb, oracle_lib as ol
from test_lowering_fuzz import random_scene
from test_gpu_parity import _render
seed, ext = int(sys.argv[1]), bool(int(sys.argv[2]))
scene = lb.Scene.from_string(random_scene(seed + (100 if ext else 0), ext, fixed_head=False))
w, h = 200, 112
want = ol.port_render(scene, w, h)
for kw in [dict(variant=1, prune_bounds=0, guarded_fastpath=0), dict(variant=1, prune_bounds=0, guarded_fastpath=0, skip_black_miss=0, cull_backfacing=0, shadow_early_out=0)]:
    got = _render(lb, scene, w, h, options=lb.Options.default(**kw))
    bad = np.nonzero(got["rgba"] != want["rgba"])
    print(kw, "differing pixels", len(bad[0]), "dist equal", np.array_equal(got["dist"].view(np.uint32), want["dist"].view(np.uint32)),
          "nshadow equal", None)
    for y, x in list(zip(*bad))[:8]:
        print("  px", x, y, "gpu %08x" % got["rgba"][y, x], "cpu %08x" % want["rgba"][y, x], "id", got["id"][y, x], want["id"][y, x], "dist", got["dist"][y, x])
