"""Renders N frames of one scene: the short command ncu wraps (see profiles/README.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import loltracer_b200 as lb

name, w, h, variant, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
arith = int(sys.argv[6]) if len(sys.argv) > 6 else 0
if name.startswith("synthetic"):
    from loltracer_b200 import scenegen
    scene = lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg")))
else:
    scene = lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol"))
extra = dict(kv.split("=") for kv in os.environ.get("LOL_OPTS", "").split(",") if kv)
kw = dict(variant=variant, arith=arith)
kw.update({k: int(v) for k, v in extra.items()})  # LOL_OPTS may override the positional variant
r = lb.Renderer(scene, lb.Options.default(**kw))
frame = torch.zeros((h, w), dtype=torch.int32, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st = torch.cuda.current_stream().cuda_stream
r.render_device(frame.data_ptr(), w, h, stream=st)
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    r.render_device(frame.data_ptr(), w, h, stream=st)
e1.record()
torch.cuda.synchronize()
print(f"{name} {w}x{h} variant {variant}: {e0.elapsed_time(e1) / n:.3f} ms/frame, {r.kernel_info()}")
