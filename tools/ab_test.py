"""A/B timing of kernel options on one GPU:  python tools/ab_test.py [WxH] scene[,scene..] "k=v,k=v" "k=v" ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import loltracer_b200 as lb
from loltracer_b200 import scenegen

size = sys.argv[1]
w, h = (int(x) for x in size.split("x"))
names = sys.argv[2].split(",")
configs = sys.argv[3:] or [""]
frame = torch.zeros((h, w), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for name in names:
    scene = (lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg"))) if name.startswith("synthetic") else
             lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol")))
    for cfg in configs:
        kw = dict(kv.split("=") for kv in cfg.split(",") if kv)
        r = lb.Renderer(scene, lb.Options.default(**{k: int(v) for k, v in kw.items()}))
        n = 3 if name.startswith("synthetic") else 20
        for _ in range(2):
            r.render_device(frame.data_ptr(), w, h, stream=st)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                r.render_device(frame.data_ptr(), w, h, stream=st)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / n)
        info = r.kernel_info()
        print(f"{name:10s} {size} [{cfg or 'default':40s}] {best:9.3f} ms  {w * h / best / 1e3:9.1f} Mrays/s  "
              f"regs {info['regs']} local {info['local_bytes']}", flush=True)
        r.close()
