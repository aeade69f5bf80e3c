"""Kernel time of ONE rank's shard on one GPU (what a rank of an N-GPU job runs):
    python tools/shard_timing.py scene4 3840x2160 8        # rank 0 of 8
LOLB200_LPT=0 turns the longest-first chunk order off (A/B)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import loltracer_b200 as lb
from loltracer_b200 import scenegen
name, size, world = sys.argv[1], sys.argv[2], int(sys.argv[3])
w, h = (int(x) for x in size.split("x"))
scene = (lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg"))) if name.startswith("synthetic")
         else lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol")))
r = lb.Renderer(scene)
st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(lb.shard_pixels(w, h, world) if world > 1 else w * h, dtype=torch.int32, device="cuda")
for rank in (0, world - 1) if world > 1 else (0,):
    shard = lb.Shard(rank=rank, world=world) if world > 1 else None
    for _ in range(12):
        r.render_device(buf.data_ptr(), w, h, shard=shard, pitch_px=w, stream=st)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(32):
            r.render_device(buf.data_ptr(), w, h, shard=shard, pitch_px=w, stream=st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 32)
    print(f"{name} {size} rank {rank}/{world} LPT={os.environ.get('LOLB200_LPT', '1')}: {best:.4f} ms/frame", flush=True)
