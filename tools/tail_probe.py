"""Kernel time and TAIL of one rank's shard on one GPU (what a rank of an N-GPU job runs), per option set:
    python tools/tail_probe.py scene4 3840x2160 8 "" "variant=4" "variant=4,defer_cap_primary=24,defer_cap_shadow=12"
tail = first moment a warp finds the work queue dry -> last warp's exit (the kernel's own global-timer probes);
span = first CTA's start -> last warp's exit.  LOLB200_LPT=0 turns the longest-first chunk order off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import loltracer_b200 as lb
from loltracer_b200 import scenegen
name, size, world = sys.argv[1], sys.argv[2], int(sys.argv[3])
configs = sys.argv[4:] or [""]
w, h = (int(x) for x in size.split("x"))
scene = (lb.Scene.from_string(scenegen.synthetic_scene_text(csg=name.endswith("csg"))) if name.startswith("synthetic")
         else lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol")))
st = torch.cuda.current_stream().cuda_stream
buf = torch.zeros(lb.shard_pixels(w, h, world) if world > 1 else w * h, dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for cfg in configs:
    kw = {k: int(v) for k, v in (kv.split("=") for kv in cfg.split(",") if kv)}
    r = lb.Renderer(scene, lb.Options.default(**kw))
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
        # flushed L2 + probes
        n = 8
        init = torch.tensor([-1, 0, -1], dtype=torch.int64, device="cuda")
        probes = torch.zeros((n, 3), dtype=torch.int64, device="cuda")
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for i in range(n):
            probes[i].copy_(init)
            flush.fill_(i)
            ev[i][0].record()
            r.render_device(buf.data_ptr(), w, h, shard=shard, pitch_px=w, stream=st,
                            aux=lb.Aux(launch_timing=probes[i].data_ptr()))
            ev[i][1].record()
        torch.cuda.synchronize()
        p = probes.cpu().numpy().view(np.uint64).astype(np.float64)
        cold = sum(a.elapsed_time(b) for a, b in ev) / n
        print(f"{name} {size} rank {rank}/{world} [{cfg or 'default':45s}] back-to-back {best:.4f} ms, flushed {cold:.4f} ms, "
              f"tail {np.mean(p[:, 1] - p[:, 0]) / 1e3:6.1f} us, span {np.mean(p[:, 1] - p[:, 2]) / 1e3:7.1f} us, "
              f"regs {r.kernel_info()['regs']}", flush=True)
    r.close()
