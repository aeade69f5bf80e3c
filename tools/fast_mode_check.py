"""How far is the fast-arithmetic build from the exact one?  (GPU vs GPU; the exact build is
bit-identical to the oracle in distance/id.)  python tools/fast_mode_check.py [WxH]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import loltracer_b200 as lb
w, h = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "3840x2160").split("x"))
st = torch.cuda.current_stream().cuda_stream
def render(scene, **kw):
    r = lb.Renderer(scene, lb.Options.default(**kw))
    f = torch.zeros((h, w), dtype=torch.int32, device="cuda"); i = torch.zeros((h, w), dtype=torch.int32, device="cuda")
    r.render_device(f.data_ptr(), w, h, aux=lb.Aux(id=i.data_ptr()), stream=st)
    for _ in range(3): r.render_device(f.data_ptr(), w, h, stream=st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): r.render_device(f.data_ptr(), w, h, stream=st)
    e1.record(); torch.cuda.synchronize()
    return f.cpu().numpy().view(np.uint32), i.cpu().numpy().view(np.uint32), e0.elapsed_time(e1) / 10
for name in ("scene", "scene2", "scene3", "scene4"):
    scene = lb.Scene.from_file(os.path.join(ROOT, "tests", "golden", "scenes", name + ".lol"))
    ef, ei, et = render(scene, arith=0)
    ff, fi, ft = render(scene, arith=1)
    hit_e, hit_f = ei != 0, fi != 0
    agree = (hit_e == hit_f)
    both = hit_e & hit_f
    err = np.zeros(ef.shape, np.int32)
    for s in (16, 8, 0):
        err = np.maximum(err, np.abs(((ef >> s) & 255).astype(np.int32) - ((ff >> s) & 255).astype(np.int32)))
    hist = np.bincount(err[both].ravel(), minlength=4)
    print(f"{name}: exact {et:.3f} ms, fast {ft:.3f} ms ({et / ft:.2f}x); mask agreement {agree.mean() * 100:.5f} % "
          f"({(~agree).sum()} px); RGB err on agreeing hits: max {err[both].max()}, >1: {(err[both] > 1).sum()} px, "
          f"hist0..3 {hist[:4].tolist()}; id agreement {(ei == fi).mean() * 100:.5f} %", flush=True)
