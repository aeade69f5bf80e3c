import os, sys, time, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import loltracer_b200 as lb
tf, ms = lb.measure_fp32_peak(0)
print('fp32 peak TFLOP/s', tf, ms)
for name in ['scene','scene2','scene3','scene4']:
    scene = lb.Scene.from_file(f'/root/repo/tests/golden/scenes/{name}.lol')
    for variant in [1]:
        r = lb.Renderer(scene, lb.Options.default(variant=variant, counters=0))
        print(name, r.kernel_info())
        w,h = 3840,2160
        frame = torch.zeros((h,w), dtype=torch.int32, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3): r.render_device(frame.data_ptr(), w, h, stream=st)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): r.render_device(frame.data_ptr(), w, h, stream=st)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1)/10
        print(f'{name} v{variant} 4K: {t:.3f} ms  {w*h/t/1e3:.1f} Mrays/s', flush=True)
        rc = lb.Renderer(scene, lb.Options.default(variant=variant, counters=1))
        rc.render_device(frame.data_ptr(), w, h, stream=st); torch.cuda.synchronize()
        print('   counters', rc.read_counters())
        host = np.zeros((h,w), np.uint32)
        for _ in range(2): r.render_host(host.ctypes.data, w, h)
        t0=time.perf_counter()
        for _ in range(5): r.render_host(host.ctypes.data, w, h)
        print(f'   host e2e {(time.perf_counter()-t0)/5*1e3:.3f} ms')
