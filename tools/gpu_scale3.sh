#!/bin/bash
# Two-GPU check of the end-of-round kernel (gpurun --gpus 2): multi-GPU tests, N = 2 bench (peer gather) and the C host.
tag=${1:-r01n2}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_backend.py tests/test_gpu_sharding.py -q > $out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > $out/scene4_n2.json 2> $out/scene4_n2.err; echo "bench n2 rc=$?"
python -c "import json; d=json.loads(open('$out/scene4_n2.json').read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_frame'],4), d['e2e'].get('host_frame_equals_single_gpu'), d.get('sharded_frame_equals_single_gpu'))"
H=loltracer_b200/backend/build/lol_headless_b200
for m in host peer; do
  $H 4 tests/golden/scenes/scene4.lol --gpus 2 --gather $m --size 3840x2160 --frames 20 --warmup 5 > $out/headless_g2_$m.log 2>&1
  echo "headless --gpus 2 --gather $m rc=$? $(grep -h 'min \|hash' $out/headless_g2_$m.log | tr '\n' ' ')"
done
