#!/bin/bash
# Multi-GPU round 2 (gpurun --gpus 8): multi-GPU tests, then e2e paths at N = 2, 4, 8 and the C host.
tag=${1:-scale2}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_backend.py tests/test_gpu_sharding.py -x -q > $out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_multi.log
port=29600
run() { # N, name, extra args...
  n=$1; name=$2; shift 2
  port=$((port+1))
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > $out/$name.json 2> $out/$name.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n "$@" > $out/$name.json 2> $out/$name.err
  fi
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads(open('$out/$name.json').read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_frame'],4), d['e2e'].get('host_frame_equals_single_gpu'), d.get('sharded_frame_equals_single_gpu'))" 2>&1)"
}
for n in 2 4 8; do
  run $n scene4_n$n --steps 50 --warmup 5 --no-cpu-baseline
done
run 8 scene4_n8_gathercopy --steps 50 --warmup 5 --no-cpu-baseline --e2e-path gather-then-copy
run 8 synthetic_n8 --scene synthetic --steps 3 --warmup 3 --no-cpu-baseline
run 8 scene_n8 --scene scene --steps 50 --warmup 5 --no-cpu-baseline
H=loltracer_b200/backend/build/lol_headless_b200
for g in 2 4 8; do
  for m in host peer; do
    $H 4 tests/golden/scenes/scene4.lol --gpus $g --gather $m --size 3840x2160 --frames 20 --warmup 5 > $out/headless_g${g}_$m.log 2>&1
    echo "headless --gpus $g --gather $m rc=$? $(grep -h 'min \|hash' $out/headless_g${g}_$m.log | tr '\n' ' ')"
  done
done
