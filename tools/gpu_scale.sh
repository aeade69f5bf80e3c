#!/bin/bash
# Multi-GPU round (gpurun --gpus 8): the GPU tests that need >1 device, then the bench at
# N = 1, 2, 4, 8 launched as the driver launches it.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale.sh r01_scale'
tag=${1:-scale}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/gpus.txt
nvidia-smi topo -m >> $out/gpus.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q  > $out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_multi.log
port=29500
run() { # N, name, extra args...
  n=$1; name=$2; shift 2
  port=$((port+1))
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 "$@" > $out/$name.json 2> $out/$name.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n "$@" > $out/$name.json 2> $out/$name.err
  fi
  echo "$name rc=$? $(python -c "import json,sys; d=json.loads(open('$out/$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('sharded_frame_equals_single_gpu'))" 2>&1)"
}
for n in 1 2 4 8; do
  run $n scene4_peer_n$n --steps 50 --warmup 5 --no-cpu-baseline
done
for n in 2 8; do
  run $n scene4_nccl_n$n --steps 50 --warmup 5 --no-cpu-baseline --gather nccl
done
for n in 1 8; do
  run $n synthetic_n$n --scene synthetic --steps 3 --warmup 3 --no-cpu-baseline
done
run 8 orbit_n8 --workload orbit --steps 2 --warmup 3 --no-cpu-baseline
run 8 scene_n8 --scene scene --steps 50 --warmup 5 --no-cpu-baseline
# the single-process C host (renderer.h backend) driving all GPUs
H=loltracer_b200/backend/build/lol_headless_b200
if [ -x $H ]; then
  for g in 1 8; do
    $H 4 tests/golden/scenes/scene4.lol --gpus $g --size 3840x2160 --frames 20 --warmup 5 > $out/headless_g$g.log 2>&1
    echo "headless --gpus $g rc=$?"; tail -2 $out/headless_g$g.log
  done
fi
