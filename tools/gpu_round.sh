#!/bin/bash
# One gpurun call: GPU tests, the bench (both arms), the ncu launch list of the bench
# command and one --set full capture of lol_render with its generated source imported.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01f'
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out/src
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 600 python bench.py --all-scenes > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; echo "ref rc=$?"
timeout 600 python bench.py --scene synthetic --steps 5 --warmup 3 > $out/bench_synthetic.json 2>> $out/bench.err; echo "synthetic rc=$?"
timeout 600 python bench.py --scene synthetic_csg --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_synthetic_csg.json 2>> $out/bench.err; echo "csg rc=$?"
cat $out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
export LOLB200_DUMP_DIR=$PWD/$out/src
for scene in scene4 synthetic; do
  python tools/profile_one.py $scene 3840 2160 0 5 > $out/plain_$scene.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:lol_render -s 2 -c 1 -f \
      -o $out/prof_$scene python tools/profile_one.py $scene 3840 2160 0 5 > $out/ncu_$scene.log 2>&1
  echo "ncu $scene rc=$?"; cat $out/plain_$scene.log
done
