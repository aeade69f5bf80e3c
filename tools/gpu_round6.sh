#!/bin/bash
# End-of-round check: GPU tests, smoke, the bench (both arms), C4 and C5 workloads
tag=${1:-r01m}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
timeout 600 python bench.py --all-scenes > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; echo "ref rc=$?"
timeout 600 python bench.py --scene synthetic --steps 5 --warmup 3 > $out/bench_synthetic.json 2>> $out/bench.err; echo "synthetic rc=$?"
timeout 600 python bench.py --scene synthetic_csg --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_synthetic_csg.json 2>> $out/bench.err; echo "csg rc=$?"
cat $out/bench.json
tail -3 $out/bench.err
timeout 300 python bench.py --workload orbit --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_orbit.json 2>> $out/bench.err; echo "orbit rc=$?"
