#!/bin/bash
# GPU tests + bench lines (default, synthetic, synthetic_csg) + the fast-arithmetic tolerance check.
tag=${1:-r01d}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
timeout 600 python bench.py --all-scenes > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"; cat $out/bench.json
timeout 600 python bench.py --scene synthetic --steps 5 --warmup 3 > $out/bench_synthetic.json 2>> $out/bench.err; echo "synthetic rc=$?"; cat $out/bench_synthetic.json
timeout 600 python bench.py --scene synthetic_csg --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_synthetic_csg.json 2>> $out/bench.err; echo "csg rc=$?"; cat $out/bench_synthetic_csg.json
timeout 600 python tools/fast_mode_check.py 3840x2160 > $out/fast_mode_check.txt 2>&1; cat $out/fast_mode_check.txt
timeout 300 python tools/ab_test.py 3840x2160 synthetic,synthetic_csg "variant=1" "variant=3" > $out/ab_synth_4k.txt 2>&1; cat $out/ab_synth_4k.txt
