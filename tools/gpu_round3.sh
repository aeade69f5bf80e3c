#!/bin/bash
# GPU tests + the packed-pairs A/B:  gpurun --timeout 900 -- 'bash tools/gpu_round3.sh r01i'
tag=${1:-r01i}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 300 python tools/ab_test.py 3840x2160 scene3,scene4,synthetic,synthetic_csg "pack_pairs=0" "pack_pairs=1" "pack_pairs=1,min_blocks=3" 2>&1 | tee $out/ab_pack_pairs.txt
