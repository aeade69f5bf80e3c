#!/bin/bash
# GPU tests + the packed-pairs A/B:  gpurun --timeout 900 -- 'bash tools/gpu_round3.sh r01j'
tag=${1:-r01j}
out=gpurun_out/$tag
mkdir -p $out
timeout 700 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 300 python tools/ab_test.py 3840x2160 scene,scene2,scene3,scene4 "" "pack_pairs=2" "pack_pairs=3" 2>&1 | tee $out/ab_pack_pairs.txt
timeout 300 python tools/ab_test.py 3840x2160 synthetic,synthetic_csg "pack_pairs=0" "pack_pairs=1" "pack_pairs=3" 2>&1 | tee -a $out/ab_pack_pairs.txt
