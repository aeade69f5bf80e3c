#!/bin/bash
# the new GPU tests + A/B of the shared first step and the division pre-test
tag=${1:-r01k}
out=gpurun_out/$tag
mkdir -p $out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "shared_first or cameras_outside or options_do_not or counters_match" > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 300 python tools/ab_test.py 3840x2160 scene,scene2,scene3,scene4 "share_first_step=0,shadow_div_pretest=0" "share_first_step=1,shadow_div_pretest=0" "share_first_step=0,shadow_div_pretest=1" "" 2>&1 | tee $out/ab_first_step_div_pretest.txt
timeout 300 python tools/ab_test.py 3840x2160 synthetic "share_first_step=0,shadow_div_pretest=0" "" 2>&1 | tee -a $out/ab_first_step_div_pretest.txt
