#!/bin/bash
# One parametrised script for every gpurun call (replaces the per-call gpu_round*.sh / gpu_scale*.sh).
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh <tag> <stage> [<stage> ...]'
#
# Output goes to gpurun_out/<tag>/.  Stages (run in the order given):
#   info              GPU name, clocks, topology
#   tests[:<-k expr>] pytest -m gpu (optionally -k <expr>)
#   smoke             __graft_entry__.smoke()
#   bench             bench.py (default line, N=1)
#   bench_ref         bench.py --impl reference (short)
#   bench:<name>:<args with + for spaces>   any other bench.py line, e.g. bench:c4:--scene+synthetic+--steps+5
#   launches          ncu launch list (gpu__time_duration) of the default bench command
#   ncu:<name>:<scene>[:<LOL_OPTS>]         ncu --set full of lol_render (4K, generated source imported)
#   ab:<name>:<scenes>:<optsA>|<optsB>|...   tools/ab_test.py at 4K
#   scale[:<Ns>]      bench.py at N = 1 2 4 8 (or the list given, e.g. scale:2,8), launched as the driver does
#   multi_tests       the multi-GPU tests (tests/test_gpu_backend.py, tests/test_gpu_sharding.py)
#   headless[:<Ns>]   the renderer.h drop-in (headless main.c twin) with --gpus N, host and peer gathers
#   py:<name>:<script+args>                  any python script under tools/
#   sh:<name>:<command with + for spaces>    any shell command (environment variables in front of a tool)
tag=${1:-session}
shift
out=gpurun_out/$tag
mkdir -p $out/src
port=29700

line() { # last JSON line of a bench output, abridged
  python - "$1" <<'EOF'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e = d.get("e2e", {})
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(e.get("value", 0), 1),
          "frac", round(d.get("roofline", {}).get("frac", 0), 4), "equal", d.get("sharded_frame_equals_single_gpu"),
          e.get("host_frame_equals_single_gpu"))
except Exception as ex:
    print("no JSON line:", ex)
EOF
}

bench_n() { # N name args...
  n=$1; name=$2; shift 2
  port=$((port+1))
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 "$@" > $out/$name.json 2> $out/$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port $port bench.py --gpus $n "$@" > $out/$name.json 2> $out/$name.err
  fi
  echo "$name rc=$? $(line $out/$name.json)"
}

for stage in "$@"; do
  IFS=':' read -r kind a b c <<< "$stage"
  case $kind in
    info)
      nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt
      nvidia-smi -L >> $out/gpu.txt; nvidia-smi topo -m >> $out/gpu.txt 2>&1 ;;
    tests)
      if [ -n "$a" ]; then timeout 1500 python -m pytest tests -m gpu -q -k "$a" > $out/pytest_gpu.log 2>&1
      else timeout 1500 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; fi
      echo "pytest rc=$?" | tee -a $out/pytest_gpu.log; tail -5 $out/pytest_gpu.log ;;
    multi_tests)
      timeout 600 python -m pytest tests/test_gpu_backend.py tests/test_gpu_sharding.py -q > $out/pytest_multi.log 2>&1
      echo "multi pytest rc=$?"; tail -3 $out/pytest_multi.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log ;;
    bench)
      if [ -n "$a" ]; then bench_n 1 bench_$a ${b//+/ }; else bench_n 1 bench; cat $out/bench.json; fi ;;
    bench_ref)
      timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
      echo "reference arm rc=$?"; cat $out/bench_reference.json ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
          --log-file $out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1
      echo "ncu launches rc=$?" ;;
    ncu)
      LOLB200_DUMP_DIR=$PWD/$out/src LOL_OPTS=$c python tools/profile_one.py $b 3840 2160 0 5 > $out/plain_$a.log 2>&1 && \
      LOLB200_DUMP_DIR=$PWD/$out/src LOL_OPTS=$c timeout 400 ncu --set full --clock-control none --import-source on \
          -k regex:lol_re -s 2 -c 2 -f -o $out/prof_$a python tools/profile_one.py $b 3840 2160 0 5 > $out/ncu_$a.log 2>&1
      echo "ncu $a rc=$?"; cat $out/plain_$a.log ;;
    ab)
      IFS='|' read -ra sets <<< "$c"
      timeout 600 python tools/ab_test.py 3840x2160 $b "${sets[@]}" 2>&1 | tee $out/ab_$a.txt ;;
    scale)
      for n in ${a:-1,2,4,8}; do :; done
      for n in $(echo ${a:-1,2,4,8} | tr ',' ' '); do bench_n $n scene4_n$n --steps 50 --warmup 5 --no-cpu-baseline; done ;;
    headless)
      H=loltracer_b200/backend/build/lol_headless_b200
      for g in $(echo ${a:-1,2,4,8} | tr ',' ' '); do
        for m in host peer nccl; do
          [ $g -eq 1 ] && [ $m != host ] && continue
          timeout 120 $H 8 tests/golden/scenes/scene4.lol --gpus $g --gather $m --size 3840x2160 --frames 30 --warmup 5 \
              > $out/headless_g${g}_$m.log 2>&1
          echo "headless --gpus $g --gather $m rc=$? $(grep -h 'min \|hash' $out/headless_g${g}_$m.log | tr '\n' ' ')"
        done
      done ;;
    py)
      timeout 900 python tools/${b//+/ } > $out/$a.txt 2>&1; echo "$a rc=$?"; tail -40 $out/$a.txt ;;
    sh) # sh:<name>:<shell command with + for spaces> (e.g. an environment variable in front of a tool)
      timeout 900 bash -c "${b//+/ }" > $out/$a.txt 2>&1; echo "$a rc=$?"; tail -40 $out/$a.txt ;;
    *) echo "unknown stage $stage" ;;
  esac
done
ls $out
