#!/bin/bash
# ncu evidence of the current kernels: launch list of the bench command, full captures of lol_render
tag=${1:-r01l}
out=gpurun_out/$tag
mkdir -p $out/src
export LOLB200_DUMP_DIR=$PWD/$out/src
cap() { # name scene opts
  LOL_OPTS=$3 python tools/profile_one.py $2 3840 2160 0 5 > $out/plain_$1.log 2>&1 && \
  LOL_OPTS=$3 timeout 300 ncu --set full --clock-control none --import-source on -k regex:lol_render -s 2 -c 1 -f \
      -o $out/prof_$1 python tools/profile_one.py $2 3840 2160 0 5 > $out/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?"; cat $out/plain_$1.log
}
cap scene4 scene4 ""
cap scene4_packed scene4 "pack_pairs=2"
cap synthetic synthetic ""
unset LOLB200_DUMP_DIR
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_plain.json 2> $out/bench.err; echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la $out
