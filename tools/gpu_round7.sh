#!/bin/bash
# ncu --set full of lol_render on the three small example scenes (end-of-round kernel)
tag=${1:-r01o}
out=gpurun_out/$tag
mkdir -p $out
for scene in scene scene2 scene3; do
  python tools/profile_one.py $scene 3840 2160 0 5 > $out/plain_$scene.log 2>&1 && \
  timeout 200 ncu --set full --clock-control none -k regex:lol_render -s 2 -c 1 -f \
      -o $out/prof_$scene python tools/profile_one.py $scene 3840 2160 0 5 > $out/ncu_$scene.log 2>&1
  echo "ncu $scene rc=$?"; cat $out/plain_$scene.log
done
