#!/bin/bash
# A/B of two builds of the library on ONE box (new = the tree's build, old = $AB_OLD, a library built from another revision;
# *.so files are git-ignored but travel with gpurun), alternating.  Usage: [AB_OLD=path] [AB_RUNS="new old"] tools/ab/run_ab.sh <tag>
tag=${1:-ab}
L=focused-attention-vit_b200/libfavit_b200.so
cp $L /tmp/new.so
for v in ${AB_RUNS:-new old new2 old2}; do
  case $v in new*) cp /tmp/new.so $L;; old*) cp ${AB_OLD:-tools/ab/libfavit_b200_old.so} $L;; esac
  timeout 120 python bench.py --steps 20 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - <<P
import json
for l in open("gpurun_out/${tag}_$v.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline_mhla"]["frac"], d["clocks"]["sm_mhz"])
P
done
cp /tmp/new.so $L
