#!/bin/bash
# usage: sweep.sh tag "ENV=.. ENV=.." ...
cd /root/repo
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err
  echo "cfg $i: $cfg"
  python profiles/tools/bench_summary.py gpurun_out/sweep_$i.json
  tail -2 gpurun_out/sweep_$i.err
done
