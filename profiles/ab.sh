#!/bin/bash
# A/B of kernel variants on one box.  usage: bash profiles/ab.sh TAG "name|ENV|bench args" ...
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r name envs args <<< "$spec"
  echo "== $name"
  env $envs python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $args > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err || tail -5 gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}: %.4g w-s/s  stage_ms %s" % (d["value"], {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}))
except Exception as e:
    print("${name}: failed", e)
PY
done
