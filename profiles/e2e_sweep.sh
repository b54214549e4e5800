#!/bin/bash
# End-to-end column-block sweep on one GPU: bash profiles/e2e_sweep.sh TAG "name|chunks|streams" ...
TAG=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r name chunks streams <<< "$spec"
  python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-parity-sample --e2e-chunks $chunks --e2e-compute-streams $streams \
      > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err || tail -3 gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}: chunks ${chunks} streams ${streams}: e2e %.4g  device %.4g  ratio %.3f" % (d["e2e"]["value"], d["value"], d["e2e"]["value"]/d["value"]))
except Exception as e:
    print("${name}: failed", e)
PY
done
