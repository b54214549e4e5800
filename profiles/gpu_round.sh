#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu --set full of the three fused kernels.
# usage (here): gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh TAG'
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1000 -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --worlds 65536 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline \
    > gpurun_out/${TAG}_launches.log 2>&1
for k in gs prepare finish; do
  ncu --set full --clock-control none --import-source on -k regex:k_fused_${k} --launch-skip 240 --launch-count 1 \
      -o gpurun_out/${TAG}_${k} -f python bench.py --worlds 65536 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline \
      > gpurun_out/${TAG}_ncu_${k}.log 2>&1
done
ls -la gpurun_out
