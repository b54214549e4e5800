#!/bin/bash
# One GPU session: parity tests, bench line, ncu launch list, ncu --set full of the three fused kernels
# AT THE BENCH SIZE (262144 worlds, the staggered-episode mix the bench times).
# usage (here): gpurun --timeout 2400 -- 'bash profiles/gpu_round.sh TAG [notests]'
TAG=${1:-r02}
W=${W:-262144}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
  python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
  tail -5 gpurun_out/${TAG}_pytest.log
fi
python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
cat gpurun_out/${TAG}_bench.json
BARGS="--worlds $W --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-sample"
# launch list: every kernel of a few timed steps (the first 1500 launches are the priming of the episodes)
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py $BARGS > gpurun_out/${TAG}_launches.log 2>&1
FP64=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,sm__inst_executed_pipe_fp64.sum
for k in gs prepare finish; do
  ncu --set full --metrics $FP64 --clock-control none --import-source on -k regex:k_fused_${k} --launch-skip 240 --launch-count 1 \
      -o gpurun_out/${TAG}_${k} -f python bench.py $BARGS > gpurun_out/${TAG}_ncu_${k}.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu_${k}.log
done
ls -la gpurun_out | tail -12
