# end-to-end block layouts with the final kernels and prioritised compute streams (sync | queued)
for cfg in "auto" "1,1,6,1,1" "1,2,4,2,1" "1,1,3,3,1,1" "1,3,3,1" "1,1,1,1,2,2,1,1" "2,3,3,2"; do
  n=$(echo $cfg | tr ',' '_')
  python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-parity-sample --e2e-chunks $cfg > gpurun_out/r05k_$n.json 2> gpurun_out/r05k_$n.err
  python -c "
import json
d=json.load(open('gpurun_out/r05k_$n.json')); e=d['e2e']; print('$cfg', d['value'], e['value'], e.get('queued',{}).get('value'))"
done
