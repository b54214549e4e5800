for cfg in "auto" "8" "4" "1,2,2,2,1" "1,1,2,2,2,2,1,1" "2,2,2,2,1,1" "1,1,1,2,2,2,1,1,1"; do
  n=$(echo $cfg | tr ',' '_')
  python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-parity-sample --e2e-chunks $cfg > gpurun_out/r04k_$n.json 2> gpurun_out/r04k_$n.err
  python -c "
import json
d=json.load(open('gpurun_out/r04k_$n.json')); e=d['e2e']; print('$cfg', d['value'], e['value'], e.get('queued',{}).get('value'))"
done
