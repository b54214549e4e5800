V=$PWD/arboris-python_b200/arboris_b200/lib/variants
for rep in 1 2; do
for n in g32 g64 g128 auto; do
  if [ $n = auto ]; then L=""; else L="ARB_B200_LIB=$V/libarboris_b200_$n.so"; fi
  env $L python bench.py --steps 100 --warmup 8 --no-cpu-baseline --no-parity-sample > gpurun_out/r02t_${n}_$rep.json 2> gpurun_out/r02t_${n}_$rep.err
  python -c "
import json; d=json.load(open('gpurun_out/r02t_${n}_$rep.json')); print('$n rep $rep: device %.4g e2e %.4g gs %.3f' % (d['value'], d['e2e']['value'], d['roofline']['stage_ms']['gs']))"
done; done
