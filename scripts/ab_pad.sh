for pad in 0 22000 41000 80000; do
  QVM_B200_SMEM_PAD=$pad python bench.py --qubits 30 --depth 40 --no-e2e --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/pad_$pad.json 2>&1
  echo "pad=$pad $(python3 -c "
import json
d=json.loads(open('gpurun_out/pad_$pad.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value'],1),'frac',round(r['frac'],3),'launch ms',round(r['avg_launch_ms'],2))")"
done
