for cfg in "11 4" "11 3" "12 4" "12 5" "10 4" "10 3" "11 5"; do
  set -- $cfg
  QVM_B200_TILE_BITS=$1 QVM_B200_LOW_BITS=$2 python bench.py --qubits 30 --depth 40 --no-e2e --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/ab2_$1_$2.json 2>&1
  echo "$1 $2 $(python3 -c "
import json
d=json.loads(open('gpurun_out/ab2_$1_$2.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value'],1),'frac',round(r['frac'],3),'launch ms',round(r['avg_launch_ms'],2),'launches',r['launches_per_step'])")"
done
