# tile-size sweep at 30 qubits, specialised kernels on (steady state): QVM_B200_TILE_BITS x QVM_B200_LOW_BITS
for cfg in ${SWEEP:-"11 4" "12 4" "12 5" "13 4" "13 5" "10 4"}; do
  set -- $cfg
  QVM_B200_TILE_BITS=$1 QVM_B200_LOW_BITS=$2 timeout 300 python bench.py --qubits ${Q:-30} --depth 40 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ab3_$1_$2.json 2>gpurun_out/ab3_$1_$2.err
  echo "$1 $2 $(python3 -c "
import json
d=json.loads(open('gpurun_out/ab3_$1_$2.json').read().strip().splitlines()[-1]); r=d['roofline']
print('value',round(d['value'],1),'frac',round(r['frac'],3),'launch ms',round(r['avg_launch_ms'],2),'launches',r['launches_per_step'],'spec share',r['specialised_launch_share'],'clk',d['clocks']['sm_mhz'])" 2>&1 | tail -1)"
done
