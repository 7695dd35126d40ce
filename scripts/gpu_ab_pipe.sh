# A/B of the pipelined (persistent, cp.async-prefetching) specialised kernels at 30 qubits
QVM_B200_SPEC_PIPE=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "specialised or relabel or engine_with" 2>&1 | tail -3
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --qubits ${Q:-30} --depth 40 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ab6_$name.json 2>gpurun_out/ab6_$name.err
  python3 - "$name" <<'P'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ab6_%s.json'%name).read().strip().splitlines()[-1]); r=d['roofline']
    print('RESULT',name,'value',round(d['value'],1),'frac',round(r['frac'],3),'launch ms',round(r['avg_launch_ms'],2),'launches',r['launches_per_step'],'spec',r['specialised_launch_share'],'clk',d['clocks']['sm_mhz'])
except Exception as e:
    print('RESULT',name,'failed',e); print(open('gpurun_out/ab6_%s.err'%name).read()[-600:])
P
}
run base_t11 QVM_B200_SPEC_PIPE=0
run pipe_t11 QVM_B200_SPEC_PIPE=1
run pipe_t10 QVM_B200_SPEC_PIPE=1 QVM_B200_TILE_BITS=10
run pipe_t12 QVM_B200_SPEC_PIPE=1 QVM_B200_TILE_BITS=12
