# A/B of the specialised kernels at 30 qubits: relabelling on/off, resident CTAs, tile size, streaming loads
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r1f.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --qubits ${Q:-30} --depth 40 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ab5_$name.json 2>gpurun_out/ab5_$name.err
  python3 - "$name" <<'P'
import json,sys
name=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/ab5_%s.json'%name).read().strip().splitlines()[-1]); r=d['roofline']
    print('RESULT',name,'value',round(d['value'],1),'frac',round(r['frac'],3),'launch ms',round(r['avg_launch_ms'],2),'launches',r['launches_per_step'],'spec',r['specialised_launch_share'],'clk',d['clocks']['sm_mhz'],d['clocks']['reasons'])
except Exception as e:
    print('RESULT',name,'failed',e)
P
}
run norelabel_t11 QVM_B200_RELABEL=0
run relabel_t11_occ4 QVM_B200_SPEC_OCC=4
run relabel_t11_occ5 QVM_B200_SPEC_OCC=5
run relabel_t11_occ6 QVM_B200_SPEC_OCC=6
run relabel_t12_occ4 QVM_B200_SPEC_OCC=4 QVM_B200_TILE_BITS=12
run relabel_t12_occ6 QVM_B200_SPEC_OCC=6 QVM_B200_TILE_BITS=12
run relabel_t11_occ4_stream QVM_B200_SPEC_OCC=4 QVM_B200_SPEC_STREAM=1
run relabel_t12_occ6_stream QVM_B200_SPEC_OCC=6 QVM_B200_TILE_BITS=12 QVM_B200_SPEC_STREAM=1
run relabel_t10_occ4 QVM_B200_SPEC_OCC=4 QVM_B200_TILE_BITS=10
