timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --qubits 30 --depth 40 --no-e2e --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/quick_q30.json 2>gpurun_out/quick_q30.err
python3 - <<'P'
import json
d=json.loads(open('gpurun_out/quick_q30.json').read().strip().splitlines()[-1]); r=d['roofline']
print('RESULT value',round(d['value'],1),'frac',round(r['frac'],3),'launch ms',round(r['avg_launch_ms'],2),'launches',r['launches_per_step'],'clk',d['clocks']['sm_mhz'])
P
ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum --clock-control none -k regex:spec_kernel --launch-skip 30 -c 24 --csv --log-file gpurun_out/quick_conflicts.csv python bench.py --qubits 28 --depth 40 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
python3 - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/quick_conflicts.csv')))
i=[k for k,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[i]; d=collections.defaultdict(dict)
for r in rows[i+1:]:
    if len(r)==len(h): d[r[0]][r[h.index('Metric Name')]]=float(r[h.index('Metric Value')].replace(',',''))
tot=collections.Counter()
for k,v in d.items():
    for a,b in v.items(): tot[a]+=b
print({a:round(b/len(d),1) for a,b in tot.items()}, len(d))
P
