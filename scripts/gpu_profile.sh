# Round bench line + ncu launch list + one full capture of the dominant kernel (run under gpurun, 1 GPU)
tag=${1:-r1e}
python bench.py > gpurun_out/bench_${tag}_q33.json 2> gpurun_out/bench_${tag}_q33.err; tail -c 3000 gpurun_out/bench_${tag}_q33.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}_q30.csv \
    python bench.py --qubits 30 --depth 40 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spec_kernel --launch-skip 40 -c 2 -f -o gpurun_out/prof_${tag}_spec \
    python bench.py --qubits 28 --depth 40 --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
