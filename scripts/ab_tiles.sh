for cfg in "12 5" "11 5" "10 5" "9 5" "12 4" "11 4" "10 4" "10 3" "11 3"; do
  set -- $cfg
  QVM_B200_TILE_BITS=$1 QVM_B200_LOW_BITS=$2 python bench.py --qubits 30 --depth 40 --no-e2e --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/ab_$1_$2.json 2>&1
done
cat gpurun_out/ab_*.json
