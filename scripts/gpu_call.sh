set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
bash scripts/profile_round.sh r01h
for spec in "FMR1 100000" "FMR1_MGG 100000" "DM2 100000" "CAN 50000" "C9ORF72_1000 2000"; do
  set -- $spec
  python bench.py --no-cpu-baseline --steps 3 --locus $1 --reads $2 2> gpurun_out/bench_$1.err | tail -1 > gpurun_out/bench_$1.json
  cut -c1-160 gpurun_out/bench_$1.json
done
python scripts/bench_aux.py > gpurun_out/bench_aux.log 2>&1
