set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
python scripts/bench_aux.py > gpurun_out/bench_aux.log 2>&1; cut -c1-220 gpurun_out/bench_aux.log
bash scripts/profile_round.sh r01i
