set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/r01j_bench.json 2> gpurun_out/r01j_bench.err; cut -c1-300 gpurun_out/r01j_bench.json; tail -n 2 gpurun_out/r01j_bench.err
