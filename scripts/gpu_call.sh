set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
python scripts/bench_aux.py > gpurun_out/bench_aux_nb5.log 2>&1; grep normalize gpurun_out/bench_aux_nb5.log | cut -c1-210
WSTR_LIB=$PWD/warpstr_b200/libwarpstr_b200.nb4.so python scripts/bench_aux.py > gpurun_out/bench_aux_nb4.log 2>&1; grep normalize gpurun_out/bench_aux_nb4.log | cut -c1-210
WSTR_LIB=$PWD/warpstr_b200/libwarpstr_b200.nb3.so python scripts/bench_aux.py > gpurun_out/bench_aux_nb3.log 2>&1; grep normalize gpurun_out/bench_aux_nb3.log | cut -c1-210
