set -x
mkdir -p gpurun_out
python scripts/time_fill.py > gpurun_out/tf_w12.log 2>&1; tail -n 1 gpurun_out/tf_w12.log
WSTR_LIB=$PWD/warpstr_b200/libwarpstr_b200.w16.so python scripts/time_fill.py > gpurun_out/tf_w16.log 2>&1; tail -n 1 gpurun_out/tf_w16.log
WSTR_LIB=$PWD/warpstr_b200/libwarpstr_b200.w16.so python bench.py --no-cpu-baseline > gpurun_out/bench_w16.json 2> gpurun_out/bench_w16.err; cut -c1-200 gpurun_out/bench_w16.json; tail -n 2 gpurun_out/bench_w16.err
WSTR_LIB=$PWD/warpstr_b200/libwarpstr_b200.w16.so timeout 300 python -m pytest tests/test_gpu_dtw.py tests/test_gpu_scale.py -x -q > gpurun_out/pytest_w16.log 2>&1; tail -n 2 gpurun_out/pytest_w16.log
