set -x
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv
./scripts/ubench/fp64_pipe > gpurun_out/ubench_fp64.txt 2>&1
python scripts/time_fill.py > gpurun_out/exp1.txt 2>&1
for v in notb k8b3 k8b3notb; do WSTR_LIB=$PWD/warpstr_b200/libwarpstr_b200.$v.so python scripts/time_fill.py --tag $v >> gpurun_out/exp1.txt 2>&1; done
cat gpurun_out/exp1.txt
ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -o gpurun_out/r01b_fill_full python scripts/time_fill.py --reads 30000 --reps 1 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
