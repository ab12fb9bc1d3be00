ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -o gpurun_out/r01c_fill_full python scripts/time_fill.py --reads 30000 --reps 1 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
WSTR_LIB=$PWD/warpstr_b200/libwarpstr_b200.notb.so ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -o gpurun_out/r01c_fill_notb python scripts/time_fill.py --reads 30000 --reps 1 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
