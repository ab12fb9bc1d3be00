# ncu evidence for profiles/: launch list of the bench command + one full capture of the fill kernel
set -x
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 2 -c 1 -o gpurun_out/${TAG}_fill_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_fill_full.log 2>&1
tail -c 200 gpurun_out/${TAG}_fill_full.log
python bench.py 2>&1 | tail -1 > gpurun_out/${TAG}_bench.json
cat gpurun_out/${TAG}_bench.json | cut -c1-600
