# ncu evidence for profiles/ (run under gpurun, one GPU): the launch list of the bench command, full captures
# of the DP fill kernel on the four automaton layouts, of the catch-all kernel, of the mid-stage and side
# kernels and of the FP64 add probe that is the roofline's denominator.  Reports are summarised on the box
# (gpurun brings back at most 64 MiB); only the HD fill report itself is kept.   bash scripts/profile_round.sh r02
set -x
TAG=${1:-r02}
WHAT=${2:-all}
Q="--no-cpu-baseline --parity-reads 0 --legs= --no-e2e-variants"
mkdir -p gpurun_out /tmp/ncu
if [ "$WHAT" = all ] || [ "$WHAT" = list ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 $Q > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
fi
full() {   # name, kernel regex, launches to skip, launches to take, command...
    local name=$1 rx=$2 skip=$3 take=$4; shift 4
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $take -f -o /tmp/ncu/${TAG}_$name "$@" \
        > /tmp/ncu/${TAG}_$name.log 2>&1
    tail -c 200 /tmp/ncu/${TAG}_$name.log
    ncu -i /tmp/ncu/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null
    for ((k = 0; k < take; k++)); do
        python scripts/ncu_regions.py /tmp/ncu/${TAG}_$name.ncu-rep "#$k" > gpurun_out/${TAG}_${name}_k$k.txt 2>&1
    done
}
if [ "$WHAT" = all ] || [ "$WHAT" = fill ]; then
full fill_HD   dtw_fill_kernel 2 2 python bench.py --reads 40000 --steps 1 --warmup 1 $Q
cp /tmp/ncu/${TAG}_fill_HD.ncu-rep gpurun_out/
full fill_DM2  dtw_fill_kernel 4 4 python bench.py --locus DM2 --reads 40000 --steps 1 --warmup 1 $Q
full fill_CAN  dtw_fill_kernel 4 4 python bench.py --locus CAN --reads 30000 --steps 1 --warmup 1 $Q
full fill_C4   dtw_fill_kernel 2 2 python bench.py --locus C9ORF72_1000 --reads 2000 --steps 1 --warmup 1 $Q
full fill_any  dtw_fill_any    2 2 python bench.py --reads 20000 --generic-only --steps 1 --warmup 1 $Q
fi
if [ "$WHAT" = all ] || [ "$WHAT" = rest ]; then
full mid_HD    mid_            5 5 python bench.py --reads 40000 --steps 1 --warmup 1 $Q
full mid_C4    mid_            5 5 python bench.py --locus C9ORF72_1000 --reads 2000 --steps 1 --warmup 1 $Q
full probe     fp64_add_probe  1 1 python bench.py --reads 5000 --steps 1 --warmup 1 $Q
full aux       "normalize_kernel|pore_lookup|dequantize" 2 4 python scripts/bench_aux.py
fi
if [ "$WHAT" = aux ]; then
full aux       normalize_kernel 2 1 env AUX_ONLY_NORM=1 python scripts/bench_aux.py
fi
du -sh gpurun_out
