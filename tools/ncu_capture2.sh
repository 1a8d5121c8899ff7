#!/bin/bash
# round-2 captures of the batch path (after the same commands exited 0 without ncu):
#   gpurun --timeout 1500 -- 'bash tools/ncu_capture2.sh r2'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
B="python bench.py --batch 4096 --steps 1 --warmup 1 --no-sweep --no-cpu-baseline"
$B > $out/${tag}_capture_plain.json 2> $out/${tag}_capture_plain.err || exit 1
cap() {   # name, regex, skip
    ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 -c 1 -f -o $out/${tag}_ncu_$1 $B > $out/${tag}_ncu_$1.log 2>&1
    ncu -i $out/${tag}_ncu_$1.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$1.csv 2>/dev/null
    ncu -i $out/${tag}_ncu_$1.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $out/${tag}_ncu_source_$1.csv.gz
    rm -f $out/${tag}_ncu_$1.ncu-rep
}
cap k_msm_lut '^k_msm_lut' 30
cap k_tr_squeeze '^k_tr_squeeze' 20
# launch list of one small step (cold-cache, serialised: compare SHARES)
python bench.py --batch 512 --steps 1 --warmup 3 --no-sweep --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $out/${tag}_launches_bench_b512.csv \
    python bench.py --batch 512 --steps 1 --warmup 3 --no-sweep --no-cpu-baseline > $out/${tag}_launches.log 2>&1
ls -la $out | tail -12
