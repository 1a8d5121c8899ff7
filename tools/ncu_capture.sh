#!/bin/bash
# ncu --set full captures of the hot kernels (one launch each), run on the GPU box through gpurun:
#   gpurun --timeout 1200 -- 'bash tools/ncu_capture.sh r2'
# Each capture only after the same command exited 0 without ncu.  Raw-page CSVs land in gpurun_out/.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
set -x
python tools/sweep.py 18 > $out/${tag}_sweep18.jsonl 2> $out/${tag}_sweep18.err || exit 1
python tools/kernel_bench.py 512 1024 261 9 1 > $out/${tag}_kernel_bench.log 2>&1 || exit 1
cap() {   # name, regex, skip, command...
    name=$1; rx=$2; skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c 1 -f -o $out/${tag}_ncu_$name "$@" > $out/${tag}_ncu_$name.log 2>&1
    ncu -i $out/${tag}_ncu_$name.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$name.csv 2>/dev/null
    ncu -i $out/${tag}_ncu_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $out/${tag}_ncu_source_$name.csv.gz
    rm -f $out/${tag}_ncu_$name.ncu-rep      # gpurun_out/ is capped at 64 MiB
}
cap k_pair_fold '^k_pair_fold' 0 python tools/sweep.py 18
cap k_pip_accum '^k_pip_accum' 0 python tools/sweep.py 18
cap k_fold_dots '^k_fold_dots' 0 python tools/sweep.py 18
cap k_pip_reduce1 '^k_pip_reduce1' 0 python tools/sweep.py 18
cap k_msm_gens 'k_msm_gens$' 0 python tools/kernel_bench.py 512 1024 261 9 1
ls -la $out
