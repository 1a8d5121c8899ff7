#!/bin/bash
# final round-2 captures (after the same commands exited 0 without ncu):
#   gpurun --timeout 900 -- 'bash tools/ncu_capture3.sh r2b'
# k_msm_lut of the batch bench (this build: one addition site, independent 977-products in the Fq reduction) and
# k_tr_squeeze_coop of a lone proof (tools/latency_probe.py).
tag=${1:-r2b}
out=gpurun_out
mkdir -p $out
B="python bench.py --batch 4096 --steps 1 --warmup 1 --no-sweep --no-cpu-baseline"
L="python tools/latency_probe.py --lut-gb 48"
$B > $out/${tag}_capture_plain.json 2> $out/${tag}_capture_plain.err || exit 1
$L > $out/${tag}_latency_probe.txt 2>&1 || exit 1
cap() {   # name, regex, skip, command...
    name=$1; rx=$2; skip=$3; shift 3
    ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip -c 1 -f -o $out/${tag}_ncu_$name "$@" > $out/${tag}_ncu_$name.log 2>&1
    ncu -i $out/${tag}_ncu_$name.ncu-rep --page raw --csv > $out/${tag}_ncu_full_$name.csv 2>/dev/null
    ncu -i $out/${tag}_ncu_$name.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $out/${tag}_ncu_source_$name.csv.gz
    rm -f $out/${tag}_ncu_$name.ncu-rep
}
if [ "$2" != "more" ]; then
    cap k_msm_lut '^k_msm_lut' 30 $B
    cap k_tr_squeeze_coop '^k_tr_squeeze_coop' 20 $L
    ls -la $out | tail -8
fi
# the next kernels of the batch step by share (run with: bash tools/ncu_capture3.sh r2b more)
if [ "$2" = "more" ]; then
    cap k_trrp_phase2 '^k_trrp_phase2' 4 $B
    cap k_batch_to_affine '^k_batch_to_affine' 8 $B
    ls -la $out | tail -8
fi
