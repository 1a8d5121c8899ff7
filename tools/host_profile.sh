#!/bin/bash
# Flat CPU profile of the host side of the batch prover / verifier (dev tool, run on the GPU box through gpurun):
# rebuilds the host layer with the SIGPROF sampler (-DBPPP_SAMPLER -g) into the scratch copy, runs bench.py under a
# 4-core limit and resolves the samples to symbols and source lines.   bash tools/host_profile.sh [cores]
cores=${1:-4}
set -e
cd "$(dirname "$0")/.."
g++ -O3 -g -fno-omit-frame-pointer -funroll-loops -std=c++17 -fPIC -march=x86-64-v2 -Wno-unknown-pragmas -DBPPP_SAMPLER -c bulletproofspp_b200/csrc/host/rp_host.cpp -o build/rp_host_prof.o
/usr/local/cuda/bin/nvcc -shared -o bulletproofspp_b200/lib/libbppp_b200.so build/capi.o build/rp_host_prof.o -lpthread -ldl
mkdir -p gpurun_out
BPPP_SAMPLE=gpurun_out/host_samples.txt taskset -c 0-$((cores-1)) python bench.py --no-cpu-baseline --no-sweep --steps 4 --warmup 3 --host-threads $cores > gpurun_out/host_profile_bench.json 2>/dev/null
python tools/sample_report.py gpurun_out/host_samples.txt 45 > gpurun_out/host_profile.txt
# source lines of the hottest addresses inside the library
grep libbppp_b200 gpurun_out/host_samples.txt | awk '{print $2}' | cut -d@ -f1 | sort | uniq -c | sort -rn | head -60 > gpurun_out/host_hot_pcs.txt
awk '{print "0x"$2}' gpurun_out/host_hot_pcs.txt | addr2line -e bulletproofspp_b200/lib/libbppp_b200.so -f -C -i | paste - - | head -150 > gpurun_out/host_hot_lines.txt || true
head -60 gpurun_out/host_profile.txt
rm -f gpurun_out/host_samples.txt
