#!/bin/bash
# N independent end-to-end scans at the same time, one per GPU; the library's batch trace of GPU 0 goes to gpurun_out/e2e_n_trace.log
n=${1:-8}; reads=${2:-20000000}
python -c "import sys; sys.path.insert(0,'.'); import bench; bench.ensure_bam($reads, 2, True)" > /dev/null 2>&1
for i in $(seq 1 $((n-1))); do CUDA_VISIBLE_DEVICES=$i python tools/e2e_trace.py $reads > /dev/null 2> gpurun_out/e2e_n_$i.log & done
CUDA_VISIBLE_DEVICES=0 BAMSCAN_TRACE=1 python tools/e2e_trace.py $reads > /dev/null 2> gpurun_out/e2e_n_trace.log
wait
grep -h "rep " gpurun_out/e2e_n_trace.log gpurun_out/e2e_n_[1-9].log
