#!/bin/bash
# compute-sanitizer evidence of HEAD (round 2): memcheck + racecheck over the persistent kernels (both execution modes),
# the visual map, the batched path, and -- with 2 GPUs -- the fused NVLink exchange.  Summaries go to gpurun_out/;
# the judged copies are profiles/r02_sanitizer_*.txt.
#   usage (on the GPU box): bash profiles/run_sanitizer.sh [2]      (argument 2: also the 2-GPU multirank tests)
mkdir -p gpurun_out
SAN="compute-sanitizer --target-processes all --print-limit 5"
T1="tests/test_gpu_parity.py -k update"
T2="tests/test_vmap_select.py tests/test_gpu_batch.py tests/test_gpu_edge_sizes.py tests/test_gpu_scan_order.py"
for tool in memcheck racecheck; do
  timeout 900 $SAN --tool $tool python -m pytest $T1 -m gpu -q -x > gpurun_out/san_${tool}_updates.txt 2>&1
  timeout 900 $SAN --tool $tool python -m pytest $T2 -m gpu -q -x > gpurun_out/san_${tool}_vmap_batch.txt 2>&1
done
if [ "$1" == "2" ]; then
  for tool in memcheck racecheck; do
    timeout 900 $SAN --tool $tool python -m pytest tests/test_multirank.py -m gpu -q -x -k p2p > gpurun_out/san_${tool}_p2p_n2.txt 2>&1
  done
fi
for f in gpurun_out/san_*.txt; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $f | tail -6; done
