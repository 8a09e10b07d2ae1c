#!/bin/bash
# tools/pcie_probe.sh <ngpus> [variants...]: runs tools/bin/pcie_probe on GPUs 0..n-1 at the same time
n=${1:-1}; shift
v=${@:-pinned wc thp sm h2d both}
start=$(( $(date +%s) + 6 ))
for g in $(seq 0 $((n-1))); do tools/bin/pcie_probe $g $start 1.5 $v > gpurun_out/pcie_probe_n${n}_g$g.log 2>&1 & done
wait
cat gpurun_out/pcie_probe_n${n}_g*.log | sort -k3,3 -k2,2n | awk '{print} {s[$3]+=$4} END {for (k in s) printf "SUM over %s processes, %-6s %8.2f GB/s\n", "'$n'", k, s[k]}'
