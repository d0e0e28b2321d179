#!/bin/bash
# round 2, run 24: does DRAM deliver scattered 4 KB blocks (the KV pages' access pattern) as fast as a sequential stream?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 tools/probes/bulk_probe > gpurun_out/r2x_bulk_probe_scatter.log 2>&1; echo "probe exit $?"; cat gpurun_out/r2x_bulk_probe_scatter.log
