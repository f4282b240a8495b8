#!/bin/bash
# Build (if needed) and run the issue-rate probes on the GPU box; writes gpurun_out/issue_probe.txt
# (cycle table followed by the SASS opcode mix of every probe loop).
set -e
cd "$(dirname "$0")"
if [ ! -x issue_probe ] || [ issue_probe.cu -nt issue_probe ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o issue_probe issue_probe.cu
fi
mkdir -p ../../gpurun_out
{
  ./issue_probe ${1:-20000}
  echo
  echo "# SASS opcode mix of each probe loop (cuobjdump -sass; n = instructions per iteration incl. 3 loop-control)"
  cuobjdump -sass issue_probe | python3 sass_mix.py -
} | tee ../../gpurun_out/issue_probe.txt
