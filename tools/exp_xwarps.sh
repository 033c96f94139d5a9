#!/bin/bash
# Developer experiment (one gpurun call): rebuild the grid sweep with 1..4 exchange warps and time it on three shapes.
for x in 1 2 3 4; do
  touch fasta-python_b200/csrc/dense_gsweep.cu
  FB200_NVCC_DEFS="-DFB200_GS_XWARPS=$x" python fasta-python_b200/build.py > /dev/null 2>&1 || echo "build failed for $x"
  for shape in "40000 100000" "5000 100000" "100000 20000"; do
    echo "xwarps=$x shape=$shape: $(python tools/microbench.py $shape 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['sweep_ms'], d['sweep_avg_ms'], round(d['sweep_dram_GBs']), d['sweep_plan'])")"
  done
done
