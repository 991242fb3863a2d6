#!/bin/bash
# pipeline-depth sensitivity of the conv main loop (AVC_T2_NSTAGE=2 vs 3) and hs=1 with bulk weight copies (bit 4096)
set -u
O=gpurun_out; mkdir -p $O
export DIAG_PROBES=0,496,4096
for cfg in "0 0" "2 0" "0 1"; do
  set -- $cfg
  AVC_T2_NSTAGE=$1 AVC_T2_HS=$2 timeout 100 python tools/diag_ablate.py 2>&1 | grep "conv5\|status" | cut -c1-300 | sed "s/^/nstage=$1 hs=$2: /"
done
