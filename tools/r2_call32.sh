#!/bin/bash
# Round-2 GPU call 32 (1 GPU): CTA shapes around the start of a strict batch — the first four waves of the per-key kernel as 128- or
# 384-thread CTAs, the signature / message kernels under them as 128- / 64- / 32-thread CTAs.
set -u
O=gpurun_out/r2c32; mkdir -p $O
B200_PROBE_CTA=1 timeout 600 python tools/probe_split.py 4096 512 > $O/cta_shapes.txt 2>&1
cat $O/cta_shapes.txt
