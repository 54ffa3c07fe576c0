#!/bin/bash
# Round-2 GPU call 4 (1 GPU): K1 inline vs by-value-call products x 8 / 12 warps; where the signature/message kernels go
# relative to K1 now that K1's CTA leaves room on the SM; incremental-root phase trace; bench + ncu of the lazy build.
set -u
O=gpurun_out/r2c4; mkdir -p $O
for lib in ethereum_consensus_b200/libb200_consensus.so ethereum_consensus_b200/libb200_consensus_call.so; do
  for v in 0 7; do echo "== $lib"; B200_LIB=$PWD/$lib B200_G1_VARIANT=$v timeout 300 python tools/tune_k1.py; done
done > $O/k1_variants.txt 2>&1
for so in 1 0 2; do
  echo "== B200_SMALL_ORDER=$so"
  B200_SMALL_ORDER=$so B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single 2> /tmp/so.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tuples/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'K1', round(d['roofline']['kernel_ms'],2))"
  grep "b200 bls" /tmp/so.err | sed -n "5,5p"
done > $O/small_order.txt 2>&1
timeout 600 python tools/probe_incremental.py > $O/incremental_trace.txt 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" >> $O/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_g1_validate -s 2 -c 1 -o $O/k1_lazy \
   python bench.py --steps 1 --warmup 3 --skip-ssz --skip-strong --skip-single > /dev/null 2> $O/k1_ncu.err
ls -la $O
