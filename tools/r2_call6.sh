#!/bin/bash
# Round-2 GPU call 6 (1 GPU): signature/message kernels UNDER the per-key kernel with packed CTAs; small-batch behaviour.
set -u
O=gpurun_out/r2c6; mkdir -p $O
run() {  # $1 = label, rest = env assignments
  echo "== $*"
  env "$@" B200_BLS_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --skip-ssz --skip-strong --skip-single --tuples ${TUPLES:-4096} 2> /tmp/so.err \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tuples/s', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'K1', round(d['roofline']['kernel_ms'],2))"
  grep "b200 bls" /tmp/so.err | sed -n "5,5p"
}
{
TUPLES=4096 run B200_SMALL_ORDER=1
TUPLES=4096 run B200_SMALL_ORDER=0 B200_SMALL_CTA=512
TUPLES=4096 run B200_SMALL_ORDER=0 B200_SMALL_CTA=256
TUPLES=4096 run B200_SMALL_ORDER=0 B200_SMALL_CTA=128
TUPLES=256 run B200_SMALL_ORDER=1
TUPLES=256 run B200_SMALL_ORDER=0 B200_SMALL_CTA=256
TUPLES=256 run B200_SMALL_ORDER=0 B200_SMALL_CTA=32
} > $O/small_order.txt 2>&1
ls -la $O
