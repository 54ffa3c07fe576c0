#!/bin/bash
# Round-2 GPU call 3 (2 GPUs): K1 variants (lazy [0,2p) arithmetic vs canonical; 8..11 warps/SM), shuffling parity, the
# 2-rank NCCL-in-library test, bench at N=2.
set -u
O=gpurun_out/r2c3; mkdir -p $O
for lib in ethereum_consensus_b200/libb200_consensus.so ethereum_consensus_b200/libb200_consensus_canon.so; do
  for v in 0 9 8 10; do echo "== $lib"; B200_LIB=$PWD/$lib B200_G1_VARIANT=$v timeout 300 python tools/tune_k1.py; done
done > $O/k1_variants.txt 2>&1
timeout 900 python -m pytest tests/test_shuffle_gpu.py tests/test_bls_gpu.py tests/test_config_scale_gpu.py -x -q --durations=5 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
ls -la $O
