#!/bin/bash
# Round-2 GPU call 33 (1 GPU): last check of the final build — BLS parity tests incl. BASELINE configs[1] at full scale, smoke().
set -u
O=gpurun_out/r2c33; mkdir -p $O
timeout 900 python -m pytest tests/test_bls_gpu.py tests/test_config_scale_gpu.py -k "not configs2 and not configs3 and not two_rank" -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
tail -3 $O/pytest.log; tail -2 $O/smoke.log
