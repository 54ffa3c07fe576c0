"""A/B of the chunked strict pipeline on the bench workload (T x K, C-oracle signed): device ms per batch for each knob set.
   python tools/probe_chunks.py [T] [K]     (GPU box; prints one line per configuration)"""
import os
import sys
sys.path.insert(0, ".")
import numpy as np
import bench
from ethereum_consensus_b200 import crypto, _lib
from tests import workloads

T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 512
orc_bls, _ = bench.load_oracles()
w = workloads.make_bls_workload(orc_bls, T, K, 0, threads=len(os.sched_getaffinity(0)))
_lib.init(0)
import torch
pks, off, msgs, sigs = (bench.pin(w[k]) for k in ("pks", "off", "msgs", "sigs"))
want = w["expect"].tolist()
DEFAULTS = {"bls_chunks": 1, "bls_chunk_min_tuples": 2048, "bls_chunk_alt": 1, "bls_chunk_k1_cta": 128, "vm_cta": 32, "vm_team16_max": 2048}
CONFIGS = [{"bls_chunks": 1}, {}, {"bls_chunk_k1_cta": 384}, {"bls_chunk_alt": 0}, {"bls_chunks": 2}, {"bls_chunks": 8}, {"bls_chunks": 16},
           {"bls_chunks": 8, "vm_team16_max": 0}, {"bls_chunks": 4, "vm_team16_max": 0}, {"bls_chunks": 4, "vm_cta": 64}, {"bls_chunks": 4, "vm_cta": 32},
           {"bls_chunks": 1, "vm_cta": 64}, {"bls_chunks": 1, "vm_cta": 32}, {"bls_chunks": 8, "vm_cta": 64}]
if os.environ.get("B200_PROBE_QUICK"):
    CONFIGS = [{}]
if T < 2048:
    for c in CONFIGS:
        c.setdefault("bls_chunk_min_tuples", 2)
for cfg in CONFIGS:
    full = dict(DEFAULTS, **cfg)
    for k, v in full.items():
        crypto.tune(k, v)
    ms, dom = [], []
    for i in range(6):
        got = crypto.fast_aggregate_verify_batch(pks, off, msgs, sigs)
        assert got.tolist() == want, cfg
        if i >= 2:
            ms.append(crypto.last_kernel_ms()); dom.append(crypto.last_dominant_kernel_ms())
    print(f"T={T} K={K} {cfg}: device {min(ms):.2f} ms (median {sorted(ms)[len(ms)//2]:.2f}), per-key kernels {min(dom):.2f} ms", flush=True)
reg = crypto.Registry(bench.pin(w["registry"]))
same = (w["kind"] != 4) & (w["kind"] != 5)
for cfg in (({},) if os.environ.get("B200_PROBE_QUICK") else ({"vm_cta": 128}, {"vm_cta": 64}, {"vm_cta": 32})):
    for k, v in dict(DEFAULTS, **cfg).items():
        crypto.tune(k, v)
    ms = []
    for i in range(6):
        got = reg.verify_batch(w["idx"], off, msgs, sigs)
        assert got[same].tolist() == w["expect"][same].tolist()
        if i >= 2:
            ms.append(crypto.last_kernel_ms())
    print(f"registry T={T} {cfg}: device {min(ms):.2f} ms", flush=True)
