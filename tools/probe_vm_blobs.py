"""A/B of pairing-VM schedules (tools/gen_pairing_vm.py --blob) on the bench workload, registry mode (no per-key kernel):
per blob, parity against the constructed expectation, then device ms of the whole step and the Miller / final split.
   python tools/probe_vm_blobs.py DIR [T]      (GPU box)"""
import os
import re
import subprocess
import sys
sys.path.insert(0, ".")
import numpy as np

if os.environ.get("_VM_PROBE_CHILD") != "1":
    # the per-phase split comes from the library's stderr trace: run the measurement in a child with the trace captured
    env = dict(os.environ, _VM_PROBE_CHILD="1", B200_BLS_TRACE="1")
    p = subprocess.run([sys.executable, __file__, *sys.argv[1:]], env=env, capture_output=True, text=True)
    trace = [ln for ln in p.stderr.splitlines() if ln.startswith("[b200 bls]")]
    k = 0
    for ln in p.stdout.splitlines():
        m = re.match(r"^RESULT (\d+) (.*)$", ln)
        if m:
            n = int(m.group(1))
            last = trace[k + n - 1] if k + n - 1 < len(trace) else ""
            k += n
            sp = re.search(r"K2 ([\d.]+) \| wait\(streamB\) ([\d.]+) \| miller ([\d.]+) \| final ([\d.]+)", last)
            print(m.group(2), f"| K2 {sp.group(1)} G2-wait {sp.group(2)} miller {sp.group(3)} final {sp.group(4)}" if sp else "", flush=True)
        else:
            print(ln, flush=True)
    if p.returncode:
        print(p.stderr[-2000:])
    sys.exit(p.returncode)

import bench
from ethereum_consensus_b200 import crypto, _lib
from tests import workloads

blob_dir = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
K = 512
orc_bls, _ = bench.load_oracles()
w = workloads.make_bls_workload(orc_bls, T, K, 0, threads=len(os.sched_getaffinity(0)))
_lib.init(0)
off, msgs, sigs = (bench.pin(w[k]) for k in ("off", "msgs", "sigs"))
reg = crypto.Registry(bench.pin(w["registry"]))
same = (w["kind"] != 4) & (w["kind"] != 5)
want = w["expect"][same].tolist()
team = 16 if T * 2 <= 2048 else 8     # the library's own rule (vm_team16_max = 2048 teams)
names = sorted(f for f in os.listdir(blob_dir) if f.startswith(f"t{team}_") and f.endswith(".bin"))
for name in names:
    blob = np.fromfile(os.path.join(blob_dir, name), dtype=np.uint32)
    crypto.vm_load_programs(blob)
    ms = []
    n = 6
    for i in range(n):
        got = reg.verify_batch(w["idx"], off, msgs, sigs)
        assert got[same].tolist() == want, name
        if i >= 2:
            ms.append(crypto.last_kernel_ms())
    print(f"RESULT {n} T={T} team={team} {name}: miller {blob[2]} rounds / {blob[3]} slots, final {blob[10]} / {blob[11]}: step {min(ms):.2f} ms", flush=True)
