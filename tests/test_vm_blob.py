"""Run-time loadable pairing-VM schedules (b200_vm_load_programs): the generator's blob for the default parameters is word
for word the compiled-in program (CPU), and loading schedules — the same one, an alternative one, malformed ones — never
changes a verdict (GPU)."""
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GEN = ROOT / "tools" / "gen_pairing_vm.py"


def _blob(tmp_path, *args):
    out = tmp_path / "prog.bin"
    subprocess.run([sys.executable, str(GEN), *args, "--blob", str(out)], check=True, capture_output=True, cwd=ROOT)
    return np.fromfile(out, dtype=np.uint32)


def _header_words(path, name):
    text = path.read_text()
    body = re.search(r"static const uint32_t %s\[\d+\] = \{(.*?)\};" % name, text, re.S).group(1)
    return np.array([int(x, 16) for x in re.findall(r"0x([0-9a-f]{8})u", body)], dtype=np.uint32)


@pytest.mark.parametrize("team,suffix,header", [(8, "", "pairing_vm_prog.cuh"), (16, "16", "pairing_vm_prog16.cuh")])
def test_default_blob_is_the_compiled_in_program(tmp_path, team, suffix, header):
    blob = _blob(tmp_path, str(team))
    hdr = ROOT / "ethereum_consensus_b200" / "csrc" / header
    text = hdr.read_text()
    const = lambda n: int(re.search(r"constexpr int %s = (\d+);" % n, text).group(1))  # noqa: E731
    outs = lambda n: [int(x) for x in re.search(r"constexpr int %s\[6\] = \{(.*?)\};" % n, text).group(1).split(",")]  # noqa: E731
    assert blob[0] == 0xB200564D and blob[1] == team
    mr, ms, fr, fs = int(blob[2]), int(blob[3]), int(blob[10]), int(blob[11])
    assert (mr, ms) == (const(f"kMillerRounds{suffix}"), const(f"kMillerSlots{suffix}"))
    assert (fr, fs) == (const(f"kFinalRounds{suffix}"), const(f"kFinalSlots{suffix}"))
    assert blob[4:10].tolist() == outs(f"kMillerOut{suffix}") and blob[12:18].tolist() == outs(f"kFinalOut{suffix}")
    assert np.array_equal(blob[18:18 + mr * team], _header_words(hdr, f"h_miller_code{suffix}"))
    assert np.array_equal(blob[18 + mr * team:], _header_words(hdr, f"h_final_code{suffix}"))


@pytest.mark.gpu
def test_loading_schedules_never_changes_a_verdict(engine, tmp_path):
    import json
    from ethereum_consensus_b200 import crypto, _lib
    golden = json.loads((ROOT / "tests" / "golden" / "bls_cases.json").read_text())
    cases = [c for c in golden["fast_aggregate_verify"] if len(c["msg"]) == 64] * 3
    pks = np.frombuffer(b"".join(bytes.fromhex(p) for c in cases for p in c["pks"]), dtype=np.uint8)
    off = np.cumsum([0] + [len(c["pks"]) for c in cases]).astype(np.uint32)
    msgs = np.frombuffer(b"".join(bytes.fromhex(c["msg"]) for c in cases), dtype=np.uint8)
    sigs = np.frombuffer(b"".join(bytes.fromhex(c["sig"]) for c in cases), dtype=np.uint8)
    want = [c["code"] for c in cases]
    defaults = {t: _blob(tmp_path, str(t)) for t in (8, 16)}
    try:
        for t, args in ((8, ("8", "--mix-light", "--heavy-min", "8")), (16, ("16", "--mix-light", "--heavy-min", "4")), (8, ("8", "64", "64"))):
            crypto.vm_load_programs(_blob(tmp_path, *args))
            for team16_max in (0, 1 << 20):     # force the team size the blob was made for, and the other one
                crypto.tune("vm_team16_max", team16_max)
                assert crypto.fast_aggregate_verify_batch(pks, off, msgs, sigs).tolist() == want, (args, team16_max)
        bad = defaults[8].copy()
        for mutate in (lambda b: b.__setitem__(0, 1), lambda b: b.__setitem__(1, 12), lambda b: b.__setitem__(3, 300),
                       lambda b: b.__setitem__(20, int(b[20]) | 0xff), lambda b: b.__setitem__(5, 255)):
            b = bad.copy()
            mutate(b)
            with pytest.raises(_lib.EngineError):
                crypto.vm_load_programs(b)
        with pytest.raises(_lib.EngineError):
            crypto.vm_load_programs(bad[:-1])
        assert crypto.fast_aggregate_verify_batch(pks, off, msgs, sigs).tolist() == want     # rejected blobs left the programs alone
    finally:
        for t in (8, 16):
            crypto.vm_load_programs(defaults[t])
        crypto.tune("vm_team16_max", 2048)
