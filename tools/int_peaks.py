import sys, json
sys.path.insert(0, '.')
from ethereum_consensus_b200 import _lib, crypto
_lib.init(0)
names = {0: "IMAD.WIDE.U32", 1: "IMAD.U32 (lo)", 2: "SHF/LOP3/IADD3 mix", 3: "IMAD.HI.U32", 4: "IADD3", 5: "52x52 product via 2 DFMA + DADD + int64 accumulate", 6: "DFMA (with int glue)", 7: "IMAD.WIDE.U32 (multiplicand from another chain, register multiplier)", 8: "IMAD.WIDE.U32 (immediate multiplier)"}
out = {names[k]: crypto.measure_int_peak(k) for k in range(9)}
print(json.dumps(out, indent=1))
