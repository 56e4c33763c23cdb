"""Per-kernel resource usage and instruction mix of dfpsr_b200/libdfpsr_b200.so from cuobjdump (no GPU needed):
registers / stack / shared memory per kernel (-res-usage) and counts of the SASS mnemonics that characterise each kernel
(global / shared / local loads and stores, integer multiply-add, float add / multiply, conversions, atomics, bulk copies, tensor-core ops).
usage: python tools/sass_summary.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dfpsr_b200", "libdfpsr_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage, name = {}, None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups())
demangled = dict(zip(usage.keys(), subprocess.run(["c++filt"] + list(usage.keys()), capture_output=True, text=True).stdout.splitlines()))
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
mix, current = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        current = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and current:
        op = m.group(1)
        mix[current][op.split(".")[0]] += 1
groups = [("LDG", ["LDG"]), ("STG", ["STG"]), ("LDS/STS", ["LDS", "STS"]), ("LDL/STL (spills)", ["LDL", "STL"]), ("IMAD/IADD3/LOP3/PRMT/SHF", ["IMAD", "IADD3", "LOP3", "PRMT", "SHF", "LEA"]),
          ("FADD/FMUL/FFMA", ["FADD", "FMUL", "FFMA"]), ("MUFU", ["MUFU"]), ("F2I/I2F/F2F", ["F2I", "I2F", "F2F", "I2FP", "F2IP"]), ("DADD/DMUL/DFMA", ["DADD", "DMUL", "DFMA"]),
          ("ATOM/RED", ["ATOM", "ATOMS", "ATOMG", "RED"]), ("SHFL/VOTE", ["SHFL", "VOTE", "MATCH"]), ("UBLKCP/UTMA*/SYNCS", ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS"]), ("UTC*MMA/LDTM", ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM"])]
lines = ["# SASS summary of dfpsr_b200/libdfpsr_b200.so (sm_100a), produced by tools/sass_summary.py", "",
         "Compiled with `-gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false` (bit-exact parity needs the reference's rounding).",
         "Tensor-core and tensor-map instructions are absent on purpose: nothing on this path is a dense contraction, and the one bulk-copy use that was tried",
         "(stored interpolation checkpoints, DFPSR_CHK_BULK) measured slower than vector loads (profiles/r2_tma_experiment.md).", "",
         "| kernel | registers | stack B | static shared B | instructions | " + " | ".join(g for g, _ in groups) + " |", "|---|---|---|---|---|" + "---|" * len(groups)]
for key in sorted(usage, key=lambda k: demangled[k]):
    reg, stack, shared, local = usage[key]
    counts = mix.get(key, {})
    total = sum(counts.values())
    short = demangled[key].replace("(anonymous namespace)::", "").replace("dfpsr::", "").split("(")[0].replace("void ", "")
    lines.append(f"| {short} | {reg} | {stack} | {shared} | {total} | " + " | ".join(str(sum(counts.get(op, 0) for op in ops)) for _, ops in groups) + " |")
text = "\n".join(lines) + "\n"
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text)
