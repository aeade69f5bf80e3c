"""Lists the loops of a cubin's SASS (backward branches) with their instruction counts and
opcode mix: the CPU-side feedback for an issue-bound kernel (no GPU needed).
    python tools/sass_loops.py file.cubin [min_len]"""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
min_len = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ins = []  # (addr, text)
for line in out.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
print(f"{len(ins)} instructions")
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?0x([0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr_index:
        j = addr_index[tgt]
        n = i - j + 1
        if n < min_len:
            continue
        mix = collections.Counter()
        for _, tt in ins[j : i + 1]:
            tt = re.sub(r"^@!?U?P\d+\s+", "", tt)
            mix[tt.split()[0].split(".")[0]] += 1
        top = ", ".join(f"{k} {v}" for k, v in mix.most_common(14))
        print(f"loop 0x{tgt:04x}..0x{a:04x}: {n} instr | {top}")
