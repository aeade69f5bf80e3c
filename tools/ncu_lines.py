"""Top source lines of a kernel in an .ncu-rep (captured with --import-source on): warp instructions executed,
lanes per instruction, stall samples -- the per-line view behind the summaries under profiles/.
    python tools/ncu_lines.py gpurun_out/x/prof.ncu-rep [kernel] [top N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 else "lol_render"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Line No")
col = {n: i for i, n in enumerate(hdr)}
I, T, S = col["Instructions Executed"], col["Thread Instructions Executed"], col["# Samples"]
lines = []
for r in rows:
    if len(r) == len(hdr) and r[0].isdigit():
        try:
            lines.append((int(r[0]), r[1], int(r[I]), int(r[T]), int(r[S])))
        except ValueError:
            pass
tot_i = sum(x[2] for x in lines) or 1
tot_t = sum(x[3] for x in lines)
tot_s = sum(x[4] for x in lines) or 1
print(f"{kernel}: {tot_i} warp instructions, {tot_t / tot_i:.2f} lanes per instruction, {tot_s} samples")
print(f"{'line':>6} {'% inst':>7} {'lanes':>6} {'% samples':>9}  source")
for ln, src, i, t, s in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{ln:6d} {100 * i / tot_i:7.2f} {t / max(i, 1):6.1f} {100 * s / tot_s:9.2f}  {src.strip()[:110]}")
