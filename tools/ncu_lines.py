"""Per-SOURCE-LINE sample counts of an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {h: k for k, h in enumerate(hdr)}
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        samp = int(r[ix["# Samples"]])
        inst = int(r[ix["Instructions Executed"]])
        thr = int(r[ix["Thread Instructions Executed"]])
    except Exception:
        continue
    lines.append((samp, inst, thr, cur_file.split("/")[-1], r[0], r[1].strip()[:90]))
tot = sum(l[0] for l in lines) or 1
toti = sum(l[1] for l in lines) or 1
print("total samples", tot, "warp-instr", toti)
for samp, inst, thr, f, ln, src in sorted(lines, key=lambda l: -l[0])[:top]:
    print("%5.1f%% samp %6d | inst %5.1f%% %9d thr/inst %4.1f | %s:%s  %s" %
          (100.0 * samp / tot, samp, 100.0 * inst / toti, inst, thr / max(inst, 1), f, ln, src))
