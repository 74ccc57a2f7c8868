"""Attribute executed instructions / stall samples of an ncu capture to CUDA source lines.

ncu's CLI prints per-SASS-instruction metrics but no per-line roll-up; nvdisasm -g knows the line of every SASS
instruction.  Both list the kernel's instructions in address order, so they are zipped.
  usage: ncu_lines.py <ncu source csv (--page source --csv)> <cubin> <function substring> [top N]
"""
import csv, re, subprocess, sys, collections

src_csv, cubin, fn = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# find the function section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and fn in l and l.rstrip().endswith(":"))
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or (l.startswith(".text.") and l.rstrip().endswith(":")):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
if len(lines) != len(body):
    print(f"warning: {len(lines)} SASS instructions in the cubin vs {len(body)} in the capture (different build?)")
n = min(len(lines), len(body))
inst = collections.Counter(); samp = collections.Counter()
for (f, ln), r in zip(lines[:n], body[:n]):
    inst[(f, ln)] += int(r[ix["Instructions Executed"]]); samp[(f, ln)] += int(r[ix["# Samples"]] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
text = {}
for f in {k[0] for k in inst}:
    try:
        text[f] = open(f"/root/repo/phaneron_b200/csrc/{f}").read().splitlines()
    except Exception:
        text[f] = []
print(f"total warp-instr {ti/1e6:.1f}M, samples {ts}")
print("  %instr  %samples  file:line  source")
for k, v in inst.most_common(top):
    f, ln = k
    t = text.get(f, [])
    print(f"  {100*v/ti:6.2f}  {100*samp[k]/max(ts,1):7.2f}  {f}:{ln:<4d} {t[ln-1].strip()[:110] if 0 < ln <= len(t) else ''}")
