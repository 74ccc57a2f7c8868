"""group an `ncu --page source --csv` dump into runs of SASS lines with equal execution count"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
blocks = []; cur = None
for i, r in enumerate(body):
    n = int(r[ix["Instructions Executed"]])
    if cur and cur[2] == n: cur[1] = i
    else:
        cur = [i, i, n]; blocks.append(cur)
tot = sum(int(r[ix["Instructions Executed"]]) for r in body)
ts = sum(int(r[ix["# Samples"]] or 0) for r in body)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
print("lines            n   exec/line(M) total(M)  %instr  %samples")
for b in blocks:
    cnt = b[1] - b[0] + 1; t = cnt * b[2]
    samp = sum(int(body[k][ix["# Samples"]] or 0) for k in range(b[0], b[1] + 1))
    if t / tot > thr or samp / ts > thr:
        print(f"{b[0]:5d}-{b[1]:5d} {cnt:5d}  {b[2]/1e6:9.3f}  {t/1e6:8.2f}  {100*t/tot:5.1f}%  {100*samp/ts:5.1f}%   {body[b[0]][ix['Source']].strip()[:70]}")
