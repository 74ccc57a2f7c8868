"""summarise an `ncu --page source --csv` dump: opcode mix, stall mix, hottest SASS regions"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
tot_thr = sum(int(r[ix["Thread Instructions Executed"]]) for r in body)
print(f"SASS lines {len(body)}  warp-instr {tot_inst/1e6:.1f}M  thread-instr {tot_thr/1e9:.2f}G  avg active lanes {tot_thr/max(tot_inst,1):.1f}")
ops = collections.Counter(); opl = collections.Counter()
for r in body:
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(r[ix["Instructions Executed"]])
print("opcode mix (warp-instr, %):")
for op, n in ops.most_common(28):
    print(f"  {op:10s} {n/1e6:8.2f}M {100*n/tot_inst:5.1f}%")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st = collections.Counter()
for r in body:
    for h in stalls:
        st[h] += int(r[ix[h]] or 0)
tot = sum(st.values())
print("stall samples (all):", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in st.most_common(10)))
# hottest 25 instructions by samples
samp = sorted(body, key=lambda r: -int(r[ix["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
ts = sum(int(r[ix["# Samples"]] or 0) for r in body)
print("hottest instructions by samples:")
for r in samp:
    top = max(stalls, key=lambda h: int(r[ix[h]] or 0))
    print(f"  {100*int(r[ix['# Samples']])/ts:5.2f}%  exec {int(r[ix['Instructions Executed']])/1e6:6.2f}M  {top[6:]:12s} {r[ix['Source']].strip()[:90]}")
