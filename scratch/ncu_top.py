"""top SASS instructions by stall samples from `ncu -i rep --page source --csv --kernel-name regex:K`"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
i_s = hdr.index("# Samples"); i_src = hdr.index("Source"); i_ex = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
body = rows[2:]
tot = sum(int(r[i_s]) for r in body)
print("total samples", tot, "instructions", len(body), "executed", sum(int(r[i_ex]) for r in body))
agg = {}
for r in body:
    for i, h in stall_cols:
        try: agg[h] = agg.get(h, 0) + int(r[i])
        except: pass
print(sorted(agg.items(), key=lambda kv: -kv[1])[:10])
order = sorted(range(len(body)), key=lambda k: -int(body[k][i_s]))[:top]
for k in sorted(order):
    r = body[k]
    why = sorted([(int(r[i]), h) for i, h in stall_cols if r[i] not in ("", "0")], reverse=True)[:3]
    print(k, r[i_s].rjust(5), r[i_ex].rjust(6), r[i_src].strip()[:70].ljust(70), why)
