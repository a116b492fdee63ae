import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value"); ui=h.index("Metric Unit")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    u=r[ui]
    if u=="ns": v/=1000
    elif u=="ms": v*=1000
    d[r[ki].split("(")[0].replace("wva::","")].append(v)
print(" | ".join(f"{k} {sorted(v)[len(v)//2]:.1f}" for k,v in d.items() if k.startswith("grid") or k.startswith("build")))
