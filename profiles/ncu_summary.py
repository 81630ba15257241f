import csv, re, collections, sys
path=sys.argv[1]; frac_lo=float(sys.argv[2]); frac_hi=float(sys.argv[3])
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
rows=list(csv.DictReader(lines))
n=len(rows)
seg=rows[int(n*frac_lo):int(n*frac_hi)]
names=collections.defaultdict(lambda:[0,0.0])
for row in seg:
    k=row['Kernel Name']
    m=re.search(r'(k_\w+|Device\w+Kernel\w*)', k)
    k=m.group(1) if m else k[:60]
    v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit.startswith('n'): v/=1e3
    elif unit.startswith('m'): v*=1e3
    names[k][0]+=1; names[k][1]+=v
tot=sum(v[1] for v in names.values())
print(f"launches: {len(seg)}  total device time: {tot:.1f} us")
for k,v in sorted(names.items(), key=lambda kv:-kv[1][1])[:40]:
    print(f"{v[1]:10.1f} us {v[0]:5d}  {100*v[1]/tot:5.1f}%  avg {v[1]/v[0]:8.1f} us  {k}")
