import csv, sys, glob
files = sys.argv[1:]
tab = {}
for f in files:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    for r in rows:
        tab.setdefault(r[-3] if False else r[12], {})[f] = r[14]
for k, d in tab.items():
    print(f"{k[:80]:80s} " + " ".join(f"{d.get(f, ''):>14s}" for f in files))
