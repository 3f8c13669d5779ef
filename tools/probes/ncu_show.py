"""Side-by-side table of `ncu --csv --metrics` logs (one column per kernel launch).  usage: ncu_show.py file.csv [...]"""
import csv, sys
cols, tab = [], {}
for f in sys.argv[1:]:
    for r in csv.reader(open(f)):
        if len(r) > 14 and r[0].isdigit():
            key = (f.split("/")[-1].replace(".csv", ""), r[0], r[4].split("(")[0].replace("void ", "")[:28])
            if key not in cols: cols.append(key)
            tab.setdefault(r[12], {})[key] = r[14]
print(" " * 72 + " ".join(f"{c[0][-14:]:>16s}" for c in cols))
print(" " * 72 + " ".join(f"{c[2][-16:]:>16s}" for c in cols))
for k, d in tab.items():
    print(f"{k[:72]:72s}" + " ".join(f"{d.get(c, ''):>16s}" for c in cols))
