import csv, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith("==")]
rows=list(csv.DictReader(lines))
v=[float(r["Metric Value"].replace(",","")) for r in rows]
print(sys.argv[1], "fill avg", round(sum(v)/len(v)/1e6,4), "ms n", len(v))
