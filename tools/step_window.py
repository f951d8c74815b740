"""Print '<skip> <count>' for ncu: the launches of ONE evaluation in the middle of a bench run, from its launch list."""
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 4 and r[0].isdigit()]
pat = re.compile(sys.argv[2])
names = [r[4] for r in rows if pat.search(r[4])]
evals = sum('vx_search_kernel' in n for n in names)
per = len(names) // max(evals, 1)
print(per * (evals // 2), per)
