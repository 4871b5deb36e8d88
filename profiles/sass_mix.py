"""Aggregate an `ncu --page source --csv` SASS dump by opcode: python profiles/sass_mix.py report.csv [top]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if 'Instructions Executed' in r)
ii, si = hdr.index('Instructions Executed'), hdr.index('Source')
mix = collections.Counter(); static = collections.Counter()
for r in rows:
    if len(r) != len(hdr) or r is hdr: continue
    try: n = int(r[ii])
    except ValueError: continue
    op = r[si].strip().split()
    op = [t for t in op if not t.startswith('@')]
    name = op[0].rstrip(';') if op else '?'
    key = '.'.join(name.split('.')[:2]) if name.split('.')[0] in ('LDG','LD','LDS','STS','ST','STG','MUFU','LDL','STL','SHFL','IMAD','ATOMS','RED','LDC') else name.split('.')[0]
    mix[key] += n; static[key] += 1
tot = sum(mix.values())
print(f"total warp-instructions executed: {tot}   static SASS instructions: {sum(static.values())}")
for k, v in mix.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print(f"{k:14s} {v:>14d} {100*v/tot:6.2f}%   static {static[k]}")
