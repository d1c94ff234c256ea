"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line (stall samples, executed instrs)."""
import csv
import sys


def num(x):
    try:
        return int(x)
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, agg = None, {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 7 and r[0].isdigit() and r[2] == '-':
        key = (cur, int(r[0]), r[1].strip()[:120])
        a = agg.setdefault(key, [0, 0])
        a[0] += num(r[6]); a[1] += num(r[7])
tot = sum(v[0] for v in agg.values())
print('total samples', tot, 'total warp-instr', sum(v[1] for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%6d %5.1f%% %9d  %s:%d  %s' % (v[0], 100.0 * v[0] / max(tot, 1), v[1], k[0], k[1], k[2]))
