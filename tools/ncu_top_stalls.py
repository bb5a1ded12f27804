"""Print the instructions with the most warp-stall samples from
`ncu -i X.ncu-rep --page source --csv --print-source sass > f.csv`:  python ncu_top_stalls.py f.csv [top_n] [kernel_index]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else None
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
for ki, st0 in enumerate(starts):
    if which is not None and ki != which:
        continue
    end = starts[ki + 1] if ki + 1 < len(starts) else len(rows)
    hdr = rows[st0 + 1]
    iS, iSrc, iEx = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows[st0 + 2:end] if len(r) == len(hdr)]
    tot = sum(int(r[iS]) for r in data)
    print('=== kernel', ki, rows[st0][1][:110])
    print('total samples', tot, 'instructions', len(data), 'warp-instr executed', sum(int(r[iEx]) for r in data))
    agg = {}
    for r in data:
        for j in stalls:
            if r[j]:
                agg[hdr[j]] = agg.get(hdr[j], 0) + int(r[j])
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    if n <= 0:
        continue
    top = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:n]
    for i in sorted(top):
        r = data[i]
        st = sorted([(int(r[j]), hdr[j][6:]) for j in stalls if r[j] and int(r[j]) > 0], reverse=True)[:3]
        print(i, r[iS], r[iEx], r[iSrc].strip()[:80], st)
