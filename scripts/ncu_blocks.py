"""Basic-block view of an `ncu --page source --csv` export: instructions per row tile and stall samples.
usage: python scripts/ncu_blocks.py gpurun_out/prof_X_src.csv [rows=4194304] [min_instr_per_tile=40]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4194304
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 40
hdr = rows[1]; data = rows[2:]
ia, isrc, ie, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tiles = N // 128
blocks = []
for r in data:
    try: e = int(r[ie]); s = int(r[isamp])
    except Exception: continue
    if blocks and blocks[-1]['e'] == e:
        b = blocks[-1]; b['n'] += 1; b['s'] += s; b['last'] = r[isrc]
    else:
        blocks.append(dict(a=r[ia], e=e, n=1, s=s, first=r[isrc], last=r[isrc]))
tot = sum(b['e'] * b['n'] for b in blocks)
print('total warp instructions', tot, 'per tile', round(tot / tiles, 1), 'samples', sum(b['s'] for b in blocks))
for b in blocks:
    w = b['e'] * b['n'] / tiles
    if w > thr or b['s'] > 300:
        print(f"{b['a'][-5:]} n={b['n']:3d} exec/tile={b['e']/tiles:7.2f} instr/tile={w:8.1f} samp={b['s']:6d}  {b['first'].strip()[:38]:38s} .. {b['last'].strip()[:30]}")
