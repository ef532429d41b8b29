"""Instructions executed per CUDA source line from `ncu --page source --csv --print-source cuda,sass` output.
    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv; python profiles/by_line.py /tmp/src.csv [n_warp_tiles]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 312648.0
cur = None; hdr = None
inst = collections.Counter(); smp = collections.Counter(); text = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 8 or r[2] != "-": continue
    try: ln = int(r[0]); a = int(r[7]); b = int(r[6])
    except ValueError: continue
    inst[(cur, ln)] += a; smp[(cur, ln)] += b; text[(cur, ln)] = r[1]
tot = sum(inst.values()); ts = sum(smp.values())
print("total warp instructions", tot, "samples", ts)
for k, v in inst.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 50):
    print(f"{k[0]:>14s}:{k[1]:<4d} {v / unit:7.1f}/warp-tile {100 * v / tot:5.1f}%  smp {100 * smp[k] / ts:5.1f}%  {text[k].strip()[:100]}")
