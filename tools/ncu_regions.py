"""Per-region SIMD efficiency from an ncu source page: splits the SASS of the trace kernel at landmark
instructions (node LDG.128s, triangle LDG.128s, ATOMG refill, STG epilogue) and reports instructions executed,
average active threads and stall samples per region.  usage: python tools/ncu_regions.py rep.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw[1:]))
hdr = rows[0]; rows = rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
tot_i = sum(num(r[ix["Instructions Executed"]]) for r in rows); tot_t = sum(num(r[ix["Thread Instructions Executed"]]) for r in rows)
tot_s = sum(num(r[ix["# Samples"]]) for r in rows)
print(f"total warp-inst {tot_i:.3e}  thread-inst {tot_t:.3e}  avg threads {tot_t / tot_i:.2f}  samples {tot_s:.0f}")
# landmarks
ldg128 = [i for i, r in enumerate(rows) if "LDG.E.128" in r[ix["Source"]]]
print("LDG.128 at", ldg128)
bounds = sorted(set([0] + [ldg128[0] - 40, ldg128[0], ldg128[4] + 1, ldg128[5] - 12, ldg128[7] + 1] + ([ldg128[8] - 5] if len(ldg128) > 8 else []) + [len(rows)]))
names = {}
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = rows[a:b]
    i_ = sum(num(r[ix["Instructions Executed"]]) for r in seg); t_ = sum(num(r[ix["Thread Instructions Executed"]]) for r in seg)
    s_ = sum(num(r[ix["# Samples"]]) for r in seg)
    if i_ == 0: continue
    print(f"sass[{a:4d}:{b:4d}] n={b - a:4d}  warp-inst {i_:.3e} ({100 * i_ / tot_i:5.1f}%)  avg threads {t_ / i_:5.2f}  samples {100 * s_ / tot_s:5.1f}%   first: {seg[0][ix['Source']][:60]}")
