import csv, subprocess, sys
rep = sys.argv[1]; step = int(sys.argv[2]) if len(sys.argv) > 2 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw[1:]))
hdr = rows[0]; rows = rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
I = [num(r[ix["Instructions Executed"]]) for r in rows]; T = [num(r[ix["Thread Instructions Executed"]]) for r in rows]
S = [num(r[ix["# Samples"]]) for r in rows]
ti, tt, ts = sum(I), sum(T), sum(S)
print(f"total warp-inst {ti:.3e} avg threads {tt/ti:.2f} samples {ts:.0f} nSASS {len(rows)}")
# segment at large changes of execution count
seg_start = 0
def flush(a, b):
    i_ = sum(I[a:b]); t_ = sum(T[a:b]); s_ = sum(S[a:b])
    if i_ / ti > 0.004:
        key = [rows[k][ix["Source"]].split()[0] for k in range(a, b) if any(m in rows[k][ix["Source"]] for m in ("LDG", "ATOM", "VOTE", "WARPSYNC", "LDS.64", "STG", "MUFU.RCP", "BSYNC"))]
        from collections import Counter
        c = Counter(key)
        print(f"sass[{a:4d}:{b:4d}] n={b-a:4d} warp-inst {i_:.3e} ({100*i_/ti:5.1f}%) per-inst exec {i_/(b-a):.3e} avg thr {t_/max(i_,1):5.2f} samples {100*s_/ts:5.1f}%  {dict(c)}")
for k in range(1, len(rows) + 1):
    if k == len(rows) or abs(I[k] - I[seg_start]) > 0.15 * max(I[seg_start], 1.0) and (k - seg_start) >= 6:
        flush(seg_start, k); seg_start = k
