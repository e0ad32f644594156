"""Condenses the round-2 schedule sweeps (profiles/r2_sweep_*.json, written by tools/r2_sweep.py) into one table:
best time per scene / query / schedule with its knobs.  usage: python tools/sweep_table.py > profiles/r2_sweeps.md"""
import glob, json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WHAT = {"a": "first cooperative kernel (k_trace_coop v1), 7 CTAs/SM", "a_mb8": "same, every kernel at 8 CTAs/SM (64 registers)",
        "c": "re-fill threshold sweep (coop_incoherent)", "d": "scan-based pair push + slot schedule (64 slots, 128-pair list)",
        "e_s48": "slot schedule, 48 slots / 64-pair list", "e_s64l64": "slot schedule, 64 slots / 64-pair list", "e_s40": "slot schedule, 40 slots / 64-pair list",
        "f_base": "pair list 128 (coop kernels)", "f_p64": "pair list 64 (coop kernels)", "g": "small launch (640x360), tail flush",
        "h_signbits": "node test with sign-bit miss detection (RT_NODE_SIGNBITS), all defaults otherwise", "h_signbits_ffma2": "same + packed FFMA2 plane arithmetic (RT_NODE_FFMA2)",
        "i_refill": "re-fill threshold of the coherent cooperative schedule", "j_base": "FINAL defaults (same box and run as j_signbits)",
        "j_signbits": "RT_NODE_SIGNBITS build, same box and run as j_base: rejected",
        "k_base": "63-bit Morton keys, 8 sort passes (same box and run as k_keys)", "k_keys": "48-bit Morton keys, 6 sort passes: identical nodes / triangles per ray",
        "m_k16": "48-bit Morton keys, 6 sort passes (same box and run as m_k13)", "m_k13": "39-bit Morton keys (13 per axis), 5 sort passes - the default: identical nodes / triangles per ray",
        "n_base": "final defaults (same box and run as n_l9 / n_c8)", "n_l9": "any / count / first kernels at 9 CTAs per SM (56 registers): rejected",
        "n_c8": "pooled closest-hit kernel at 8 CTAs per SM (64 registers): rejected",
        "q_nofill": "before the root-frame pre-test (-DRT_ROOT_AT_FILL=0), same box and run as q_fill",
        "q_fill": "root-frame pre-test at pool fill (default): rays that miss the box around the root's children never take a lane",
        "r_refill": "re-fill threshold of coop_incoherent with the root-frame test in place: heightfields 24-26, soup 26-28 -> default 26",
        "s_pairs": "pair-list threshold of coop_incoherent with the root-frame test in place: 24 beats 16 by 1-2.5 % on the heightfields, 0.5 % on the soup -> default 24",
        "o_refill": "re-fill threshold of coop_incoherent, final kernels: the soup prefers 28-30, the heightfield 24-26, all within 2 % (default stays 24)"}
print("# Round-2 schedule sweeps (one B200, CUDA events, L2 flushed, min of 5)\n")
print("Scenes: config2 = 327 680-tri icosphere x 3840x2160 camera; ico8 = 1.31 M-tri icosphere, same camera; hf4m = 4.19 M-tri heightfield x 20 M random rays;")
print("soup1m = 1 M-tri soup x 10 M random rays; hf16m = 16.8 M-tri heightfield x 30 M random rays; small = config-2 mesh x 640x360 camera.")
print("Every row was checked bit-identical to the first schedule of its sweep (`identical_to_first`).\n")
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_sweep_*.json"))):
    tag = os.path.basename(path)[len("r2_sweep_"):-5]
    d = json.load(open(path))
    rows = [r for r in (d["rows"] if isinstance(d, dict) else d) if "ms" in r]
    print(f"## `{os.path.basename(path)}` — {WHAT.get(tag, tag)}\n")
    print("| scene | query | schedule | best ms | Mrays/s | frac | tri_threshold | refill | nodes/ray | tris/ray |\n|---|---|---|---|---|---|---|---|---|---|")
    best = {}
    for r in rows:
        k = (r["scene"], r["query"], r["schedule"])
        if k not in best or r["ms"] < best[k]["ms"]:
            best[k] = r
    for k in sorted(best, key=lambda k: (k[0], k[1], best[k]["ms"])):
        r = best[k]
        print(f"| {k[0]} | {k[1]} | {k[2]} | {r['ms']:.3f} | {r['mrays_s']:.0f} | {r['frac']:.3f} | {r['tri_threshold']} | {r.get('refill_threshold', 0)} | {r['nodes_per_ray']:.2f} | {r['tris_per_ray']:.2f} |")
    print()
