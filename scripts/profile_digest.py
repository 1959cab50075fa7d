"""Digest an ncu launch list holding gpu__time_duration + dram bytes into per-kernel totals (JSON + text)."""
import collections, csv, json, re, sys

path, out_json = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ki, mi, vi, ui, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    d = per.setdefault(r[ii], {"name": re.sub(r'\(.*', '', r[ki]).replace('void avcer::', '').replace('avcer::', '')})
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    if r[mi].startswith('gpu__time'):
        d["us"] = v / 1000 if u == 'ns' else (v * 1000 if u == 'ms' else v)
    else:
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        d[r[mi].split('.')[0]] = v * mult
agg = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_read": 0.0, "dram_write": 0.0})
for d in per.values():
    # per exact kernel (template arguments included: bench.py's roofline names ONE kernel), plus one family entry for every
    # tcgen05 contraction kernel (generic one- / two-SM GEMM, halo 3x3 conv, fused stem + pool)
    keys = [d["name"].strip()]
    if d["name"].startswith(("tc_gemm", "conv3x3_kernel", "stem_pool")):
        keys.append("family:tcgen05_contractions")
    for k in keys:
        a = agg[k]
        a["launches"] += 1; a["us"] += d.get("us", 0); a["dram_read"] += d.get("dram__bytes_read", 0); a["dram_write"] += d.get("dram__bytes_write", 0)
tot = sum(a["us"] for k, a in agg.items() if not k.startswith("family:"))
res = {}
print(f"total {tot:.0f} us, {len(per)} launches (ncu: cold-cache, serialised -- compare shares, not absolutes)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    res[k] = dict(a, share=a["us"] / tot, dram_bytes_per_launch=(a["dram_read"] + a["dram_write"]) / a["launches"])
    print(f"{k[:60]:60s} {a['us']:10.0f} us {100 * a['us'] / tot:5.1f}%  n={a['launches']:5d}  dram r/w {a['dram_read'] / 1e9:7.2f}/{a['dram_write'] / 1e9:6.2f} GB  ({res[k]['dram_bytes_per_launch'] / 1e6:7.2f} MB/launch)")
json.dump(res, open(out_json, "w"), indent=1)
