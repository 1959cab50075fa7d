"""Pick the judge-relevant metrics out of an `ncu --page raw --csv` dump (one entry per profiled launch; the second of
each pair launched by scripts/prof_kernels.py is the warm one)."""
import csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = {
    "us": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
}
units = rows[1]
out = []
for r in data:
    e = {"kernel": r[col["Kernel Name"]][:70], "id": int(r[col["ID"]])}
    for k, m in want.items():
        if m not in col:                     # some ncu versions prefix section metrics ("FBSP.TriageCompute.dram__throughput...")
            m = next((h for h in hdr if h.endswith("." + m)), m)
        if m in col and r[col[m]] not in ("", "no data", "n/a"):
            v = float(r[col[m]].replace(",", ""))
            u = units[col[m]]
            if k.startswith("dram_") and k != "dram_pct":
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if k == "us":
                v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
            e[k] = v
    out.append(e)
for e in out:
    gbs = (e.get("dram_read", 0) + e.get("dram_write", 0)) / max(e["us"], 1e-9) / 1e3          # bytes / us -> GB/s
    e["dram_gbs"] = gbs
    print(f"{e['id']:3d} {e['kernel'][:58]:58s} {e['us']:8.1f} us  tensor {e.get('tensor_pipe_pct', 0):5.1f}%  dram {gbs:6.0f} GB/s ({gbs / 6452.8:4.2f} of copy peak)  "
          f"lts {e.get('lts_pct', 0):5.1f}%  issue {e.get('issue_active_pct', 0):5.1f}%  R/W {e.get('dram_read', 0) / 1e6:7.1f}/{e.get('dram_write', 0) / 1e6:7.1f} MB")
json.dump(out, open(sys.argv[2], "w"), indent=1)
