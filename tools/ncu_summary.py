#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small CSV + per-kernel means for profiles/.

    python tools/ncu_summary.py gpurun_out/<tag>/prof.ncu-rep profiles/<name>

writes <name>.csv (one row per profiled launch) and prints per-kernel means; also updates
profiles/traffic.json ("<kernel>@<hii>" -> dram bytes per launch) when --hii is given.
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

WANT = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
]


def to_bytes(val, unit):
    v = float(val)
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    hii = None
    if "--hii" in sys.argv:
        hii = int(sys.argv[sys.argv.index("--hii") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(a), b) for a, b in WANT if a in hdr]
    out_rows = []
    for r in rows[2:]:
        d = {}
        for i, name in cols:
            v = r[i]
            if name == "kernel":
                v = v.split("(")[0].replace("void ", "")
            elif name.startswith("dram_") and name.endswith("_MB"):
                v = "%.3f" % (to_bytes(v, units[i]) / 1e6)
            elif name == "time_us":
                u = units[i]
                f = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                v = "%.2f" % (float(v) * f)
            d[name] = v
        out_rows.append(d)
    names = [b for _, b in cols]
    with open(out + ".csv", "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=names)
        w.writeheader()
        w.writerows(out_rows)
    agg = defaultdict(list)
    for d in out_rows:
        agg[d["kernel"]].append(d)
    traffic = {}
    print("| kernel | launches | time us | DRAM rd MB | DRAM wr MB | DRAM % | SM % | L2 hit % | regs | grid x block |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for k, v in agg.items():
        m = lambda key: sum(float(x[key]) for x in v) / len(v)
        print(f"| `{k}` | {len(v)} | {m('time_us'):.1f} | {m('dram_read_MB'):.1f} | {m('dram_write_MB'):.1f} | "
              f"{m('dram_pct'):.1f} | {m('sm_pct'):.1f} | {m('l2_hit_pct'):.1f} | {v[0]['regs']} | {v[0]['grid']} x {v[0]['block']} |")
        base = k.split("<")[0]
        traffic[base] = (m("dram_read_MB") + m("dram_write_MB")) * 1e6
    if hii:
        tp = Path(out).parent / "traffic.json"
        cur = json.loads(tp.read_text()) if tp.exists() else {}
        for k, b in traffic.items():
            cur[f"{k}@{hii}"] = b
        tp.write_text(json.dumps(cur, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
