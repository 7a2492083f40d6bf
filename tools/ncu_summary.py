#!/usr/bin/env python
"""Condense `ncu --set full` reports (gpurun_out/*.ncu-rep) into profiles/: one JSON + one markdown table per round.
usage: python tools/ncu_summary.py <tag> <rep> [<rep> ...]     (run where ncu is installed; no GPU needed)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dyn",
    "smsp__inst_executed.sum": "warp_instructions",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
}


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    out = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = {"report": os.path.basename(rep), "kernel": r[hdr.index("Kernel Name")][:160]}
            for k, name in KEYS.items():
                if k in hdr:
                    i = hdr.index(k)
                    try:
                        if name.startswith("dram_r") or name.startswith("dram_w"):
                            rec[name + "_bytes"] = to_bytes(r[i], units[i])
                        elif name == "time_us":
                            rec[name] = to_us(r[i], units[i])
                        else:
                            rec[name] = float(r[i].replace(",", ""))
                    except ValueError:
                        rec[name] = r[i]
            if "dram_read_bytes" in rec and "dram_write_bytes" in rec:
                rec["dram_bytes_per_launch"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
                if rec.get("time_us"):
                    rec["dram_gbs"] = rec["dram_bytes_per_launch"] / rec["time_us"] / 1e3
            out.append(rec)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md"), "w") as f:
        f.write(f"# ncu --set full summaries ({tag}); times are under the profiler (cold cache, serialised)\n\n")
        f.write("| kernel | grid x block | regs | time us | DRAM MB (r+w) | DRAM GB/s | tensor pipe % | issue active % | warps active % |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in out:
            f.write(f"| `{r['kernel'][:90]}` | {int(r.get('grid', 0))} x {int(r.get('block', 0))} | {int(r.get('regs', 0))} | {r.get('time_us', 0):.1f} | "
                    f"{r.get('dram_bytes_per_launch', 0) / 1e6:.1f} | {r.get('dram_gbs', 0):.0f} | {r.get('tensor_pipe_pct', 0):.1f} | "
                    f"{r.get('issue_active_pct', 0):.1f} | {r.get('warps_active_pct', 0):.1f} |\n")
    print(f"wrote profiles/{tag}_ncu_full.json/.md with {len(out)} kernels")


if __name__ == "__main__":
    main()
