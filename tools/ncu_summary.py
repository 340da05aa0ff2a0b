#!/usr/bin/env python
"""Turn an `ncu --set full` report (gpurun_out/*.ncu-rep) into (1) a text summary of the metrics DESIGN.md quotes, under
profiles/, and (2) an entry of profiles/ncu_traffic.json -- the per-launch DRAM bytes bench.py reports as
roofline.traffic, keyed by kernel + workload@gpus and pinned to the sha1 of the kernel's source file so that a stale
capture is never reported as current.

    python tools/ncu_summary.py gpurun_out/r02_allpairs_c3.ncu-rep allpairs_fast_kernel c3@1 rust_exp_b200/csrc/nb_allpairs.cu profiles/r02_ncu_full_allpairs_fast_c3.txt
"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "nvlrx__bytes.sum", "nvltx__bytes.sum", "lts__t_sectors_srcunit_tex_aperture_peer", "smsp__thread_inst_executed_per_inst_executed.ratio")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, kernel, key, src, out_txt = sys.argv[1:6]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines, traffic = [], {}
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        if kernel not in name:
            continue
        lines.append(f"== {name}  grid={vals[hdr.index('Grid Size')]} block={vals[hdr.index('Block Size')]}")
        for h, u, v in zip(hdr, units, vals):
            if any(h.startswith(k) for k in KEEP):
                lines.append(f"   {h} = {v} {u}")
            if h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                traffic[h] = float(v) * UNIT.get(u, 1.0)
        break
    if not lines:
        raise SystemExit(f"no kernel matching {kernel!r} in {rep}")
    open(os.path.join(ROOT, out_txt), "w").write("\n".join(lines) + "\n")
    tab_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tab = json.load(open(tab_path)) if os.path.exists(tab_path) else {}
    tab.setdefault(kernel, {})[key] = {
        "dram_bytes_read": traffic.get("dram__bytes_read.sum"), "dram_bytes_write": traffic.get("dram__bytes_write.sum"),
        "capture": out_txt, "source_file": src, "source_sha1": hashlib.sha1(open(os.path.join(ROOT, src), "rb").read()).hexdigest()}
    json.dump(tab, open(tab_path, "w"), indent=1, sort_keys=True)
    print(f"wrote {out_txt} and {tab_path}[{kernel}][{key}]")


if __name__ == "__main__":
    main()
