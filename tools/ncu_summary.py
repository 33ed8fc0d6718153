"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into the per-kernel lines kept under profiles/.

    python tools/ncu_summary.py gpurun_out/r2a_targets.raw.csv profiles/r2a_ncu_summary.csv

One row per captured launch: duration, DRAM bytes (read + write), achieved DRAM GB/s and its fraction of
the measured copy peak (MEASURED_PEAKS.json), L2 / L1 hit rates, issue-slot utilisation, pipe utilisation
(fp64, fma, alu, lsu, xu), shared-memory bank conflicts, occupancy, registers, top stall reasons (warps stalled per issue-active cycle)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ("duration_us", "gpu__time_duration.sum", 1e-3),
    ("dram_read_bytes", "dram__bytes_read.sum", 1.0),
    ("dram_write_bytes", "dram__bytes_write.sum", 1.0),
    ("dram_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("lts_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct", 1.0),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct", 1.0),
    ("issue_active_pct", "sm__inst_issued.avg.pct_of_peak_sustained_active", 1.0),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("pipe_fp64_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1.0),
    ("pipe_fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1.0),
    ("pipe_alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1.0),
    ("pipe_lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", 1.0),
    ("pipe_xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", 1.0),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1.0),
    ("smem_bank_conflict_wavefronts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
    ("achieved_occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("registers", "launch__registers_per_thread", 1.0),
    ("smem_per_block", "launch__shared_mem_per_block_dynamic", 1.0),
    ("warp_insts", "smsp__inst_executed.sum", 1.0),
]
STALLS = "smsp__average_warp_latency_issue_stalled_"     # smsp__average_warps_issue_stalled_*_per_issue_active.ratio


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    col = {}
    for key, frag, _ in WANT:
        idx = [i for i, h in enumerate(hdr) if h == frag] or [i for i, h in enumerate(hdr) if h.endswith(frag)]
        col[key] = idx[0] if idx else None
    stall_cols = [(i, h) for i, h in enumerate(hdr) if re.search(r"average_warps_issue_stalled_.*_per_issue_active\.ratio$", h)]
    name_i, grid_i, block_i = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    out = [["launch", "kernel", "grid", "block"] + [k for k, _, _ in WANT] +
           ["dram_bytes", "dram_gbps", "dram_frac_of_measured_peak", "top_stalls"]]
    for n, r in enumerate(rows[2:]):
        vals = []
        for key, _, mul in WANT:
            i = col[key]
            v = num(r[i]) if i is not None else None
            if v is not None and key == "duration_us":
                u = units[i]
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            if v is not None and key in ("dram_read_bytes", "dram_write_bytes"):
                u = units[col[key]]
                v = v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            vals.append(v)
        d = dict(zip([k for k, _, _ in WANT], vals))
        dram = (d["dram_read_bytes"] or 0) + (d["dram_write_bytes"] or 0)
        gbps = dram / (d["duration_us"] * 1e-6) / 1e9 if d["duration_us"] else None
        st = sorted(((num(r[i]) or 0.0, re.sub(r".*issue_stalled_(.*)_per_issue_active\.ratio", r"\1", h)) for i, h in stall_cols),
                    reverse=True)[:3]
        name = re.sub(r"\(.*", "", r[name_i]).replace("void ", "")
        out.append([n, name, r[grid_i], r[block_i]] + ["" if v is None else f"{v:.6g}" for v in vals] +
                   [f"{dram:.6g}", "" if gbps is None else f"{gbps:.5g}", "" if gbps is None else f"{gbps / peak:.4f}",
                    " ".join(f"{nm}={v:.2f}" for v, nm in st)])
    with open(dst, "w", newline="") as f:
        csv.writer(f).writerows(out)
    for row in out:
        print(",".join(str(x) for x in row[:6]), row[-4:], sep=" | ")


if __name__ == "__main__":
    main()
