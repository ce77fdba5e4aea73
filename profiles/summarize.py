#!/usr/bin/env python
"""Turn an .ncu-rep (captured on the B200 with `ncu --set full --import-source on`) into the text summary
committed next to it:  python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>.txt
Runs in the CPU-only build container (ncu can read reports without a GPU)."""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__cycles_elapsed.max",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        print(f"== launch {n}: {r[hdr.index('Kernel Name')]}")
        for k in KEYS:
            if k in hdr:
                print(f"   {k:78s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                  if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        for v, h in sorted(stalls, reverse=True)[:5]:
            print(f"   stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:8.2f} warps/issue")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    hdr, cur, agg = None, "", {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 6 and r[0] == "Line No":
            hdr = r
            i_s = hdr.index("# Samples")
        elif hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
            try:
                key = (cur, int(r[0]), r[1].strip()[:110])
                agg[key] = agg.get(key, 0) + int(r[i_s])
            except ValueError:
                pass
    tot = sum(agg.values()) or 1
    print(f"== warp-stall samples by source line (all launches, {tot} samples)")
    for (f, l, src), v in sorted(agg.items(), key=lambda x: -x[1])[:16]:
        print(f"   {100 * v / tot:5.1f}%  {f}:{l:<4d} {src}")


if __name__ == "__main__":
    main(sys.argv[1])
