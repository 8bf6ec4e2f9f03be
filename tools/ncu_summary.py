#!/usr/bin/env python3
"""Condense an .ncu-rep (ncu --set full) into the few lines the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("sm__warps_active.avg.per_cycle_active", "resident warps/SM (avg)"),
    ("sm__maximum_warps_avg_per_active_cycle", "resident warps/SM (theoretical)"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank-conflict wavefronts"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 (lts) throughput %"),
    ("lts__t_bytes.sum", "L2 bytes (lts__t_bytes)"),
    ("lts__t_sectors_op_read.sum", "L2 read sectors"),
    ("lts__t_sectors_op_write.sum", "L2 write sectors"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->L1 read bytes"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "L1->L2 write bytes"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/tex throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "LSU writeback %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_xu.sum", "XU (MUFU) instructions"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")])
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {label:42s} {vals[i]} {units[i]}")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio"):
                st.append((float(vals[i].replace(",", "")), h[len(STALLS):-len("_per_issue_active.ratio")]))
        print("  stalls per issued instruction:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
        print()


if __name__ == "__main__":
    main()
