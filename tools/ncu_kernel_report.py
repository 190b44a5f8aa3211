"""Compact per-kernel digest of an `ncu --page raw --csv` export: duration, occupancy, pipe utilisation, stall mix,
L1TEX / L2 sector counts.  Usage: python tools/ncu_kernel_report.py raw.csv [kernel-substring ...]"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"), ("sm__cycles_elapsed.max", "cycles"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("smsp__inst_executed.sum", "warp insts"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "lsu pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "l1 wavefronts"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1 wavefronts shared"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1 ld sectors"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum", "l1 ld hit"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1 red sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "red insts"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "lts red sectors"), ("lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed", "lts red %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "lts read sectors"), ("lts__t_sectors_srcunit_tex_op_read.avg.pct_of_peak_sustained_elapsed", "lts read %"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "lts write sectors"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    pats = sys.argv[2:]
    for r in rows[2:]:
        nm = r[ci["Kernel Name"]]
        if pats and not any(p in nm for p in pats):
            continue
        print("=====", nm[:100])
        for k, label in KEYS:
            if k in ci and r[ci[k]] not in ("", "n/a"):
                print(f"  {label:24s} {r[ci[k]]} {units[ci[k]]}")
        st = [(float(r[i].replace(",", "")), h[len(STALL):-len("_per_issue_active.ratio")]) for h, i in ci.items()
              if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        print("  stalls (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))


if __name__ == "__main__":
    main()
