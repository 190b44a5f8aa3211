"""Turn ncu CSV exports into the small text summaries committed under profiles/.
  launches: python tools/ncu_summary.py launches gpurun_out/launches.csv
  raw:      ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw raw.csv
"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg"]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0][:70], [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>10s} {'avg_us':>9s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} {n:5d} {t / 1e3:10.1f} {t / n / 1e3:9.2f} {t / tot * 100:5.1f}%")
    print(f"{'TOTAL':70s} {sum(a[0] for a in agg.values()):5d} {tot / 1e3:10.1f}")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:60]
        if name in seen:
            continue
        seen.add(name)
        print("==", name)
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
              for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        print("  top stalls (warps per issue):", ", ".join(f"{k}={v:.2f}" for k, v in sorted(st, key=lambda x: -x[1])[:6]))


def traffic(path):
    """JSON for bench.py (profiles/rNN_traffic.json): per kernel and launch, the DRAM bytes and the L1TEX / L2 sector counts
    that its L2 / L1TEX / atomic fractions are computed from."""
    import json
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    names = {"field_fwd_kernel<1, 1>": "usl_field_fwd", "field_fwd_kernel<(bool)1, (bool)1>": "usl_field_fwd", "field_bwd2_kernel": "usl_field_bwd",
             "sdf_query_grid_kernel": "usl_sdf_query_grid", "field_fwd_kernel<0, 0>": "usl_field_fwd_render",
             "field_fwd_kernel<(bool)0, (bool)0>": "usl_field_fwd_render", "field_fwd_kernel<1, 0>": "usl_field_fwd_tracking",
             "field_fwd_kernel<(bool)1, (bool)0>": "usl_field_fwd_tracking"}
    get = lambda r, k: float(r[ci[k]].replace(",", "")) if k in ci and r[ci[k]] not in ("", "n/a") else 0.0
    out = {"_source": f"ncu --set full capture, {path} (per launch)"}
    for r in rows[2:]:
        kn = r[ci["Kernel Name"]]
        key = next((v for k, v in names.items() if k in kn), None)
        if key is None or key in out:
            continue
        unit = rows[1][ci["dram__bytes_read.sum"]] if "dram__bytes_read.sum" in ci else "byte"
        mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        tunit = rows[1][ci["gpu__time_duration.sum"]] if "gpu__time_duration.sum" in ci else "us"
        tmul = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(tunit, 1.0)
        out[key] = {
            "duration_us_under_ncu": get(r, "gpu__time_duration.sum") * tmul,
            "dram_bytes": (get(r, "dram__bytes_read.sum") + get(r, "dram__bytes_write.sum")) * mul,
            "lts_sectors_read": get(r, "lts__t_sectors_srcunit_tex_op_read.sum"), "lts_sectors_write": get(r, "lts__t_sectors_srcunit_tex_op_write.sum"),
            "lts_sectors_red": get(r, "lts__t_sectors_srcunit_tex_op_red.sum"),
            "l1_sector_lookups_ld": get(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
            "l1_sectors_red": get(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum"),
            "warp_instructions": get(r, "smsp__inst_executed.sum"), "registers": get(r, "launch__registers_per_thread"),
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](sys.argv[2])
