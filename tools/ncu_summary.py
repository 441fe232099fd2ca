"""Summarise `ncu -i report --page raw --csv` (one row per profiled launch) into a compact table.
python tools/ncu_summary.py raw.csv > profiles/...txt"""
import csv
import sys

KEEP = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smemKB"),
        ("dram__bytes_read.sum", "dramR_MB"), ("dram__bytes_write.sum", "dramW_MB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu%"),
        ("sm__cycles_elapsed.avg.per_second", "sm_GHz"), ("lts__t_sector_hit_rate.pct", "L2hit%")]


def to_unit(v, unit, want):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    scale = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}
    if want == "us" and unit in scale:
        return f"{x * scale[unit]:.1f}"
    if want.endswith("_MB"):
        f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        return f"{x * f:.1f}"
    if want == "smemKB":
        f = {"byte/block": 1e-3, "Kbyte/block": 1.0}.get(unit, 1.0)
        return f"{x * f:.0f}"
    if want == "sm_GHz":
        f = {"hz": 1e-9, "Hz": 1e-9, "Mhz": 1e-3, "MHz": 1e-3, "Ghz": 1.0, "GHz": 1.0, "cycle/nsecond": 1.0, "cycle/second": 1e-9}.get(unit, 1.0)
        return f"{x * f:.2f}"
    if want.endswith("%"):
        return f"{x:.1f}"
    return f"{x:.0f}"


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx["Kernel Name"]
    cols = [(m, lab) for m, lab in KEEP if m in idx]
    print("kernel".ljust(34) + " ".join(lab.rjust(11) for _, lab in cols))
    for r in rows[2:]:
        if len(r) <= name_i:
            continue
        nm = r[name_i].split("(")[0].replace("fd::<unnamed>::", "").replace("void ", "")[-33:]
        print(nm.ljust(34) + " ".join(to_unit(r[idx[m]], units[idx[m]], lab).rjust(11) for m, lab in cols))


if __name__ == "__main__":
    main(sys.argv[1])
