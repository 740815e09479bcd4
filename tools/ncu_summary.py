"""Compact per-kernel table from `ncu -i X.ncu-rep --page raw --csv` output.   python tools/ncu_summary.py raw.csv > summary.md"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("dram__bytes_read.sum.per_second", "DRAM rd"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for n, r in enumerate(data):
        name = r[ix["Kernel Name"]]
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*$", "", name).replace("nxdi::", "")
        if name.startswith("at::"):
            name = "(torch) " + name.split("<")[0][4:]
        cells = []
        for key, _ in COLS:
            if key not in ix:
                cells.append("-")
                continue
            v, u = r[ix[key]], units[ix[key]]
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.1f}" if abs(f) < 1000 and f != int(f) else f"{int(f)}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip().replace("%", "").strip() if u not in ("", "%") else v)
        print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
