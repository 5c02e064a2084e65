"""Summaries of the round-2 ncu captures (scripts/ncu_r2.sh), written next to the raw CSVs:
    python profiles/summarize_ncu.py conv  <conv_ncu csv> <launches per forward>   # per-launch table of the LAST forward
    python profiles/summarize_ncu.py full  <*_raw.csv> ...                         # key metrics of --set full captures
"""
import collections
import csv
import re
import sys


def read_rows(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def conv(path, per_forward):
    rows = read_rows(path)
    by = collections.OrderedDict()
    for r in rows:
        d = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        try:
            d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            d[r["Metric Name"]] = None
    launches = list(by.values())[-per_forward:]
    T = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    tot_ns = sum(l["gpu__time_duration.sum"] for l in launches)
    dram = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in launches)
    l2sm = sum(l["l1tex__m_xbar2l1tex_read_bytes.sum"] or 0 for l in launches)
    tw = sum(l[T] * l["gpu__time_duration.sum"] for l in launches) / tot_ns
    print(f"# {path}: last forward = {len(launches)} launches, {tot_ns / 1e6:.3f} ms under ncu (cold cache, serialised)")
    print(f"# time-weighted tensor pipe active {tw:.1f} %; DRAM traffic {dram / 1e9:.3f} GB = {dram / len(launches) / 1e6:.1f} MB per "
          f"launch; L2->SM (xbar2l1tex) {l2sm / 1e9:.2f} GB")
    print(f"{'#':>3s} {'kernel':28s} {'grid':>6s} {'us':>8s} {'tensor%':>8s} {'dramR MB':>9s} {'dramW MB':>9s} {'L2->SM MB':>10s} {'DRAM GB/s':>10s}")
    for i, l in enumerate(launches):
        k = re.search(r"(conv_\w+<[^>]*>|conv3x3_c64_halo_kernel|conv1x1_b2b_kernel<[^>]*>)", l["name"])
        ns = l["gpu__time_duration.sum"]
        rd, wr = l["dram__bytes_read.sum"], l["dram__bytes_write.sum"]
        print(f"{i:3d} {k.group(1) if k else l['name'][:28]:28s} {l['grid'].split(',')[0].strip('('):>6s} {ns / 1e3:8.1f} {l[T]:8.1f} "
              f"{rd / 1e6:9.1f} {wr / 1e6:9.1f} {(l['l1tex__m_xbar2l1tex_read_bytes.sum'] or 0) / 1e6:10.1f} {(rd + wr) / ns:10.0f}")


KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__cycles_active.avg", "active cycles/SMSP"),
]


def full(paths):
    for path in paths:
        rows = read_rows(path)
        if not rows:
            continue
        hdr = rows[0]  # units row in --page raw --csv
        print(f"== {path}")
        for r in rows[1:]:
            name = re.sub(r"\(.*", "", r.get("Kernel Name", ""))
            print(f"-- launch {r.get('ID')}: {name}  grid {r.get('Grid Size')} block {r.get('Block Size')}")
            for key, label in KEYS:
                if key in r and r[key] not in ("", None):
                    print(f"   {label:24s} {r[key]:>16s} {hdr.get(key, '')}")


if __name__ == "__main__":
    try:
        if sys.argv[1] == "conv":
            conv(sys.argv[2], int(sys.argv[3]))
        else:
            full(sys.argv[2:])
    except BrokenPipeError:
        pass
