"""Condense `ncu -i X.ncu-rep --page raw --csv` output into the JSON kept under profiles/.

    ncu -i gpurun_out/prof_tile.ncu-rep --page raw --csv > gpurun_out/prof_tile.raw.csv
    python tools/ncu_summary.py gpurun_out/prof_tile.raw.csv [more.raw.csv ...] \
        --out profiles/ncu_full_summary_r1.json --traffic profiles/traffic.json

One entry per captured launch with the metrics DESIGN.md / profiles/README.md quote; --traffic
writes dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged per kernel family
(`bench.py` reports it as roofline.traffic).
"""
import argparse
import csv
import json
import re

KEEP = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct",
    "launch__grid_size",
    "launch__block_size",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
UNIT_TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def family(kernel: str) -> str:
    m = re.search(r"(k_[a-z0-9_]+)", kernel)
    name = m.group(1) if m else kernel
    return "k_advect4" if name.startswith("k_advect4") else name


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic")
    args = ap.parse_args()
    entries = []
    for path in args.csv:
        rows = list(csv.reader(open(path)))
        head, units = rows[0], rows[1]
        col = {}
        for i, h in enumerate(head):
            col.setdefault(h, i)
            col.setdefault(h.split(".", 2)[-1] if h.count(".") > 2 else h, i)  # "SECTION.Group.metric" -> metric
        for r in rows[2:]:
            e = {"kernel": r[col["Kernel Name"]], "source": path.split("/")[-1]}
            for k in KEEP:
                i = col.get(k)
                if i is None:
                    cand = [j for j, h in enumerate(head) if h.endswith(k)]
                    i = cand[0] if cand else None
                if i is None or r[i] == "":
                    continue
                try:
                    e[k] = [float(r[i].replace(",", "")), units[i]]
                except ValueError:
                    e[k] = [r[i], units[i]]
            entries.append(e)
    json.dump(entries, open(args.out, "w"), indent=1)
    print(f"{len(entries)} launches -> {args.out}")
    if args.traffic:
        acc = {}
        for e in entries:
            if "dram__bytes_read.sum" not in e or "dram__bytes_write.sum" not in e:
                continue
            b = sum(e[k][0] * UNIT_TO_BYTES.get(e[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            acc.setdefault(family(e["kernel"]), []).append(b)
        traffic = {k: int(sum(v) / len(v)) for k, v in acc.items()}
        json.dump(traffic, open(args.traffic, "w"), indent=1)
        print("traffic per launch:", traffic)
    for e in entries:
        t = e.get("gpu__time_duration.sum", ["?", ""])
        print(f"  {e['kernel'][:60]:60s} {t[0]} {t[1]}  ipc={e.get('smsp__issue_active.avg.per_cycle_active', ['?'])[0]}")


if __name__ == "__main__":
    main()
