"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/prof_x.ncu-rep [...]   -> markdown table rows on stdout
  python tools/ncu_summary.py list gpurun_out/launches.csv           -> per-kernel share table on stdout
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
]


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


def rep(paths):
    print("| report | kernel | " + " | ".join(m[1] for m in METRICS) + " |")
    print("|---|---|" + "---|" * len(METRICS))
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            cells = []
            for key, _ in METRICS:
                i = col.get(key)
                cells.append("-" if i is None else f"{r[i]} {units[i]}".strip())
            print(f"| {p.split('/')[-1]} | `{short(r[col['Kernel Name']])}` | " + " | ".join(cells) + " |")


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        k = short(row["Kernel Name"])[:80]
        tot[k] += float(row["Metric Value"].replace(",", ""))
        cnt[k] += 1
    s = sum(tot.values())
    print(f"launches: {sum(cnt.values())}   total device time: {s / 1e3:.1f} us\n")
    print("| kernel | launches | total us | share % | avg us |")
    print("|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        print(f"| `{k}` | {cnt[k]} | {v / 1e3:.1f} | {100 * v / s:.1f} | {v / cnt[k] / 1e3:.1f} |")


if __name__ == "__main__":
    {"rep": rep, "list": lambda a: launches(a[0])}[sys.argv[1]](sys.argv[2:])
