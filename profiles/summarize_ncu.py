"""Turn an ncu report (--set full) / launch list (--metrics gpu__time_duration.sum --csv) into the
markdown / csv summaries committed under profiles/.  Runs on the CPU box:
  python profiles/summarize_ncu.py full gpurun_out/prof_r1e.ncu-rep > profiles/r1e_ncu_full_one_layer.md
  python profiles/summarize_ncu.py launches gpurun_out/r1e_launches.csv > profiles/r1e_ncu_launch_shares.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    ("time us", "gpu__time_duration.sum"),
    ("dram rd MB", "dram__bytes_read.sum"),
    ("dram wr MB", "dram__bytes_write.sum"),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("xu (ex2) pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("smem lsu wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("smem tensor wavefronts %", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("sm MHz", "gpc__cycles_elapsed.max.per_second"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(k for k, _ in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for n, r in enumerate(rows[2:]):
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        vals = []
        for label, key in KEYS:
            v = r[ix[key]] if key in ix else ""
            try:
                f = float(v.replace(",", ""))
                u = units[ix[key]]
                if key.startswith("dram__bytes"):
                    f = f * {"Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
                if key == "gpu__time_duration.sum":
                    f = f * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(u, 1.0)
                if key.endswith("per_second"):
                    f = f * {"Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6}.get(u, 1.0)
                v = f"{f:.1f}"
            except ValueError:
                pass
            vals.append(v)
        print(f"| {n} | `{name}` | " + " | ".join(vals) + " |")


def launches(path):
    rows, hdr = [], None
    for line in open(path):
        if line.startswith('"ID"'):
            hdr = next(csv.reader([line]))
        elif line.startswith('"') and hdr:
            rows.append(dict(zip(hdr, next(csv.reader([line])))))
    agg, tot = collections.OrderedDict(), 0.0
    for d in rows:
        t = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
        name = d["Kernel Name"].split("(")[0].replace("void ", "")[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print("| kernel | launches | total us | avg us | share % |")
    print("|---|---|---|---|---|")
    for k, (n, t) in agg.items():
        print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f} |")
    print(f"| **all** | {sum(n for n, _ in agg.values())} | {tot:.1f} | | 100 |")


if __name__ == "__main__":
    {"full": full, "launches": launches}[sys.argv[1]](sys.argv[2])
