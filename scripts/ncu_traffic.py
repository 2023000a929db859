"""Summarise .ncu-rep captures into (a) a per-launch text table for profiles/ and (b) profiles/r02_dram_traffic.json,
the table bench.py reads `roofline.traffic` from (DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum).

    python scripts/ncu_traffic.py <config key> <file.ncu-rep | raw-page .csv | .csv.gz> [<out.txt>]

Runs in the build container (ncu -i needs no GPU)."""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = {"gpu__time_duration.sum": "us", "dram__bytes_read.sum": "rd_MB", "dram__bytes_write.sum": "wr_MB",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor%",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram%", "launch__grid_size": "grid",
        "launch__registers_per_thread": "regs", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm%",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "lsu_smem_wavefronts",
        "lts__t_sector_hit_rate.pct": "l2hit%"}


def to_float(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    return x * scale.get(unit, 1.0)


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("avc::", "").replace("void ", "")
    return name


def main(key, path, out_txt=None):
    if path.endswith(".gz"):
        import gzip
        raw = gzip.open(path, "rt").read()
    elif path.endswith(".csv"):
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kname = col["Kernel Name"]
    lines = []
    agg = {}
    fmt = "{:<58} {:>9} {:>9} {:>9} {:>8} {:>7} {:>6} {:>5} {:>7}"
    lines.append(fmt.format("kernel", "us", "rd_MB", "wr_MB", "tensor%", "dram%", "grid", "regs", "l2hit%"))
    for r in rows[2:]:
        if len(r) <= kname:
            continue
        get = lambda m: to_float(r[col[m]], units[col[m]]) if m in col else None
        name = short(r[kname])
        us, rd, wr = get("gpu__time_duration.sum"), get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
        f = lambda v, d=1: "-" if v is None else f"{v:.{d}f}"
        lines.append(fmt.format(name[:58], f(us), f(rd), f(wr),
                                f(get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
                                f(get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
                                f(get("launch__grid_size"), 0), f(get("launch__registers_per_thread"), 0),
                                f(get("lts__t_sector_hit_rate.pct"))))
        base = name.split("<")[0]
        a = agg.setdefault(base, dict(launches=0, bytes=0.0, us=0.0))
        a["launches"] += 1
        a["bytes"] += ((rd or 0.0) + (wr or 0.0)) * 1e6
        a["us"] += us or 0.0
    text = "\n".join(lines)
    print(text)
    if out_txt:
        with open(out_txt, "w") as f:
            f.write(f"# {key}: ncu --set full --clock-control none ({os.path.basename(path)}); per-launch times are "
                    "serialised / cold-cache: compare shares, not absolutes\n" + text + "\n")
    tpath = os.path.join(ROOT, "profiles", "r02_dram_traffic.json")
    table = {}
    if os.path.exists(tpath):
        with open(tpath) as f:
            table = json.load(f)
    table[key] = {k: dict(bytes_per_launch=v["bytes"] / v["launches"], launches=v["launches"], us_total=v["us"],
                          source=f"profiles/{os.path.basename(out_txt)}" if out_txt else os.path.basename(path))
                  for k, v in agg.items()}
    with open(tpath, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
