"""Per-kernel DRAM traffic and duration from an .ncu-rep (one `ncu --set full` capture) -> JSON.
usage: python tools/ncu_traffic.py report.ncu-rep out.json
Keys: "<kernel name> grid(<grid>)"; values: launches, mean duration (us), mean dram read / write bytes per launch."""
import csv
import json
import re
import subprocess
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    agg = {}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "").replace("caae::", "")
        key = f"{name} grid{r[col['Grid Size']]}"
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        du = units[col["gpu__time_duration.sum"]]
        dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(du, 1.0)
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        a = agg.setdefault(key, {"launches": 0, "duration_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
        a["launches"] += 1; a["duration_us"] += dur_us; a["dram_read_bytes"] += rd; a["dram_write_bytes"] += wr
    for a in agg.values():
        n = a["launches"]
        for k in ("duration_us", "dram_read_bytes", "dram_write_bytes"):
            a[k] = a[k] / n
        a["dram_bytes"] = a["dram_read_bytes"] + a["dram_write_bytes"]
    json.dump({"source": rep.split("/")[-1], "note": "ncu --set full --clock-control none, per-launch means (cold cache, serialised)",
               "kernels": agg}, open(out, "w"), indent=1, sort_keys=True)
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["duration_us"]):
        print(f"{a['duration_us']:9.1f} us  {a['dram_bytes'] / 1e6:9.2f} MB  x{a['launches']:2d}  {k}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
