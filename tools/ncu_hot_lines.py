"""Hot source lines of one kernel from an .ncu-rep captured with --import-source on (needs -lineinfo).
usage: python tools/ncu_hot_lines.py report.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys


def main(rep, kernel, top=40):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv", "--kernel-name",
                          f"regex:{kernel}", "-c", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue
        try:
            ie = int(r[hdr.index("Instructions Executed")]); sm = int(r[hdr.index("# Samples")])
            th = int(r[hdr.index("Thread Instructions Executed")])
        except ValueError:
            continue
        data.append((ie, sm, th, cur_file, r[0], r[1]))
    tot = sum(d[0] for d in data) or 1; tots = sum(d[1] for d in data) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for ie, sm, th, f, ln, src in sorted(data, reverse=True)[:top]:
        print(f"{100 * ie / tot:5.1f}% inst {100 * sm / tots:5.1f}% smp  act {th / max(ie, 1):4.1f}  {f}:{ln}: {src.strip()[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
