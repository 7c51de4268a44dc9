"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list into per-kernel totals.
usage: python tools/ncu_launch_summary.py gpurun_out/launches.csv profiles/r01_ncu_launches_x.txt "<command line>" """
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    cmd = sys.argv[3] if len(sys.argv) > 3 else ""
    tot = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    with open(src, newline="") as f:
        rows = [r for r in csv.reader(l for l in f if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
    for r in rows[1:]:
        if len(r) <= iv or r[iv] in ("", "n/a"):
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
        t = tot[name]
        t[0] += 1
        t[1] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
        n += 1
    total = sum(v[1] for v in tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  {cmd}\n")
        f.write("# per-kernel totals over the captured launches (cold-cache, serialised: compare SHARES, not absolutes)\n")
        f.write(f"# {n} launches captured, {total:.1f} ms\n")
        for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:70s} x{c:5d} {ms:10.3f} ms {100 * ms / total:6.2f} %\n")


if __name__ == "__main__":
    main()
