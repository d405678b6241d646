#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list of scripts/one_pass.py into the per-kernel table committed under profiles/ and
into profiles/traffic.json (DRAM bytes of the dominant op, read by bench.py).

    python scripts/summarize_ncu.py gpurun_out/launches.csv profiles/r01_launches_per_kernel.txt [per_rep]

per_rep = launches of one repetition of scripts/one_pass.py (9 SGA forward + 16 SGA backward
+ 6 LGA2 = 31); the LAST repetition in the list is summarised.
"""
import collections
import csv
import json
import os
import re
import sys

F = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
T = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    per_rep = int(sys.argv[3]) if len(sys.argv) > 3 else 31
    voxels = int(sys.argv[4]) if len(sys.argv) > 4 else 32 * 192 * 240 * 624
    what = sys.argv[5] if len(sys.argv) > 5 else "`python scripts/one_pass.py`: the %d launches of its last repetition, one 920M-voxel sample" % per_rep
    write_traffic = len(sys.argv) <= 5
    lines = [l for l in open(src) if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        k = (int(r["ID"]), r["Kernel Name"])
        per.setdefault(k, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * \
            (T.get(r["Metric Unit"], None) if r["Metric Name"].startswith("gpu__time") else F[r["Metric Unit"]])
    items = list(per.items())
    last = items[-per_rep:]                 # the last repetition (warm)
    out = ["# ncu launch list of " + what,
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
           "%-4s %-58s %9s %9s %9s %8s" % ("id", "kernel", "ms", "rd GB", "wr GB", "GB/s")]
    tot = collections.Counter()
    sga = {"ms": 0.0, "bytes": 0.0}
    agg = collections.OrderedDict()
    for (i, name), m in last:
        short = re.sub(r"\(.*", "", name).replace("void ganet::", "")
        ms = m["gpu__time_duration.sum"]
        rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        out.append("%-4d %-58s %9.3f %9.2f %9.2f %8.0f" % (i, short[:58], ms, rd / 1e9, wr / 1e9, (rd + wr) / ms / 1e6))
        a = agg.setdefault(short, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += rd + wr
        if "lga" not in short:
            sga["ms"] += ms; sga["bytes"] += rd + wr
        tot["ms"] += ms
    out.append("")
    out.append("%-58s %6s %9s %7s %9s" % ("kernel (aggregated)", "calls", "ms", "share", "GB"))
    for k, (c, ms, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-58s %6d %9.3f %6.1f%% %9.2f" % (k[:58], c, ms, 100 * ms / tot["ms"], b / 1e9))
    out.append("total %.2f ms; SGA launches %.2f ms, %.1f GB DRAM = %.1f B/voxel (algorithmic 23.25)"
               % (tot["ms"], sga["ms"], sga["bytes"] / 1e9, sga["bytes"] / voxels))
    open(dst, "w").write("\n".join(out) + "\n")
    if not write_traffic:
        print("\n".join(out[-8:]))
        return
    tj = {"op": "SGA forward+backward, all launches of one call", "dram_bytes_per_voxel": sga["bytes"] / voxels,
          "dram_bytes_per_call": sga["bytes"], "voxels_per_call": voxels, "source": os.path.basename(dst)}
    json.dump(tj, open(os.path.join(os.path.dirname(dst) or ".", "traffic.json"), "w"), indent=1)
    print("\n".join(out[-8:]))


if __name__ == "__main__":
    main()
