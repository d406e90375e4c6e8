#!/usr/bin/env python
"""Launch list of ONE benchmark step with DRAM bytes, from an ncu CSV log.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:'(^|[ :])k_[a-z]' -c 600 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline
    python tools/launches_with_traffic.py gpurun_out/launches.csv profiles/r02_launches.txt profiles/ncu_traffic.json

The run has 4 identical steps (3 warm-up + 1 timed); the last quarter of the launches is the timed step.
"""
import collections
import csv
import json
import sys

MAIN = ["k_ir_fft", "k_x_fft", "k_cmac", "k_cmac_static", "k_ifft_ola", "k_amb_partial", "k_mix"]


def main(src, txt, js):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, mi, vi, ui, ii = (hdr.index(n) for n in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(int(r[ii]), {"k": r[ki].split("(")[0].replace("void ", "")})
        v = float(r[vi].replace(",", ""))
        if r[mi].startswith("gpu__time"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)  # -> us
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1.0)
        d[r[mi]] = v
    ids = list(launches)
    step = ids[3 * len(ids) // 4:]
    tot = collections.OrderedDict()
    for i in step:
        d = launches[i]
        t = tot.setdefault(d["k"], [0, 0.0, 0.0, 0.0])
        t[0] += 1
        t[1] += d["gpu__time_duration.sum"]
        t[2] += d["dram__bytes_read.sum"]
        t[3] += d["dram__bytes_write.sum"]
    T = sum(t[1] for t in tot.values())
    out = ["ncu launch list of one benchmark step (C5, 64 scenes): %d launches, %.2f ms under ncu (per-launch times are cold-cache"
           % (len(step), T / 1e3),
           "and serialised: compare SHARES with bench.py's CUDA-event times, not absolutes), DRAM %.1f GB read + %.1f GB written"
           % (sum(t[2] for t in tot.values()) / 1e9, sum(t[3] for t in tot.values()) / 1e9), "",
           "%-24s %8s %10s %7s %10s %10s %8s" % ("kernel", "launches", "total us", "share", "read GB", "write GB", "TB/s")]
    for k, t in sorted(tot.items(), key=lambda x: -x[1][1]):
        out.append("%-24s %8d %10.1f %7.3f %10.2f %10.2f %8.2f" % (k, t[0], t[1], t[1] / T, t[2] / 1e9, t[3] / 1e9,
                                                                 (t[2] + t[3]) / t[1] / 1e6 if t[1] > 0 else 0.0))
    out += ["", "every launch of the step, in order:"]
    for i in step:
        d = launches[i]
        out.append("  %-24s %9.1f us  read %7.3f GB  write %7.3f GB" % (d["k"], d["gpu__time_duration.sum"],
                                                                       d["dram__bytes_read.sum"] / 1e9, d["dram__bytes_write.sum"] / 1e9))
    open(txt, "w").write("\n".join(out) + "\n")
    if js:
        old = json.load(open(js))
        new = {"note": old.get("note", ""), "round": 2}
        for k in MAIN:
            if k in tot:
                n, us, rd, wr = tot[k]
                new[k] = {"launches_in_step": n, "dram_read_bytes": rd / n, "dram_write_bytes": wr / n,
                          "duration_us_under_ncu": us / n, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr}
        json.dump(new, open(js, "w"), indent=1)
    print("\n".join(out[:16]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
