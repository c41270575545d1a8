"""Summarise an ncu launch list (the `--metrics gpu__time_duration.sum --clock-control none --csv` pass of
/opt/skills/guides/B200_PROFILING.md over bench.py) into profiles/<round>_ncu_launch_list_summary.csv:
    python tools/launch_list_summary.py gpurun_out/launches.csv profiles/r2_ncu_launch_list_summary.csv [last_n]
Only libspconv's own kernels are kept (torch / cuDNN / NCCL kernels of the set-up are dropped); with `last_n` only
the last n such launches (one steady step).  Per-launch times under ncu are cold-cache and serialised: the SHARES
are what bench.py's per-kernel step shares must agree with, not the absolute times."""
import csv
import re
import sys

FOREIGN = ("at::", "cudnn", "cutlass", "nccl", "cublas", "elementwise", "sm80_", "sm90_", "sm100_")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    last = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(l for l in open(src, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = []
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "").replace("spc::<", "").replace("spc::", "")
        if any(o in name for o in FOREIGN):
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
        launches.append((name, us))
    if last:
        launches = launches[-last:]
    agg = {}
    for n, us in launches:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# %s ; %d launches of libspconv kernels kept%s; per-launch times are cold-cache/serialised: compare SHARES\n"
                % (" ".join(sys.argv[4:]) or "ncu --metrics gpu__time_duration.sum --clock-control none", len(launches),
                   " (the last step)" if last else ""))
        f.write("kernel,launches,total_us,share_pct,avg_us\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.2f,%.2f\n" % (n, c, t, 100 * t / tot, t / c))
    print(open(dst).read())


if __name__ == "__main__":
    main()
