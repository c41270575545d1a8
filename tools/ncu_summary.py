"""Summarise ncu reports (gpurun_out/*.ncu-rep, read here without a GPU) into small tracked files under profiles/:
    python tools/ncu_summary.py profiles/r2_ncu_summary.csv gpurun_out/a.ncu-rep[:label] gpurun_out/b.ncu-rep[:label] ...
One row per captured launch: duration, DRAM bytes read / written, L2->SM bytes, tensor-pipe and DRAM utilisation,
registers, shared memory -- the numbers DESIGN.md / bench.py's roofline.traffic quote."""
import csv
import io
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__warps_active.avg.per_cycle_active": "warps_active",
}
TENSOR = "sm__pipe_tensor_cycles_active"


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    if len(rd) < 3:
        return
    hdr, units = rd[0], rd[1]
    for vals in rd[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")][:90] if "Kernel Name" in hdr else ""}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                d[WANT[h]] = "%s %s" % (v, u) if u and WANT[h] in ("dram_read", "dram_write", "l2_to_sm_read", "duration_us") else v
            elif TENSOR in h and "pct_of_peak_sustained_elapsed" in h and "tensor_pct" not in d:
                d["tensor_pct"] = v
        yield d


def main():
    dst, srcs = sys.argv[1], sys.argv[2:]
    cols = ["label", "kernel"] + list(dict.fromkeys(WANT.values())) + ["tensor_pct"]
    with open(dst, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for s in srcs:
            path, _, label = s.partition(":")
            for d in rows_of(path):
                d["label"] = label or path.split("/")[-1]
                w.writerow({k: d.get(k, "") for k in cols})
    print(open(dst).read())


if __name__ == "__main__":
    main()
