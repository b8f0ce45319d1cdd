"""Summarise an ncu report (.ncu-rep brought back from the GPU box) into a small CSV/markdown pair.

    python profiles/summarize_ncu.py gpurun_out/prof_r1a.ncu-rep profiles/r1a_ncu_summary
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [hdr.index("Kernel Name")] + [hdr.index(w) for w in WANT if w in hdr]
    names = ["kernel"] + [f"{hdr[i]} [{units[i]}]" for i in cols[1:]]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(names)
        for r in data:
            w.writerow([r[i] for i in cols])
    with open(out + ".md", "w") as f:
        f.write(f"ncu --set full --clock-control none summary of `{rep}` (one row per profiled launch)\n\n")
        f.write("| kernel | ms | DRAM rd GB | DRAM wr GB | DRAM % | L1tex % | LTS % | L1 hit % | warps act % | regs |\n|---|---|---|---|---|---|---|---|---|---|\n")
        g = lambda r, k: r[hdr.index(k)] if k in hdr else ""  # noqa: E731
        for r in data:
            f.write("| " + " | ".join([r[hdr.index("Kernel Name")][:60].replace("|", "/"), g(r, WANT[0]), g(r, WANT[1]), g(r, WANT[2]),
                                       g(r, WANT[3])[:5], g(r, WANT[4])[:5], g(r, WANT[5])[:5], g(r, WANT[6])[:5], g(r, WANT[7])[:5],
                                       g(r, WANT[8])]) + " |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
