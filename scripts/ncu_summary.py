"""Key metrics of every kernel in one or more .ncu-rep files (ncu -i ... --page raw --csv), as a small CSV for profiles/.
Usage: python scripts/ncu_summary.py out.csv a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def main(out, reps):
    rows_out = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        h, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"report": rep.split("/")[-1], "kernel": r[h.index("Kernel Name")][:100]}
            for w in WANT:
                if w in h:
                    d[w + " [" + units[h.index(w)] + "]"] = r[h.index(w)]
            rows_out.append(d)
    keys = []
    for d in rows_out:
        for k in d:
            if k not in keys:
                keys.append(k)
    with open(out, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(rows_out)
    for d in rows_out:
        print({k.split(" [")[0].split(".")[0][-28:]: v for k, v in d.items()})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
