"""gpurun_out/r02_*_raw.csv (scratch/profile_r02.sh) -> profiles/r02_step_kernels_ncu_summary.csv, r02_launch_shares.csv."""
import csv, collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"\((int|bool)\)", "", name)
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("spb::", "")
    return name.strip()


def summary(paths, out):
    rows = []
    for path in paths:
        if not os.path.exists(path):
            continue
        rd = list(csv.reader(l for l in open(path) if l.startswith('"')))
        hdr = rd[0]
        ci = {k: i for i, k in enumerate(hdr)}
        seen = set()
        for r in rd[2:]:
            key = (short(r[ci["Kernel Name"]]), r[ci["Grid Size"]])
            if key in seen:
                continue
            seen.add(key)
            f = lambda k: float(r[ci[k]]) if r[ci[k]] not in ("", "n/a") else float("nan")
            ms = f("gpu__time_duration.sum")
            rd_gb, wr_gb = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
            rows.append([key[0], key[1], r[ci["Block Size"]], int(f("launch__registers_per_thread")), "%.4f" % ms,
                         "%.4f" % rd_gb, "%.4f" % wr_gb, "%.0f" % ((rd_gb + wr_gb) / ms * 1e3),
                         "%.3f" % f("sm__cycles_elapsed.avg.per_second"),
                         "%.1f" % f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         "%.1f" % f("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                         "%.1f" % f("sm__warps_active.avg.pct_of_peak_sustained_active"),
                         "%.1f" % f("smsp__thread_inst_executed_per_inst_executed.ratio"),
                         "%.1f" % f("lts__t_sector_hit_rate.pct")])
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "grid", "block", "regs", "duration_ms", "dram_read_GB", "dram_write_GB", "dram_GBps",
                    "sm_clock_GHz", "tensor_pipe_active_pct", "sm_throughput_pct", "warps_active_pct",
                    "threads_per_inst", "l2_hit_pct"])
        w.writerows(rows)
    return rows


def shares(path, out):
    rd = list(csv.reader(l for l in open(path) if l.startswith('"')))
    ci = {k: i for i, k in enumerate(rd[0])}
    tot = collections.OrderedDict()
    for r in rd[1:]:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = short(r[ci["Kernel Name"]])
        v = float(r[ci["Metric Value"]]) / 1e6
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    s = sum(v for _, v in tot.values())
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "launches", "total_ms", "share"])
        for k, (n, v) in sorted(tot.items(), key=lambda x: -x[1][1]):
            w.writerow([k, n, "%.3f" % v, "%.4f" % (v / s)])


if __name__ == "__main__":
    rows = summary([os.path.join(G, "r02_step_kernels_raw.csv"), os.path.join(G, "r02_xgate_kernels_raw.csv")],
                   os.path.join(P, "r02_step_kernels_ncu_summary.csv"))
    for r in rows:
        print(r)
    shares(os.path.join(G, "r02_launches_raw.csv"), os.path.join(P, "r02_launch_shares.csv"))
