"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.

  python tools/summarize_profiles.py launches <launches.csv> <out_prefix>
  python tools/summarize_profiles.py full <prof.ncu-rep | raw.csv> <out.csv> [case names...]
"""
import collections
import csv
import gzip
import re
import shutil
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def to_ns(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)


def launches(path, prefix):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    with gzip.open(prefix + "_launch_list.csv.gz", "wt") as f:
        f.writelines(lines)
    names = [r["Kernel Name"] for r in rows]
    adam = [i for i, n in enumerate(names) if "FusedAd" in n or "adamw_multi" in n]   # last kernel of a step
    groups = []
    for i in adam:
        if groups and i - groups[-1][-1] <= 3:
            groups[-1].append(i)
        else:
            groups.append([i])
    a, b = (groups[-2][-1] + 1, groups[-1][-1] + 1) if len(groups) >= 2 else (0, len(rows))
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[a:b]:
        n = r["Kernel Name"]
        key = n[:70] if "gd4d" in n else re.sub(r"<.*", "", n)[:70]
        tot[key] += to_ns(r["Metric Value"], r["Metric Unit"])
        cnt[key] += 1
    T = sum(tot.values())
    with open(prefix + "_step_shares.txt", "w") as f:
        f.write(f"# one full training step (rows {a}..{b} of the launch list, {b - a} launches), ncu "
                f"gpu__time_duration.sum, serialised + cold cache: compare SHARES, not absolutes\n")
        f.write(f"# sum of kernel durations: {T / 1e6:.3f} ms\n")
        mine = sum(v for k, v in tot.items() if "gd4d" in k)
        f.write(f"# share of libgd4d_xview.so kernels: {100 * mine / T:.1f}%\n")
        for k, v in tot.most_common(40):
            f.write(f"{v / 1e6:9.3f} ms {100 * v / T:5.1f}%  n={cnt[k]:4d}  avg={v / cnt[k] / 1e3:8.1f} us  {k}\n")
    print(open(prefix + "_step_shares.txt").read())


def full(rep, out, cases):
    """``rep``: an .ncu-rep, or the raw CSV already exported next to the GPU (`ncu -i rep --page raw --csv`;
    a --set full report of 25 kernels is tens of MB, the CSV a few hundred KB)."""
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["case", "slot"] + [f"{hdr[i]} [{units[i]}]" for i in idx])
        # a backward call of the sorted path is five kernels: scan / scatter / owner / finish belong to the
        # slot their emit kernel (items<.., false>) opened; every other xview kernel opens its own slot.
        # Slots are named by walking REPS repetitions of each case group (fwd, bwd[, bwdS]) in launch order.
        ki = hdr.index("Kernel Name")
        follower = re.compile(r"xview_bwd_(scan|scatter|owner)_kernel|xview_bwd_items_kernel<[^>]*(true|\(bool\)1|, 1)>")
        groups, i = [], 0
        while i < len(cases):                                   # [[fwd, bwd], [fwd, bwd, bwdS], ...]
            j = i + 1
            while j < len(cases) and cases[j].rsplit("_", 1)[0] == cases[i].rsplit("_", 1)[0]:
                j += 1
            groups.append(cases[i:j])
            i = j
        slots = -1
        names = []
        for r in data:
            if not follower.search(r[ki]):
                slots += 1
            names.append(slots)
        nslots = slots + 1
        reps = max(1, nslots // max(1, len(cases)))
        order = [c for g in groups for _ in range(reps) for c in g] if cases else []
        for r, sl in zip(data, names):
            case = order[sl] if sl < len(order) else ""
            w.writerow([case, sl] + [r[i][:90] for i in idx])
    print(open(out).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4:])
