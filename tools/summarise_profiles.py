"""Turn the ncu outputs a `tools/profile.sh` run left under gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarise_profiles.py r01

* gpurun_out/launches.csv  (ncu --metrics gpu__time_duration.sum --csv)  -> profiles/<round>_launches.md
  per-kernel launch count, total/mean duration and SHARE of the summed device time (cold-cache, serialised: only shares are
  comparable with bench.py's CUDA-event numbers);
* gpurun_out/prof_<kernel>.ncu-rep (ncu --set full)  -> profiles/<round>_<kernel>_full.md
  the metrics the roofline discussion needs, per captured launch: duration, DRAM bytes read/written, DRAM/L2/SM throughput in %
  of peak, achieved occupancy, registers, local-memory traffic, the top stall reasons.
"""
import csv
import io
import os
import re
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")


def read_ncu_csv(text):
    lines = text.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


def short(name):
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name.split("(")[0][-48:]


def launches(tag):
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in read_ncu_csv(open(path, errors="replace").read()) if r.get("Metric Name") == "gpu__time_duration.sum"]
    per = defaultdict(list)
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)  # -> microseconds
        per[short(r["Kernel Name"])].append(v)
    total = sum(sum(v) for v in per.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 3 --no-vector`\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised; compare the\n")
        f.write("SHARE column with `roofline.kernel_time_share` of the bench line, not the absolute times.\n\n")
        f.write(f"{len(rows)} launches, {total / 1e3:.2f} ms summed device time\n\n| kernel | launches | total ms | mean us | max us | share |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v) / 1e3:.3f} | {sum(v) / len(v):.1f} | {max(v):.1f} | {sum(v) / total:.3f} |\n")
    print("wrote", f"{tag}_launches.md")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), blocks"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "local/global LSU wavefronts"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads"),
    ("smsp__inst_executed_op_local_st.sum", "local stores"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
]


def full(tag):
    for fn in sorted(os.listdir(SRC)):
        if not fn.endswith(".ncu-rep"):
            continue
        rep = os.path.join(SRC, fn)
        try:
            text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=600).stdout
            rows = read_ncu_csv(text)
        except Exception as e:  # noqa: BLE001
            print("skip", fn, e)
            continue
        if len(rows) < 2:
            print("skip (empty)", fn)
            continue
        units, data = rows[0], rows[1:]
        name = fn[: -len(".ncu-rep")].replace("prof_", "")
        with open(os.path.join(OUT, f"{tag}_{name}_full.md"), "w") as f:
            f.write(f"# {tag}: `ncu --set full --clock-control none` of `{name}` ({len(data)} launches captured)\n\n")
            f.write("Times under ncu are serialised replays: use them for ratios (bytes, hit rates, stalls), not as bench numbers.\n\n")
            f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|---|" + "---|" * len(data) + "\n")
            f.write("| kernel | | " + " | ".join(short(r.get("Kernel Name", "")) for r in data) + " |\n")
            for key, label in WANT:
                if key not in data[0]:
                    continue
                f.write(f"| {label} (`{key}`) | {units.get(key, '')} | " + " | ".join(r[key] for r in data) + " |\n")
        print("wrote", f"{tag}_{name}_full.md")


def traffic(tag):
    """gpurun_out/traffic.csv (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over tools/prof_keyword.py) ->
    profiles/<tag>_traffic.json: mean DRAM bytes per launch of every kernel (all launches of the run); bench.py copies the
    dominant kernel's figure into roofline.traffic."""
    import json
    path = os.path.join(SRC, "traffic.csv")
    if not os.path.exists(path):
        return
    rows = read_ncu_csv(open(path, errors="replace").read())
    per = defaultdict(lambda: defaultdict(float))
    cnt = defaultdict(set)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows:
        k = short(r["Kernel Name"])
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            per[k][m] += float(r["Metric Value"].replace(",", "")) * scale.get(r.get("Metric Unit", "byte"), 1.0)
            cnt[k].add(r["ID"])
    out = {k: {"launches": len(cnt[k]), "dram_bytes_per_launch": (v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) / max(1, len(cnt[k])),
               "dram_read_per_launch": v["dram__bytes_read.sum"] / max(1, len(cnt[k])), "dram_write_per_launch": v["dram__bytes_write.sum"] / max(1, len(cnt[k]))}
           for k, v in per.items()}
    json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python tools/prof_keyword.py (3 identical batches of cfg2)",
               "kernels": out}, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    print("wrote", f"{tag}_traffic.json")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    full(tag)
    traffic(tag)
