"""gpurun_out/ev2/ (tools/collect_profiles_r2.sh) -> the committed summaries under profiles/r2/.

    python tools/summarise_profiles_r2.py

* ncu_forward_b1_r2.txt / ncu_config3_b8_r2.txt: one row per profiled launch (ncu --set full, every kernel of one
  forward at B = 1; the cost-volume / set-conv / set-upconv / predictor kernels at B = 8 = BASELINE.json configs[2]):
  duration, grid, registers, tensor-pipe active %, SM and DRAM throughput %, DRAM bytes;
* ncu_index_*_r2.txt: the details pages of the index op captures (configs[0]);
* traffic_r2.json: dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed the way bench.py looks it up;
* launches_r2.csv, bench / index / train lines: copied."""
import csv
import glob
import gzip
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "ev2")
DST = os.path.join(ROOT, "profiles", "r2")


def rows_of(path):
    """ncu --page raw --csv: list of dicts (one per launch), numbers as floats where possible."""
    with open(path) as f:
        rd = list(csv.reader(f))
    hdr = rd[0]
    out = []
    for r in rd[2:]:                      # rd[1] holds the units
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        out.append(d)
    return out, dict(zip(hdr, rd[1]))


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return float("nan")


def short(name):
    name = re.sub(r"\(.*$", "", name).replace("elo::", "").replace("void ", "")
    return name[:58]


def table(path, title, out_path, tags=None):
    rows, units = rows_of(path)
    lines = [title, "", "%-3s %-58s %9s %7s %5s %8s %8s %8s %10s %10s" % (
        "#", "kernel", "dur [us]", "grid", "regs", "tensor%", "SM thr%", "DRAM%", "DRAM rd MB", "DRAM wr MB")]
    traffic = {}
    for i, r in enumerate(rows):
        dur = num(r.get("gpu__time_duration.sum"))
        if units.get("gpu__time_duration.sum", "").startswith("ns") or units.get("gpu__time_duration.sum", "") == "nsecond":
            dur /= 1e3
        rd_b, wr_b = num(r.get("dram__bytes_read.sum")), num(r.get("dram__bytes_write.sum"))
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        rd_b *= scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
        wr_b *= scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
        lines.append("%-3d %-58s %9.2f %7s %5s %8.1f %8.1f %8.1f %10.2f %10.2f" % (
            i, short(r["Kernel Name"]), dur, r.get("launch__grid_size", "?"), r.get("launch__registers_per_thread", "?"),
            num(r.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")),
            num(r.get("sm__throughput.avg.pct_of_peak_sustained_elapsed")),
            num(r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")), rd_b / 1e6, wr_b / 1e6))
        if tags and i < len(tags) and tags[i]:
            traffic[tags[i]] = rd_b + wr_b
    open(out_path, "w").write("\n".join(lines) + "\n")
    return rows, traffic


def main():
    os.makedirs(DST, exist_ok=True)
    traffic = {}
    p = os.path.join(SRC, "forward_b1_metrics.csv")
    if os.path.exists(p):
        rows, _ = table(p, "ncu --set full, every kernel launch of ONE forward (B = 1, 64x1800, second forward of the process; "
                           "cold-cache, serialised)", os.path.join(DST, "ncu_forward_b1_r2.txt"))
        # tag the launches the way bench.py names kernels: order of appearance per kernel name, level by level
        order = {"group_mlp_max_tc_kernel": ["sa3", "l3", "l2", "l1", "l0"], "cost_volume_1_tc_kernel": ["l2o", "l2", "l1", "l0"],
                 "cost_volume_2_tc_kernel": ["l2o", "l2", "l1", "l0"], "row_mlp_tc_kernel": ["l3", "l2", "l1", "l0"]}
        api = {"group_mlp_max_tc_kernel": "elo_group_mlp_max", "cost_volume_1_tc_kernel": "elo_cost_volume_1",
               "cost_volume_2_tc_kernel": "elo_cost_volume_2", "row_mlp_tc_kernel": "elo_row_mlp"}
        seen = {}
        rr, units = rows_of(p)
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        # the capture window starts inside one forward (its level-0 kernels) and runs into the next one, which begins
        # with gt_pose_kernel: read the levels in execution order from there, level 0 from the rows before it
        start = next((i for i, r in enumerate(rr) if "gt_pose_kernel" in r["Kernel Name"]), 0)
        rr = rr[start:] + rr[:start]
        for r in rr:
            for k in order:
                if k in r["Kernel Name"]:
                    i = seen.get(k, 0)
                    seen[k] = i + 1
                    if i < len(order[k]):
                        b = num(r["dram__bytes_read.sum"]) * scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0) + \
                            num(r["dram__bytes_write.sum"]) * scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
                        traffic["%s[%s]" % (api[k], order[k][i])] = b
    p = os.path.join(SRC, "config3_b8_metrics.csv")
    if os.path.exists(p):
        table(p, "ncu --set full, BASELINE.json configs[2]: B = 8 frame pairs, the cost-volume / set-conv / set-upconv / "
                 "predictor kernels of one forward", os.path.join(DST, "ncu_config3_b8_r2.txt"))
    for r in ("index_7x25", "index_7x25_storewarp", "index_11x41"):
        p = os.path.join(SRC, r + "_metrics.csv")
        if os.path.exists(p):
            rr, units = rows_of(p)
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            b = num(rr[0]["dram__bytes_read.sum"]) * scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0) + \
                num(rr[0]["dram__bytes_write.sum"]) * scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
            if r == "index_7x25":
                traffic["elo_fused_conv_select_k[config1]"] = b
            traffic["index:" + r] = b
        d = os.path.join(SRC, r + "_details.txt.gz")
        if os.path.exists(d):
            with gzip.open(d, "rt") as f:
                open(os.path.join(DST, "ncu_%s_r2.txt" % r), "w").write(f.read())
    json.dump({"dram_bytes_per_launch": traffic, "source": "ncu --set full captures of tools/collect_profiles_r2.sh (B = 1)"},
              open(os.path.join(ROOT, "profiles", "traffic_r2.json"), "w"), indent=1)
    for name in ("launches.csv", "bench_k20.json", "bench_ref_k20.json", "bench_default.json", "bench_serial.json",
                 "bench_b8.json", "bench_128x2048.json", "index_bench.jsonl", "train_b8.json", "train_b32.json",
                 "bench_serial.err", "sanitize_index.log", "sanitize_storewarp.log", "sanitize_forward.log"):
        s = os.path.join(SRC, name)
        if os.path.exists(s):
            dst = name.replace(".csv", "_r2.csv").replace(".json", "_r2.json") if not name.endswith(".jsonl") else name.replace(".jsonl", "_r2.jsonl")
            dst = dst.replace("bench_serial.err", "kernel_times_serial_r2.txt").replace(".log", "_r2.txt")
            shutil.copy(s, os.path.join(DST, dst))
    print(sorted(os.listdir(DST)))


if __name__ == "__main__":
    main()
