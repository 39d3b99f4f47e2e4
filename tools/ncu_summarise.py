"""Summarise an ncu report (read here, no GPU needed): per-launch key metrics as CSV, and the mean DRAM traffic per
launch of every kernel as JSON (bench.py reads profiles/traffic.json for the roofline `traffic` field).

    python tools/ncu_summarise.py gpurun_out/prof.ncu-rep profiles/r1_ncu_<name>.csv [profiles/traffic.json]
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out_csv = sys.argv[1], sys.argv[2]
    traffic_json = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    with open(out_csv, "w") as f:
        f.write(f"# ncu --set full --clock-control none, summarised from {rep.split('/')[-1]}\n")
        w = csv.writer(f)
        w.writerow(["kernel"] + [hdr[i] for i in idx])
        w.writerow([""] + [units[i] for i in idx])
        for d in data:
            w.writerow([re.sub(r"\(.*", "", d[name_i]).replace("void ", "")] + [d[i] for i in idx])
    if traffic_json:
        ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        acc = {}
        for d in data:
            k = re.sub(r"[<(].*", "", d[name_i]).replace("void ", "").replace("kp::", "")
            if re.search(r"<\(int\)\d+, \(bool\)1>|<\d+, 1>", d[name_i]):
                k += "_dense"  # the dense (unary block) instantiation of kp_fwd / kp_dw is reported separately
            b = float(d[ri].replace(",", "")) * UNIT.get(units[ri], 1.0) + float(d[wi].replace(",", "")) * UNIT.get(units[wi], 1.0)
            a = acc.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += b
        try:
            cur = json.load(open(traffic_json))
        except Exception:
            cur = {}
        for k, (n, b) in acc.items():
            cur[k] = {"launches": n, "dram_bytes_per_launch": b / n, "source": rep.split("/")[-1]}
        json.dump(cur, open(traffic_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
