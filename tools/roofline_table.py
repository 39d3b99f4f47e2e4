"""Roofline table of one bench line: per hand-written kernel, launches and device time per step, algorithmic bytes
 where defined, achieved GB/s against the measured HBM peak, and the tensor-pipe figure of the
KPConv contraction. Reads a bench JSON line and MEASURED_PEAKS.json; writes markdown to stdout.

    python tools/roofline_table.py profiles/r1_bench_v10_final.json > profiles/r1_roofline_table.md
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    d = json.load(open(sys.argv[1]))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", d["roofline"]["peak"]))
    tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    k = d["kernels"]
    print(f"# Roofline table — {os.path.basename(sys.argv[1])}\n")
    print(f"{d['config']['workload']}; {d['ms_per_step']:.3f} ms/step device-resident, {d['e2e']['ms_per_step']:.3f} ms/step "
          f"end to end; peaks: HBM {hbm:.0f} GB/s (measured copy), bf16 {tf:.0f} TFLOP/s sustained (MEASURED_PEAKS.json).\n")
    print("Times are CUDA-event durations recorded inside the library around each launch (eager profile leg of bench.py,")
    print("same kernels and shapes as the graphed step; event overhead inflates the 3-5 us kernels).\n")
    print("| kernel | launches/step | ms/step | us/launch | algorithmic MB/step | achieved GB/s | frac of HBM peak |")
    print("|---|---|---|---|---|---|---|")
    tot = 0.0
    for name, v in sorted(k.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        if "tags" in v:  # an aggregate of other rows (kp_fwd_kernel = forward + dX launches of the same kernel)
            name += " = " + " + ".join(v["tags"])
        else:
            tot += v["ms_per_step"]
        mb = v.get("algorithmic_mb_per_step")
        g = v.get("achieved_gbs")
        print(f"| `{name}` | {v['launches_per_step']:.0f} | {v['ms_per_step']:.3f} | "
              f"{1e3 * v['ms_per_step'] / max(v['launches_per_step'], 1):.1f} | {'%.1f' % mb if mb else '—'} | "
              f"{'%.0f' % g if g else '—'} | {'%.3f' % (g / hbm) if g else '—'} |")
    print(f"| sum | | {tot:.3f} | | | | |\n")
    f = k.get("kp_fwd", {})
    if "contraction_tflops" in f:
        print(f"KPConv forward contraction (2·Nq·K·Cin·Cout flops per call, TF32 operands): {f['contraction_tflops']:.1f} TFLOP/s "
              f"over the whole kernel time (gather included) = {f['contraction_tflops'] / tf:.3f} of the sustained bf16 peak "
              f"(TF32 runs at half the bf16 rate).")
    r = d["roofline"]
    print(f"\n`roofline` of the line: `{r['kernel']}`, {r['achieved']:.0f} GB/s of {r['peak']:.0f} = {r['frac']:.3f}; "
          f"DRAM traffic per launch from the ncu capture {r['traffic'] / 1e6 if r.get('traffic') else float('nan'):.1f} MB against "
          f"{r['bytes_per_launch'] / 1e6:.1f} MB algorithmic: no wasted re-reads; the kernels are latency / issue bound at these "
          f"sizes (profiles/README.md).")


if __name__ == "__main__":
    main()
