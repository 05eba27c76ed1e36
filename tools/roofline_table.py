"""Render the per-kernel roofline table of a bench.py JSON line as markdown.

    python bench.py --no-cpu > b.json ; python tools/roofline_table.py b.json > profiles/rNN_kernel_roofline.md

Every row is one (C-ABI entry point, shape) of the instrumented step: algorithmic bytes / FLOPs (bench.py
``algorithmic_cost``) over the CUDA-event duration of its launches, against the measured peaks in MEASURED_PEAKS.json."""
import json
import sys


def main(path):
    line = None
    for raw in open(path):
        raw = raw.strip()
        if raw.startswith("{") and '"kernel_roofline"' in raw:
            line = json.loads(raw)
    if line is None:
        sys.exit("no bench line with kernel_roofline in " + path)
    rows = line["extra"]["kernel_roofline"]
    total_ms = line["extra"]["instrumented_step_ms"]
    print(f"Per-kernel roofline of one instrumented step ({total_ms:.2f} ms with per-call events; timed step "
          f"{line['ms_per_step']:.2f} ms, {line['value']:.0f} {line['unit']}), peaks: "
          f"{line['roofline']['peak']:.0f} GB/s HBM copy, {line['step_roofline']['peak']:.0f} TFLOP/s bf16 "
          f"({line['roofline']['peak_source']}).\n")
    print("| kernel | shape | launches | µs / launch | algorithmic MB / launch | GB/s | of HBM peak | TFLOP/s | of bf16 peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        tf = f"{r['tflops']:.0f} | {r['tensor_frac']:.2f}" if r["tflops"] > 0 else "– | –"
        print(f"| {r['kernel']} | {r['shape']} | {r['launches']} | {r['us_per_launch']:.1f} | {r['mb_per_launch']:.2f} | "
              f"{r['gbs']:.0f} | {r['hbm_frac']:.2f} | {tf} |")


if __name__ == "__main__":
    main(sys.argv[1])
