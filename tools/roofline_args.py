#!/usr/bin/env python
"""tools/roofline_args.py — the --beta / --pi a B200 run needs for the reference's plots/plot_roofline.py.

    python tools/roofline_args.py            # prints e.g.  --beta 3113.0 --pi 792857.1

The reference's driver (and ours, host/perf_timer.hpp) reports performance in flops per HOST
time-stamp-counter cycle, and plot_roofline.py draws its roof from --beta (bytes/cycle, default 24)
and --pi (flops/cycle, default 4) — an Apple M1's numbers (plots/plot_roofline.py:597-598).  For a
JSON written by the untouched plots/run_benchmark.py driving host/SparseGEMM.out on a B200 box the
roof is the GPU's, expressed per host TSC cycle:
    beta = measured HBM bytes/s  / TSC Hz        pi = measured dense bf16 flops/s / TSC Hz
(MEASURED_PEAKS.json when present, else the fallback of /opt/skills/guides/B200_PROFILING.md; the
TSC rate is measured with a few lines of C around a 0.2 s sleep).  `--pi-fp32` prints the CUDA-core
fp32 roof instead (148 SMs x 128 lanes x 2 flops x SM clock) for the kernels that do not use the
tensor cores (gather, code_gemv).
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TSC_C = r"""
#include <stdio.h>
#include <time.h>
#include <x86intrin.h>
int main(void) {
    struct timespec a, b, d = {0, 200000000};
    clock_gettime(CLOCK_MONOTONIC, &a);
    unsigned long long t0 = __rdtsc();
    nanosleep(&d, 0);
    unsigned long long t1 = __rdtsc();
    clock_gettime(CLOCK_MONOTONIC, &b);
    double s = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
    printf("%.0f\n", (double)(t1 - t0) / s);
    return 0;
}
"""


def tsc_hz():
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "tsc.c"), os.path.join(d, "tsc")
        open(src, "w").write(TSC_C)
        subprocess.run(["gcc", "-O1", "-o", exe, src], check=True)
        return float(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"] * 1e9, d["bf16_tflops"] * 1e12, d.get("sm_max_mhz", 1965.0) * 1e6, "MEASURED_PEAKS.json"
    return 6.65e12, 1.8e15, 1965e6, "fallback (B200_PROFILING.md)"


def main():
    hz = tsc_hz()
    hbm, tensor, sm_hz, src = peaks()
    pi = 148 * 128 * 2 * sm_hz if "--pi-fp32" in sys.argv else tensor
    print(f"--beta {hbm / hz:.1f} --pi {pi / hz:.1f}")
    print(f"# host TSC {hz / 1e9:.3f} GHz; HBM {hbm / 1e9:.0f} GB/s, {'fp32 CUDA cores' if '--pi-fp32' in sys.argv else 'dense bf16'} "
          f"{pi / 1e12:.0f} TFLOP/s ({src})", file=sys.stderr)


if __name__ == "__main__":
    main()
