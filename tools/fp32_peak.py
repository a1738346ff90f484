#!/usr/bin/env python
"""Measured fp32 pipe peaks of this GPU (csrc/mppi_microbench.cu through the C ABI): FFMA / FFMA2 /
unfused FMUL+FADD / FMUL2+FADD2 full-chip throughput and dependent-issue latencies.
   python tools/fp32_peak.py [--out gpurun_out/fp32_peak.json]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mppi_playground_b200 import _capi  # noqa: E402


def measure(device: int = 0) -> dict:
    rep = _capi.MppiFp32Report()
    _capi.check(_capi.load().mppi_fp32_microbench(device, C.byref(rep)))
    return {n: getattr(rep, n) for n, _ in rep._fields_ if n != "reserved"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fp32_peak.json"))
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    r = measure(a.device)
    r["derived_fma_peak_tflops_at_measured_clock"] = r["sms"] * 128 * 2 * r["sm_clock_mhz"] * 1e6 / 1e12
    print(json.dumps(r, indent=1))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(r, open(a.out, "w"), indent=1)
