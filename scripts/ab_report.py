#!/usr/bin/env python3
"""Reads the bench lines scripts/gpu_r02_ab.sh leaves in gpurun_out/ (ab_v<mask>.json, ab_prio*.json, ab_wave*.json,
ab_stride4128.json) and prints one table: step time against the default build of the same box, the two FFT kernels'
times, reproducibility, parity -- and which cuts clear the bar for promotion (>= 1 % faster step, bitwise reproducible,
parity within 1e-4 where it was measured).  No GPU needed: run it here on the merged gpurun_out/.

  python scripts/ab_report.py [gpurun_out]
"""
import glob
import json
import os
import sys


def load(path):
    try:
        return json.load(open(path))
    except Exception:
        return None


def kernel_ms(d, name):
    for k in d.get("roofline", {}).get("kernels", []):
        if k["kernel"].startswith(name):
            return k["avg_ms"]
    return float("nan")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    base = load(os.path.join(out, "ab_v0.json"))
    if not base:
        print("no %s/ab_v0.json: run scripts/gpu_r02_ab.sh first" % out)
        return 1
    rows = []
    for path in sorted(glob.glob(os.path.join(out, "ab_*.json"))):
        name = os.path.basename(path)[3:-5]
        if name.startswith("stft_") or name == "v0":
            continue
        d = load(path)
        if not d or "ms_per_step" not in d:
            rows.append((name, None))
            continue
        rows.append((name, d))
    b_ms = base["ms_per_step"]
    print("default: %.2f ms/step, %.0f songs/s, stft8192 %.2f ms, pvoc512 %.2f ms" %
          (b_ms, base["value"], kernel_ms(base, "stft8192"), kernel_ms(base, "pvoc512")))
    print("| run | ms/step | vs default | stft8192 ms | pvoc512 ms | reproducible | parity | promote |")
    print("|---|---|---|---|---|---|---|---|")
    for name, d in sorted(rows, key=lambda r: (r[1] or {}).get("ms_per_step", 1e9)):
        if d is None:
            print("| %s | unreadable | | | | | | no |" % name)
            continue
        gain = 1.0 - d["ms_per_step"] / b_ms
        repro = d.get("bitwise_reproducible_across_steps")
        par = (d.get("cpu_baseline") or {}).get("parity_within_1e-4")
        ok = gain >= 0.01 and repro is True and par is not False
        print("| %s | %.2f | %+.1f %% | %.2f | %.2f | %s | %s | %s |" %
              (name, d["ms_per_step"], -100.0 * gain, kernel_ms(d, "stft8192"), kernel_ms(d, "pvoc512"), repro,
               "n/a" if par is None else par, "YES" if ok else "no"))
    for v in ("0", "256"):
        d = load(os.path.join(out, "ab_stft_v%s.json" % v))
        if d:
            print("STFT micro-benchmark, variant %s: %s" % (v, json.dumps({k: d[k] for k in d if k in ("value", "unit", "roofline", "cufft")})[:400]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
