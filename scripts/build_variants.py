#!/usr/bin/env python3
"""Builds experimental variants of libbliss_b200.so (same sources, different -D tuning knobs) into
bliss-rs_b200/variants/ so that one gpurun call can A/B them:  BLISS_B200_SO=<path> python bench.py ..."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bliss-rs_b200", "csrc")
OUT = os.path.join(ROOT, "bliss-rs_b200", "variants")
SOURCES = ["spectral.cu", "tempo.cu", "chroma.cu", "finalize.cu", "distance.cu", "gather.cu", "wave_setup.cu", "api.cu"]
BASE = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
        "-Wno-deprecated-gpu-targets"]

VARIANTS = {
    "nopacked": ["-DBLISS_NO_PACKED_FP"],  # scalar FADD/FMUL/FFMA butterflies instead of the f32x2 forms
    "stride4104": ["-DBLISS_CH_STRIDE=4104"],  # round 1's pitch (rows 32 bytes off the 128-byte lines); 4128 is the default now
    # chroma_pipe_kernel's cp.async ring: 3 stages (3 CTAs per SM) is the default
    "k5p4": ["-DK5P_STAGES_N=4", "-DK5P_SWIZZLE=0"],
    "k5p2": ["-DK5P_STAGES_N=2", "-DK5P_SWIZZLE=0"],
    # rows of exactly 64 B (chunk swizzle instead of padding) at four CTAs per SM are the default now; the steps there:
    "k5p3pad": ["-DK5P_STAGES_N=3", "-DK5P_SWIZZLE=0"],                       # round 2's first layout (pitch 80 B, 3 CTAs)
    "k5p3sw": ["-DK5P_STAGES_N=3", "-DK5P_SWIZZLE=1", "-DK5P_MIN_BLOCKS=3"],  # swizzle alone
    "k5p4sw": ["-DK5P_STAGES_N=4", "-DK5P_SWIZZLE=1"],                        # ... a fourth stage
    "k5p2sw4": ["-DK5P_STAGES_N=2", "-DK5P_SWIZZLE=1", "-DK5P_MIN_BLOCKS=4"],  # two stages at four CTAs
    # timedomain_kernel's register cap (65536 / (128 x min blocks)): does it fit beside three chroma-STFT CTAs?
    "td48": ["-DBLISS_TD_MIN_BLOCKS=10"],
    "td40": ["-DBLISS_TD_MIN_BLOCKS=12"],
    "td32": ["-DBLISS_TD_MIN_BLOCKS=16"],
}


def main():
    os.makedirs(OUT, exist_ok=True)
    want = sys.argv[1:] or list(VARIANTS)
    for name, defs in VARIANTS.items():
        if name not in want:
            continue
        objs = []
        for src in SOURCES:
            o = os.path.join(OUT, "%s_%s.o" % (name, src[:-3]))
            cmd = BASE + defs + (["-fmad=false"] if src == "tempo.cu" else []) + ["-c", os.path.join(CSRC, src), "-o", o]
            subprocess.check_call(cmd)
            objs.append(o)
        so = os.path.join(OUT, "libbliss_b200_%s.so" % name)
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs + ["-lcudart"])
        for o in objs:
            os.remove(o)
        print(so)


if __name__ == "__main__":
    sys.exit(main())
