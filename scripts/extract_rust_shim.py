#!/usr/bin/env python3
"""Writes the Rust side of the boundary (INTEGRATION.md's ```rust blocks) as source files under integration/rust/,
so that a maintainer of the crate can copy them instead of cutting them out of the document.  INTEGRATION.md stays
the single source of truth: tests/test_host_abi.py checks the files are what this script produces.
No Rust toolchain exists in the build image: the files are shipped uncompiled."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# (file, heading the block sits under) in document order
TARGETS = ["b200.rs", "song_mod_analyze_with_options.rs", "decoder_analyze_paths_with_options.rs", "b200_s16.rs", "b200_pcm.rs"]


def blocks():
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.findall(r"```rust\n(.*?)```", md, flags=re.S)


def render():
    bl = blocks()
    assert len(bl) == len(TARGETS), (len(bl), TARGETS)
    out = {}
    for name, body in zip(TARGETS, bl):
        out[name] = ("// Extracted from INTEGRATION.md by scripts/extract_rust_shim.py -- edit the document, not this file.\n"
                     "// Uncompiled: the build image of this repository has no Rust toolchain.\n" + body)
    return out


def main():
    d = os.path.join(ROOT, "integration", "rust")
    os.makedirs(d, exist_ok=True)
    for name, text in render().items():
        open(os.path.join(d, name), "w").write(text)
        print("wrote", os.path.join("integration", "rust", name))


if __name__ == "__main__":
    sys.exit(main())
