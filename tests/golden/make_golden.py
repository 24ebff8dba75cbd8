#!/usr/bin/env python3
"""Regenerates tests/golden/golden.npz from the reference's own fixtures.

Run in the build container only (needs /root/reference); the GPU box and the
test-suite read the committed golden.npz and never touch /root/reference.

What goes in (all are *fixtures of the reference's test-suite*, no source code):
  * pcm_s16_mono        - data/s16_mono_22_5kHz.flac decoded bit-exactly by the
                          small FLAC decoder below (int16; tests divide by 32768).
                          Checked here against the adler32 the reference asserts
                          for ffmpeg's f32le output (src/song/decoder/ffmpeg.rs:455-462).
  * pcm_piano           - data/piano.wav (mono s16 22050 Hz) via the wave module,
                          adler32 checked against src/song/decoder/ffmpeg.rs:524-527.
  * every data/*.npy the hot-path tests of the reference read.
  * expected_* vectors  - the literal expected values of the reference's tests,
                          each with its file:line.
"""
import os
import struct
import sys
import wave
import zlib

import numpy as np

REF = "/root/reference"
DATA = os.path.join(REF, "data")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.npz")


class BitReader:
    def __init__(self, data: bytes, pos: int = 0):
        self.data = data + b"\0" * 16  # pad: final frame bit reads may look ahead
        self.pos = pos * 8

    def read(self, n: int) -> int:
        v = 0
        while n > 0:
            byte = self.data[self.pos >> 3]
            avail = 8 - (self.pos & 7)
            take = min(avail, n)
            shift = avail - take
            v = (v << take) | ((byte >> shift) & ((1 << take) - 1))
            self.pos += take
            n -= take
        return v

    def read_signed(self, n: int) -> int:
        v = self.read(n)
        if v >= 1 << (n - 1):
            v -= 1 << n
        return v

    def read_unary(self) -> int:
        n = 0
        while self.read(1) == 0:
            n += 1
        return n

    def align(self):
        self.pos = (self.pos + 7) & ~7

    def read_utf8(self) -> int:
        b0 = self.read(8)
        if b0 < 0x80:
            return b0
        n = 0
        while b0 & (0x80 >> n):
            n += 1
        v = b0 & ((1 << (7 - n)) - 1)
        for _ in range(n - 1):
            v = (v << 6) | (self.read(8) & 0x3F)
        return v


def _residual(br: BitReader, blocksize: int, order: int):
    method = br.read(2)
    assert method in (0, 1)
    plen = 4 if method == 0 else 5
    esc = (1 << plen) - 1
    porder = br.read(4)
    nparts = 1 << porder
    out = []
    for p in range(nparts):
        n = (blocksize >> porder) - (order if p == 0 else 0)
        k = br.read(plen)
        if k == esc:
            bits = br.read(5)
            out.extend(br.read_signed(bits) if bits else 0 for _ in range(n))
        else:
            for _ in range(n):
                q = br.read_unary()
                r = br.read(k) if k else 0
                u = (q << k) | r
                out.append((u >> 1) ^ -(u & 1))
    return out


_FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _subframe(br: BitReader, blocksize: int, bps: int):
    assert br.read(1) == 0
    typ = br.read(6)
    wasted = 0
    if br.read(1):
        wasted = br.read_unary() + 1
        bps -= wasted
    if typ == 0:
        s = [br.read_signed(bps)] * blocksize
    elif typ == 1:
        s = [br.read_signed(bps) for _ in range(blocksize)]
    elif 8 <= typ <= 12:
        order = typ - 8
        s = [br.read_signed(bps) for _ in range(order)]
        res = _residual(br, blocksize, order)
        c = _FIXED[order]
        for r in res:
            s.append(r + sum(c[j] * s[-1 - j] for j in range(order)))
    elif typ >= 32:
        order = typ - 31
        s = [br.read_signed(bps) for _ in range(order)]
        prec = br.read(4) + 1
        shift = br.read_signed(5)
        coef = [br.read_signed(prec) for _ in range(order)]
        res = _residual(br, blocksize, order)
        for r in res:
            s.append(r + (sum(coef[j] * s[-1 - j] for j in range(order)) >> shift))
    else:
        raise ValueError("reserved subframe type %d" % typ)
    if wasted:
        s = [x << wasted for x in s]
    return s


def decode_flac(path: str, expect):
    """-> int64 array [channels, n_frames]; `expect` = (rate, channels, bits per sample) of the stream"""
    data = open(path, "rb").read()
    assert data[:4] == b"fLaC"
    pos = 4
    total = None
    while True:
        hdr = data[pos]
        ln = int.from_bytes(data[pos + 1:pos + 4], "big")
        body = data[pos + 4:pos + 4 + ln]
        if hdr & 0x7F == 0:
            v = int.from_bytes(body[10:18], "big")
            rate = v >> 44
            ch = ((v >> 41) & 7) + 1
            bps = ((v >> 36) & 31) + 1
            total = v & ((1 << 36) - 1)
            assert (rate, ch, bps) == tuple(expect), (rate, ch, bps)
        pos += 4 + ln
        if hdr & 0x80:
            break
    out = [[] for _ in range(ch)]
    br = BitReader(data, pos)
    bs_table = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608}
    while len(out[0]) < total:
        assert br.read(14) == 0x3FFE, "lost frame sync"
        br.read(1)
        br.read(1)
        bs_code = br.read(4)
        sr_code = br.read(4)
        ch_code = br.read(4)
        br.read(3)
        br.read(1)
        assert ch_code <= 10
        br.read_utf8()
        if bs_code == 6:
            blocksize = br.read(8) + 1
        elif bs_code == 7:
            blocksize = br.read(16) + 1
        elif bs_code >= 8:
            blocksize = 256 << (bs_code - 8)
        else:
            blocksize = bs_table[bs_code]
        if sr_code == 12:
            br.read(8)
        elif sr_code in (13, 14):
            br.read(16)
        br.read(8)  # crc8
        if ch_code < 8:      # independent channels
            subs = [_subframe(br, blocksize, bps) for _ in range(ch_code + 1)]
        elif ch_code == 8:   # left + side
            left = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            subs = [left, [a - b for a, b in zip(left, side)]]
        elif ch_code == 9:   # side + right
            side = _subframe(br, blocksize, bps + 1)
            right = _subframe(br, blocksize, bps)
            subs = [[a + b for a, b in zip(side, right)], right]
        else:                # mid + side
            mid = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            subs = [[], []]
            for m, d in zip(mid, side):
                m = (m << 1) | (d & 1)
                subs[0].append((m + d) >> 1)
                subs[1].append((m - d) >> 1)
        assert len(subs) == ch
        for dst, sub in zip(out, subs):
            dst.extend(sub)
        br.align()
        br.read(16)  # crc16
    return np.array([c[:total] for c in out], dtype=np.int64)


def decode_flac_mono16(path: str) -> np.ndarray:
    pcm = decode_flac(path, (22050, 1, 16))[0]
    assert pcm.min() >= -32768 and pcm.max() <= 32767
    return pcm.astype(np.int16)


def adler32_f32(pcm_s16: np.ndarray) -> int:
    f = (pcm_s16.astype(np.float32) / np.float32(32768.0)).astype("<f4")
    return zlib.adler32(f.tobytes()) & 0xFFFFFFFF


def main():
    g = {}
    pcm = decode_flac_mono16(os.path.join(DATA, "s16_mono_22_5kHz.flac"))
    a = adler32_f32(pcm)
    assert a == 0x5E01930B, hex(a)  # src/song/decoder/ffmpeg.rs:455-462
    g["pcm_s16_mono"] = pcm

    with wave.open(os.path.join(DATA, "piano.wav"), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (1, 2, 22050)
        piano = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()
    a = adler32_f32(piano)
    assert a == 0xDE831E82, hex(a)  # src/song/decoder/ffmpeg.rs:524-527
    dec = np.load(os.path.join(DATA, "librosa-decoded.npy"))
    assert np.array_equal(dec, piano.astype(np.float32) / np.float32(32768.0))
    g["pcm_piano"] = piano

    # data/s16_stereo_22_5kHz.flac: the decoder test src/song/decoder/ffmpeg.rs:447-452 asserts the adler32 of
    # ffmpeg's mono f32le output, i.e. of swresample's s16 -> flt conversion followed by its stereo -> mono
    # matrix c*L + c*R, c = (float)sqrt(1/2) (no rate change: the file runs at 22 050 Hz).  Both channels of
    # this file are identical, so the hash pins the two constants, not the rounding order of the products.
    st = decode_flac(os.path.join(DATA, "s16_stereo_22_5kHz.flac"), (22050, 2, 16))
    assert st.min() >= -32768 and st.max() <= 32767
    st = np.ascontiguousarray(st.T.astype(np.int16))  # interleaved frames [n, 2]
    k, c = np.float32(1.0 / 32768.0), np.float32(np.sqrt(0.5))
    mono = (st[:, 0].astype(np.float32) * k) * c + (st[:, 1].astype(np.float32) * k) * c
    a = zlib.adler32(mono.astype("<f4").tobytes()) & 0xFFFFFFFF
    assert a == 0x1D7B2D6D, hex(a)  # src/song/decoder/ffmpeg.rs:447-452
    g["pcm_s16_stereo"] = st
    g["adler32_stereo_downmix"] = np.uint32(a)

    for name in ["chroma-filter", "chroma-interval", "chroma", "interval-feature-matrix",
                 "librosa-stft", "pitch-tuning", "spectrum-chroma-mags",
                 "spectrum-chroma-pitches", "spectrum-chroma"]:
        g[name.replace("-", "_")] = np.load(os.path.join(DATA, name + ".npy"))

    # src/song/mod.rs:553-580 (v2, tol 1e-5) and :593-617 (v1)
    g["expected_analysis_v2"] = np.array([
        0.3846389, -0.849141, -0.75481045, -0.8790748, -0.63258266, -0.7258959,
        -0.7757379, -0.8146726, 0.2716726, 0.25779057, -0.34292513, -0.62803423,
        -0.28095096, 0.08686459, 0.24446082, -0.5723257, 0.23292065, 0.19981146,
        -0.58594406, -0.06784296, -0.06000763, -0.58485717, -0.07880378], dtype=np.float32)
    g["expected_analysis_v1"] = np.array([
        0.3846389, -0.849141, -0.75481045, -0.8790748, -0.63258266, -0.7258959,
        -0.7757379, -0.8146726, 0.2716726, 0.25779057, -0.35661936, -0.63578653,
        -0.29593682, 0.06421304, 0.21852458, -0.581239, -0.9466835, -0.9481153,
        -0.9820945, -0.95968974], dtype=np.float32)
    # src/chroma.rs:498-509
    g["expected_chroma_interval_features"] = np.array([
        0.03860284, 0.02185281, 0.04224379, 0.06385278, 0.07311148, 0.02512566,
        0.00319899, 0.00311308, 0.00107433, 0.00241861])
    # src/utils.rs:255-513: the 256-magnitude frame of test_geometric_mean (a test
    # vector embedded in the reference's test module), expected 0.0025750597 +- 1e-8
    import re
    src = open(os.path.join(REF, "src", "utils.rs")).read()
    blk = src[src.index("let input = ["):]
    blk = blk[:blk.index("];")]
    vals = [float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:e[-+]?\d+)?", blk)]
    assert len(vals) == 256, len(vals)
    g["geometric_mean_input"] = np.array(vals, dtype=np.float32)
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
