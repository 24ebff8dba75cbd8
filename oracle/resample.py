"""TEST INFRASTRUCTURE (like everything under oracle/): the sample-rate conversion of the decode-side feed, restated
in numpy.  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this.

PARITY UNPINNED.  The reference resamples inside its decoders with third-party libraries that are not part of its
tree and cannot run here: swresample behind src/song/decoder/ffmpeg.rs:36-109 (ffmpeg-next 8.1.0) and rubato 3.0.0's
synchronous FFT resampler behind src/song/decoder/symphonia.rs:304-404.  The two do not agree sample for sample (the
reference's own tests compare its decoders through tolerances), and the reference holds no golden vector of a
resampled signal.  What is restated here is the published algorithm the CUDA kernel (bliss-rs_b200/csrc/wave_setup.cu
resample_kernel) implements -- the rational polyphase resampler of scipy.signal.resample_poly(x, 22050, rate), default
Kaiser(5.0) window -- so that the kernel has a CPU checker (tests/test_resample.py pins THIS file against scipy itself).
From the reference: only the output length, ceil(22050 / rate x n) (src/song/decoder/symphonia.rs:379-380).
"""
import math

import numpy as np

SAMPLE_RATE = 22050  # src/lib.rs:143


def resampled_len(n, rate):
    """src/song/decoder/symphonia.rs:379-380: (resampler.resample_ratio() * samples.len() as f64).ceil()"""
    return int(n) if rate == SAMPLE_RATE else int(math.ceil(float(SAMPLE_RATE) / float(rate) * float(n)))


def design(rate):
    """(up, down, pre_remove, h'): scipy/signal/_signaltools.py resample_poly + firwin (pass_zero, scale=True):
    h[k] = f_c sinc(f_c (k - half)) kaiser_5(k), k = 0 .. 2 half, half = 10 max(up, down), f_c = 1 / max(up, down),
    scaled to unit sum, times up; `down - half mod down` zeros in front make the delay a whole number of outputs."""
    g = math.gcd(SAMPLE_RATE, int(rate))
    up, down = SAMPLE_RATE // g, int(rate) // g
    m = max(up, down)
    half = 10 * m
    k = np.arange(2 * half + 1, dtype=np.float64)
    h = (1.0 / m) * np.sinc((k - half) / m)
    h = h * (np.i0(5.0 * np.sqrt(np.maximum(0.0, 1.0 - ((k - half) / half) ** 2))) / np.i0(5.0))
    h = h / h.sum() * up
    pre_pad = down - half % down
    return up, down, (half + pre_pad) // down, np.concatenate([np.zeros(pre_pad), h])


def resample(x, rate):
    """y[j] = sum_i h'[(j + pre_remove) down - up i] x[i], j < resampled_len: f64 sums of the f32-rounded coefficients
    the device holds, rounded to f32 at the end"""
    x = np.asarray(x, np.float32)
    n = x.size
    if rate == SAMPLE_RATE:
        return x.copy()
    n_out = resampled_len(n, rate)
    up, down, pre_remove, h = design(rate)
    hf = h.astype(np.float32).astype(np.float64)
    taps = -(-hf.size // up)
    tab = np.zeros((up, taps))
    for p in range(up):
        row = hf[p::up]
        tab[p, :row.size] = row
    q = (np.arange(n_out, dtype=np.int64) + pre_remove) * down
    i0, ph = q // up, q % up
    xp = np.concatenate([np.zeros(taps), x.astype(np.float64), np.zeros(max(0, int(i0.max()) + 1 - n) if n_out else 0)])
    y = np.zeros(n_out)
    for t in range(taps):
        y += tab[ph, t] * xp[i0 - t + taps]
    return y.astype(np.float32)
