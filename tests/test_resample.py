"""oracle/resample.py (the CPU checker of the device resampler) against the public implementation of the algorithm
it restates, scipy.signal.resample_poly.  The resampler's parity against the REFERENCE is unpinned (swresample and
rubato are third-party code outside the reference's tree; see oracle/resample.py): this file pins the checker, the GPU
suite (tests/test_gpu_parity.py::test_resample_*) pins the kernel against the checker."""
import math

import numpy as np
import pytest

from oracle import resample as R

RATES = (44100, 48000, 32000, 96000, 88200, 11025, 8000, 16000, 24000, 22051, 12345)


@pytest.mark.parametrize("rate", RATES)
def test_oracle_resampler_is_scipys_resample_poly(rate):
    signal = pytest.importorskip("scipy.signal")
    rng = np.random.default_rng(rate)
    for n in (1, 7, 500, 4001):
        x = rng.standard_normal(n).astype(np.float32)
        want = signal.resample_poly(x.astype(np.float64), 22050, rate)
        got = R.resample(x, rate)
        m = min(want.size, got.size)
        assert abs(want.size - got.size) <= 1 and m > 0
        # coefficients rounded to f32 (what the device holds) and an f32 result: a few ulp of the peak
        assert np.abs(want[:m] - got[:m]).max() <= 1e-6 * max(1.0, float(np.abs(x).max())), (rate, n)


def test_output_length_is_the_symphonia_decoders():
    """src/song/decoder/symphonia.rs:379-380: ceil(ratio x n) in f64"""
    assert R.resampled_len(0, 44100) == 0 and R.resampled_len(1, 44100) == 1 and R.resampled_len(2, 44100) == 1
    assert R.resampled_len(3, 44100) == 2 and R.resampled_len(1000, 22050) == 1000
    assert R.resampled_len(48000, 48000) == 22050 and R.resampled_len(48001, 48000) == 22051
    assert R.resampled_len(8000, 8000) == 22050 and R.resampled_len(441, 11025) == 882
    for rate in RATES:
        for n in (1, 999, 123457):
            exact = -(-n * 22050 // rate)
            assert 0 <= R.resampled_len(n, rate) - exact <= 1   # f64 rounding of the ratio may add one sample


def test_filter_design_properties():
    for rate in (44100, 48000, 8000):
        up, down, pre_remove, h = R.design(rate)
        g = math.gcd(22050, rate)
        assert (up, down) == (22050 // g, rate // g)
        assert abs(h.sum() - up) < 1e-9                      # unit DC gain after the zero-stuffing
        nz = h[np.flatnonzero(h)[0]:]
        assert np.allclose(nz, nz[::-1], atol=1e-15)         # linear phase
        # a constant comes out as the constant away from the edges
        y = R.resample(np.ones(4000, np.float32), rate)
        assert np.abs(y[200:-200] - 1.0).max() < 2e-3
