"""Pins the CPU oracle (oracle/) against every golden vector / known answer the
reference's own test-suite holds for the analysis + distance path (SURVEY.md
section 4 / 8c).  Each test cites the reference test it mirrors and uses the
reference's own tolerance.  CPU only.
"""
import numpy as np
import pytest

from oracle import oracle as O


# ---- song/mod.rs -------------------------------------------------------------

def test_analyze_golden_v2(pcm_song, golden):
    # src/song/mod.rs:553-591 test_analyze, tol 1e-5
    rc, v = O.analyze(pcm_song, 2)
    assert rc == 0
    assert np.abs(v - golden["expected_analysis_v2"]).max() < 1e-5


def test_analyze_golden_v1(pcm_song, golden):
    # src/song/mod.rs:593-633 test_analyze_with_options, tol 1e-5
    rc, v = O.analyze(pcm_song, 1)
    assert rc == 0 and v.shape == (20,)
    assert np.abs(v - golden["expected_analysis_v1"]).max() < 1e-5


def test_analyze_too_short():
    # src/song/mod.rs:543-550 test_analysis_too_small / :417-430
    assert O.analyze(np.zeros(0, np.float32))[0] == 1
    assert O.analyze(np.zeros(8191, np.float32))[0] == 1
    assert O.analyze(np.zeros(8192, np.float32))[0] == 0


def test_analyze_batch_threads(pcm_song, pcm_piano):
    st, out = O.analyze_batch([pcm_song, pcm_piano, pcm_song[:5000], pcm_piano], 2, n_threads=3)
    assert list(st) == [0, 0, 1, 0]
    assert np.array_equal(out[0], O.analyze(pcm_song)[1])
    assert np.array_equal(out[1], out[3])


# ---- utils.rs ----------------------------------------------------------------

def test_compute_stft(pcm_piano, golden):
    # src/utils.rs:529-541 test_compute_stft: stft(piano, 2048, 512) vs librosa, tol 1e-4
    S = O.stft(pcm_piano, 2048, 512)
    exp = golden["librosa_stft"]
    assert S.shape == exp.shape == (1025, 253)
    assert np.abs(S - exp).max() < 1e-4


def test_reflect_pad():
    # src/utils.rs:544-551
    x = np.arange(100, dtype=np.float32)
    out = O.reflect_pad(x, 3)
    exp = np.r_[[3, 2, 1], np.arange(100), [98, 97, 96]].astype(np.float32)
    assert np.array_equal(out, exp)


def test_geometric_mean(golden):
    # src/utils.rs:238-514 test_geometric_mean
    assert O.geometric_mean(np.array([0, 1, 2, 3, 4, 5, 6, 7], np.float32)) == 0.0
    assert abs(2.0 - O.geometric_mean([4.0, 2.0, 1.0, 4.0, 2.0, 1.0, 2.0, 2.0])) < 1e-4
    assert abs(3.668016172818685 - O.geometric_mean([256.0, 4.0, 2.0, 1.0, 4.0, 2.0, 1.0, 2.0])) < 1e-4
    sub = np.array([4.0, 2.0, 1.0, 4.0, 2.0, 1.0, 2.0, 1.0e-40], np.float32)
    assert abs(1.8340080864093417e-05 - O.geometric_mean(sub)) < 1e-4
    big = np.float32(2.0) ** 65
    assert abs(O.geometric_mean(np.full(256, big, np.float32)) / big - 1) < 1e-5
    assert abs(0.0025750597 - O.geometric_mean(golden["geometric_mean_input"])) < 1e-8


def test_number_crossings_and_zcr(pcm_song):
    # src/timbral.rs:270-286 test_zcr_boundaries
    assert O.zcr(np.zeros(1024, np.float32)) == -1.0
    x = np.tile(np.array([-1.0, 1.0], np.float32), 512)
    assert abs(0.9980469 - O.zcr(x)) < 0.001
    # src/timbral.rs:288-297 test_zcr (chunks_exact(128) -> sums per chunk)
    n = (pcm_song.size // 128) * 128
    total = sum(O.number_crossings(pcm_song[i:i + 128]) for i in range(0, n, 128))
    val = 2.0 * (total / n) - 1.0
    assert abs(-0.85036 - val) < 0.001


# ---- timbral.rs --------------------------------------------------------------

def _spectral(pcm):
    nfr = pcm.size // 128  # the unit tests feed chunks_exact(HOP_SIZE)
    c, r, f = O.timbral_frames(pcm, nfr)
    return O.summarise(c, 0), O.summarise(r, 0), O.summarise(f, 1)


def test_spectral_descriptors(pcm_song):
    cent, roll, flat = _spectral(pcm_song)
    # src/timbral.rs:402-419 test_spectral_centroid, tol 1e-4
    assert np.abs(cent - [-0.75483, -0.87916887]).max() < 1e-4
    # src/timbral.rs:383-399 test_spectral_roll_off, tol 1e-2
    assert np.abs(roll - [-0.6326486, -0.7260933]).max() < 1e-2
    # src/timbral.rs:345-361 test_spectral_flatness, tol 1e-2
    assert np.abs(flat - [-0.77610075, -0.8148179]).max() < 1e-2


def test_spectral_boundaries():
    # src/timbral.rs:299-317, :363-381, :421-441: one all-zero hop -> (-1, -1)
    z = np.zeros(128, np.float32)
    c, r, f = O.timbral_frames(z, 1)
    for v, kind in ((c, 0), (r, 0), (f, 1)):
        assert np.abs(O.summarise(v, kind) - [-1.0, -1.0]).max() < 1e-7


# ---- misc.rs -----------------------------------------------------------------

def test_loudness(pcm_song):
    # src/misc.rs:85-95 test_loudness (chunks_exact), tol 1e-2
    assert np.abs(O.loudness(pcm_song, chunks_exact=True) - [0.271263, 0.2577181]).max() < 0.01


def test_loudness_boundaries():
    # src/misc.rs:98-122
    assert np.abs(O.loudness(np.zeros(1024, np.float32)) - [-1, -1]).max() < 1e-7
    assert np.abs(O.loudness(np.ones(1024, np.float32)) - [1, -1]).max() < 1e-7
    assert np.abs(O.loudness(-np.ones(1024, np.float32)) - [1, -1]).max() < 1e-7


# ---- temporal.rs -------------------------------------------------------------

def test_tempo_real(pcm_song):
    # src/temporal.rs:102-109 (chunks_exact(256)), tol 1e-2
    v = O.tempo(pcm_song, n_frames=pcm_song.size // 256, silence_len=256)
    assert abs(0.378605 - v) < 0.01


def test_tempo_artificial():
    # src/temporal.rs:122-138: 60 BPM click track -> -0.416853
    one = np.r_[np.zeros(22000), np.ones(100)].astype(np.float32)
    x = np.tile(one, 100)
    v = O.tempo(x, n_frames=x.size // 256, silence_len=256)
    assert abs(-0.416853 - v) < 0.01


def test_tempo_boundaries():
    # src/temporal.rs:141-162
    assert O.tempo(np.zeros(1024, np.float32), n_frames=1, silence_len=1024) == -1.0
    one = np.r_[np.zeros(6989), np.ones(20)].astype(np.float32)
    x = np.tile(one, 500)
    v = O.tempo(x, n_frames=x.size // 256, silence_len=256)
    assert abs(0.86 - v) < 0.01


# ---- chroma.rs ---------------------------------------------------------------

def test_chroma_interval_features(golden):
    # src/chroma.rs:498-509
    f = O.chroma_interval_features(golden["chroma"])
    assert np.abs(f - golden["expected_chroma_interval_features"]).max() < 1e-8


def test_extract_interval_features(golden):
    # src/chroma.rs:512-540, tol 1e-7
    out = O.extract_interval_features(golden["chroma_interval"])
    assert np.abs(out - golden["interval_feature_matrix"]).max() < 1e-7


def test_normalize_feature_sequence():
    # src/chroma.rs:543-557
    a = np.array([[0.1, 0.3, 0.4, 0.0], [1.1, 0.53, 1.01, 0.0]])
    exp = np.array([[0.08333333, 0.36144578, 0.28368794, 0.0],
                    [0.91666667, 0.63855422, 0.71631206, 0.0]])
    assert np.abs(O.normalize_feature_sequence(a) - exp).max() < 1e-7


def test_chroma_desc(pcm_song):
    # src/chroma.rs:571-593 (v2, first 10) and :595-619 (v1), tol 1e-7
    v2, tuning = O.chroma(pcm_song, 2)
    exp = [-0.34292513, -0.62803423, -0.28095096, 0.08686459, 0.24446082, -0.5723257,
           0.23292065, 0.19981146, -0.58594406, -0.06784296]
    assert np.abs(v2[:10] - np.array(exp, np.float32)).max() < 1e-7
    v1, _ = O.chroma(pcm_song, 1)
    exp1 = [-0.35661936, -0.63578653, -0.29593682, 0.06421304, 0.21852458, -0.581239,
            -0.9466835, -0.9481153, -0.9820945, -0.95968974]
    assert np.abs(v1 - np.array(exp1, np.float32)).max() < 1e-7
    # src/chroma.rs:657-665 test_estimate_tuning_decode
    assert abs(-0.04999999999999999 - tuning) < 1e-6


def test_chroma_end_result_on_silence():
    # src/chroma.rs:816-866 test_end_result_edge_cases, the digital-silence row (data/silence.ogg decodes to zeros):
    # every chroma column sums to 0 -> E = exp(0) L1-normalised = 1/12 -> all interval classes equal within their
    # group -> 2/sqrt(6) - 1 six times, 2/sqrt(4) - 1 = 0 four times; tol 1e-7 as in the reference
    v2, tuning = O.chroma(np.zeros(22050 * 3, np.float32), 2)
    exp = np.array([-0.18350339] * 6 + [0.0] * 4, np.float32)
    assert np.abs(v2[:10] - exp).max() < 1e-7 and tuning == 0.0


def test_chroma_stft_decode(pcm_song, golden):
    # src/chroma.rs:623-639, tol 1e-7
    S = O.stft(pcm_song, 8192, 2205)
    c = O.chroma_stft(S, 8192, -0.04999999999999999)
    assert c.shape == golden["chroma"].shape
    assert np.abs(c - golden["chroma"]).max() < 1e-7


def test_estimate_tuning(golden):
    # src/chroma.rs:642-648
    assert abs(-0.09999999999999998 - O.estimate_tuning(golden["spectrum_chroma"], 2048)) < 1e-6
    # src/chroma.rs:651-654 empty fix
    assert O.estimate_tuning(np.zeros((8192, 1)), 8192) == 0.0


def test_hz_to_octs_inplace():
    # src/utils.rs:517-525
    got = O.hz_to_octs([32.0, 64.0, 128.0, 256.0], 0.5, 10)
    assert np.abs(got - np.array([0.16864029, 1.16864029, 2.16864029, 3.16864029])).max() < 1e-4


def test_pitch_tuning(golden):
    # src/chroma.rs:668-679
    assert O.pitch_tuning(golden["pitch_tuning"], 0.05) == -0.1
    assert O.pitch_tuning(np.zeros(0), 0.05) == 0.0


def test_pip_track(golden):
    # src/chroma.rs:682-702, compared sorted, tol 1e-8
    p, m = O.pip_track(golden["spectrum_chroma"], 2048)
    assert p.size == golden["spectrum_chroma_pitches"].size == 772
    assert np.abs(np.sort(p) - golden["spectrum_chroma_pitches"]).max() < 1e-8
    assert np.abs(np.sort(m) - golden["spectrum_chroma_mags"]).max() < 1e-8


def test_chroma_filter(golden):
    # src/chroma.rs:705-714, tol 1e-9
    f = O.chroma_filter(2048, -0.1)
    assert np.abs(f - golden["chroma_filter"]).max() < 1e-9


# ---- playlist.rs / lib.rs ----------------------------------------------------

A20 = [1.0] * 19 + [0.0]


def test_euclidean_distance():
    # src/playlist.rs:1079-1092 (exact equality in the reference)
    b = [0.0] * 16 + [1.0, 0.0, 0.0, 0.0]
    assert O.euclidean_distance(A20, b) == np.float32(4.242640687119285)
    assert O.euclidean_distance([0.5] * 20, [0.5] * 20) == 0.0


def test_cosine_distance():
    # src/playlist.rs:1094-1108
    b = [0.0] * 16 + [1.0, 0.0, 0.0, 0.0]
    assert O.cosine_distance(A20, b) == np.float32(0.7705842661294382)
    assert O.cosine_distance([0.5] * 20, [0.5] * 20) == 0.0


def test_mahalanobis_distance():
    # src/playlist.rs:1009-1024
    b = [1.0] + [0.0] * 15 + [1.0, 0.0, 0.0, 0.0]
    m = np.diag([1.0, 1.0] + [0.0] * 18).astype(np.float32)
    assert O.mahalanobis_distance(A20, b, m) == 1.0


def test_distance_metric_features_version():
    # src/lib.rs:273-291 (exact equality in the reference)
    assert O.default_distance(np.zeros(20), np.ones(20), 1) == np.float32(4.47213595)
    assert O.default_distance(np.zeros(23), np.ones(23), 2) == np.float32(3.4999998)
    # src/lib.rs:262-271
    assert O.feature_weights(1).shape == (20, 20) and O.feature_weights(2).shape == (23, 23)


def test_closest_to_songs_order():
    # src/playlist.rs:1026-1076 test_mahalanobis_distance_with_songs
    first = np.ones(23, np.float32)
    second = np.array([1.5, 5, 6, 5, 6, 6] + [1.0] * 17, np.float32)
    third = np.array([5.0] + [1.0] * 22, np.float32)
    m = np.diag([1.0] + [0.0] * 22).astype(np.float32)
    order, keys = O.closest_to_songs([first], [third, second], m)
    assert list(order) == [1, 0]
    # stable on ties (sort_by_cached_key is stable, src/playlist.rs:267-268)
    order, _ = O.closest_to_songs([first], [third, second, third, second], m)
    assert list(order) == [1, 3, 0, 2]


def test_song_to_song_chain():
    # src/playlist.rs:272-326 semantics: greedy nearest neighbour from the seed
    pts = np.array([[0.0], [10.0], [1.0], [3.0], [2.5]], np.float32)
    order = O.song_to_song(pts[:1], pts[1:])
    assert list(order) == [1, 3, 2, 0]


def test_fft_matches_numpy():
    rng = np.random.default_rng(0)
    for n in (512, 2048, 8192):
        z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        ref = np.fft.fft(z.astype(np.complex128))
        err = np.abs(O.fft(z) - ref).max() / np.abs(ref).max()
        assert err < 2e-6


# ---- song/decoder/*.rs: sample format + down-mix of sources already at 22 050 Hz ----------

def _adler32_f32(x):
    import zlib
    return zlib.adler32(np.asarray(x, "<f4").tobytes()) & 0xFFFFFFFF


def test_pcm_to_mono_decoder_hashes(golden):
    # src/song/decoder/ffmpeg.rs:454-462 test_decode_mono, :447-452 test_resample_stereo, :523-527 test_decode_wav:
    # the adler32 of the decoder's f32le output
    assert _adler32_f32(O.pcm_to_mono(golden["pcm_s16_mono"])) == 0x5E01930B
    assert _adler32_f32(O.pcm_to_mono(golden["pcm_piano"])) == 0xDE831E82
    st = golden["pcm_s16_stereo"]
    assert st.shape[1] == 2
    mono = O.pcm_to_mono(st)
    assert _adler32_f32(mono) == 0x1D7B2D6D == int(golden["adler32_stereo_downmix"])
    # "averaging the channels and multiplying by the square root of 2" (src/song/decoder/symphonia.rs:260-287) agrees
    # with it within the f32::EPSILON mean difference of compare_ffmpeg_to_symphonia_for_all_test_songs (:703-709)
    f = st.astype(np.float32) / np.float32(32768.0)
    sym = (f[:, 0] + f[:, 1]) * np.float32(np.sqrt(2.0)) / np.float32(2.0)
    assert np.abs(sym - mono).mean() < np.finfo(np.float32).eps


def test_pcm_to_mono_formats():
    rng = np.random.default_rng(5)
    k15, k31 = np.float32(2.0 ** -15), np.float32(2.0 ** -31)
    c = np.float32(np.sqrt(0.5))
    s16 = rng.integers(-32768, 32768, (1000, 2), dtype=np.int16)
    want = (s16[:, 0].astype(np.float32) * k15) * c + (s16[:, 1].astype(np.float32) * k15) * c
    assert np.array_equal(O.pcm_to_mono(s16), want)
    s32 = rng.integers(-2 ** 31, 2 ** 31, (1000, 1), dtype=np.int64).astype(np.int32)
    assert np.array_equal(O.pcm_to_mono(s32), s32[:, 0].astype(np.float32) * k31)
    # more than two channels: chunk.iter().sum::<f32>() / n (src/song/decoder/symphonia.rs:289-299)
    f6 = rng.standard_normal((777, 6)).astype(np.float32)
    acc = np.zeros(777, np.float32)
    for ch in range(6):
        acc = acc + f6[:, ch]
    assert np.array_equal(O.pcm_to_mono(f6), acc / np.float32(6))
    assert np.array_equal(O.pcm_to_mono(f6[:, 0].copy()), f6[:, 0])
