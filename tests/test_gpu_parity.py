"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference's
golden vectors.  Needs a B200: every test is marked gpu.

Tolerances (stated per BASELINE.json north_star: "feature vectors within 1e-4 rel"):
features live in [-1, 1], so the bar used is |gpu - oracle| <= 1e-4 * max(1, |oracle|) per
feature; the golden clip is additionally held to the reference's own 1e-5.
Discrete stages (roll-off bin, tuning bin, BPM list) are compared exactly or by mismatch rate.
"""
import numpy as np
import pytest
import torch

import bliss_rs_b200 as B
from bliss_rs_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-4

# This suite runs on a B200.  The CPU suite also runs a slice of it WITHOUT a GPU against the host-emulated TEST
# build of the library (tests/test_host_abi.py::test_c_abi_on_the_host_emulated_library: the same sources compiled
# with g++ against tests/cpu_emul/cuda_on_cpu, BLISS_B200_SO pointing at it): "device" buffers are host tensors then
# and the synthetic extras shrink, because every CUDA thread is a fiber there.
DEV = "cuda" if torch.cuda.is_available() else "cpu"
EMULATED = DEV == "cpu"


def _sync():
    if DEV == "cuda":
        torch.cuda.synchronize()


def _stream():
    return torch.cuda.current_stream().cuda_stream if DEV == "cuda" else None


def _extra_tracks(seed, count, seconds, step):
    """synthetic tracks that widen a test's batch (one short one under the host emulation)"""
    if EMULATED:
        count, seconds = 1, 3
    return [synth.gen_track(seed, i, 22050 * seconds + step * i, device=DEV).cpu().numpy() for i in range(count)]


def _close(got, want, tol=TOL):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.abs(got - want) <= tol * np.maximum(1.0, np.abs(want))


def _report(name, got, want):
    err = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))
    print("%s: max abs err %.3e (feature %d)" % (name, err.max(), int(err.argmax())))


@pytest.fixture(scope="module", autouse=True)
def _init():
    B.native.init(0)
    yield


# ---------------------------------------------------------------- golden clip
def test_golden_clip_v2(pcm_song, golden):
    a = B.Song.analyze(pcm_song).as_arr1()
    _report("gpu vs reference golden (v2)", a, golden["expected_analysis_v2"])
    assert np.abs(a - golden["expected_analysis_v2"]).max() < 1e-5  # src/song/mod.rs:582-591
    rc, o = O.analyze(pcm_song, 2)
    assert _close(a, o).all()


def test_golden_clip_v1(pcm_song, golden):
    opts = B.AnalysisOptions(features_version=B.FeaturesVersion.Version1)
    a = B.Song.analyze_with_options(pcm_song, opts)
    assert a.features_version == B.FeaturesVersion.Version1 and a.as_arr1().shape == (20,)
    assert np.abs(a.as_arr1() - golden["expected_analysis_v1"]).max() < 1e-5  # src/song/mod.rs:619-633


def test_too_short_is_an_error(pcm_song):
    # src/song/mod.rs:543-550
    for n in (0, 1, 8191):
        with pytest.raises(B.AnalysisError) as e:
            B.Song.analyze(pcm_song[:n])
        assert "empty or too short song." in str(e.value)
    B.Song.analyze(pcm_song[:8192])


# ---------------------------------------------------------------- stage by stage
def _stage_checks(pcm, label):
    rc, feats, t = B.native.analyze_taps(pcm, 2)
    assert rc == 0
    c, r, f = O.timbral_frames(pcm)
    tv, flux, thr, bpms = O.tempo(pcm, taps=True)
    # timbral per-frame values
    assert t["centroid"].shape == c.shape
    ce = np.abs(t["centroid"] - c) / np.maximum(1.0, np.abs(c))
    print(label, "centroid max rel err %.2e" % ce.max())
    assert ce.max() < 1e-4
    roll_mismatch = np.mean(t["rolloff"] != r)
    print(label, "rolloff mismatching frames: %.4f%%" % (100 * roll_mismatch))
    assert roll_mismatch < 2e-3 and np.abs(t["rolloff"] - r).max() <= 2 * 22050 / 512 + 1e-3
    # flatness = geometric / arithmetic mean: its weakest bins sit at the f32 FFT's own noise
    # floor (eps * frame peak) in BOTH implementations, so tonal frames get an absolute floor
    fe = np.abs(t["flatness"] - f) / (2e-4 * np.abs(f) + 1e-5)
    print(label, "flatness max err / (2e-4*|f| + 1e-5) = %.2f" % fe.max())
    assert fe.max() < 1.0
    # tempo chain
    fl = np.abs(t["flux"] - flux) / np.maximum(1e-3, np.abs(flux).max())
    print(label, "flux max err (rel. to max) %.2e" % fl.max())
    assert fl.max() < 1e-5
    te = np.abs(t["thresholded"] - thr) / np.maximum(1e-3, np.abs(thr).max())
    assert te.max() < 1e-5
    assert len(t["bpms"]) == len(bpms), (len(t["bpms"]), len(bpms))
    if len(bpms):
        assert np.abs(np.sort(t["bpms"]) - np.sort(bpms)).max() < 1e-2
    # time domain
    assert int(t["zero_crossings"][0]) == O.number_crossings(pcm)
    # chroma chain
    S = O.stft(pcm, 8192, 2205)  # [bins, frames]
    se = np.abs(t["stft8192"].T - S).max() / S.max()
    print(label, "stft8192 max err (rel. to max) %.2e" % se)
    assert se < 2e-6
    ochroma, otuning, ocm = O.chroma(pcm, 2, want_chroma=True)
    assert abs(t["tuning"][0] - otuning) < 1e-9, (t["tuning"][0], otuning)
    p, m = O.pip_track(S, 8192)
    assert abs(int(t["n_peaks"][0]) - p.size) <= max(2, p.size // 2000)
    # pip_track's own arithmetic (src/chroma.rs:269-331; the reference pins pitches and magnitudes at 1e-8 sorted,
    # :682-702): the oracle run on the DEVICE's magnitudes must give the device's candidates
    pd, md = O.pip_track(t["stft8192"].T.astype(np.float64), 8192)
    assert pd.size == int(t["n_peaks"][0]) == t["peak_pitches"].size
    if pd.size:
        pe = np.abs(np.sort(t["peak_pitches"]) - np.sort(pd)).max()
        me = np.abs(np.sort(t["peak_mags"]) - np.sort(md)).max() / max(1e-30, np.abs(md).max())
        print(label, "pip_track on the device's magnitudes: pitch err %.2e Hz, mag err (rel. to max) %.2e" % (pe, me))
        assert pe < 1e-8 and me < 1e-12
    # loudness chunks (level_lin per chunk, tail chunk kept: src/misc.rs:12-18, src/song/mod.rs:478)
    ms = np.array([np.mean(pcm[i:i + 1024].astype(np.float64) ** 2) for i in range(0, pcm.size, 1024)])
    assert t["loudness_chunks"].shape == ms.shape and np.abs(t["loudness_chunks"] - ms).max() <= 1e-6 * max(ms.max(), 1e-30)
    ch = np.abs(t["chroma"].T - ocm).max()
    print(label, "chroma max abs err %.2e" % ch)
    assert ch < 2e-5
    oif = O.chroma_interval_features(ocm)
    assert np.abs(t["interval_features"] - oif).max() < 1e-6
    rc, o = O.analyze(pcm, 2)
    _report(label + " features", feats, o)
    assert _close(feats, o).all()


def test_stages_golden_clip(pcm_song):
    _stage_checks(pcm_song, "golden")


def test_stages_piano(pcm_piano):
    _stage_checks(pcm_piano, "piano")


def test_stages_synthetic_music():
    x = synth.gen_track(7, 3, 22050 * 20).numpy()
    _stage_checks(x, "synth")


def test_digital_silence_transitions():
    # sound -> exact zeros -> sound: frames that are exactly silent sit next to loud ones, which is
    # where packing two frames into one complex FFT would leak rounding noise
    x = synth.gen_track(5, 2, 22050 * 12).numpy()
    x[22050 * 3:22050 * 5] = 0.0
    x[22050 * 8 + 77:22050 * 9 + 1234] = 0.0
    x[-30000:] = 0.0
    _stage_checks(x, "silence-gaps")


# ---------------------------------------------------------------- batches
def test_ragged_batch_with_bad_songs(pcm_song, pcm_piano):
    songs = [pcm_song, pcm_piano[:5000], pcm_piano, np.zeros(0, np.float32), pcm_song[:100003],
             synth.gen_track(1, 0, 22050 * 9 + 17).numpy()]
    res = B.analyze_batch(songs)
    ost, ofe = O.analyze_batch(songs, 2, n_threads=4)
    for i, (r, st) in enumerate(zip(res, ost)):
        if st == 1:
            assert isinstance(r, B.AnalysisError)
        else:
            assert isinstance(r, B.Analysis), r
            _report("batch song %d" % i, r.as_arr1(), ofe[i])
            assert _close(r.as_arr1(), ofe[i]).all()
    # order independence / no cross-talk between songs of a wave
    res2 = B.analyze_batch(list(reversed(songs)))
    for a, b in zip(res, reversed(res2)):
        if isinstance(a, B.Analysis):
            assert np.array_equal(a.as_arr1(), b.as_arr1())


def test_synthetic_corpus_parity_and_flip_rate():
    n_tracks = 12
    lens = [22050 * 30 + 1000 * i for i in range(n_tracks)]
    songs = [synth.gen_track(11, i, n).numpy() for i, n in enumerate(lens)]
    st, feats = B.native.analyze_batch(songs, 2)
    ost, ofe = O.analyze_batch(songs, 2, n_threads=8)
    assert (st == 0).all() and (ost == 0).all()
    err = np.abs(feats - ofe)
    print("corpus max abs err per feature:", np.array2string(err.max(axis=0), precision=1))
    tempo_flips = int((err[:, 0] > 1e-3).sum())
    print("tempo decisions differing: %d / %d" % (tempo_flips, n_tracks))
    assert tempo_flips == 0
    assert _close(feats, ofe).all()


def test_version1_batch(pcm_song, pcm_piano):
    st, feats = B.native.analyze_batch([pcm_song, pcm_piano], 1)
    ost, ofe = O.analyze_batch([pcm_song, pcm_piano], 1)
    assert feats.shape == (2, 20) and _close(feats, ofe).all()


def test_small_workspace_forces_many_waves(pcm_song, pcm_piano):
    songs = [pcm_song, pcm_piano, pcm_song[:60000], pcm_piano[:90001]] * 3
    st0, f0 = B.native.analyze_batch(songs, 2)
    B.native.check(B.native.lib().bliss_b200_set_workspace_limit(24 << 20))
    try:
        st1, f1 = B.native.analyze_batch(songs, 2)
    finally:
        B.native.check(B.native.lib().bliss_b200_set_workspace_limit(60 << 30))
    assert np.array_equal(st0, st1) and np.array_equal(f0, f1)


# ---------------------------------------------------------------- edge cases of the reference tests
def test_silence_and_constant():
    z = np.zeros(22050 * 3, np.float32)
    a = B.Song.analyze(z).as_arr1()
    rc, o = O.analyze(z, 2)
    _report("silence", a, o)
    assert a[0] == -1.0 and a[1] == -1.0  # no beats (temporal.rs:66-70), no crossings
    assert np.allclose(a[2:8], -1.0) and np.allclose(a[8:10], -1.0)  # timbral.rs / misc.rs boundaries
    assert _close(a, o).all()
    # the reference's own known answer for silence (src/chroma.rs:816-866), its tolerance
    assert np.abs(a[10:20] - np.array([-0.18350339] * 6 + [0.0] * 4, np.float32)).max() < 1e-7
    ones = np.ones(22050 * 2, np.float32)
    a = B.Song.analyze(ones).as_arr1()
    rc, o = O.analyze(ones, 2)
    assert abs(a[8] - 1.0) < 1e-6 and abs(a[9] + 1.0) < 1e-6  # misc.rs:98-122
    assert _close(a, o).all()


def test_white_noise_and_alternating():
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(22050 * 10) * 0.2).astype(np.float32)
    a = B.Song.analyze(x).as_arr1()
    rc, o = O.analyze(x, 2)
    _report("white noise", a, o)
    ok = _close(a, o)
    assert ok[1:].all(), (a, o)   # everything but tempo must agree; noise has no stable beat
    alt = np.tile(np.array([-1.0, 1.0], np.float32), 22050)
    a = B.Song.analyze(alt).as_arr1()
    rc, o = O.analyze(alt, 2)
    assert _close(a[1:], o[1:]).all() and a[1] > 0.99  # zcr ~ 1 (timbral.rs:279-285)


@pytest.mark.parametrize("zeros,ones,reps,expected", [(22000, 100, 100, -0.416853), (6989, 20, 500, 0.86)])
def test_tempo_click_tracks(zeros, ones, reps, expected):
    # src/temporal.rs:122-138 and :141-162 (here through the analyze framing)
    one = np.r_[np.zeros(zeros), np.ones(ones)].astype(np.float32)
    x = np.tile(one, reps)
    a = B.Song.analyze(x).as_arr1()
    o = O.tempo(x)
    print("click track tempo gpu %.6f oracle %.6f expected %.6f" % (a[0], o, expected))
    assert abs(a[0] - o) < 1e-4 and abs(a[0] - expected) < 0.02


# ---------------------------------------------------------------- full-size property checks
def test_three_minute_track_matches_oracle():
    n = 3969000  # BASELINE config: 3 min at 22 050 Hz
    x = synth.gen_track(3, 1, n).numpy()
    a = B.Song.analyze(x).as_arr1()
    rc, o = O.analyze(x, 2)
    _report("3-min track", a, o)
    assert _close(a, o).all()
    # gain invariance of every descriptor except loudness (linearity of the STFT chain)
    b = B.Song.analyze((x * np.float32(0.5)).astype(np.float32)).as_arr1()
    keep = [i for i in range(23) if i not in (8, 9)]
    assert np.abs(a[keep] - b[keep]).max() < 2e-4
    # idempotence: same input, same bits
    assert np.array_equal(a, B.Song.analyze(x).as_arr1())


# ---------------------------------------------------------------- BASELINE.json config 5 shape
def test_large_batch_is_bitwise_reproducible():
    """1024 tracks keep every SM busy with several kernels at once: results must not depend on how
    warps happen to be scheduled (a barrier-free read of beat-tracker state once flipped ~2 % of tempos)."""
    dev = torch.device("cuda", 0)
    n = 1024
    pcm, offs, lens = synth.gen_corpus_flat(4242, list(range(n)), [22050 * 30] * n, device=dev)
    runs = []
    for _ in range(3):
        out = torch.zeros((n, 23), device=dev)
        st = B.native.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, out.data_ptr())
        _sync()
        assert (st == 0).all()
        runs.append(out)
    for r in runs[1:]:
        diff = (r != runs[0])
        assert not diff.any(), "%d values differ between runs (columns %s)" % (
            int(diff.sum()), diff.any(0).nonzero().flatten().tolist())
    # and the batch agrees with the same songs analysed in small groups
    small = torch.zeros((16, 23), device=dev)
    B.native.analyze_batch_device(pcm.data_ptr(), offs[500:516], lens[500:516], 2, small.data_ptr())
    _sync()
    assert torch.equal(small, runs[0][500:516])
    rc, o = O.analyze(pcm[offs[1000]:offs[1000] + lens[1000]].cpu().numpy(), 2)
    assert _close(runs[0][1000].cpu().numpy(), o).all()


def test_mixed_duration_corpus_and_playlist_order():
    """30 s .. 10 min tracks (Zipf-like), one call; then playlist-from-seed = closest_to_songs
    ([seed], all, v2 metric) must come out in the oracle's order (stable on ties)."""
    secs = [30, 30, 45, 600, 60, 30, 210, 120, 30, 75, 30, 300]
    songs = [synth.gen_track(23, i, 22050 * s_ + 37 * i, device=DEV).cpu().numpy() for i, s_ in enumerate(secs)]
    st, feats = B.native.analyze_batch(songs, 2)
    ost, ofe = O.analyze_batch(songs, 2, n_threads=12)
    assert (st == 0).all() and (ost == 0).all()
    err = np.abs(feats - ofe)
    print("mixed-duration corpus: max abs err %.2e (song %d, feature %d)" % (err.max(), *np.unravel_index(err.argmax(), err.shape)))
    assert _close(feats, ofe).all()
    w = O.feature_weights(2)
    order, keys = B.native.closest_to_songs(feats[:1], feats, 0, w)
    oorder, okeys = O.closest_to_songs(ofe[:1], ofe, w)
    assert order[0] == 0 and list(order) == list(oorder)
    # and the sharding used at N>1 keeps every song exactly once, balanced by duration
    from bliss_rs_b200 import multigpu as M
    shards = M.shard_longest_first([len(s_) for s_ in songs], 4)
    assert sorted(i for sh in shards for i in sh) == list(range(len(songs)))


def test_song_longer_than_2_pow_24_samples():
    """utils.rs:30 computes the chroma frame count in f32 ((len as f32 / 2205.).ceil()), exact only below
    2^24 samples (~12.7 min).  A 13.2-minute track exercises the same f32 formula on both sides."""
    n = (1 << 24) + 700001
    x = synth.gen_track(29, 5, n, device=DEV).cpu().numpy()
    a = B.Song.analyze(x).as_arr1()
    rc, o = O.analyze(x, 2)
    _report("13-min track", a, o)
    assert rc == 0 and _close(a, o).all()


# ---------------------------------------------------------------- device-resident API + STFT micro-benchmark path
def test_device_api_matches_host_api(pcm_song, pcm_piano):
    songs = [pcm_song, pcm_piano, pcm_song[:7000], pcm_piano[:50001]]
    offs, total = [], 0
    for s in songs:
        offs.append(total)
        total += (len(s) + 3) // 4 * 4
    flat = torch.zeros(total, dtype=torch.float32, device=DEV)
    for o, s in zip(offs, songs):
        flat[o:o + len(s)] = torch.from_numpy(s).to(DEV)
    out = torch.full((len(songs), 23), 7.0, dtype=torch.float32, device=DEV)
    st = B.native.analyze_batch_device(flat.data_ptr(), offs, [len(s) for s in songs], 2, out.data_ptr(),
                                       _stream())
    _sync()
    hst, hfe = B.native.analyze_batch(songs, 2)
    assert list(st) == list(hst) == [0, 0, 1, 0]
    got = out.cpu().numpy()
    assert np.array_equal(got[[0, 1, 3]], hfe[[0, 1, 3]]) and (got[2] == 0).all()


def test_s16_ingest_is_the_f32_path_on_converted_samples(golden, pcm_song, pcm_piano):
    """bliss_b200_analyze_batch_s16: the decoder's 16-bit samples, converted on the device exactly as ffmpeg's
    s16 -> flt step does (x / 32768; the fixture reproduces the reference's decoder test bit for bit)."""
    s16 = [golden["pcm_s16_mono"], golden["pcm_piano"], golden["pcm_s16_mono"][:5000], golden["pcm_piano"][:60001]]
    st16, f16 = B.native.analyze_batch_s16(s16, 2)
    st32, f32 = B.native.analyze_batch([x.astype(np.float32) / np.float32(32768.0) for x in s16], 2)
    assert list(st16) == [0, 0, 1, 0] and np.array_equal(st16, st32)
    assert np.array_equal(f16[st16 == 0], f32[st32 == 0])
    assert np.abs(f16[0] - golden["expected_analysis_v2"]).max() < 1e-5  # src/song/mod.rs:553-591
    st1, f1 = B.native.analyze_batch_s16(s16[:2], 1)
    assert np.abs(f1[0] - golden["expected_analysis_v1"]).max() < 1e-5


def test_kernel_implementations_agree(pcm_song, pcm_piano):
    """bliss_b200_set_variant: the previous implementation of every reworked kernel is still in the library.
    Tuning select, chroma contraction, autocorrelation and beat-tracker CTA width must reproduce the current
    kernels BIT FOR BIT; the other cuts of the 8192-point FFT (previous epilogue, round-1 kernel, one column per thread) round
    differently in the last place and must agree to 1e-5 with identical tuning / tempo decisions (so must the round-1
    chroma STFT behind bit 32768 and the one-column one behind bit 131072)."""
    songs = [pcm_song, pcm_piano] + _extra_tracks(77, 6, 40, 101)
    try:
        B.native.set_variant(0)
        st0, f0 = B.native.analyze_batch(songs, 2)
        assert (st0 == 0).all()
        for mask in (2, 4, 8, 16, 2 | 4 | 8 | 16):
            B.native.set_variant(mask)
            st, f = B.native.analyze_batch(songs, 2)
            assert np.array_equal(f, f0), "variant %d differs in columns %s" % (mask, np.nonzero((f != f0).any(0))[0])
        for mask in (1, 31, STFT_V1, STFT_V3):
            B.native.set_variant(mask)
            st, f = B.native.analyze_batch(songs, 2)
            assert (st == 0).all()
            assert np.abs(f - f0).max() < 1e-5, (mask, np.abs(f - f0).max(0))
            assert np.array_equal(f[:, 0], f0[:, 0])  # tempo does not depend on the chroma STFT at all
    finally:
        B.native.set_variant(0)


def test_cue_style_subslices_of_one_buffer(pcm_song):
    """BlissCueFile::get_songs (src/cue.rs:208-243) analyses sub-slices of ONE decoded buffer cut at
    (start_s * 22050) as usize -- arbitrary, unaligned sample offsets; they may even overlap."""
    d = torch.from_numpy(np.ascontiguousarray(pcm_song)).to(DEV)
    cuts = [(0, 88201), (88201, 176403), (100003, 230001), (176403, len(pcm_song))]
    offs = [a for a, b in cuts]
    lens = [b - a for a, b in cuts]
    out = torch.zeros((len(cuts), 23), dtype=torch.float32, device=DEV)
    st = B.native.analyze_batch_device(d.data_ptr(), offs, lens, 2, out.data_ptr(), None)
    _sync()
    assert (st == 0).all()
    got = out.cpu().numpy()
    for i, (a, b) in enumerate(cuts):
        rc, want = O.analyze(pcm_song[a:b], 2)
        assert rc == 0 and _close(got[i], want).all(), (i, np.abs(got[i] - want).max())


def test_stft512_magnitudes(pcm_song):
    x = pcm_song[:60000]
    n_t = (len(x) - 512) // 256 + 1
    d = torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    mags = torch.zeros((n_t, 257), dtype=torch.float32, device=DEV)
    fo = B.native.stft512_mag_device(d.data_ptr(), [0], [len(x)], mags.data_ptr(), None)
    _sync()
    assert list(fo) == [0, n_t]
    want = O.tempo_norms(x)
    err = np.abs(mags.cpu().numpy() - want).max() / want.max()
    print("stft512 max err (rel. to max) %.2e" % err)
    assert err < 1e-6


# ---------------------------------------------------------------- distances (bit-exact vs reference tests / oracle)
def test_distance_known_answers():
    a20 = np.array([1.0] * 19 + [0.0], np.float32)
    b = np.array([0.0] * 16 + [1.0, 0.0, 0.0, 0.0], np.float32)
    assert B.playlist.euclidean_distance(a20, b) == float(np.float32(4.242640687119285))  # playlist.rs:1079-1092
    assert B.playlist.euclidean_distance([0.5] * 20, [0.5] * 20) == 0.0
    assert B.playlist.cosine_distance(a20, b) == float(np.float32(0.7705842661294382))    # playlist.rs:1094-1108
    b2 = np.array([1.0] + [0.0] * 15 + [1.0, 0.0, 0.0, 0.0], np.float32)
    m = np.diag([1.0, 1.0] + [0.0] * 18).astype(np.float32)
    assert B.playlist.mahalanobis_distance(a20, b2, m) == 1.0                              # playlist.rs:1009-1024
    # lib.rs:273-291
    assert B.FeaturesVersion.Version1.distance_metric()(np.zeros(20), np.ones(20)) == float(np.float32(4.47213595))
    assert B.FeaturesVersion.Version2.distance_metric()(np.zeros(23), np.ones(23)) == float(np.float32(3.4999998))
    # song/mod.rs:772-807
    s1 = B.Song(analysis=B.Analysis(np.zeros(20), B.FeaturesVersion.Version1))
    s2 = B.Song(analysis=B.Analysis(np.ones(20), B.FeaturesVersion.Version1))
    assert s1.distance(s2) == float(np.float32(4.472136))


def test_distance_matrix_bit_exact_vs_oracle():
    rng = np.random.default_rng(0)
    for dim, ver in ((23, 2), (20, 1)):
        a = rng.uniform(-1, 1, (70, dim)).astype(np.float32)
        b = rng.uniform(-1, 1, (133, dim)).astype(np.float32)
        a[3] = b[5]  # an exact duplicate
        a[4] = b[6] + np.float32(1e-4)  # a near duplicate (dedup threshold territory)
        w = O.feature_weights(ver)
        full = (w + 0.01 * rng.uniform(0, 1, w.shape)).astype(np.float32)
        full = ((full + full.T) / 2).astype(np.float32)
        for name, metric, m in (("weights", 0, w), ("euclid", 0, None), ("full", 0, full), ("cosine", 2, None)):
            got = B.native.distance_matrix(a, b, metric, m)
            want = np.zeros_like(got)
            for i in range(a.shape[0]):
                for j in range(b.shape[0]):
                    want[i, j] = (O.cosine_distance(a[i], b[j]) if metric == 2 else
                                  O.euclidean_distance(a[i], b[j]) if m is None else
                                  O.mahalanobis_distance(a[i], b[j], m))
            assert np.array_equal(got, want), (name, dim, np.abs(got - want).max())
        assert B.native.distance_matrix(a, b, 0, None)[3, 5] == 0.0


def test_closest_to_songs_and_song_to_song():
    # playlist.rs:1026-1076 test_mahalanobis_distance_with_songs
    first = B.Song(path="first", analysis=B.Analysis(np.ones(23)))
    second = B.Song(path="second", analysis=B.Analysis(np.array([1.5, 5, 6, 5, 6, 6] + [1.0] * 17)))
    third = B.Song(path="third", analysis=B.Analysis(np.array([5.0] + [1.0] * 22)))
    dist = B.playlist.mahalanobis_distance_builder(np.diag([1.0] + [0.0] * 22))
    assert [s.path for s in B.playlist.closest_to_songs([first], [third, second], dist)] == ["second", "third"]
    rng = np.random.default_rng(1)
    seeds = rng.uniform(-1, 1, (3, 23)).astype(np.float32)
    cands = rng.uniform(-1, 1, (2000, 23)).astype(np.float32)
    cands[100] = cands[50]
    cands[1999] = cands[50]  # ties must keep input order (stable sort, playlist.rs:267-268)
    w = O.feature_weights(2)
    order, keys = B.native.closest_to_songs(seeds, cands, 0, w)
    oorder, okeys = O.closest_to_songs(seeds, cands, w)
    assert np.array_equal(keys, okeys) and np.array_equal(order, oorder)
    pos = {int(v): i for i, v in enumerate(order)}
    assert pos[50] < pos[100] < pos[1999]
    so = B.native.song_to_song(seeds[:1], cands[:300], 0, w)
    assert np.array_equal(so, O.song_to_song(seeds[:1], cands[:300], w))
    pts = [B.Analysis(np.r_[v, np.zeros(22)]) for v in (10.0, 1.0, 3.0, 2.5)]
    chain = B.playlist.song_to_song([B.Analysis(np.zeros(23))], pts)
    assert [float(c.as_arr1()[0]) for c in chain] == [1.0, 2.5, 3.0, 10.0]


def test_dedup_playlist():
    # playlist.rs:343-402: consecutive near-duplicates (< 0.05) and same title+artist are dropped
    def song(v, title=None, artist=None):
        return B.Song(title=title, artist=artist, analysis=B.Analysis(np.full(23, v, np.float32)))
    pl = [song(0.0), song(0.001), song(0.5, "t", "a"), song(0.9, "t", "a"), song(0.9), song(0.0)]
    kept = list(B.playlist.dedup_playlist(pl))
    assert [round(float(s.analysis.as_arr1()[0]), 6) for s in kept] == [0.0, 0.5, 0.9, 0.0]


def test_pcm_feed_matches_the_decoders_conversion(golden):
    """bliss_b200_pcm_to_mono / bliss_b200_analyze_batch_pcm: interleaved s16 / s32 / f32 frames at 22 050 Hz,
    converted and down-mixed on the device like the reference's decoders do (src/song/decoder/ffmpeg.rs:36-109,
    symphonia.rs:260-300).  Bit-exact against the oracle and against the adler32 values the reference's decoder
    tests assert for ffmpeg's output (ffmpeg.rs:447-462)."""
    import zlib

    def adler(x):
        return zlib.adler32(np.asarray(x, "<f4").tobytes()) & 0xFFFFFFFF

    st = golden["pcm_s16_stereo"]
    mono = B.native.pcm_to_mono(st)
    assert adler(mono) == 0x1D7B2D6D                                   # ffmpeg.rs:447-452 test_resample_stereo
    assert adler(B.native.pcm_to_mono(golden["pcm_s16_mono"])) == 0x5E01930B  # :454-462 test_decode_mono
    rng = np.random.default_rng(11)
    cases = [rng.integers(-32768, 32768, (50001, ch), dtype=np.int16) for ch in (1, 2, 3, 6, 8)]
    cases += [rng.integers(-2 ** 31, 2 ** 31, (40003, ch), dtype=np.int64).astype(np.int32) for ch in (1, 2, 5)]
    cases += [rng.standard_normal((30002, ch)).astype(np.float32) for ch in (1, 2, 4)]
    cases += [st[:1], st[:0]]
    for a in cases:
        got, want = B.native.pcm_to_mono(a), O.pcm_to_mono(a)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (a.dtype, a.shape)
    # the analysis of such sources = the f32 path on the converted samples, bit for bit
    songs = [st, st[:100000], st[:5000], np.repeat(golden["pcm_piano"][:, None], 2, axis=1)]
    st_p, f_p = B.native.analyze_batch_pcm(songs, 22050, 2)
    st_f, f_f = B.native.analyze_batch([O.pcm_to_mono(x) for x in songs], 2)
    assert list(st_p) == [0, 0, 1, 0] and np.array_equal(st_p, st_f)
    assert np.array_equal(f_p[st_p == 0], f_f[st_f == 0])
    rc, want = O.analyze(O.pcm_to_mono(st), 2)
    assert rc == 0 and _close(f_p[0], want).all()
    s32 = [(x.astype(np.int32) << 16) for x in songs[:2]]          # the same samples as 32-bit material
    st_q, f_q = B.native.analyze_batch_pcm(s32, 22050, 1)
    st_g, f_g = B.native.analyze_batch([O.pcm_to_mono(x) for x in songs[:2]], 1)
    assert np.array_equal(f_q, f_g)
    six = [np.ascontiguousarray(np.repeat(golden["pcm_s16_mono"][:, None], 6, axis=1))]
    st_6, f_6 = B.native.analyze_batch_pcm(six, 22050, 2)
    st_m, f_m = B.native.analyze_batch([O.pcm_to_mono(six[0])], 2)
    assert np.array_equal(f_6, f_m)
    with pytest.raises(B.native.NativeError, match="sample rate"):
        B.native.analyze_batch_pcm(songs[:1], 999, 2)
    with pytest.raises(B.native.NativeError):
        B.native.analyze_batch_pcm([np.zeros((9000, 9), np.int16)], 22050, 2)


RESAMPLE_V1 = 524288  # VARIANT_RESAMPLE_V1, common.cuh


def test_resample_kernel_against_the_oracle():
    """bliss_b200_resample: the polyphase resampler of the decode-side feed (wave_setup.cu resample_kernel) against
    oracle/resample.py (scipy.signal.resample_poly's algorithm, pinned against scipy in tests/test_resample.py) --
    PARITY UNPINNED against the reference's swresample / rubato, see the header.  Same f32 coefficients, f32 FMA sums
    on the device against f64 sums in the oracle: 2e-6 x the signal's peak.  Rates: the common ones, up- and
    down-sampling, a pair with a large `up`; lengths: empty, one sample, shorter than the filter, ragged."""
    from oracle import resample as R
    rng = np.random.default_rng(5)
    cases = [(44100, 20001), (48000, 30011), (32000, 7777), (96000, 50000), (88200, 12345), (11025, 4097), (8000, 3001),
             (16000, 6000), (24000, 5000), (22051, 3000), (44100, 0), (44100, 1), (48000, 7), (8000, 2), (22050, 1234),
             (44100, 2048), (44100, 2049), (44100, 4 * 1024 * 2 + 3), (88200, 4096 * 3 + 5), (88200, 9), (44100, 45),
             (48000, 3528 * 320 // 147 * 2 + 11), (12000, 9000), (64000, 70001), (192000, 40000), (37800, 20000), (48000, 1), (29400, 30000), (58800, 40000),
             (176400, 50000), (1000, 2000), (768000, 100000)]
    if EMULATED:
        cases = [(44100, 6001), (44100, 2049), (88200, 4101), (48000, 9003), (8000, 1501), (44100, 0), (48000, 7), (22050, 100),
                 (11025, 1000), (96000, 6000), (32000, 3000), (29400, 3000), (58800, 4000)]
    for rate, n in cases:
        x = rng.standard_normal(n).astype(np.float32)
        got, want = B.native.resample(x, rate), R.resample(x, rate)
        assert got.shape == want.shape == (R.resampled_len(n, rate),), (rate, n, got.shape, want.shape)
        assert B.native.resampled_len(n, rate) == want.size
        if n:
            err = np.abs(got.astype(np.float64) - want).max()
            assert err <= 2e-6 * max(1.0, float(np.abs(x).max())), (rate, n, err)
        # the first cut of the kernel (one output per thread, filter rows from global memory) against the kernel the
        # ratio picks: the shared-table kernel (11 025 Hz here) sums in the same order (same bits); the decimation
        # kernels (44.1 / 88.2 kHz) sum phase by phase and the register-resident kernel (4 | down) in four strands
        prev = B.native.set_variant(RESAMPLE_V1)
        try:
            first_cut = B.native.resample(x, rate)
        finally:
            B.native.set_variant(prev)
        if rate in (11025, 22051, 22050):
            assert np.array_equal(first_cut.view(np.uint32), got.view(np.uint32)), (rate, n)
        else:
            assert first_cut.shape == got.shape
            assert n == 0 or np.abs(first_cut - got).max() <= 2e-6 * max(1.0, float(np.abs(x).max())), (rate, n)
    # what a resampler is for: a 1 kHz tone at 48 kHz comes out as the 1 kHz tone at 22 050 Hz, a 15 kHz tone
    # (above the new Nyquist frequency) does not come out
    t48 = np.arange(48000 if not EMULATED else 9600, dtype=np.float64) / 48000.0
    low = B.native.resample(np.sin(2 * np.pi * 1000.0 * t48).astype(np.float32), 48000)
    t22 = np.arange(low.size, dtype=np.float64) / 22050.0
    assert np.abs(low[500:-500] - np.sin(2 * np.pi * 1000.0 * t22)[500:-500]).max() < 2e-3
    high = B.native.resample(np.sin(2 * np.pi * 15000.0 * t48).astype(np.float32), 48000)
    assert np.abs(high[500:-500]).max() < 1e-2
    with pytest.raises(B.native.NativeError, match="sample rate"):
        B.native.resample(np.zeros(10, np.float32), 1 << 20)


def test_resample_feed_is_the_two_steps_fused(golden):
    """bliss_b200_analyze_batch_pcm at a rate other than 22 050 Hz: down-mix, resampler and analysis behind one copy.
    Bit-identical to analysing bliss_b200_resample(bliss_b200_pcm_to_mono(frames)); the features are those the
    oracle computes from the oracle's resampled signal (1e-4); a song that is too short AFTER the conversion is
    status 1; several songs of ragged lengths share a chunk."""
    from oracle import resample as R
    st = golden["pcm_s16_stereo"]                                     # taken as if it ran at 44 100 Hz
    piano = np.ascontiguousarray(np.repeat(golden["pcm_piano"][:, None], 2, axis=1))
    songs = [st, piano, st[:16001], st[:20000], st[:0]] if not EMULATED else [st[:60000], st[:16001], st[:0]]
    for rate in ((44100, 48000) if not EMULATED else (44100,)):
        st_p, f_p = B.native.analyze_batch_pcm(songs, rate, 2)
        monos = [B.native.resample(B.native.pcm_to_mono(x), rate) for x in songs]
        st_f, f_f = B.native.analyze_batch(monos, 2)
        assert np.array_equal(st_p, st_f) and list(st_p) == [0 if m.size >= 8192 else 1 for m in monos]
        assert np.array_equal(f_p[st_p == 0].view(np.uint32), f_f[st_f == 0].view(np.uint32))
        rc, want = O.analyze(R.resample(O.pcm_to_mono(songs[0]), rate), 2)
        _report("resampled feed %d Hz" % rate, f_p[0], want)
        assert rc == 0 and _close(f_p[0], want).all()
    if EMULATED:
        # the chunked host path with the staging ring re-used (1 MB chunks: four songs each): offsets of the input-rate and
        # the 22 050 Hz buffers, job tables per ring slot.  (On a GPU the bench's e2e_cd leg is the multi-chunk case.)
        import os
        rng = np.random.default_rng(3)
        many = [rng.integers(-20000, 20000, (n, 2), dtype=np.int16) for n in (60000, 45001, 90003, 17000, 0, 70000, 52000, 61001, 40000)]
        os.environ["BLISS_B200_CHUNK_MB"] = "1"
        try:
            st_c, f_c = B.native.analyze_batch_pcm(many, 48000, 2)
        finally:
            del os.environ["BLISS_B200_CHUNK_MB"]
        st_1, f_1 = B.native.analyze_batch([B.native.resample(B.native.pcm_to_mono(x), 48000) for x in many], 2)
        assert np.array_equal(st_c, st_1) and np.array_equal(f_c[st_c == 0].view(np.uint32), f_1[st_1 == 0].view(np.uint32))
    # mono f32 at another rate takes the direct copy into the resampler's input
    x = synth.gen_track(3, 0, 22050 * (8 if not EMULATED else 2), device=DEV).cpu().numpy()
    up = R.resample(x, 11025)                                          # taken as 11 025 Hz material: twice the samples
    st_m, f_m = B.native.analyze_batch_pcm([x[:, None]], 11025, 2)
    st_d, f_d = B.native.analyze_batch([B.native.resample(x, 11025)], 2)
    assert st_m[0] == 0 and np.array_equal(f_m, f_d) and up.size == 2 * x.size


def test_library_playlists_run_on_the_distance_kernels(tmp_path):
    """Library::playlist_from / playlist_from_custom / album_playlist_from (src/library.rs:762-876): the stored songs
    come back from SQLite, the orderings come from the device (closest_to_songs, song_to_song, closest_album_to_group,
    dedup) and must be the orderings the oracle's f32 distances give."""
    L = B.library
    rng = np.random.default_rng(9)
    lib = L.Library(str(tmp_path / "songs.db"))
    rows = (rng.random((40, 23), dtype=np.float32) * 2 - 1).astype(np.float32)
    rows[7] = rows[3] + np.float32(1e-3)      # a near-duplicate of song 3: closer than the 0.05 de-duplication threshold
    for i, r in enumerate(rows):
        song = B.Song(path="/music/%02d" % i, title="t%d" % i, artist="a%d" % (i % 5), album="album%d" % (i // 4),
                      track_number=i % 4 + 1, disc_number=1, analysis=B.Analysis(r), duration=1.0)
        lib.store_song(L.LibrarySong(song, {"n": i}))
    seeds = ["/music/03", "/music/11"]
    others = [i for i in range(40) if i not in (3, 11)]
    order, _ = O.closest_to_songs(rows[[3, 11]], rows[others], np.eye(23, dtype=np.float32))
    want = [3, 11] + [others[i] for i in order]
    got = lib.playlist_from_custom(seeds, B.playlist.euclidean_distance, B.playlist.closest_to_songs, False)
    assert [s.bliss_song.path for s in got] == ["/music/%02d" % i for i in want] and got[5].extra_info == {"n": want[5]}
    dedup = [s.bliss_song.path for s in lib.playlist_from(["/music/03"])]   # one seed: its near-duplicate comes right behind it
    rest = [i for i in range(40) if i != 3]
    order1, _ = O.closest_to_songs(rows[[3]], rows[rest], np.eye(23, dtype=np.float32))
    kept, last = [], None
    for i in [3] + [rest[k] for k in order1]:   # dedup_playlist_custom_distance, src/playlist.rs:367-402, on the oracle's distances
        if last is not None and np.float32(O.euclidean_distance(rows[last], rows[i])) < np.float32(0.05):
            continue
        kept.append(i)
        last = i
    assert dedup == ["/music/%02d" % i for i in kept] and len(kept) == 39 and "/music/07" not in dedup
    chain = lib.playlist_from_custom(seeds, B.playlist.euclidean_distance, B.playlist.song_to_song, False)
    want_chain = [3, 11] + [others[i] for i in O.song_to_song(rows[[3, 11]], rows[others], np.eye(23, dtype=np.float32))]
    assert [s.bliss_song.path for s in chain] == ["/music/%02d" % i for i in want_chain]
    with pytest.raises(B.ProviderError, match="has not been analyzed"):
        lib.playlist_from(["/music/nope"])
    # album playlist: the album itself by (disc, track), then the closest albums by the distance of the mean analyses
    albums = lib.album_playlist_from("album2", 2)
    means = np.stack([rows[4 * a:4 * a + 4].mean(axis=0, dtype=np.float32) for a in range(10)])
    d = [np.float32(O.euclidean_distance(means[2], means[a])) for a in range(10)]
    nearest = sorted((a for a in range(10) if a != 2), key=lambda a: d[a])[:2]
    assert [s.bliss_song.album for s in albums] == ["album2"] * 4 + ["album%d" % nearest[0]] * 4 + ["album%d" % nearest[1]] * 4
    assert [s.bliss_song.track_number for s in albums] == [1, 2, 3, 4] * 3
    with pytest.raises(B.ProviderError, match="target album was not found"):
        lib.album_playlist_from("nope", 1)
    assert lib.delete_paths(["/music/00", "/music/01", "/music/zz"]) == 2 and lib.delete_paths([]) == 0
    lib.close()


def test_wav_files_through_the_decoder_pipeline(tmp_path, golden, pcm_piano):
    """File -> features: WAV files through WavDecoder.analyze_paths (decoding threads, one batcher, packed frames
    converted, down-mixed and -- the 44.1 kHz one -- resampled on the device).  The 16-bit mono file is
    data/piano.wav's content, so its row must be Song::analyze of the decoder's f32 samples (ffmpeg.rs:523-527 pins
    those) bit for bit; the stereo and 24-bit files must equal the analysis of the oracle's conversion of the same
    frames; the 44.1 kHz file that of bliss_b200_resample's output; an unreadable file and a short one are error
    items of the same call."""
    import wave

    def write(name, frames, width, rate=22050):
        frames = np.asarray(frames)
        with wave.open(str(tmp_path / name), "wb") as w:
            w.setnchannels(1 if frames.ndim == 1 else frames.shape[1])
            w.setsampwidth(width)
            w.setframerate(rate)
            if width == 3:
                w.writeframes(np.ascontiguousarray(frames.astype("<i4").reshape(-1).view(np.uint8).reshape(-1, 4)[:, :3]).tobytes())
            else:
                w.writeframes(frames.astype("<i2").tobytes())
        return str(tmp_path / name)

    s16 = golden["pcm_piano"]
    st = np.stack([s16, np.roll(s16, 7) // 2], 1).astype(np.int16)
    s24 = s16.astype(np.int32) * 256 + 37
    paths = [write("piano.wav", s16, 2), write("stereo.wav", st, 2), write("piano24.wav", s24, 3),
             write("cd.wav", s16, 2, rate=44100), str(tmp_path / "junk.wav"), write("short.wav", s16[:4000], 2)]
    (tmp_path / "junk.wav").write_bytes(b"RIFF....WAVEnothing")
    got = dict(B.WavDecoder.analyze_paths_with_options(paths, B.AnalysisOptions(number_cores=2)))
    assert len(got) == 6
    assert np.array_equal(got[paths[0]].analysis.as_arr1(), B.Song.analyze(pcm_piano).as_arr1())
    assert np.array_equal(got[paths[1]].analysis.as_arr1(), B.Song.analyze(O.pcm_to_mono(st)).as_arr1())
    assert np.array_equal(got[paths[2]].analysis.as_arr1(), B.Song.analyze(O.pcm_to_mono(s24 * 256)).as_arr1())
    cd = B.native.resample(O.pcm_to_mono(s16), 44100)
    assert cd.size == (s16.size + 1) // 2 and abs(got[paths[3]].duration - s16.size / 44100.0) < 1e-9
    assert np.array_equal(got[paths[3]].analysis.as_arr1(), B.Song.analyze(cd).as_arr1())
    assert isinstance(got[paths[4]], B.DecodingError) and isinstance(got[paths[5]], B.AnalysisError)
    assert abs(got[paths[0]].duration - s16.size / 22050.0) < 1e-9
    # a CUE sheet over one of the files (src/cue.rs:208-243): both tracks are slices of ONE decoded buffer analysed in one
    # call; each must be Song::analyze of that slice of the decoder's samples, bit for bit
    (tmp_path / "stereo.cue").write_text('PERFORMER "P"\nTITLE "T"\nFILE "stereo.wav" WAVE\n  TRACK 01 AUDIO\n    TITLE "one"\n'
                                         '    INDEX 01 0:00:00\n  TRACK 02 AUDIO\n    TITLE "two"\n    INDEX 01 0:02:37\n')
    cut = int(np.float32(np.float32(2) + np.float32(37 * 1_000_000_000 // 75) / np.float32(1e9)) * np.float32(22050))
    tracks = B.BlissCue(B.WavDecoder).songs_from_path(str(tmp_path / "stereo.cue"))
    mono = O.pcm_to_mono(st)
    assert [t.title for t in tracks] == ["one", "two"] and 54000 < cut < 55000
    assert np.array_equal(tracks[0].analysis.as_arr1(), B.Song.analyze(mono[:cut]).as_arr1())
    assert np.array_equal(tracks[1].analysis.as_arr1(), B.Song.analyze(mono[cut:]).as_arr1())
    # ... and on into the reference's on-disk format (src/library.rs:500-529, 1544-1670): files -> decoder threads ->
    # GPU batches -> SQLite rows; the two refused files are rows of the failed-song kind, a second run is a no-op
    lib = B.library.Library(str(tmp_path / "songs.db"), decoder=B.WavDecoder)
    assert lib.update_library(paths) == (4, 2)
    stored = {s.bliss_song.path: s.bliss_song.analysis.as_arr1() for s in lib.songs_from_library()}
    assert sorted(stored) == sorted(paths[:4])
    for p in paths[:4]:
        assert np.array_equal(stored[p], got[p].analysis.as_arr1())
    assert sorted(f.song_path for f in lib.get_failed_songs()) == sorted(paths[4:])
    lib.close()


# The kernel cuts behind BLISS_B200_VARIANT bits 64 ... 8192 were written after round 1's GPU budget was spent, passed
# on the driver's B200 at the end of round 1 and won the A/B of round 2 (profiles/ab_r02.md): they are the default
# now (mask 0) and a set bit switches BACK to the kernel measured in round 1.  The tests below compare the two
# implementations of each cut whichever way round the bit reads; they are plain tests (no xfail) run in a child
# process with a time limit.


STFT_V3 = 131072  # BLISS_B200_VARIANT bit: stft8192v3_kernel (256 threads, one column per thread) instead of stft8192v2_kernel
STFT_V1 = 32768  # BLISS_B200_VARIANT bit: the round-1 stft8192_kernel (its cuts are bits 64 / 128 / 4096 / 8192) instead of stftv2
PVOC_V1 = 16384  # BLISS_B200_VARIANT bit: the round-1 pvoc512_kernel (its cuts are bits 512 / 1024 / 2048) instead of pvoc512v2_kernel


def experimental(fn):
    """... and each of them runs in a CHILD pytest process with a time limit: a kernel that has never met hardware may
    fault (a sticky CUDA error would fail every later test of this process) or hang (and take the whole GPU run
    with it); the child is killed at the limit and the test reports xfail.  The child runs the body for real
    (--runxfail), its output is passed on."""
    import functools
    import os
    import subprocess
    import sys

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if os.environ.get("BLISS_B200_EXPERIMENTAL_CHILD") == "1":
            return fn(*args, **kwargs)
        out = subprocess.run([sys.executable, "-m", "pytest", "%s::%s" % (os.path.abspath(__file__), fn.__name__), "-q", "-s", "-m", "gpu",
                              "-p", "no:cacheprovider", "--runxfail", "--tb=short"],
                             env=dict(os.environ, BLISS_B200_EXPERIMENTAL_CHILD="1"), capture_output=True, text=True, timeout=900,
                             cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        print(out.stdout[-3000:])
        assert out.returncode == 0 and " passed" in out.stdout, out.stdout[-3000:] + out.stderr[-1000:]

    return wrapper


@experimental
def test_experimental_pass1_variants_agree(pcm_song, pcm_piano):
    """BLISS_B200_VARIANT bits 64 / 128: pass 1 of stft8192_kernel with product twiddles (4 table loads instead
    of 15) and with the Hann window synthesised from the thread's phase (no window loads).  Candidates for the
    data-pipe-bound kernel, not the default: they round differently in the last place (window coefficients differ
    from the reference's f32 table by <= 2.4e-7) and must agree with the measured kernel to 1e-5."""
    songs = [pcm_song, pcm_piano] + _extra_tracks(78, 6, 35, 211)
    try:
        B.native.set_variant(STFT_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        assert (st0 == 0).all()
        for mask in (64, 128, 64 | 128):
            B.native.set_variant(STFT_V1 | mask)
            st, f = B.native.analyze_batch(songs, 2)
            assert (st == 0).all()
            assert np.abs(f - f0).max() < 1e-5, (mask, np.abs(f - f0).max(0))
            assert np.array_equal(f[:, :10], f0[:, :10])  # only the chroma features depend on the chroma STFT
    finally:
        B.native.set_variant(0)



@experimental
def test_experimental_stft_pair_kernel(pcm_song, pcm_piano):
    """BLISS_B200_VARIANT bit 256: the STFT micro-benchmark with hop-256 frames j, j+1 sharing one FFT
    (stft512_pairs_kernel) against the oracle's PVocTempo norms and against the default kernel, on several
    songs of ragged lengths in one call (odd frame counts, item boundaries)."""
    songs = [pcm_song[:60000], pcm_piano[:8192 + 256 * 7 + 13], pcm_song[1000:1000 + 256 * 300 + 511], pcm_piano]
    flat = np.concatenate([np.pad(x, (0, (-len(x)) % 4)) for x in songs]).astype(np.float32)
    offs = np.cumsum([0] + [len(x) + (-len(x)) % 4 for x in songs[:-1]]).tolist()
    lens = [len(x) for x in songs]
    n_t = [(n - 512) // 256 + 1 for n in lens]
    d = torch.from_numpy(flat).to(DEV)
    try:
        outs = {}
        for mask in (0, 256):
            B.native.set_variant(mask)
            mags = torch.full((sum(n_t), 257), -1.0, dtype=torch.float32, device=DEV)
            fo = B.native.stft512_mag_device(d.data_ptr(), offs, lens, mags.data_ptr(), None)
            _sync()
            assert list(fo) == np.cumsum([0] + n_t).tolist()
            outs[mask] = mags.cpu().numpy()
        for i, x in enumerate(songs):
            want = O.tempo_norms(x)
            lo = int(np.sum(n_t[:i]))
            for mask in (0, 256):
                got = outs[mask][lo:lo + n_t[i]]
                assert (got >= 0).all(), "variant %d left magnitudes of song %d unwritten" % (mask, i)
                err = np.abs(got - want).max() / want.max()
                assert err < 1e-6, (mask, i, err)
    finally:
        B.native.set_variant(0)


@experimental
def test_experimental_pvoc_twiddle_variant(pcm_song, pcm_piano):
    """BLISS_B200_VARIANT bit 512: pvoc512_kernel with the phase-A twiddles formed from four per-lane registers
    (no twiddle loads in the frame loop).  Rounds differently in the last place: timbral descriptors within 1e-5
    of the measured kernel, parity with the oracle within the usual bar, everything behind the chroma STFT
    untouched."""
    songs = [pcm_song, pcm_piano] + _extra_tracks(79, 6, 30, 173)
    try:
        B.native.set_variant(PVOC_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        B.native.set_variant(PVOC_V1 | 512)
        st, f = B.native.analyze_batch(songs, 2)
        assert (st0 == 0).all() and (st == 0).all()
        assert np.abs(f[:, 1:10] - f0[:, 1:10]).max() < 1e-5, np.abs(f - f0).max(0)
        assert np.array_equal(f[:, 10:], f0[:, 10:])
        assert (np.abs(f[:, 0] - f0[:, 0]) < 1e-5).sum() >= len(songs) - 1  # a tempo decision may sit on an edge
        for i in (0, 1):
            rc, want = O.analyze(songs[i], 2)
            assert rc == 0 and _close(f[i], want).all(), (i, np.abs(f[i] - want).max())
    finally:
        B.native.set_variant(0)


@experimental
def test_experimental_pvoc_pair_descriptors_are_bit_identical(pcm_song, pcm_piano):
    """BLISS_B200_VARIANT bit 1024: both frames' descriptors reduced by transposed butterflies and finished once
    per pair, magnitudes by MUFU.SQRT alone on data kept 2^30 above its level.  Same arithmetic, same trees, exact
    power-of-two scalings: every feature is expected to equal the measured kernel's bit for bit (reported; held
    to 1e-6); together with bit 512 (product twiddles) within 1e-5."""
    songs = [pcm_song, pcm_piano, np.concatenate([np.zeros(30000, np.float32), pcm_piano[:40000], np.zeros(5000, np.float32)])]
    songs += _extra_tracks(80, 5, 25, 97)
    try:
        B.native.set_variant(PVOC_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        _, _, taps0 = B.native.analyze_taps(pcm_song, 2)
        B.native.set_variant(PVOC_V1 | 1024)
        st, f = B.native.analyze_batch(songs, 2)
        _, _, taps = B.native.analyze_taps(pcm_song, 2)
        assert (st0 == 0).all() and (st == 0).all()
        # Expected: identical bits.  That rests on MUFU.SQRT being invariant under the exact 2^60 scaling of its
        # argument (same mantissa path), which this first run on hardware establishes: report it, and hold the
        # features to 1e-6 either way.
        same = {k: bool(np.array_equal(taps[k], taps0[k])) for k in ("centroid", "rolloff", "flatness", "flux")}
        print("pair-descriptor variant: features bit-identical = %s, per-frame taps bit-identical = %s"
              % (np.array_equal(f, f0), same))
        assert np.abs(f[:, 1:] - f0[:, 1:]).max() < 1e-6, np.abs(f - f0).max(0)
        assert (np.abs(f[:, 0] - f0[:, 0]) < 1e-6).sum() >= len(songs) - 1
        assert np.array_equal(f[:, 10:], f0[:, 10:])
        for k in ("centroid", "flatness", "flux"):
            assert np.allclose(taps[k], taps0[k], rtol=2e-6, atol=1e-9), k
        assert (taps["rolloff"] != taps0["rolloff"]).sum() <= 2  # a discrete bin: an ulp may move a frame by one bin
        B.native.set_variant(PVOC_V1 | 1024 | 512)
        st, f = B.native.analyze_batch(songs, 2)
        assert np.abs(f[:, 1:10] - f0[:, 1:10]).max() < 1e-5 and np.array_equal(f[:, 10:], f0[:, 10:])
    finally:
        B.native.set_variant(0)


@experimental
def test_experimental_pvoc_tile_padding_is_bit_identical(pcm_song, pcm_piano):
    """BLISS_B200_VARIANT bit 2048: the natural-order tile of pvoc512_kernel padded k + (k >> 4) instead of
    k + (k >> 3) (conflict-free stores).  Only shared-memory addresses change: every feature bit for bit."""
    songs = [pcm_song, pcm_piano] + _extra_tracks(81, 4, 20, 59)
    try:
        B.native.set_variant(PVOC_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        B.native.set_variant(PVOC_V1 | 2048)
        st, f = B.native.analyze_batch(songs, 2)
        assert (st0 == 0).all() and (st == 0).all()
        assert np.array_equal(f, f0), np.abs(f - f0).max(0)
    finally:
        B.native.set_variant(0)


@experimental
def test_experimental_fft8192_buffer_layout_is_bit_identical(pcm_song, pcm_piano):
    """BLISS_B200_VARIANT bit 4096: stft8192_kernel's FFT buffer without the per-16 padding (conflict-free mirror
    loads in the pair epilogue).  Only shared-memory addresses change: every feature and every magnitude bit for bit."""
    songs = [pcm_song, pcm_piano] + _extra_tracks(82, 4, 20, 61)
    try:
        B.native.set_variant(STFT_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        _, _, taps0 = B.native.analyze_taps(pcm_piano, 2)
        B.native.set_variant(STFT_V1 | 4096)
        st, f = B.native.analyze_batch(songs, 2)
        _, _, taps = B.native.analyze_taps(pcm_piano, 2)
        assert (st0 == 0).all() and (st == 0).all()
        assert np.array_equal(f, f0), np.abs(f - f0).max(0)
        assert np.array_equal(taps["stft8192"], taps0["stft8192"])
    finally:
        B.native.set_variant(0)


@experimental
def test_experimental_odd_frame_rotation(pcm_song, pcm_piano):
    """BLISS_B200_VARIANT bit 8192: chroma frames that start on an odd sample (every second one: the hop is 2205) are
    transformed rotated by one sample, which makes their sample pairs aligned 64-bit loads; |DFT| is unchanged by a
    circular shift, so only rounding moves: magnitudes within 2e-6 of the oracle, features within 1e-5 of the
    measured kernel, nothing outside the chroma features touched."""
    songs = [pcm_song, pcm_piano] + _extra_tracks(83, 4, 20, 67)
    try:
        B.native.set_variant(STFT_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        for mask in (8192, 8192 | 4096 | 128 | 64):
            B.native.set_variant(STFT_V1 | mask)
            st, f = B.native.analyze_batch(songs, 2)
            assert (st == 0).all() and np.abs(f - f0).max() < 1e-5, (mask, np.abs(f - f0).max(0))
            assert np.array_equal(f[:, :10], f0[:, :10])
        B.native.set_variant(STFT_V1 | 8192)
        _, _, taps = B.native.analyze_taps(pcm_piano, 2)
        S = O.stft(pcm_piano, 8192, 2205)
        assert np.abs(taps["stft8192"].T - S).max() / S.max() < 2e-6
    finally:
        B.native.set_variant(0)


@experimental
def test_pvoc512v2_against_the_round1_kernel(pcm_song, pcm_piano):
    """pvoc512v2_kernel (mask 0) against pvoc512_kernel (bit 16384): the same FFT pair packing, another split of the
    32-point stage, register mirrors instead of the natural-order tile, descriptors reduced four lanes per frame.
    Per-frame taps within the bars of the stage tests, features within 1e-5, everything behind the chroma STFT
    untouched; parity with the oracle through the usual bar."""
    silence = np.concatenate([np.zeros(30000, np.float32), pcm_piano[:40000], np.zeros(5000, np.float32)])
    songs = [pcm_song, pcm_piano, silence] + _extra_tracks(83, 6, 30, 173)
    try:
        B.native.set_variant(PVOC_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        _, _, taps0 = B.native.analyze_taps(pcm_song, 2)
        B.native.set_variant(0)
        st, f = B.native.analyze_batch(songs, 2)
        _, _, taps = B.native.analyze_taps(pcm_song, 2)
        assert (st0 == 0).all() and (st == 0).all()
        assert np.abs(f[:, 1:10] - f0[:, 1:10]).max() < 1e-5, np.abs(f - f0).max(0)
        assert np.array_equal(f[:, 10:], f0[:, 10:])
        assert (np.abs(f[:, 0] - f0[:, 0]) < 1e-5).sum() >= len(songs) - 1  # a tempo decision may sit on an edge
        assert np.allclose(taps["centroid"], taps0["centroid"], rtol=1e-5, atol=1e-3)
        assert np.allclose(taps["flatness"], taps0["flatness"], rtol=2e-4, atol=1e-5)
        assert np.abs(taps["flux"] - taps0["flux"]).max() <= 1e-5 * np.abs(taps0["flux"]).max()
        assert np.mean(taps["rolloff"] != taps0["rolloff"]) < 2e-3
        for i in (0, 1, 2):
            rc, want = O.analyze(songs[i], 2)
            assert rc == 0 and _close(f[i], want).all(), (i, np.abs(f[i] - want).max())
    finally:
        B.native.set_variant(0)


@experimental
def test_stft8192v2_against_the_round1_kernel(pcm_song, pcm_piano):
    """stft8192v2_kernel (mask 0: bulk-copy staging, frames rotated to a 16-byte boundary, mirror pairs in registers,
    bulk-stored magnitude rows) against stft8192_kernel (bit 32768).  Magnitudes of both within 2e-6 of the
    oracle's STFT, the same pip_track candidates (count, and sorted pitches / magnitudes against the oracle's f64
    interpolation), identical tuning, features within 1e-5, nothing outside the chroma features touched.  Songs sit
    at every offset mod 4 of one device buffer (CUE-style slices), so all four rotations and both edge paths run."""
    songs = [pcm_song, pcm_piano, pcm_song[1:], pcm_song[2:90001], pcm_piano[3:]] + _extra_tracks(84, 4, 25, 61)
    try:
        B.native.set_variant(STFT_V1)
        st0, f0 = B.native.analyze_batch(songs, 2)
        _, _, taps0 = B.native.analyze_taps(pcm_piano, 2)
        S = O.stft(pcm_piano, 8192, 2205)
        assert np.abs(taps0["stft8192"].T - S).max() / S.max() < 2e-6
        for mask in (STFT_V3, 0):  # the 256-thread / one-column cut of the same design, then the default (128 threads)
            B.native.set_variant(mask)
            st, f = B.native.analyze_batch(songs, 2)
            _, _, taps = B.native.analyze_taps(pcm_piano, 2)
            assert (st0 == 0).all() and (st == 0).all()
            assert np.abs(f - f0).max() < 1e-5, (mask, np.abs(f - f0).max(0))
            assert np.array_equal(f[:, :10], f0[:, :10])
            assert np.abs(taps["stft8192"].T - S).max() / S.max() < 2e-6
            assert taps["tuning"] == taps0["tuning"]
            assert taps["n_peaks"] == taps0["n_peaks"]
        # unaligned slices of ONE device buffer through the device API: offsets 0, 1, 2, 3 mod 4
        import torch
        flat = torch.from_numpy(np.concatenate([pcm_song, pcm_piano])).to(DEV)
        offs = [0, 1, 2, 3, len(pcm_song) + 1, len(pcm_song) + 6]
        lens = [50001, 50002, 60003, 70000, 40001, 44444]
        out = torch.zeros((len(offs), 23), dtype=torch.float32, device=DEV)
        B.native.analyze_batch_device(flat.data_ptr(), offs, lens, 2, out.data_ptr(), None)
        _sync()
        got = out.cpu().numpy()
        host = flat.cpu().numpy()
        for i, (o, l) in enumerate(zip(offs, lens)):
            rc, want = O.analyze(host[o:o + l], 2)
            assert rc == 0 and _close(got[i], want).all(), (i, np.abs(got[i] - want).max())
    finally:
        B.native.set_variant(0)


def test_device_chroma_filter_tables():
    """chroma_filter_table_kernel: the device's filterbank (the f64 table the f32 contraction weights are rounded
    from) against the oracle's chroma_filter for tunings {-0.5, -0.05, 0, 0.49} at the reference's own 1e-9
    (src/chroma.rs:705-714, data/chroma-filter.npy pins the oracle in tests/test_oracle_golden.py)."""
    for idx in (0, 45, 50, 99):
        tuning = (-50.0 + idx) / 100.0
        got = B.native.chroma_filter(idx)
        want = O.chroma_filter(8192, tuning)
        assert got.shape == want.shape == (12, 4097)
        assert np.abs(got - want).max() < 1e-9, (idx, np.abs(got - want).max())


def test_packed_distance_kernel_is_bit_identical_to_the_scalar_one():
    """distance_matrix_diag_kernel (f32x2: two columns per thread, products and sums rounded separately -- the add is
    an FMA by a run-time 1.0 so that ptxas cannot contract it) against the round-1 scalar kernel (bit 65536) and the
    oracle, bit for bit, for the v2 weights, the identity (v1) and an arbitrary diagonal metric, on shapes that leave
    partial tiles."""
    rng = np.random.default_rng(11)
    for dim, m in ((23, B.native.feature_weights(2)), (20, None), (23, np.diag(rng.uniform(0.1, 2.0, 23)).astype(np.float32)),
                   (20, B.native.feature_weights(1))):
        rows = (rng.uniform(-1, 1, (131, dim))).astype(np.float32)
        cols = (rng.uniform(-1, 1, (517, dim))).astype(np.float32)
        rows[3] = cols[5]          # an exact zero distance
        cols[7] = cols[5] + np.float32(1e-6)
        try:
            B.native.set_variant(65536)
            want = B.native.distance_matrix(rows, cols, m=m)
        finally:
            B.native.set_variant(0)
        got = B.native.distance_matrix(rows, cols, m=m)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dim, np.abs(got - want).max())
        ref = np.array([[O.mahalanobis_distance(r, c, m if m is not None else np.eye(dim, dtype=np.float32)) for c in cols[:40]]
                        for r in rows[:9]], np.float32)
        assert np.array_equal(got[:9, :40].view(np.uint32), ref.view(np.uint32))


def test_bench_corpus_256_three_minute_tracks_against_the_oracle():
    """SURVEY section 8(d) config 2 at full size: 256 of the 3-minute bench tracks (the generator and seed of bench.py)
    through the device path against the oracle on all host threads: max / median error per feature, tempo and tuning
    flips.  Bar: every feature within 1e-4 (relative to max(1, |oracle|)); at most one tempo decision on an edge."""
    import os
    n_tracks, n = 256, 3 * 60 * 22050
    pcm, offs, lens = synth.gen_corpus_flat(20260925, list(range(n_tracks)), [n] * n_tracks, device=DEV)
    out = torch.zeros((n_tracks, 23), dtype=torch.float32, device=DEV)
    st = B.native.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    feats = out.cpu().numpy()
    songs = [pcm[o:o + n].cpu().numpy() for o in offs]
    del pcm
    ost, ofe = O.analyze_batch(songs, 2, n_threads=os.cpu_count() or 8)
    assert (st == 0).all() and (ost == 0).all()
    err = np.abs(feats - ofe)
    rel = err / np.maximum(1.0, np.abs(ofe))
    print("256 x 3-min tracks: max abs err per feature   ", np.array2string(err.max(axis=0), precision=1))
    print("256 x 3-min tracks: median abs err per feature", np.array2string(np.median(err, axis=0), precision=1))
    tempo_flips = int((err[:, 0] > 1e-3).sum())
    chroma_off = int((err[:, 10:].max(axis=1) > 1e-3).sum())  # a flipped tuning bin moves every chroma feature
    print("tempo decisions differing: %d / %d, tuning decisions differing: %d" % (tempo_flips, n_tracks, chroma_off))
    assert tempo_flips <= 1 and chroma_off == 0
    ok = rel <= TOL
    ok[err[:, 0] > 1e-3, 0] = True
    assert ok.all(), (np.argwhere(~ok)[:5], rel.max())


def test_chroma_stft_frame_edges_and_alignments():
    """stft8192v2_kernel's staging decisions, swept: an interior frame is bulk-copied from the 16-byte boundary below
    its first sample (rotation r = start & 3) unless the copy would run past the song's end, in which case -- like the
    reflect-padded frames at both ends -- it is filled by hand.  Songs of lengths around every boundary of that logic
    (8192 + k, and lengths that put the last interior frame 0..4 samples short of the end), at every offset mod 4 of
    one device buffer, against the oracle; and the same through the one-column kernel."""
    rng = np.random.default_rng(5)
    lens = [8192 + k for k in range(0, 9)] + [4096 + 2205 * 3 + 4096 + k for k in range(-5, 6)] + \
           [4096 + 2205 * 7 + 4096 + k for k in (-3, -2, -1, 0, 1, 2, 3)] + [30011, 44100 + 1, 65537, 2205 * 20, 2205 * 20 + 1]
    songs = [(0.3 * rng.standard_normal(n) + 0.2 * np.sin(2 * np.pi * 440.0 * np.arange(n) / 22050.0)).astype(np.float32) for n in lens]
    flat, offs, pos = [], [], 0
    for i, x in enumerate(songs):
        pad = (i % 4)  # arbitrary alignment: consecutive songs start at every offset mod 4
        flat.append(np.zeros(pad, np.float32))
        pos += pad
        offs.append(pos)
        flat.append(x)
        pos += len(x)
    buf = torch.from_numpy(np.concatenate(flat + [np.zeros(8, np.float32)])).to(DEV)
    want = [O.analyze(x, 2) for x in songs]
    try:
        for mask in (0, STFT_V3):
            B.native.set_variant(mask)
            out = torch.zeros((len(songs), 23), dtype=torch.float32, device=DEV)
            st = B.native.analyze_batch_device(buf.data_ptr(), offs, lens, 2, out.data_ptr(), None)
            _sync()
            got = out.cpu().numpy()
            assert (st == 0).all()
            for i, (rc, w) in enumerate(want):
                assert rc == 0
                ok = _close(got[i], w)
                assert ok[10:].all(), (mask, lens[i], offs[i] % 4, np.abs(got[i] - w)[10:].max())   # the chroma features
                assert ok[:10].all(), (mask, lens[i], np.abs(got[i] - w)[:10].max())
    finally:
        B.native.set_variant(0)
