"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/bliss_b200.h declares, the host mirror keeps the reference's API
behaviour, and nothing silently falls back to the CPU when no GPU is present."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import bliss_rs_b200 as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as G
    G.build()
    hdr = open(os.path.join(ROOT, "include", "bliss_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(bliss_b200_[a-z0-9_]+)\s*\(", hdr)))
    out = subprocess.check_output(["nm", "-D", "--defined-only", B.native.SO_PATH]).decode()
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert declared and set(declared) <= exported
    assert sorted(B.native.SYMBOLS) == declared


def test_sm100a_code_in_library():
    out = subprocess.check_output(["cuobjdump", "--list-elf", B.native.SO_PATH]).decode()
    assert "sm_100a" in out


def test_feature_count_and_weights_need_no_device():
    assert B.native.feature_count(2) == 23 and B.native.feature_count(1) == 20
    assert B.native.feature_count(7) == 0
    w = B.FeaturesVersion.Version2.feature_weights()  # src/lib.rs:262-271 test_dimensions_weights
    assert w.shape == (23, 23) and B.FeaturesVersion.Version1.feature_weights().shape == (20, 20)
    assert w[0, 0] == 0.25 and w[1, 1] == 1.0 and w[10, 10] == np.float32(3.0 / 13.0)
    assert np.count_nonzero(w - np.diag(np.diag(w))) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(B.native.NativeError):
        B.Song.analyze(np.zeros(10000, np.float32))
    with pytest.raises(B.native.NativeError):
        B.playlist.euclidean_distance(np.zeros(20), np.ones(20))


def test_analysis_type_behaviour():
    # Analysis::new length check, src/song/mod.rs:326-339
    with pytest.raises(B.ProviderError):
        B.Analysis([0.0] * 5, B.FeaturesVersion.Version2)
    a = B.Analysis(np.arange(23) / 10.0, B.FeaturesVersion.LATEST)
    # index access, src/song/mod.rs:693-697
    assert a[B.AnalysisIndex.Tempo] == 0.0 and abs(a[B.AnalysisIndex.Chroma13] - 2.2) < 1e-6
    with pytest.raises(RuntimeError):  # panics in the reference, src/song/mod.rs:276-278
        a[B.AnalysisIndexv1.Tempo]
    v1 = B.Analysis(np.zeros(20), B.FeaturesVersion.Version1)
    with pytest.raises(RuntimeError):  # src/song/mod.rs:365-367
        a.distance(v1)
    assert a.as_vec()[1] == pytest.approx(0.1) and a.as_arr1().dtype == np.float32
    assert "Analysis (Version 2)" in repr(a) and "Tempo" in repr(a)
    # impl Debug for Analysis: the exact strings of the reference's tests (src/song/mod.rs:712-735), values taken from them
    v2 = [0.3846389, -0.849141, -0.7548105, -0.8790748, -0.63258266, -0.7258959, -0.775738, -0.8146726, 0.2716726, 0.25779057,
          -0.34292513, -0.62803423, -0.28095096, 0.08686459, 0.24446082, -0.5723257, 0.23292065, 0.19981146, -0.58594406,
          -0.06784296, -0.06000763, -0.58485717, -0.07880378]
    assert repr(B.Analysis(v2)) == (
        "Analysis (Version 2) { Tempo: 0.3846389, Zcr: -0.849141, MeanSpectralCentroid: -0.7548105, StdDeviationSpectralCentroid: "
        "-0.8790748, MeanSpectralRolloff: -0.63258266, StdDeviationSpectralRolloff: -0.7258959, MeanSpectralFlatness: -0.775738, "
        "StdDeviationSpectralFlatness: -0.8146726, MeanLoudness: 0.2716726, StdDeviationLoudness: 0.25779057, Chroma1: -0.34292513, "
        "Chroma2: -0.62803423, Chroma3: -0.28095096, Chroma4: 0.08686459, Chroma5: 0.24446082, Chroma6: -0.5723257, Chroma7: "
        "0.23292065, Chroma8: 0.19981146, Chroma9: -0.58594406, Chroma10: -0.06784296, Chroma11: -0.06000763, Chroma12: -0.58485717, "
        "Chroma13: -0.07880378 } /* [0.3846389, -0.849141, -0.7548105, -0.8790748, -0.63258266, -0.7258959, -0.775738, -0.8146726, "
        "0.2716726, 0.25779057, -0.34292513, -0.62803423, -0.28095096, 0.08686459, 0.24446082, -0.5723257, 0.23292065, 0.19981146, "
        "-0.58594406, -0.06784296, -0.06000763, -0.58485717, -0.07880378] */")
    v1 = v2[:10] + [-0.35661936, -0.63578653, -0.29593682, 0.06421304, 0.21852458, -0.581239, -0.9466835, -0.9481153, -0.9820945, -0.95968974]
    r1 = repr(B.Analysis(v1, B.FeaturesVersion.Version1))
    assert r1.startswith("Analysis (Version 1) { Tempo: 0.3846389, ") and "Chroma10: -0.95968974 } /* [0.3846389, " in r1
    assert r1.endswith("-0.9820945, -0.95968974] */") and "Chroma11" not in r1
    from bliss_rs_b200.song import _f32_debug
    assert [_f32_debug(x) for x in (1, -0.0, 1.5e-5, 1e-7, 3e20, 0.0001, 123456.7)] == ["1.0", "-0.0", "1.5e-5", "1e-7", "3e20", "0.0001", "123456.7"]
    # FeaturesVersion::try_from, src/lib.rs:195-207
    assert B.FeaturesVersion.try_from(1) == B.FeaturesVersion.Version1
    with pytest.raises(B.ProviderError):
        B.FeaturesVersion.try_from(3)
    assert len(B.AnalysisIndex) == B.NUMBER_FEATURES == 23 and len(B.AnalysisIndexv1) == 20


def test_error_strings():
    assert str(B.AnalysisError("empty or too short song.")) == \
        "error happened while analyzing file - empty or too short song."
    L = B.native.load()
    assert L.bliss_b200_strerror(1) == b"empty or too short song."


def test_decoder_trait_batches_errors_as_items(monkeypatch):
    calls = []

    class Dec(B.Decoder):
        BATCH_SONGS = 2

        @classmethod
        def decode(cls, path):
            if path == "bad":
                raise B.DecodingError("nope")
            return B.PreAnalyzedSong(path=path, title=path, sample_array=np.zeros(9000, np.float32))

    def fake_batch(arrays, opts=None):
        calls.append(len(arrays))
        return [B.Analysis(np.zeros(23)) for _ in arrays]

    monkeypatch.setattr(B.song, "analyze_batch", fake_batch)
    one_core = B.AnalysisOptions(number_cores=1)  # one decoding thread: the order of arrival is determined
    got = list(Dec.analyze_paths_with_options(["a", "bad", "b", "c"], one_core))
    assert [p for p, _ in got] == ["bad", "a", "b", "c"]
    assert isinstance(got[0][1], B.DecodingError) and isinstance(got[1][1], B.Song)
    assert calls == [2, 1]
    assert list(Dec.analyze_paths([])) == []


def test_decoder_trait_decodes_on_number_cores_threads_while_the_batcher_runs(monkeypatch):
    """Decoder::analyze_paths_with_options (src/song/decoder.rs:278-332): min(cores, number_cores) decoding threads
    on contiguous chunks, ONE batcher making the GPU calls while they keep decoding, a bounded hand-over queue,
    every path answered exactly once, failures of decode() that are not BlissErrors re-raised to the consumer."""
    import threading
    import time
    lock, state = threading.Lock(), {"now": 0, "peak": 0, "threads": set(), "batch_threads": set(), "overlap": False}

    class Dec(B.Decoder):
        BATCH_SONGS = 4

        @classmethod
        def decode(cls, path):
            with lock:
                state["now"] += 1
                state["peak"] = max(state["peak"], state["now"])
                state["threads"].add(threading.get_ident())
            time.sleep(0.01)
            with lock:
                state["now"] -= 1
            if path.startswith("bad"):
                raise B.DecodingError(path)
            if path == "bug":
                raise ZeroDivisionError("decode() is broken")
            return B.PreAnalyzedSong(path=path, sample_array=np.full(9000, float(path), np.float32))

    def fake_batch(arrays, opts=None):
        with lock:
            state["batch_threads"].add(threading.get_ident())
            state["overlap"] = state["overlap"] or state["now"] > 0
        assert 1 <= len(arrays) <= Dec.BATCH_SONGS
        time.sleep(0.005)
        return [B.AnalysisError("empty or too short song.") if a[0] == 7 else B.Analysis(np.full(23, a[0])) for a in arrays]

    monkeypatch.setattr(B.song, "analyze_batch", fake_batch)
    monkeypatch.setattr(B.song.os, "cpu_count", lambda: 8)
    paths = [str(i) for i in range(40)] + ["bad1", "bad2"]
    got = list(Dec.analyze_paths_with_options(paths, B.AnalysisOptions(number_cores=4)))
    assert sorted(p for p, _ in got) == sorted(paths)                     # every path exactly once
    for p, r in got:
        if p.startswith("bad"):
            assert isinstance(r, B.DecodingError)
        elif p == "7":
            assert isinstance(r, B.AnalysisError)                        # a rejected song is an item of its batch
        else:
            assert isinstance(r, B.Song) and r.path == p and r.analysis.as_arr1()[0] == float(p)   # rows stay with their songs
    # 42 paths / 4 cores -> chunks of 10 -> 5 chunk threads, like the reference's paths.chunks(len / cores)
    assert len(state["threads"]) == 5 and 2 <= state["peak"] <= 5
    assert len(state["batch_threads"]) == 1 and not state["batch_threads"] & state["threads"]
    assert state["overlap"]                                               # GPU calls ran while decoders were busy
    with pytest.raises(ZeroDivisionError):
        list(Dec.analyze_paths_with_options(["1", "2", "bug", "3"], B.AnalysisOptions(number_cores=2)))
    # a consumer that walks away does not leave the producers blocked on the bounded queue for ever
    before = threading.active_count()
    it = Dec.analyze_paths_with_options([str(i) for i in range(200)], B.AnalysisOptions(number_cores=4))
    next(it)
    it.close()
    deadline = time.time() + 10
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() <= before


def _write_wav(path, frames, width, rate=22050):
    import wave
    frames = np.asarray(frames)
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1 if frames.ndim == 1 else frames.shape[1])
        w.setsampwidth(width)
        w.setframerate(rate)
        if width == 3:
            b = frames.astype("<i4").reshape(-1).view(np.uint8).reshape(-1, 4)[:, :3]
            w.writeframes(np.ascontiguousarray(b).tobytes())
        else:
            w.writeframes(frames.astype({1: "u1", 2: "<i2", 4: "<i4"}[width]).tobytes())


def test_wav_decoder_keeps_the_codecs_frames(tmp_path, golden, monkeypatch):
    """bliss-rs_b200/decoder.py: a RIFF/WAVE PCM file is unpacked on the host and nothing else -- the frames reach
    bliss_b200_analyze_batch_pcm as the file holds them, with the file's sample rate; unreadable files are
    DecodingErrors (items of analyze_paths); a batch that mixes formats is split into one call per (format, channel
    count, sample rate)."""
    s16 = golden["pcm_piano"][:30000]                       # data/piano.wav is such a file (ffmpeg.rs:523-527)
    rng = np.random.default_rng(1)
    st = np.stack([s16, (s16 // 3).astype(np.int16)], 1)
    s24 = rng.integers(-(1 << 23), 1 << 23, 5000)
    u8 = rng.integers(0, 256, 5000)
    _write_wav(tmp_path / "mono16.wav", s16, 2)
    _write_wav(tmp_path / "stereo16.wav", st, 2)
    _write_wav(tmp_path / "mono24.wav", s24, 3)
    _write_wav(tmp_path / "mono8.wav", u8, 1)
    _write_wav(tmp_path / "mono32.wav", s24 * 256, 4)
    _write_wav(tmp_path / "cd.wav", s16, 2, rate=44100)
    (tmp_path / "junk.wav").write_bytes(b"not a wave file")
    f32 = (s16[:7001] / np.float32(32768.0)).astype("<f4")   # IEEE float (format tag 3), odd data size + a trailing chunk
    (tmp_path / "float.wav").write_bytes(b"RIFF" + (36 + f32.nbytes + 12).to_bytes(4, "little") + b"WAVEfmt " + (16).to_bytes(4, "little")
                                         + np.array([3, 1], "<u2").tobytes() + np.array([22050, 88200], "<u4").tobytes()
                                         + np.array([4, 32], "<u2").tobytes() + b"data" + f32.nbytes.to_bytes(4, "little")
                                         + f32.tobytes() + b"LIST" + (4).to_bytes(4, "little") + b"INFO")
    df = B.WavDecoder.decode(str(tmp_path / "float.wav")).pcm_frames
    assert df.dtype == np.float32 and df.shape == (7001, 1) and np.array_equal(df[:, 0], f32)
    d = B.WavDecoder.decode(str(tmp_path / "mono16.wav"))
    assert d.pcm_frames.dtype == np.int16 and np.array_equal(d.pcm_frames[:, 0], s16) and d.sample_array.size == 0
    assert abs(d.duration - 30000 / 22050) < 1e-9 and d.path.endswith("mono16.wav")
    assert np.array_equal(B.WavDecoder.decode(str(tmp_path / "stereo16.wav")).pcm_frames, st)
    d24 = B.WavDecoder.decode(str(tmp_path / "mono24.wav")).pcm_frames
    assert d24.dtype == np.int32 and np.array_equal(d24[:, 0], s24 * 256)   # pcm_s24le -> s32: x << 8
    assert np.array_equal(B.WavDecoder.decode(str(tmp_path / "mono32.wav")).pcm_frames[:, 0], s24 * 256)
    d8 = B.WavDecoder.decode(str(tmp_path / "mono8.wav")).pcm_frames
    assert d8.dtype == np.int16 and np.array_equal(d8[:, 0], (u8 - 128) * 256)   # pcm_u8: (x - 128) * 2^-7
    dcd = B.WavDecoder.decode(str(tmp_path / "cd.wav"))
    assert dcd.pcm_rate == 44100 and d.pcm_rate == 22050 and np.array_equal(dcd.pcm_frames[:, 0], s16)
    assert abs(dcd.duration - 30000 / 44100) < 1e-9
    _write_wav(tmp_path / "slow.wav", s16[:100], 2, rate=500)
    for bad in ("slow.wav", "junk.wav", "missing.wav"):
        with pytest.raises(B.DecodingError):
            B.WavDecoder.decode(str(tmp_path / bad))
    calls = []

    def fake_pcm(frames, sample_rate=22050, opts=None):
        calls.append(("pcm", frames[0].dtype.str, frames[0].shape[1], len(frames)) + ((sample_rate,) if sample_rate != 22050 else ()))
        return [B.Analysis(np.full(23, f.shape[1] + f.dtype.itemsize / 10)) for f in frames]

    def fake_f32(arrays, opts=None):
        calls.append(("f32", len(arrays)))
        return [B.Analysis(np.zeros(23)) for _ in arrays]

    monkeypatch.setattr(B.song, "analyze_batch_pcm", fake_pcm)
    monkeypatch.setattr(B.song, "analyze_batch", fake_f32)
    names = ["mono16.wav", "stereo16.wav", "cd.wav", "mono24.wav", "mono8.wav", "junk.wav", "mono32.wav"]
    got = dict(B.WavDecoder.analyze_paths_with_options([str(tmp_path / n) for n in names], B.AnalysisOptions(number_cores=1)))
    assert len(got) == len(names)
    tag = {n: (got[str(tmp_path / n)].analysis.as_arr1()[0] if isinstance(got[str(tmp_path / n)], B.Song) else None) for n in names}
    assert tag["junk.wav"] is None and round(float(tag["cd.wav"]), 1) == 1.2
    assert [round(float(tag[n]), 1) for n in ("mono16.wav", "stereo16.wav", "mono24.wav", "mono8.wav", "mono32.wav")] == \
        [1.2, 2.2, 1.4, 1.2, 1.4]                                       # every row back with its own song
    assert sorted(calls) == [("pcm", "<i2", 1, 1, 44100), ("pcm", "<i2", 1, 2), ("pcm", "<i2", 2, 1), ("pcm", "<i4", 1, 2)]
    # songs that carry sample_array keep going through analyze_batch, in the same batch
    calls.clear()
    mixed = [B.PreAnalyzedSong(path="a", sample_array=np.zeros(9000, np.float32)), B.WavDecoder.decode(str(tmp_path / "mono16.wav"))]
    res = B.analyze_decoded(mixed)
    assert sorted(calls) == [("f32", 1), ("pcm", "<i2", 1, 1)] and res[0].as_arr1()[0] == 0.0 and res[1].as_arr1()[0] > 1.0


CUE_SHEET = """REM GENRE Random
REM DATE 2022
REM DISCNUMBER 1
PERFORMER "Polochon_street"
TITLE "Album for CUE test"
FILE "%s" WAVE
  TRACK 01 AUDIO
    TITLE "Renaissance"
    PERFORMER "David TMX"
    INDEX 01 0:00:00
  TRACK 02 AUDIO
    TITLE "Piano"
    PERFORMER "Polochon_street"
    INDEX 01 0:11:05
  TRACK 03 AUDIO
    TITLE "Tone"
    PERFORMER "Polochon_street"
    INDEX 01 0:16:69

FILE "not-existing.wav" WAVE
  TRACK 01 AUDIO
    TITLE "Nope"
    PERFORMER "Charlie"
    INDEX 01 0:00:00
  TRACK 02 AUDIO
    TITLE "Nope"
    PERFORMER "Charlie"
    INDEX 01 0:10:00
"""  # data/testcue.cue of the reference, the audio file's name left open


def test_cue_sheet_tracks_are_slices_of_one_decoded_buffer(tmp_path, monkeypatch):
    """bliss-rs_b200/cue.py against src/cue.rs and its test (:262-420): track boundaries in f32 exactly as the
    reference computes them -- the three durations its test asserts for data/testcue.cue pin 0:11:05 -> sample 244 020
    and 0:16:69 -> 373 086 on the 496 272-sample file --, tags and paths of the songs, a missing audio file as an
    error item, every slice of the sheet in ONE analysis call, .cue paths inside Decoder::analyze_paths."""
    total = 496272
    rng = np.random.default_rng(0)
    s16 = rng.integers(-3000, 3000, total).astype(np.int16)
    _write_wav(tmp_path / "album.wav", s16, 2)
    sheet = tmp_path / "album.cue"
    sheet.write_text(CUE_SHEET % "album.wav")
    calls = []

    def fake(songs, opts=None):
        calls.append([np.asarray(p.pcm_frames)[:, 0].copy() for p in songs])
        return [B.Analysis(np.full(23, float(len(p.pcm_frames)))) for p in songs]

    monkeypatch.setattr(B.cue, "analyze_decoded", fake)
    got = B.BlissCue(B.WavDecoder).songs_from_path(str(sheet))
    assert len(got) == 4 and len(calls) == 1 and len(calls[0]) == 3
    for piece, (a, b) in zip(calls[0], ((0, 244020), (244020, 373086), (373086, total))):
        assert np.array_equal(piece, s16[a:b])
    durations = [np.float32(11.066666603), np.float32(5.853333473), np.float32(5.586666584)]   # src/cue.rs:311, 356, 402
    for i, (song, title, artist) in enumerate(zip(got, ("Renaissance", "Piano", "Tone"), ("David TMX", "Polochon_street", "Polochon_street"))):
        assert isinstance(song, B.Song) and song.path == "%s/CUE_TRACK%03d" % (sheet, i + 1)
        assert (song.title, song.artist, song.album, song.album_artist) == (title, artist, "Album for CUE test", "Polochon_street")
        assert (song.track_number, song.disc_number, song.genre) == (i + 1, 1, "Random")
        assert np.float32(song.duration) == durations[i]
        assert song.cue_info.cue_path == str(sheet) and song.cue_info.audio_file_path == str(tmp_path / "album.wav")
        assert song.analysis.as_arr1()[0] == float(len(calls[0][i]))            # rows stay with their tracks
    assert isinstance(got[3], B.DecodingError)                                   # not-existing.wav: one error for the file
    with pytest.raises(B.DecodingError):
        B.BlissCue(B.WavDecoder).songs_from_path(str(tmp_path / "nope.cue"))
    # a track that starts behind the end of the audio: an error item here (the reference's slice would panic)
    (tmp_path / "late.cue").write_text('FILE "album.wav" WAVE\n TRACK 01 AUDIO\n  INDEX 01 0:00:00\n TRACK 02 AUDIO\n  INDEX 01 9:00:00\n')
    late = B.BlissCue(B.WavDecoder).songs_from_path(str(tmp_path / "late.cue"))
    assert len(late) == 2 and all(isinstance(x, B.DecodingError) for x in late)
    # Decoder::analyze_paths: every track of a .cue path is an item under the sheet's path (src/song/decoder.rs:305-318)
    monkeypatch.setattr(B.song, "analyze_decoded", fake)
    items = list(B.WavDecoder.analyze_paths_with_options([str(tmp_path / "album.wav"), str(sheet)], B.AnalysisOptions(number_cores=2)))
    assert sorted(p for p, _ in items) == sorted([str(tmp_path / "album.wav")] + [str(sheet)] * 4)
    assert sum(isinstance(r, B.Song) for _, r in items) == 4
    # ... and in the Library: "passing vec![file.cue] will add individual tracks with the cue_info field set in the
    # database" (src/library.rs:891-892); the sheet's missing second FILE is a failed-song row under the sheet's path
    lib = B.library.Library(str(tmp_path / "songs.db"), decoder=B.WavDecoder)
    assert lib.update_library([str(sheet)]) == (3, 1)
    stored = sorted(lib.songs_from_library(), key=lambda x: x.bliss_song.path)
    assert [x.bliss_song.path for x in stored] == ["%s/CUE_TRACK%03d" % (sheet, i) for i in (1, 2, 3)]
    assert all(x.bliss_song.cue_info.cue_path == str(sheet) and x.bliss_song.cue_info.audio_file_path == str(tmp_path / "album.wav")
               for x in stored)
    assert [x.bliss_song.title for x in stored] == ["Renaissance", "Piano", "Tone"]
    assert [f.song_path for f in lib.get_failed_songs()] == [str(sheet)]
    lib.close()


def test_fft_index_logic_on_host(tmp_path):
    """tests/cpu_emul/emul_fft.cu runs the warp / CTA FFT passes of pvoc512.cuh and rfft8192.cuh (incl. the
    fused pass 3 + mirror-pair epilogue) thread by thread on the host and compares with an f64 DFT."""
    exe = str(tmp_path / "emul_fft")
    src = os.path.join(ROOT, "tests", "cpu_emul", "emul_fft.cu")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


def _build_kernel_emulation(tmp_path):
    exe = str(tmp_path / "emul_kernels")
    here = os.path.join(ROOT, "tests", "cpu_emul")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DBLISS_HOST_EMUL", "-I",
                           os.path.join(here, "cuda_on_cpu"), "-o", exe, os.path.join(here, "emul_kernels.cpp")])
    return exe


def _build_host_emulated_library(tmp_path):
    """The library's own sources (api.cu included) compiled with g++ against tests/cpu_emul/cuda_on_cpu: a TEST build
    in a temporary directory, never installed, never looked for by the package."""
    here = os.path.join(ROOT, "tests", "cpu_emul")
    csrc = os.path.join(ROOT, "bliss-rs_b200", "csrc")
    names = ["spectral", "tempo", "chroma", "finalize", "distance", "gather", "wave_setup", "api"]
    procs = [subprocess.Popen(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-x", "c++", "-DBLISS_HOST_EMUL",
                               "-I", os.path.join(here, "cuda_on_cpu"), "-I", os.path.join(ROOT, "include"), "-c",
                               os.path.join(csrc, n + ".cu"), "-o", str(tmp_path / (n + ".o"))]) for n in names]
    assert all(p.wait() == 0 for p in procs)
    so = str(tmp_path / "libbliss_b200_hostemu_TEST.so")
    subprocess.check_call(["g++", "-shared", "-o", so] + [str(tmp_path / (n + ".o")) for n in names] + ["-lpthread"])
    return so


def test_c_abi_on_the_host_emulated_library(tmp_path):
    """The whole library -- host logic of api.cu (chunked host path, waves, descriptors, raw-PCM staging, taps,
    playlist calls) AND every kernel -- as a TEST build for the host: the same sources compiled with g++ against
    tests/cpu_emul/cuda_on_cpu (kernel launches run the kernel's source thread by thread at the point of the call,
    the runtime API is a synchronous stub).  A slice of the GPU suite then runs against it through the real C ABI
    and the Python binding (BLISS_B200_SO), without a GPU: golden vector, too-short songs, 16-bit and
    interleaved-PCM ingest, WAV files through the decoder pipeline, CUE-style sub-slices through the device API, both STFT micro-benchmark kernels,
    distances and playlist orders.  This is how host-side changes made without a GPU are checked; it is not a
    product path (the product library is built by nvcc, and bliss_b200_init fails without a device)."""
    so = _build_host_emulated_library(tmp_path)
    env = dict(os.environ, BLISS_B200_SO=so, CUDA_VISIBLE_DEVICES="")
    pick = ("golden_clip_v2 or too_short or s16_ingest or pcm_feed or distance_known or distance_matrix_bit or "
            "closest_to_songs or dedup or stft512_magnitudes or experimental_stft_pair or cue_style or wav_files or library_playlists "
            "or frame_edges or chroma_filter or packed_distance or resample")
    # the C++17 host mirror (include/bliss_b200.hpp: Song, Decoder, WavDecoder, analyze_batch[_s16|_pcm], playlist) end to
    # end, started first and left running beside the Python slice
    exe = str(tmp_path / "host_mirror_emu")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"), "-o", exe, so, "-Wl,-rpath," + str(tmp_path)])
    mirror = subprocess.Popen([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=dict(env, TMPDIR=str(tmp_path)))
    workers = str(max(1, min(4, (os.cpu_count() or 2) - 1)))  # the slice's tests are independent processes' worth of work
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-q", "-m", "gpu",
                          "-p", "no:cacheprovider", "--tb=short", "-n", workers, "-k", pick],
                         capture_output=True, text=True, env=env, cwd=ROOT)
    tail = out.stdout[-1500:] + out.stderr[-500:]
    mirror_out, mirror_err = mirror.communicate(timeout=1200)
    assert out.returncode == 0, tail
    m = re.search(r"(\d+) passed", out.stdout)
    assert m and int(m.group(1)) >= 10 and "failed" not in out.stdout, tail
    assert mirror.returncode == 0 and mirror_out.strip().endswith("OK"), mirror_out + mirror_err


def test_kernel_sources_reproduce_the_reference_golden_vector(tmp_path, golden):
    """The whole path -- timedomain, pvoc512, peakpick, beattrack, stft8192, tuning_select, chroma_filter_table,
    chroma_pipe, finalize: the kernels' own source, launched in run_wave's order by tests/cpu_emul/emul_kernels.cpp on
    the host (cuda_on_cpu) -- on the reference's golden clip: its 23 expected values (src/song/mod.rs:553-580) within
    the reference's own 1e-5, once with the measured kernels and once with every experimental cut switched on."""
    from oracle import oracle as O
    exe = _build_kernel_emulation(tmp_path)
    x = golden["pcm_s16_mono"].astype(np.float32) / np.float32(32768.0)
    song = str(tmp_path / "song.f32")
    x.tofile(song)
    out = subprocess.run([exe, song, str(tmp_path), "full"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
    rc, want = O.analyze(x, 2)
    assert rc == 0
    for tag in ("default", "all_cuts"):
        f = np.fromfile(str(tmp_path / ("features_" + tag)), np.float32)
        tuning_idx, n_bpm, tempo = np.fromfile(str(tmp_path / ("misc_" + tag)), np.float32)
        assert f.shape == (23,)
        assert np.abs(f - golden["expected_analysis_v2"]).max() < 1e-5, (tag, f)
        assert np.abs(f - want).max() < 1e-5, (tag, np.abs(f - want).max())
        assert int(tuning_idx) == 45 and n_bpm > 0 and abs(tempo - want[0]) < 1e-6  # tuning -0.05 (src/chroma.rs:657-665)
    # beattrack_kernel's three autocorrelation cuts (balanced lag pairs = the default, one lag at a time, four consecutive
    # lags per thread) keep every lag's sum in the reference's order: tempo, count and every BPM of the list bit for bit
    modes = np.split(a := np.fromfile(str(tmp_path / "acf_modes_default"), np.float32), np.nonzero(a == np.float32(-12345.0))[0] + 1)[:3]
    assert len(modes) == 3 and modes[0].size > 10
    for m in modes[1:]:
        assert np.array_equal(m.view(np.uint32), modes[0].view(np.uint32))


def test_distance_kernel_sources_are_bit_exact_on_the_host(tmp_path):
    """tests/cpu_emul/emul_distance.cpp: distance_matrix_kernel<23|20>, the generic kernel, seed_distance_kernel +
    key packing and nearest_alive_kernel of distance.cu run on the host (cuda_on_cpu).  These kernels spell their
    f32 order out (__fmul_rn / __fadd_rn, ndarray's unrolled_dot), so every distance and both playlist orders --
    ties between exact duplicates included -- must equal the oracle BIT FOR BIT (src/playlist.rs:65-142, 256-326)."""
    from oracle import oracle as O
    exe = str(tmp_path / "emul_distance")
    here = os.path.join(ROOT, "tests", "cpu_emul")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DBLISS_HOST_EMUL", "-I",
                           os.path.join(here, "cuda_on_cpu"), "-o", exe, os.path.join(here, "emul_distance.cpp")])
    rng = np.random.default_rng(3)
    for dim in (23, 20, 7):
        n, n_seeds = 70, 2
        rows = (rng.random((n, dim), dtype=np.float32) * 2 - 1).astype(np.float32)
        rows[5], rows[11] = rows[9], rows[0]            # exact duplicates: ties in both orderings
        d = tmp_path / ("dim%d" % dim)
        d.mkdir()
        rows.tofile(str(d / "rows.f32"))
        out = subprocess.run([exe, str(d / "rows.f32"), str(n), str(dim), str(n_seeds), str(d)], capture_output=True, text=True)
        assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr

        def ld(name, t=np.float32):
            return np.fromfile(str(d / name), t)

        w = np.ones(dim, np.float32)
        if dim == 23:
            w[0], w[10:] = 0.25, np.float32(3.0) / np.float32(13.0)  # VERSION2_WEIGHTS, src/lib.rs:209-234
        M, full = np.diag(w).astype(np.float32), ld("full_matrix").reshape(dim, dim)
        mw, me, mf, mc = (ld("matrix_" + k).reshape(n, n) for k in ("weights", "euclidean", "full", "cosine"))
        for i in range(0, n, 7):
            for j in range(n):
                a, b = rows[i], rows[j]
                assert np.float32(O.mahalanobis_distance(a, b, M)) == mw[i, j]
                assert np.float32(O.euclidean_distance(a, b)) == me[i, j]
                assert np.float32(O.mahalanobis_distance(a, b, full)) == mf[i, j]
                c = np.float32(O.cosine_distance(a, b))
                assert c == mc[i, j] or (np.isnan(c) and np.isnan(mc[i, j]))
        if dim in (20, 23):  # FeaturesVersion::distance_metric, src/lib.rs:176-178
            got = mw if dim == 23 else me
            for i in range(0, n, 9):
                for j in range(0, n, 3):
                    assert np.float32(O.default_distance(rows[i], rows[j], 2 if dim == 23 else 1)) == got[i, j]
        order, keys = O.closest_to_songs(rows[:n_seeds], rows, M)
        assert np.array_equal(order, ld("closest_order", np.uint32)) and np.array_equal(keys, ld("closest_keys"))
        assert np.array_equal(O.song_to_song(rows[:n_seeds], rows, M), ld("chain_order", np.uint32))


def test_kernel_sources_on_a_ragged_batch(tmp_path, golden):
    """The whole path over five songs at once -- ragged lengths, one too short (status 1, src/song/mod.rs:417-430),
    one of exactly 8192 samples -- with descriptors and prefix arrays laid out as plan_wave (api.cu) lays them out:
    song lookup, item boundaries and the per-song offsets of every kernel, on the host, against the oracle; measured
    kernels and all experimental cuts."""
    from oracle import oracle as O
    exe = _build_kernel_emulation(tmp_path)
    a = golden["pcm_s16_mono"].astype(np.float32) / np.float32(32768.0)
    b = golden["pcm_piano"].astype(np.float32) / np.float32(32768.0)
    songs = [a[:90001], b[:5000], b[:70003], a[100000:100000 + 8192], a[50000:50000 + 33333]]
    flat = np.concatenate([np.pad(x, (0, (-len(x)) % 4)) for x in songs]).astype(np.float32)
    path = str(tmp_path / "batch.f32")
    flat.tofile(path)
    out = subprocess.run([exe, path, str(tmp_path), "batch"] + [str(len(x)) for x in songs], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
    for tag in ("default", "all_cuts"):
        f = np.fromfile(str(tmp_path / ("batch_features_" + tag)), np.float32).reshape(len(songs), 23)
        for i, x in enumerate(songs):
            rc, want = O.analyze(x, 2)
            if rc == 0:
                assert np.abs(f[i] - want).max() < 1e-5, (tag, i, np.abs(f[i] - want).max())
            else:
                assert rc == 1 and (f[i] == 0).all()  # the row of a rejected song is left alone


def test_kernel_sources_run_on_the_host(tmp_path, golden):
    """tests/cpu_emul/emul_kernels.cpp: the SOURCE of pvoc512_kernel, stft512_pairs_kernel, timedomain_kernel,
    pcm_to_mono_kernel and stft8192_kernel -- the measured builds and every experimental BLISS_B200_VARIANT cut --
    compiled with g++ against cuda_on_cpu/cuda_runtime.h (each CUDA thread a fiber; shuffles, reductions and
    __syncthreads() real rendezvous points) and run on a 40 000-sample clip with a stretch of digital silence.
    Outputs against the oracle with the bars of the GPU stage tests; address-only and same-tree variants must be
    bit-identical to the measured kernels (the host has no FMA contraction and an IEEE sqrt, so this says the
    variants compute the same thing, not what the device rounds to)."""
    from oracle import oracle as O
    exe = _build_kernel_emulation(tmp_path)
    x = (golden["pcm_s16_mono"][20000:60000].astype(np.float32) / np.float32(32768.0)).copy()
    x[15000:16500] = 0.0
    song = str(tmp_path / "song.f32")
    x.tofile(song)
    out = subprocess.run([exe, song, str(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr

    def ld(name, t=np.float32):
        return np.fromfile(str(tmp_path / name), t)

    def same_bits(a, b):
        return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))

    # ---- pvoc512_kernel and its cuts: per-frame descriptors + spectral flux
    c, r, f = O.timbral_frames(x)
    _, flux, _, _ = O.tempo(x, taps=True)
    for tag in ("default", "v512", "v1024", "v2048", "v3584", "v2"):
        ce, ro, fl, fx = (ld("%s_%s" % (k, tag)) for k in ("centroid", "rolloff", "flatness", "flux"))
        assert ce.shape == c.shape and fx.shape == flux.shape
        assert (np.abs(ce - c) / np.maximum(1.0, np.abs(c))).max() < 1e-4, tag
        assert np.mean(ro != r) < 2e-3 and np.abs(ro - r).max() <= 2 * 22050 / 512 + 1e-3, tag
        assert (np.abs(fl - f) / (2e-4 * np.abs(f) + 1e-5)).max() < 1.0, tag
        assert np.abs(fx - flux).max() / np.abs(flux).max() < 1e-5, tag
    for tag in ("v1024", "v2048"):   # transposed reductions / tile padding: the same arithmetic, the same trees
        for k in ("centroid", "rolloff", "flatness", "flux"):
            assert same_bits(ld("%s_%s" % (k, tag)), ld(k + "_default")), (tag, k)
    # ---- STFT micro-benchmark kernels
    want = O.tempo_norms(x)
    for tag in ("default", "v256"):
        m = ld("stft512_" + tag).reshape(-1, 257)
        assert m.shape == want.shape and (m >= 0).all(), tag
        assert np.abs(m - want).max() / want.max() < 1e-6, tag
    # ---- timedomain_kernel
    assert int(ld("zcr_count", np.uint32)[0]) == O.number_crossings(x)
    chunks = ld("loudness_chunks")
    ms = np.array([np.mean(x[i:i + 1024].astype(np.float64) ** 2) for i in range(0, x.size, 1024)])
    assert chunks.shape == ms.shape and np.abs(chunks - ms).max() <= 1e-6 * ms.max()
    # ---- pcm_to_mono_kernel: bit-exact against the oracle
    assert same_bits(ld("mono_from_s16_stereo"), O.pcm_to_mono(ld("in_s16_stereo", np.int16).reshape(-1, 2)))
    assert same_bits(ld("mono_from_s32"), O.pcm_to_mono(ld("in_s32", np.int32)))
    assert same_bits(ld("mono_from_f32x3"), O.pcm_to_mono(ld("in_f32x3").reshape(-1, 3)))
    # ---- stft8192_kernel and its cuts: magnitudes + pip-track candidates
    S = O.stft(x, 8192, 2205)
    p, _ = O.pip_track(S, 8192)
    for tag in ("default", "v64", "v128", "v4096", "v4288", "v8192", "v12480", "old_epilogue", "v2", "v3"):
        g = ld("stft8192_" + tag).reshape(-1, 4097)
        assert g.T.shape == S.shape and (g >= 0).all(), tag
        assert np.abs(g.T - S).max() / S.max() < 2e-6, tag
        assert int(ld("peaks_" + tag, np.uint32)[0]) == p.size, tag
        assert np.abs(ld("peak_pitches_" + tag, np.float64) - np.sort(p)).max() < 1e-3, tag
    assert same_bits(ld("stft8192_v4096"), ld("stft8192_default"))       # addresses only
    g0, g8 = ld("stft8192_default").reshape(-1, 4097), ld("stft8192_v8192").reshape(-1, 4097)
    changed = np.nonzero((g0 != g8).any(1))[0]  # the rotated transform touches interior frames that start on an odd sample only
    assert len(changed) > 0 and all((2205 * int(f) - 4096) % 2 == 1 and 2205 * int(f) - 4096 >= 0 for f in changed)
    # ---- and the whole path on this clip (too short for a beat: tempo = -1, src/temporal.rs:66-77)
    rc, feats = O.analyze(x, 2)
    for tag in ("default", "all_cuts"):
        got = ld("features_" + tag)
        assert rc == 0 and got.shape == (23,) and np.abs(got - feats).max() < 1e-5, tag


def test_stft_pair_kernel_design_on_host(tmp_path):
    """tests/cpu_emul/emul_stft_pairs.cu: the hop-256 pairing of stft512_pairs_kernel (window rows, slide,
    load bounds, silence, magnitudes against an f64 DFT) transcribed onto the host."""
    exe = str(tmp_path / "emul_stft_pairs")
    src = os.path.join(ROOT, "tests", "cpu_emul", "emul_stft_pairs.cu")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "-o", exe, src])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


def test_transposed_pair_reduction_tree_is_the_butterfly():
    """pair_sum / pair_prod of spectral.cu (VARIANT_PV_PAIRDESC) reduce frame A's partials into lanes 0..15 and
    frame B's into lanes 16..31 with one shuffle per step.  Their tree must be the xor-butterfly of warp_sum /
    warp_prod, value for value -- modelled here lane by lane in f32 / f64 -- so that the variant's descriptors are
    bit-identical to the default kernel's."""
    lane = np.arange(32)
    hi = (lane & 16) != 0

    def butterfly(v, op):
        v = v.copy()
        for o in (16, 8, 4, 2, 1):
            v = op(v, v[lane ^ o])
        return v

    def transposed(a, b, op):
        mine, send = np.where(hi, b, a), np.where(hi, a, b)
        mine = op(mine, send[lane ^ 16])
        for o in (8, 4, 2, 1):
            mine = op(mine, mine[lane ^ o])
        return mine

    rng = np.random.default_rng(0)
    for _ in range(500):
        a = (rng.standard_normal(32) * 10 ** rng.uniform(-3, 3)).astype(np.float32)
        b = rng.standard_normal(32).astype(np.float32)
        p = transposed(a, b, np.add)
        assert np.array_equal(p[:16].view(np.uint32), butterfly(a, np.add)[:16].view(np.uint32))
        assert np.array_equal(p[16:].view(np.uint32), butterfly(b, np.add)[16:].view(np.uint32))
        da, db = rng.uniform(1, 2, 32), rng.uniform(1, 2, 32)
        q = transposed(da, db, np.multiply)
        assert np.array_equal(q[:16], butterfly(da, np.multiply)[:16]) and np.array_equal(q[16:], butterfly(db, np.multiply)[16:])


def test_natural_order_tile_padding_wavefront_model():
    """Why VARIANT_PV_ZPOS4 and the unpadded tile of stft512_pairs_kernel exist: shared-memory wavefronts of the
    512-point kernels' natural-order tile under the usual bank model (a 64-bit warp access = two half-warp
    wavefronts, more when two lanes of a half-warp hit different 8-byte words of one bank pair).  The measured
    kernel's padding k + (k >> 3) serves its loads (bins 8 lane + i) without conflicts but gives every one of its 16
    stores a 2-way conflict -- 32 extra wavefronts per frame pair, the 35 conflict wavefronts per pair ncu reports
    (profiles/ncu_r01b_full_128songs.md: 69.8 M conflicts over 1.98 M pairs)."""
    def bitrev4(x):
        return int("{:04b}".format(x)[::-1], 2)

    def wavefronts(idx):
        tot = 0
        for h in (0, 1):
            words = {}
            for lane in range(16 * h, 16 * h + 16):
                words.setdefault(idx[lane] % 16, set()).add(idx[lane])
            tot += max(len(v) for v in words.values())
        return tot

    def bin_of(lane, q):
        return (lane & 15) + 16 * (bitrev4(q) + 16 * (lane >> 4))

    def cost(z, own):
        st = sum(wavefronts([z(bin_of(l, q)) for l in range(32)]) for q in range(16))
        ld = sum(wavefronts([z(own(l, i)) for l in range(32)]) for i in range(8))
        mi = sum(wavefronts([z((512 - own(l, i)) & 511) for l in range(32)]) for i in range(8))
        return st, ld, mi

    eight_per_lane = lambda l, i: 8 * l + i     # pvoc512_kernel
    strided = lambda l, i: l + 32 * i           # stft512_pairs_kernel
    assert cost(lambda k: k + (k >> 3), eight_per_lane) == (64, 16, 16)   # measured layout: stores 2x the ideal 32
    assert cost(lambda k: k + (k >> 4), eight_per_lane) == (32, 16, 18)   # VARIANT_PV_ZPOS4: -30 wavefronts per pair
    assert cost(lambda k: k + (k >> 3), strided) == (64, 32, 32)
    assert cost(lambda k: k, strided) == (32, 16, 16)                     # stft512_pairs_kernel: the ideal 64


def test_fft8192_buffer_layout_wavefront_model():
    """Why VARIANT_LAY16 exists: shared-memory wavefronts per frame of stft8192_kernel's FFT buffer (8-byte elements;
    logical element 256 r + 16 c + m at 273 r + LB c + m) under the same bank model.  With the measured LB = 17 the
    three passes are conflict-free but every mirror load of the pair epilogue (thread t reads the block thread 256 - t
    published) puts lanes 0 and 15 of a half-warp into one bank: 256 instead of 136 wavefronts per frame, most of
    the 202 conflict wavefronts per frame of the ncu capture (46.5 M over 230 400 frames)."""
    def wavefronts(idx):
        tot = 0
        for h in range(0, len(idx), 16):
            words = {}
            for i in idx[h:h + 16]:
                words.setdefault(i % 16, set()).add(i)
            tot += max(len(v) for v in words.values())
        return tot

    def model(lb):
        pos = lambda r, c, m: 273 * r + lb * c + m
        t256 = range(256)
        out = {"pass1 st": sum(wavefronts([pos(k1, b >> 4, b & 15) for b in t256]) for k1 in range(16)),
               "pass2 ld": sum(wavefronts([pos(b >> 4, q, b & 15) for b in t256]) for q in range(16)),
               "pass3 ld": sum(wavefronts([pos(t & 15, t >> 4, q) for t in t256]) for q in range(16)),
               "pass3 st": sum(wavefronts([pos(t & 15, t >> 4, m) for t in t256]) for m in range(8, 16))}

        def mirror(t, M):
            if t == 0:
                return pos(0, 0, (16 - M) & 15)
            tp = (256 - t) & 255
            return pos(tp & 15, tp >> 4, 15 - M)
        out["mirror ld"] = sum(wavefronts([mirror(t, M) for t in t256]) for M in range(8))
        assert len({pos(r, c, m) for r in range(16) for c in range(16) for m in range(16)}) == 4096  # injective
        return out

    assert model(17) == {"pass1 st": 256, "pass2 ld": 256, "pass3 ld": 256, "pass3 st": 128, "mirror ld": 256}
    assert model(16) == {"pass1 st": 256, "pass2 ld": 256, "pass3 ld": 256, "pass3 st": 128, "mirror ld": 136}


def test_rust_shim_files_match_integration_md():
    """integration/rust/*.rs are INTEGRATION.md's ```rust blocks (scripts/extract_rust_shim.py), and every C-ABI
    function the Rust side declares exists in the header with that name."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("extract_rust_shim", os.path.join(ROOT, "scripts", "extract_rust_shim.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hdr = open(os.path.join(ROOT, "include", "bliss_b200.h")).read()
    for name, text in mod.render().items():
        assert open(os.path.join(ROOT, "integration", "rust", name)).read() == text, name
        for fn in re.findall(r"\bfn (bliss_b200_[a-z0-9_]+)\(", text):
            assert re.search(r"\b%s\s*\(" % fn, hdr), fn


def test_variant_mask_names_match_header():
    """BLISS_B200_VARIANT bits (A/B switch back to a kernel's previous implementation) stay documented."""
    txt = open(os.path.join(ROOT, "bliss-rs_b200", "csrc", "common.cuh")).read()
    for name in ("VARIANT_OLD_EPILOGUE = 1", "VARIANT_OLD_TUNING = 2", "VARIANT_OLD_CHROMA = 4", "VARIANT_OLD_ACF = 8",
                 "VARIANT_BT512 = 16", "VARIANT_TWPROD = 64", "VARIANT_WINSYN = 128", "VARIANT_STFT_PAIRS = 256", "VARIANT_PV_TWPROD = 512", "VARIANT_PV_PAIRDESC = 1024", "VARIANT_PV_ZPOS4 = 2048", "VARIANT_LAY16 = 4096", "VARIANT_ODDSHIFT = 8192"):
        assert name in txt


def _build_cpp_mirror(tmp_path):
    exe = str(tmp_path / "host_mirror")
    libdir = os.path.dirname(B.native.SO_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"), "-o", exe, "-L" + libdir,
                           "-lbliss_b200", "-Wl,-rpath," + libdir])
    return exe


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_cpp_host_mirror_builds_and_refuses_to_run_without_gpu(tmp_path):
    """include/bliss_b200.hpp (C++17 mirror of Song / Analysis / Decoder / playlist) compiles and links against the
    C ABI; the device-free parts work and the first analysis fails loudly: there is no CPU fallback."""
    out = subprocess.run([_build_cpp_mirror(tmp_path)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.strip() == "NO_DEVICE"


def test_cpp_decoder_pipeline_under_thread_sanitizer(tmp_path):
    """tests/cpp/pipeline_tsan.cpp: Decoder::analyze_paths of the C++ mirror -- decoding threads, bounded hand-over, one
    batcher, the two failure paths -- built with -fsanitize=thread over a C ABI stubbed inside the test program: 27 000
    songs through 1 / 3 / 8 decoding threads and batch sizes 1 / 7 / 64, every path answered once, rows with their
    songs, GPU calls never overlapping, no data race reported."""
    exe = str(tmp_path / "pipeline_tsan")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-Wall", "-Wextra", "-Werror",
                            "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
                            os.path.join(ROOT, "tests", "cpp", "pipeline_tsan.cpp"), "-o", exe], capture_output=True, text=True)
    if build.returncode != 0 and "tsan" in build.stderr.lower():
        pytest.skip("no ThreadSanitizer runtime in this image")
    assert build.returncode == 0, build.stderr[-2000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("OK") and "ThreadSanitizer" not in out.stderr, out.stdout + out.stderr[-3000:]


def test_header_is_c99_and_the_library_links_from_c(tmp_path):
    """include/bliss_b200.h under gcc -std=c99 -pedantic -Werror; tests/c/abi_c99.c links against the library, uses the
    device-free entry points and sees bliss_b200_init refuse a box without a GPU (on a GPU box: one too-short song)."""
    exe = str(tmp_path / "abi_c99")
    libdir = os.path.dirname(B.native.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_c99.c"), "-o", exe, "-L" + libdir, "-lbliss_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.startswith("DEVICE_OK" if torch.cuda.is_available() else "NO_DEVICE")


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path):
    """the same program on a GPU box: Song::analyze, the Decoder batching seam (errors as items), the 16-bit
    entry point and closest_to_songs through the C++ mirror"""
    out = subprocess.run([_build_cpp_mirror(tmp_path)], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, TMPDIR=str(tmp_path)))
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert out.stdout.strip() == "OK"
