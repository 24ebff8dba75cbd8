"""ctypes binding of the CPU ORACLE (oracle/bliss_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package
(bliss-rs_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbliss_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "bliss_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libbliss_oracle.so"])
    return _SO


_lib = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.bo_analyze.argtypes = [f32p, C.c_uint64, C.c_int, f32p]
        L.bo_analyze.restype = C.c_int
        L.bo_analyze_batch.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_uint32,
                                       C.c_int, f32p, C.POINTER(C.c_int32), C.c_int]
        L.bo_analyze_batch.restype = C.c_int
        L.bo_stft_num_frames.argtypes = [C.c_uint64, C.c_uint32]
        L.bo_stft_num_frames.restype = C.c_uint32
        L.bo_reflect_pad.argtypes = [f32p, C.c_uint64, C.c_uint32, f32p]
        L.bo_stft.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint32, f64p]
        L.bo_stft.restype = C.c_uint32
        L.bo_geometric_mean.argtypes = [f32p, C.c_uint32]
        L.bo_geometric_mean.restype = C.c_float
        L.bo_number_crossings.argtypes = [f32p, C.c_uint64]
        L.bo_number_crossings.restype = C.c_uint32
        L.bo_timbral_frames.argtypes = [f32p, C.c_uint64, C.c_uint32, f32p, f32p, f32p, C.c_void_p]
        L.bo_summarise.argtypes = [f32p, C.c_uint32, C.c_int, f32p]
        L.bo_zcr.argtypes = [f32p, C.c_uint64]
        L.bo_zcr.restype = C.c_float
        L.bo_loudness.argtypes = [f32p, C.c_uint64, C.c_int, f32p]
        L.bo_tempo.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.POINTER(C.c_uint32)]
        L.bo_tempo.restype = C.c_float
        L.bo_tempo_norms.argtypes = [f32p, C.c_uint64, C.c_uint32, f32p]
        L.bo_pip_track.argtypes = [f64p, C.c_uint32, C.c_uint32, f64p, f64p]
        L.bo_pip_track.restype = C.c_uint64
        L.bo_hz_to_octs.argtypes = [f64p, C.c_uint64, C.c_double, C.c_uint32]
        L.bo_hz_to_octs.restype = None
        L.bo_pitch_tuning.argtypes = [f64p, C.c_uint64, C.c_double]
        L.bo_pitch_tuning.restype = C.c_double
        L.bo_estimate_tuning.argtypes = [f64p, C.c_uint32, C.c_uint32]
        L.bo_estimate_tuning.restype = C.c_double
        L.bo_chroma_filter.argtypes = [C.c_uint32, C.c_double, f64p]
        L.bo_chroma_stft.argtypes = [f64p, C.c_uint32, C.c_uint32, C.c_double, f64p]
        L.bo_normalize_feature_sequence.argtypes = [f64p, C.c_uint32, C.c_uint32, f64p]
        L.bo_extract_interval_features.argtypes = [f64p, C.c_uint32, f64p]
        L.bo_chroma_interval_features.argtypes = [f64p, C.c_uint32, f64p]
        L.bo_chroma_interval_features.restype = C.c_int
        L.bo_chroma_values.argtypes = [f64p, C.c_int, f32p]
        L.bo_chroma.argtypes = [f32p, C.c_uint64, C.c_int, f32p, C.POINTER(C.c_double), C.c_void_p]
        L.bo_chroma.restype = C.c_int
        for name in ("bo_euclidean_distance", "bo_cosine_distance"):
            getattr(L, name).argtypes = [f32p, f32p, C.c_uint32]
            getattr(L, name).restype = C.c_float
        L.bo_mahalanobis_distance.argtypes = [f32p, f32p, f32p, C.c_uint32]
        L.bo_mahalanobis_distance.restype = C.c_float
        L.bo_default_distance.argtypes = [f32p, f32p, C.c_int]
        L.bo_default_distance.restype = C.c_float
        L.bo_pcm_to_mono.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint32, f32p]
        L.bo_pcm_to_mono.restype = C.c_int
        L.bo_feature_weights.argtypes = [C.c_int, f32p]
        L.bo_closest_to_songs.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, C.c_uint32,
                                          C.c_void_p, u32p, C.c_void_p]
        L.bo_song_to_song.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, C.c_uint32, C.c_void_p,
                                      u32p]
        L.bo_fft.argtypes = [f32p, C.c_uint32]
        _lib = L
    return _lib


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def feature_count(version=2):
    return 20 if version == 1 else 23


def analyze(pcm, version=2):
    """Song::analyze_with_options (src/song/mod.rs:413-508). Returns (status, features)."""
    pcm = _f32(pcm)
    out = np.zeros(feature_count(version), np.float32)
    if pcm.size == 0:
        return 1, out
    rc = lib().bo_analyze(pcm, pcm.size, version, out)
    return rc, out


def analyze_batch(pcms, version=2, n_threads=1):
    pcms = [_f32(p) for p in pcms]
    n = len(pcms)
    ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pcms])
    lens = (C.c_uint64 * n)(*[p.size for p in pcms])
    out = np.zeros((n, feature_count(version)), np.float32)
    status = (C.c_int32 * n)()
    lib().bo_analyze_batch(ptrs, lens, n, version, out, status, n_threads)
    return np.array(status[:], np.int32), out


def stft(x, win, hop):
    """utils::stft -> [bins, frames] float64 like the reference."""
    x = _f32(x)
    frames = lib().bo_stft_num_frames(x.size, hop)
    out = np.zeros((frames, win // 2 + 1), np.float64)
    lib().bo_stft(x, x.size, win, hop, out)
    return out.T.copy()


def reflect_pad(x, pad):
    x = _f32(x)
    out = np.zeros(x.size + 2 * pad, np.float32)
    lib().bo_reflect_pad(x, x.size, pad, out)
    return out


def geometric_mean(v):
    v = _f32(v)
    return float(lib().bo_geometric_mean(v, v.size))


def number_crossings(x):
    x = _f32(x)
    return int(lib().bo_number_crossings(x, x.size))


def timbral_frames(x, n_frames=None, want_norms=False):
    """Per-frame centroid / rolloff / flatness (Hz, Hz, ratio)."""
    x = _f32(x)
    if n_frames is None:
        n_frames = (x.size - 512) // 128 + 1
    c = np.zeros(n_frames, np.float32)
    r = np.zeros(n_frames, np.float32)
    f = np.zeros(n_frames, np.float32)
    norms = np.zeros((n_frames, 256), np.float32) if want_norms else None
    lib().bo_timbral_frames(x, x.size, n_frames, c, r, f,
                            norms.ctypes.data if want_norms else None)
    return (c, r, f, norms) if want_norms else (c, r, f)


def summarise(v, kind):
    v = _f32(v)
    out = np.zeros(2, np.float32)
    lib().bo_summarise(v, v.size, kind, out)
    return out


def zcr(x):
    x = _f32(x)
    return float(lib().bo_zcr(x, x.size))


def loudness(x, chunks_exact=False):
    x = _f32(x)
    out = np.zeros(2, np.float32)
    lib().bo_loudness(x, x.size, int(chunks_exact), out)
    return out


def tempo(x, n_frames=None, silence_len=512, taps=False):
    x = _f32(x)
    if n_frames is None:
        n_frames = (x.size - 512) // 256 + 1
    nb = C.c_uint32(0)
    if not taps:
        return float(lib().bo_tempo(x, x.size, n_frames, silence_len, None, None, None, C.byref(nb)))
    flux = np.zeros(n_frames, np.float32)
    thr = np.zeros(n_frames, np.float32)
    bpms = np.zeros(max(n_frames, 1), np.float32)
    v = lib().bo_tempo(x, x.size, n_frames, silence_len, flux.ctypes.data, thr.ctypes.data,
                       bpms.ctypes.data, C.byref(nb))
    return float(v), flux, thr, bpms[:nb.value].copy()


def tempo_norms(x, n_frames=None):
    x = _f32(x)
    if n_frames is None:
        n_frames = (x.size - 512) // 256 + 1
    out = np.zeros((n_frames, 257), np.float32)
    lib().bo_tempo_norms(x, x.size, n_frames, out)
    return out


def _frame_major(S):
    """reference layout [bins, frames] -> contiguous [frames, bins] f64"""
    return np.ascontiguousarray(np.asarray(S, np.float64).T)


def pip_track(S, n_fft):
    Sf = _frame_major(S)
    cap = Sf.size
    p = np.zeros(cap, np.float64)
    m = np.zeros(cap, np.float64)
    cnt = lib().bo_pip_track(Sf, Sf.shape[0], n_fft, p, m)
    return p[:cnt].copy(), m[:cnt].copy()


def hz_to_octs(freqs, tuning=0.0, bins_per_octave=12):
    """utils.rs:119-129"""
    f = np.ascontiguousarray(freqs, np.float64).copy()
    lib().bo_hz_to_octs(f, f.size, float(tuning), int(bins_per_octave))
    return f


def pitch_tuning(freqs, resolution):
    f = np.ascontiguousarray(freqs, np.float64).copy()
    return float(lib().bo_pitch_tuning(f, f.size, resolution))


def estimate_tuning(S, n_fft):
    Sf = _frame_major(S)
    return float(lib().bo_estimate_tuning(Sf, Sf.shape[0], n_fft))


def chroma_filter(n_fft, tuning):
    out = np.zeros((12, n_fft // 2 + 1), np.float64)
    lib().bo_chroma_filter(n_fft, tuning, out)
    return out


def chroma_stft(S, n_fft, tuning):
    Sf = _frame_major(S)
    out = np.zeros((12, Sf.shape[0]), np.float64)
    lib().bo_chroma_stft(Sf, Sf.shape[0], n_fft, tuning, out)
    return out


def normalize_feature_sequence(a):
    a = np.ascontiguousarray(a, np.float64)
    out = np.zeros_like(a)
    lib().bo_normalize_feature_sequence(a, a.shape[0], a.shape[1], out)
    return out


def extract_interval_features(chroma):
    c = np.ascontiguousarray(chroma, np.float64)
    out = np.zeros((10, c.shape[1]), np.float64)
    lib().bo_extract_interval_features(c, c.shape[1], out)
    return out


def chroma_interval_features(chroma):
    c = np.ascontiguousarray(chroma, np.float64)
    out = np.zeros(10, np.float64)
    rc = lib().bo_chroma_interval_features(c, c.shape[1], out)
    if rc:
        raise ValueError("Tried to run the chroma descriptor on an empty array.")
    return out


def chroma_values(f10, version=2):
    out = np.zeros(13 if version != 1 else 10, np.float32)
    lib().bo_chroma_values(np.ascontiguousarray(f10, np.float64), version, out)
    return out


def chroma(x, version=2, want_chroma=False):
    x = _f32(x)
    out = np.zeros(13 if version != 1 else 10, np.float32)
    tuning = C.c_double(0)
    frames = lib().bo_stft_num_frames(x.size, 2205)
    cm = np.zeros((12, frames), np.float64) if want_chroma else None
    rc = lib().bo_chroma(x, x.size, version, out, C.byref(tuning),
                         cm.ctypes.data if want_chroma else None)
    assert rc == 0
    return (out, tuning.value, cm) if want_chroma else (out, tuning.value)


def euclidean_distance(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().bo_euclidean_distance(a, b, a.size))


def cosine_distance(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().bo_cosine_distance(a, b, a.size))


def mahalanobis_distance(a, b, m):
    a, b, m = _f32(a), _f32(b), _f32(m)
    return float(lib().bo_mahalanobis_distance(a, b, m, a.size))


def default_distance(a, b, version=2):
    return float(lib().bo_default_distance(_f32(a), _f32(b), version))


def feature_weights(version=2):
    d = feature_count(version)
    m = np.zeros((d, d), np.float32)
    lib().bo_feature_weights(version, m)
    return m


def closest_to_songs(seeds, cands, m=None):
    seeds, cands = _f32(np.atleast_2d(seeds)), _f32(np.atleast_2d(cands))
    order = np.zeros(cands.shape[0], np.uint32)
    keys = np.zeros(cands.shape[0], np.float32)
    mm = _f32(m) if m is not None else None
    lib().bo_closest_to_songs(seeds, seeds.shape[0], cands, cands.shape[0], cands.shape[1],
                              mm.ctypes.data if mm is not None else None, order, keys.ctypes.data)
    return order, keys


def song_to_song(seeds, cands, m=None):
    seeds, cands = _f32(np.atleast_2d(seeds)), _f32(np.atleast_2d(cands))
    order = np.zeros(cands.shape[0], np.uint32)
    mm = _f32(m) if m is not None else None
    lib().bo_song_to_song(seeds, seeds.shape[0], cands, cands.shape[0], cands.shape[1],
                          mm.ctypes.data if mm is not None else None, order)
    return order


def fft(z):
    z = np.ascontiguousarray(z, np.complex64).copy()
    v = z.view(np.float32)
    lib().bo_fft(v, z.size)
    return z


PCM_FORMATS = {np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.float32): 3}


def pcm_to_mono(frames):
    """interleaved [n_frames, channels] (or [n_frames]) s16 / s32 / f32 at 22 050 Hz -> mono f32, as the
    reference's decoders do it (src/song/decoder/ffmpeg.rs:36-109, symphonia.rs:260-300)"""
    a = np.ascontiguousarray(frames)
    if a.ndim == 1:
        a = a[:, None]
    out = np.zeros(a.shape[0], np.float32)
    rc = lib().bo_pcm_to_mono(a.ctypes.data, a.shape[0], PCM_FORMATS[a.dtype], a.shape[1], out)
    assert rc == 0
    return out
