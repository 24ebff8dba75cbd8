"""ctypes binding of libbliss_b200.so (include/bliss_b200.h).

There is no fallback of any kind: if the shared library is missing, or no CUDA
device is visible, every call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BLISS_B200_SO selects an experimental build of the same sources (scripts/build_variants.py); tests and
# the bench default to the in-tree product library
SO_PATH = os.environ.get("BLISS_B200_SO") or os.path.join(_HERE, "libbliss_b200.so")

N_KERNELS = 10
METRIC_MAHALANOBIS = 0
METRIC_COSINE = 2


class NativeError(RuntimeError):
    pass


class Taps(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "centroid", "rolloff", "flatness", "flux", "thresholded", "bpms", "n_bpms",
        "loudness_chunks", "zero_crossings", "stft8192", "n_peaks", "tuning", "chroma",
        "interval_features", "peak_pitches", "peak_mags")]


_lib = None
_inited_device = None

# every symbol include/bliss_b200.h declares
SYMBOLS = [
    "bliss_b200_init", "bliss_b200_init_devices", "bliss_b200_device_count", "bliss_b200_shutdown", "bliss_b200_set_workspace_limit", "bliss_b200_set_variant",
    "bliss_b200_strerror",
    "bliss_b200_last_error", "bliss_b200_feature_count", "bliss_b200_analyze", "bliss_b200_analyze_batch",
    "bliss_b200_analyze_batch_s16", "bliss_b200_analyze_batch_pcm", "bliss_b200_pcm_to_mono",
    "bliss_b200_resample", "bliss_b200_resampled_len",
    "bliss_b200_analyze_batch_device", "bliss_b200_feature_weights", "bliss_b200_distance",
    "bliss_b200_distance_matrix", "bliss_b200_distance_matrix_device", "bliss_b200_closest_to_songs",
    "bliss_b200_song_to_song", "bliss_b200_stft512_mag_device", "bliss_b200_analyze_taps", "bliss_b200_chroma_filter",
    "bliss_b200_set_profiling", "bliss_b200_get_profile", "bliss_b200_kernel_name",
    "bliss_b200_launch_count",
    "bliss_b200_gather_create", "bliss_b200_gather_connect", "bliss_b200_analyze_batch_device_scatter",
    "bliss_b200_gather_commit", "bliss_b200_gather_check", "bliss_b200_gather_set_timeout",
    "bliss_b200_gather_destroy",
]
GATHER_HANDLE_BYTES = 128


def load():
    """dlopen the library and declare the prototypes (no device needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NativeError(
            "%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(the CUDA extension is the only implementation; there is no CPU fallback)" % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp, u64p, i32p, u32p, f32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int32), \
        C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    L.bliss_b200_init.argtypes = [C.c_int]
    L.bliss_b200_shutdown.restype = None
    L.bliss_b200_set_workspace_limit.argtypes = [C.c_uint64]
    L.bliss_b200_strerror.argtypes = [C.c_int]
    L.bliss_b200_strerror.restype = C.c_char_p
    L.bliss_b200_last_error.restype = C.c_char_p
    L.bliss_b200_feature_count.argtypes = [C.c_uint16]
    L.bliss_b200_feature_count.restype = C.c_uint32
    L.bliss_b200_analyze.argtypes = [vp, C.c_uint64, C.c_uint16, vp]
    L.bliss_b200_analyze_batch.argtypes = [vp, u64p, C.c_uint32, C.c_uint16, vp, i32p]
    L.bliss_b200_analyze_batch_s16.argtypes = [vp, u64p, C.c_uint32, C.c_uint16, vp, i32p]
    L.bliss_b200_analyze_batch_pcm.argtypes = [vp, u64p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint16, vp, i32p]
    L.bliss_b200_pcm_to_mono.argtypes = [vp, C.c_uint64, C.c_int, C.c_uint32, vp]
    L.bliss_b200_resample.argtypes = [vp, C.c_uint64, C.c_uint32, vp, C.c_uint64, u64p]
    L.bliss_b200_resampled_len.argtypes = [C.c_uint64, C.c_uint32]
    L.bliss_b200_resampled_len.restype = C.c_uint64
    L.bliss_b200_analyze_batch_device.argtypes = [vp, u64p, u64p, C.c_uint32, C.c_uint16, vp, i32p, vp]
    L.bliss_b200_feature_weights.argtypes = [C.c_uint16, vp]
    L.bliss_b200_distance.argtypes = [vp, vp, C.c_uint32, C.c_int, vp, f32p]
    L.bliss_b200_distance_matrix.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp]
    L.bliss_b200_distance_matrix_device.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_int, vp,
                                                    vp, vp]
    L.bliss_b200_closest_to_songs.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp, vp]
    L.bliss_b200_song_to_song.argtypes = [vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, C.c_int, vp, vp]
    L.bliss_b200_stft512_mag_device.argtypes = [vp, u64p, u64p, C.c_uint32, vp, u64p, vp]
    L.bliss_b200_analyze_taps.argtypes = [vp, C.c_uint64, C.c_uint16, vp, C.POINTER(Taps)]
    L.bliss_b200_set_profiling.argtypes = [C.c_int]
    L.bliss_b200_get_profile.argtypes = [C.POINTER(C.c_double), u64p]
    L.bliss_b200_kernel_name.argtypes = [C.c_int]
    L.bliss_b200_kernel_name.restype = C.c_char_p
    L.bliss_b200_launch_count.restype = C.c_uint64
    L.bliss_b200_gather_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, vp, C.POINTER(vp)]
    L.bliss_b200_gather_connect.argtypes = [vp, vp]
    L.bliss_b200_analyze_batch_device_scatter.argtypes = [vp, vp, u64p, u64p, C.c_uint32, C.c_uint16, C.c_uint64,
                                                          C.c_uint64, vp, i32p, vp]
    L.bliss_b200_gather_commit.argtypes = [vp, vp, C.POINTER(vp)]
    L.bliss_b200_gather_check.argtypes = [vp]
    L.bliss_b200_gather_set_timeout.argtypes = [vp, C.c_uint64]
    L.bliss_b200_gather_destroy.argtypes = [vp]
    _lib = L
    return L


def check(rc):
    if rc < 0:
        L = load()
        raise NativeError("%s: %s" % (L.bliss_b200_strerror(rc).decode(), L.bliss_b200_last_error().decode()))
    return rc


def init(device=None):
    """bliss_b200_init on `device` (default: $LOCAL_RANK or 0).  Raises without a GPU."""
    global _inited_device
    L = load()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if _inited_device is None else _inited_device
    if _inited_device == device:
        return L
    check(L.bliss_b200_init(int(device)))
    _inited_device = device
    return L


def init_devices(n_devices=0):
    """bliss_b200_init_devices: contexts on devices 0 .. n-1 (0 = all visible); the host-buffer calls then shard one
    call's songs over every device from this one process.  Returns the number of devices in use."""
    global _inited_device
    L = load()
    n = L.bliss_b200_init_devices(int(n_devices))
    if n <= 0:
        check(n)
    _inited_device = 0
    return int(n)


def shutdown():
    """bliss_b200_shutdown: every context, stream and device buffer of this process is released."""
    global _inited_device
    load().bliss_b200_shutdown()
    _inited_device = None


def device_count():
    return int(load().bliss_b200_device_count())


def lib():
    return init()


def feature_count(version=2):
    return int(load().bliss_b200_feature_count(version))


def kernel_names():
    L = load()
    return [L.bliss_b200_kernel_name(i).decode() for i in range(N_KERNELS)]


def get_profile():
    L = lib()
    ms = (C.c_double * N_KERNELS)()
    ln = (C.c_uint64 * N_KERNELS)()
    check(L.bliss_b200_get_profile(ms, ln))
    return list(ms), list(ln)


def set_profiling(on):
    check(lib().bliss_b200_set_profiling(1 if on else 0))


def launch_count():
    return int(load().bliss_b200_launch_count())


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def analyze_batch(pcms, version=2):
    """list of 1-D float32 host arrays -> (status int32[n], features float32[n, dim])"""
    L = lib()
    pcms = [_f32c(p) for p in pcms]
    n = len(pcms)
    dim = feature_count(version)
    out = np.zeros((n, dim), np.float32)
    status = np.zeros(n, np.int32)
    if n == 0:
        return status, out
    ptrs = (C.c_void_p * n)(*[p.ctypes.data if p.size else None for p in pcms])
    lens = (C.c_uint64 * n)(*[p.size for p in pcms])
    check(L.bliss_b200_analyze_batch(ptrs, lens, n, version, out.ctypes.data,
                                     status.ctypes.data_as(C.POINTER(C.c_int32))))
    return status, out


def analyze_batch_ptrs(ptrs, lens, version, out, status):
    """raw form used by bench.py: ptrs/lens are ctypes arrays over pinned host buffers"""
    L = lib()
    check(L.bliss_b200_analyze_batch(ptrs, lens, len(lens), version, out.ctypes.data,
                                     status.ctypes.data_as(C.POINTER(C.c_int32))))


def set_variant(mask: int) -> int:
    """diagnostic: kernel implementation mask (include/bliss_b200.h); returns the previous mask"""
    L = lib()
    L.bliss_b200_set_variant.argtypes = [C.c_int]
    return int(L.bliss_b200_set_variant(int(mask)))


def analyze_batch_s16(pcms, version=2):
    """host int16 buffers (mono, 22 050 Hz): converted to f32 on the device (x / 32768)"""
    L = lib()
    pcms = [np.ascontiguousarray(p, dtype=np.int16).reshape(-1) for p in pcms]
    n = len(pcms)
    ptrs = (C.c_void_p * n)(*[p.ctypes.data if p.size else None for p in pcms])
    lens = (C.c_uint64 * n)(*[p.size for p in pcms])
    out = np.zeros((n, feature_count(version)), np.float32)
    status = np.zeros(n, np.int32)
    check(L.bliss_b200_analyze_batch_s16(ptrs, lens, n, version, out.ctypes.data,
                                         status.ctypes.data_as(C.POINTER(C.c_int32))))
    return status, out


def analyze_batch_s16_ptrs(ptrs, lens, version, out, status):
    """raw form used by bench.py: ptrs/lens are ctypes arrays over pinned host int16 buffers"""
    L = lib()
    check(L.bliss_b200_analyze_batch_s16(ptrs, lens, len(lens), version, out.ctypes.data,
                                         status.ctypes.data_as(C.POINTER(C.c_int32))))


def analyze_batch_pcm_ptrs(ptrs, n_frames, fmt, channels, sample_rate, version, out, status):
    """raw form used by bench.py: ptrs/n_frames are ctypes arrays over pinned host buffers of interleaved frames"""
    L = lib()
    check(L.bliss_b200_analyze_batch_pcm(ptrs, n_frames, len(n_frames), int(fmt), int(channels), int(sample_rate), version,
                                         out.ctypes.data, status.ctypes.data_as(C.POINTER(C.c_int32))))


PCM_S16, PCM_S32, PCM_F32 = 1, 2, 3
_PCM_FORMATS = {np.dtype(np.int16): PCM_S16, np.dtype(np.int32): PCM_S32, np.dtype(np.float32): PCM_F32}


def _frames(a):
    """[n_frames] or [n_frames, channels] int16 / int32 / float32 -> (contiguous 2-D array, format code)"""
    a = np.ascontiguousarray(a)
    if a.dtype not in _PCM_FORMATS:
        raise TypeError("PCM frames must be int16, int32 or float32, not %s" % a.dtype)
    if a.ndim == 1:
        a = a[:, None]
    if a.ndim != 2:
        raise ValueError("PCM frames must be [n_frames] or [n_frames, channels]")
    return a, _PCM_FORMATS[a.dtype]


def analyze_batch_pcm(frames, sample_rate=22050, version=2):
    """interleaved frames as the codec delivers them ([n_frames, channels] int16 / int32 / float32 arrays of ONE
    format, channel count and sample rate): sample-format conversion, down-mix and -- for a rate other than
    22 050 Hz -- the sample-rate conversion run on the device"""
    L = lib()
    arrs = [_frames(f) for f in frames]
    n = len(arrs)
    out = np.zeros((n, feature_count(version)), np.float32)
    status = np.zeros(n, np.int32)
    if n == 0:
        return status, out
    fmt, ch = arrs[0][1], arrs[0][0].shape[1]
    if any(f != fmt or a.shape[1] != ch for a, f in arrs):
        raise ValueError("one call takes one sample format and one channel count")
    ptrs = (C.c_void_p * n)(*[a.ctypes.data if a.size else None for a, _ in arrs])
    lens = (C.c_uint64 * n)(*[a.shape[0] for a, _ in arrs])
    check(L.bliss_b200_analyze_batch_pcm(ptrs, lens, n, fmt, ch, int(sample_rate), version, out.ctypes.data,
                                         status.ctypes.data_as(C.POINTER(C.c_int32))))
    return status, out


def pcm_to_mono(frames):
    """the conversion alone: what PreAnalyzedSong.sample_array holds for such a source"""
    L = lib()
    a, fmt = _frames(frames)
    out = np.zeros(a.shape[0], np.float32)
    check(L.bliss_b200_pcm_to_mono(a.ctypes.data if a.size else None, a.shape[0], fmt, a.shape[1], out.ctypes.data))
    return out


def resampled_len(n_samples, sample_rate):
    """length of an n_samples signal at sample_rate once at 22 050 Hz (src/song/decoder/symphonia.rs:379-380)"""
    return int(load().bliss_b200_resampled_len(int(n_samples), int(sample_rate)))


def resample(pcm, sample_rate):
    """mono f32 at sample_rate -> mono f32 at 22 050 Hz on the device (bliss_b200_resample: parity unpinned against
    the reference's swresample / rubato, checked against scipy.signal.resample_poly)"""
    L = lib()
    pcm = _f32c(pcm)
    n = C.c_uint64(0)
    out = np.zeros(resampled_len(pcm.size, sample_rate), np.float32)
    check(L.bliss_b200_resample(pcm.ctypes.data if pcm.size else None, pcm.size, int(sample_rate),
                                out.ctypes.data if out.size else None, out.size, C.byref(n)))
    return out[:n.value]


def analyze(pcm, version=2):
    L = lib()
    pcm = _f32c(pcm)
    out = np.zeros(feature_count(version), np.float32)
    rc = check(L.bliss_b200_analyze(pcm.ctypes.data if pcm.size else None, pcm.size, version, out.ctypes.data))
    return rc, out


def analyze_batch_device(d_pcm_ptr, offsets, n_samples, version, d_out_ptr, stream_ptr=None):
    """device-resident PCM; offsets / n_samples: sequences of ints. Returns status int32[n]."""
    L = lib()
    n = len(n_samples)
    off = (C.c_uint64 * n)(*[int(o) for o in offsets])
    ln = (C.c_uint64 * n)(*[int(v) for v in n_samples])
    status = (C.c_int32 * n)()
    check(L.bliss_b200_analyze_batch_device(d_pcm_ptr, off, ln, n, version, d_out_ptr, status, stream_ptr))
    return np.array(status[:], np.int32)


class Gather:
    """bliss_b200_gather: rows finished by this rank's analysis land in every rank's row buffer.

    create -> exchange `handle` (GATHER_HANDLE_BYTES bytes per rank) -> connect(all handles in rank order)
    -> per step: scatter(...) one or more times, commit() -> device pointer of the complete row buffer.
    """

    def __init__(self, world, rank, max_rows):
        L = lib()
        self.world, self.rank, self.max_rows = int(world), int(rank), int(max_rows)
        buf = (C.c_ubyte * GATHER_HANDLE_BYTES)()
        h = C.c_void_p()
        check(L.bliss_b200_gather_create(self.world, self.rank, self.max_rows, buf, C.byref(h)))
        self._h = h
        self.handle = bytes(buf)

    def connect(self, all_handles):
        blob = b"".join(all_handles)
        assert len(blob) == self.world * GATHER_HANDLE_BYTES
        check(lib().bliss_b200_gather_connect(self._h, blob))

    def scatter(self, d_pcm_ptr, offsets, n_samples, version, row_offset, row_stride, d_out_ptr=None,
                stream_ptr=None):
        n = len(n_samples)
        off = (C.c_uint64 * n)(*[int(o) for o in offsets])
        ln = (C.c_uint64 * n)(*[int(v) for v in n_samples])
        status = (C.c_int32 * n)()
        check(lib().bliss_b200_analyze_batch_device_scatter(self._h, d_pcm_ptr, off, ln, n, version,
                                                            int(row_offset), int(row_stride), d_out_ptr, status,
                                                            stream_ptr))
        return np.array(status[:], np.int32)

    def commit(self, stream_ptr=None):
        """epoch barrier on the stream; returns the device address of this rank's complete row buffer"""
        p = C.c_void_p()
        check(lib().bliss_b200_gather_commit(self._h, stream_ptr, C.byref(p)))
        return p.value

    def check(self):
        check(lib().bliss_b200_gather_check(self._h))

    def set_timeout_ms(self, ms):
        check(lib().bliss_b200_gather_set_timeout(self._h, int(ms)))

    def destroy(self):
        if self._h is not None:
            lib().bliss_b200_gather_destroy(self._h)
            self._h = None


def stft512_mag_device(d_pcm_ptr, offsets, n_samples, d_mags_ptr, stream_ptr=None):
    L = lib()
    n = len(n_samples)
    off = (C.c_uint64 * n)(*[int(o) for o in offsets])
    ln = (C.c_uint64 * n)(*[int(v) for v in n_samples])
    fo = (C.c_uint64 * (n + 1))()
    check(L.bliss_b200_stft512_mag_device(d_pcm_ptr, off, ln, n, d_mags_ptr, fo, stream_ptr))
    return np.array(fo[:], np.uint64)


def analyze_taps(pcm, version=2):
    """Analyse one song and return (status, features, dict of intermediate arrays)."""
    L = lib()
    pcm = _f32c(pcm)
    n = pcm.size
    out = np.zeros(feature_count(version), np.float32)
    if n < 8192:
        rc = check(L.bliss_b200_analyze_taps(pcm.ctypes.data if n else None, n, version, out.ctypes.data, None))
        return rc, out, {}
    n_s, n_t = (n - 512) // 128 + 1, (n - 512) // 256 + 1
    n_c = int(np.ceil(np.float32(n) / np.float32(2205)))
    a = {
        "centroid": np.zeros(n_s, np.float32), "rolloff": np.zeros(n_s, np.float32),
        "flatness": np.zeros(n_s, np.float32), "flux": np.zeros(n_t, np.float32),
        "thresholded": np.zeros(n_t, np.float32), "bpms": np.zeros(n_t // 16 + 16, np.float32),
        "n_bpms": np.zeros(1, np.uint32), "loudness_chunks": np.zeros((n + 1023) // 1024, np.float32),
        "zero_crossings": np.zeros(1, np.uint32), "stft8192": np.zeros((n_c, 4097), np.float32),
        "n_peaks": np.zeros(1, np.uint64), "tuning": np.zeros(1, np.float64),
        "chroma": np.zeros((n_c, 12), np.float64), "interval_features": np.zeros(10, np.float64),
        "peak_pitches": np.zeros(n_c * 714, np.float64), "peak_mags": np.zeros(n_c * 714, np.float64),
    }
    t = Taps(**{k: v.ctypes.data for k, v in a.items()})
    rc = check(L.bliss_b200_analyze_taps(pcm.ctypes.data, n, version, out.ctypes.data, C.byref(t)))
    a["bpms"] = a["bpms"][:int(a["n_bpms"][0])]
    a["peak_pitches"] = a["peak_pitches"][:int(a["n_peaks"][0])]
    a["peak_mags"] = a["peak_mags"][:int(a["n_peaks"][0])]
    return rc, out, a


def chroma_filter(tuning_index):
    """The device's chroma filterbank for tuning (-50 + tuning_index) / 100: f64 [12, 4097] (src/chroma.rs:197-267)."""
    out = np.zeros((12, 4097), np.float64)
    check(lib().bliss_b200_chroma_filter(int(tuning_index), out.ctypes.data))
    return out


def feature_weights(version=2):
    dim = feature_count(version)
    m = np.zeros((dim, dim), np.float32)
    check(load().bliss_b200_feature_weights(version, m.ctypes.data))
    return m


def _metric_args(metric, m, dim):
    if m is not None:
        m = _f32c(m)
        assert m.shape == (dim, dim)
    return metric, m, (m.ctypes.data if m is not None else None)


def distance(a, b, metric=METRIC_MAHALANOBIS, m=None):
    L = lib()
    a, b = _f32c(a), _f32c(b)
    assert a.shape == b.shape and a.ndim == 1
    metric, m, mp = _metric_args(metric, m, a.size)
    out = C.c_float(0)
    check(L.bliss_b200_distance(a.ctypes.data, b.ctypes.data, a.size, metric, mp, C.byref(out)))
    return np.float32(out.value)


def distance_matrix(rows, cols, metric=METRIC_MAHALANOBIS, m=None):
    rows, cols = _f32c(np.atleast_2d(rows)), _f32c(np.atleast_2d(cols))
    if rows.shape[0] == 0 or cols.shape[0] == 0:
        return np.zeros((rows.shape[0], cols.shape[0]), np.float32)
    _same_dim(rows, cols, "distance_matrix")
    L = lib()
    dim = rows.shape[1]
    metric, m, mp = _metric_args(metric, m, dim)
    out = np.zeros((rows.shape[0], cols.shape[0]), np.float32)
    check(L.bliss_b200_distance_matrix(rows.ctypes.data, rows.shape[0], cols.ctypes.data, cols.shape[0], dim,
                                       metric, mp, out.ctypes.data))
    return out


def distance_matrix_device(d_rows, n_rows, d_cols, n_cols, dim, d_out, metric=METRIC_MAHALANOBIS, m=None,
                           stream_ptr=None):
    L = lib()
    metric, m, mp = _metric_args(metric, m, dim)
    check(L.bliss_b200_distance_matrix_device(d_rows, n_rows, d_cols, n_cols, dim, metric, mp, d_out, stream_ptr))


def _same_dim(a, b, what):
    """the reference panics on vectors of different lengths (ndarray dot / sub): refuse them before the device reads
    `dim` floats per row of both"""
    if a.ndim != 2 or b.ndim != 2 or a.shape[1] != b.shape[1]:
        raise NativeError("%s: feature vectors of different lengths (%s vs %s) -- analyses of different "
                          "features_version?" % (what, a.shape, b.shape))


def closest_to_songs(seeds, cands, metric=METRIC_MAHALANOBIS, m=None):
    seeds, cands = _f32c(np.atleast_2d(seeds)), _f32c(np.atleast_2d(cands))
    if seeds.size == 0:  # no seed: every key is the empty sum 0, the stable order is the input order
        return np.arange(cands.shape[0], dtype=np.uint32), np.zeros(cands.shape[0], np.float32)
    _same_dim(seeds, cands, "closest_to_songs")
    L = lib()
    dim = cands.shape[1]
    metric, m, mp = _metric_args(metric, m, dim)
    order = np.zeros(cands.shape[0], np.uint32)
    keys = np.zeros(cands.shape[0], np.float32)
    check(L.bliss_b200_closest_to_songs(seeds.ctypes.data, seeds.shape[0], cands.ctypes.data, cands.shape[0],
                                        dim, metric, mp, order.ctypes.data, keys.ctypes.data))
    return order, keys


def song_to_song(seeds, cands, metric=METRIC_MAHALANOBIS, m=None):
    seeds, cands = _f32c(np.atleast_2d(seeds)), _f32c(np.atleast_2d(cands))
    if seeds.size == 0:
        raise NativeError("song_to_song needs at least one initial song")
    _same_dim(seeds, cands, "song_to_song")
    L = lib()
    dim = cands.shape[1]
    metric, m, mp = _metric_args(metric, m, dim)
    order = np.zeros(cands.shape[0], np.uint32)
    check(L.bliss_b200_song_to_song(seeds.ctypes.data, seeds.shape[0], cands.ctypes.data, cands.shape[0], dim,
                                    metric, mp, order.ctypes.data))
    return order
