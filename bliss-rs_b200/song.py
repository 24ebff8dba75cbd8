"""Mirror of src/song/mod.rs, src/song/decoder.rs and the types of src/lib.rs."""
import enum
import os
import queue
import threading
from dataclasses import dataclass, field
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as native

SAMPLE_RATE = 22050  # src/lib.rs:143
CHANNELS = 1         # src/lib.rs:140
NUMBER_FEATURES = 23  # AnalysisIndex::COUNT, src/song/mod.rs:158-167


class BlissError(Exception):
    """src/lib.rs:236-249"""


class DecodingError(BlissError):
    def __str__(self):
        return "error happened while decoding file - %s" % (self.args[0] if self.args else "")


class AnalysisError(BlissError):
    def __str__(self):
        return "error happened while analyzing file - %s" % (self.args[0] if self.args else "")


class ProviderError(BlissError):
    def __str__(self):
        return "error happened with the music library provider - %s" % (self.args[0] if self.args else "")


class FeaturesVersion(enum.IntEnum):
    """src/lib.rs:151-207"""
    Version1 = 1
    Version2 = 2

    @classmethod
    def latest(cls):
        return cls.Version2

    def feature_count(self) -> int:
        return 23 if self == FeaturesVersion.Version2 else 20

    def feature_weights(self) -> np.ndarray:
        """src/lib.rs:168-173 (VERSION2_WEIGHTS :209-234)"""
        return native.feature_weights(int(self))

    def distance_metric(self):
        """src/lib.rs:176-178: mahalanobis_distance_builder(self.feature_weights())"""
        from .playlist import mahalanobis_distance_builder
        return mahalanobis_distance_builder(self.feature_weights())

    @classmethod
    def try_from(cls, value: int):
        """src/lib.rs:195-207"""
        try:
            return cls(value)
        except ValueError:
            raise ProviderError("This features' version (%d) does not exist" % value)


FeaturesVersion.LATEST = FeaturesVersion.Version2


class AnalysisIndex(enum.IntEnum):
    """src/song/mod.rs:103-156 (FeaturesVersion::LATEST)"""
    Tempo = 0
    Zcr = 1
    MeanSpectralCentroid = 2
    StdDeviationSpectralCentroid = 3
    MeanSpectralRolloff = 4
    StdDeviationSpectralRolloff = 5
    MeanSpectralFlatness = 6
    StdDeviationSpectralFlatness = 7
    MeanLoudness = 8
    StdDeviationLoudness = 9
    Chroma1 = 10
    Chroma2 = 11
    Chroma3 = 12
    Chroma4 = 13
    Chroma5 = 14
    Chroma6 = 15
    Chroma7 = 16
    Chroma8 = 17
    Chroma9 = 18
    Chroma10 = 19
    Chroma11 = 20
    Chroma12 = 21
    Chroma13 = 22


class AnalysisIndexv1(enum.IntEnum):
    """src/song/mod.rs:169-236 (FeaturesVersion::Version1, 20 features)"""
    Tempo = 0
    Zcr = 1
    MeanSpectralCentroid = 2
    StdDeviationSpectralCentroid = 3
    MeanSpectralRolloff = 4
    StdDeviationSpectralRolloff = 5
    MeanSpectralFlatness = 6
    StdDeviationSpectralFlatness = 7
    MeanLoudness = 8
    StdDeviationLoudness = 9
    Chroma1 = 10
    Chroma2 = 11
    Chroma3 = 12
    Chroma4 = 13
    Chroma5 = 14
    Chroma6 = 15
    Chroma7 = 16
    Chroma8 = 17
    Chroma9 = 18
    Chroma10 = 19


@dataclass(frozen=True)
class AnalysisOptions:
    """src/song/mod.rs:252-269.  number_cores is kept for API parity; on this backend the
    songs of one call are analysed together on the GPU."""
    features_version: FeaturesVersion = FeaturesVersion.Version2
    number_cores: int = field(default_factory=lambda: os.cpu_count() or 1)


class Analysis:
    """src/song/mod.rs:240-371"""

    def __init__(self, analysis: Sequence[float], features_version: FeaturesVersion = FeaturesVersion.Version2):
        analysis = np.asarray(analysis, dtype=np.float32).reshape(-1)
        features_version = FeaturesVersion(features_version)
        if analysis.size != features_version.feature_count():  # Analysis::new, :326-339
            raise ProviderError("Feature count %d does not match the expected version feature count %d"
                                % (analysis.size, features_version.feature_count()))
        self.internal_analysis = analysis.copy()
        self.features_version = features_version

    @classmethod
    def new(cls, analysis, features_version):
        return cls(analysis, features_version)

    def as_arr1(self) -> np.ndarray:
        return self.internal_analysis.copy()

    def as_vec(self) -> List[float]:
        return [float(x) for x in self.internal_analysis]

    def __getitem__(self, index):
        """impl Index<AnalysisIndex> / Index<AnalysisIndexv1>: panics (here: raises) on a version mismatch,
        src/song/mod.rs:272-292"""
        want = FeaturesVersion.Version1 if isinstance(index, AnalysisIndexv1) else FeaturesVersion.Version2
        if isinstance(index, (AnalysisIndex, AnalysisIndexv1)) and self.features_version != want:
            raise RuntimeError("Tried to index features with incompatible indexes")
        return float(self.internal_analysis[int(index)])

    def distance(self, other: "Analysis") -> float:
        """src/song/mod.rs:364-370"""
        if self.features_version != other.features_version:
            raise RuntimeError("Mismatched features version between two songs or analysis")
        return float(native.distance(self.internal_analysis, other.internal_analysis,
                                     native.METRIC_MAHALANOBIS, self.features_version.feature_weights()))

    def __eq__(self, other):
        return (isinstance(other, Analysis) and self.features_version == other.features_version
                and np.array_equal(self.internal_analysis, other.internal_analysis))

    def __repr__(self):
        """impl Debug, src/song/mod.rs:294-318: the named fields, then the vector as a comment; floats as Rust's
        `{:?}` prints an f32 (shortest digits that round-trip, at least one decimal, exponent form outside
        [1e-4, 1e16))"""
        names = AnalysisIndex if self.features_version == FeaturesVersion.Version2 else AnalysisIndexv1
        vals = [_f32_debug(v) for v in self.internal_analysis]
        body = ", ".join("%s: %s" % (n.name, v) for n, v in zip(names, vals))
        return "Analysis (Version %d) { %s } /* [%s] */" % (int(self.features_version), body, ", ".join(vals))


def _f32_debug(v) -> str:
    v = np.float32(v)
    if np.isnan(v):
        return "NaN"
    if np.isinf(v):
        return "inf" if v > 0 else "-inf"
    a = np.float32(abs(v))
    if a == 0 or np.float32(1e-4) <= a < np.float32(1e16):
        return np.format_float_positional(v, unique=True, trim="0")
    return np.format_float_scientific(v, unique=True, trim="-", exp_digits=1).replace("e+", "e")


@dataclass
class Song:
    """src/song/mod.rs:45-98 (fields the analysis path touches)"""
    path: str = ""
    artist: Optional[str] = None
    album_artist: Optional[str] = None
    title: Optional[str] = None
    album: Optional[str] = None
    track_number: Optional[int] = None
    disc_number: Optional[int] = None
    genre: Optional[str] = None
    duration: float = 0.0
    analysis: Optional[Analysis] = None
    features_version: FeaturesVersion = FeaturesVersion.Version2
    cue_info: Optional[object] = None  # library.CueInfo when the song was cut out of a CUE sheet (src/song/mod.rs:71-76)

    @staticmethod
    def analyze(sample_array) -> Analysis:
        """Song::analyze, src/song/mod.rs:403-405"""
        return Song.analyze_with_options(sample_array, AnalysisOptions())

    @staticmethod
    def analyze_with_options(sample_array, analysis_options: AnalysisOptions) -> Analysis:
        """Song::analyze_with_options, src/song/mod.rs:413-508"""
        ver = FeaturesVersion(analysis_options.features_version)
        rc, feats = native.analyze(np.asarray(sample_array, dtype=np.float32), int(ver))
        _raise_for_status(rc)
        return Analysis(feats, ver)

    def distance(self, other: "Song") -> float:
        """src/song/mod.rs:519-521"""
        return self.analysis.distance(other.analysis)

    def as_ref(self):
        return self


def _raise_for_status(rc: int):
    if rc == 1:
        raise AnalysisError("empty or too short song.")  # src/song/mod.rs:426-430
    if rc != 0:
        raise AnalysisError("internal error in the B200 analysis backend (status %d)" % rc)


def analyze_batch(sample_arrays: Sequence, analysis_options: AnalysisOptions = None) -> List:
    """The batching seam: many decoded buffers, one GPU call.  Returns one entry per
    input, either an Analysis or the BlissError the reference would have produced
    (errors are items, never exceptions: src/song/decoder.rs:319-325)."""
    analysis_options = analysis_options or AnalysisOptions()
    ver = FeaturesVersion(analysis_options.features_version)
    status, feats = native.analyze_batch(sample_arrays, int(ver))
    out = []
    for st, row in zip(status, feats):
        if st == 0:
            out.append(Analysis(row, ver))
        elif st == 1:
            out.append(AnalysisError("empty or too short song."))
        else:
            out.append(AnalysisError("internal error in the B200 analysis backend (status %d)" % st))
    return out


def analyze_batch_pcm(frames: Sequence, sample_rate: int = SAMPLE_RATE, analysis_options: AnalysisOptions = None) -> List:
    """analyze_batch for sources the codec delivers as interleaved [n_frames, channels] int16 / int32 / float32
    frames: the decoders' sample-format conversion and down-mix (src/song/decoder/ffmpeg.rs:36-109,
    symphonia.rs:260-300) and, for a sample_rate other than 22 050 Hz, the conversion to 22 050 Hz
    (bliss_b200_resample: parity unpinned against swresample / rubato) run on the device.  One format, channel count
    and rate per call."""
    analysis_options = analysis_options or AnalysisOptions()
    ver = FeaturesVersion(analysis_options.features_version)
    status, feats = native.analyze_batch_pcm(frames, sample_rate, int(ver))
    return [_result_item(st, row, ver) for st, row in zip(status, feats)]


def analyze_decoded(songs: Sequence["PreAnalyzedSong"], analysis_options: AnalysisOptions = None) -> List:
    """One batch of decoded songs -> one entry per song (Analysis | BlissError).  Songs that carry sample_array go
    through analyze_batch; songs that carry packed frames are grouped by (sample format, channel count, sample rate)
    -- one bliss_b200_analyze_batch_pcm call takes one of each -- and converted on the device."""
    out: List = [None] * len(songs)
    groups = {}
    for i, p in enumerate(songs):
        f = p.pcm_frames
        key = None if f is None else (np.asarray(f).dtype.str, 1 if np.asarray(f).ndim == 1 else np.asarray(f).shape[1],
                                      int(p.pcm_rate))
        groups.setdefault(key, []).append(i)
    for key, idx in groups.items():
        if key is None:
            res = analyze_batch([songs[i].sample_array for i in idx], analysis_options)
        else:
            res = analyze_batch_pcm([songs[i].pcm_frames for i in idx], key[2], analysis_options)
        for i, r in zip(idx, res):
            out[i] = r
    return out


def pcm_to_mono(frames) -> np.ndarray:
    """PreAnalyzedSong.sample_array of such a source (src/song/decoder.rs:64)"""
    return native.pcm_to_mono(frames)


def _result_item(st, row, ver):
    if st == 0:
        return Analysis(row, ver)
    if st == 1:
        return AnalysisError("empty or too short song.")
    return AnalysisError("internal error in the B200 analysis backend (status %d)" % st)


@dataclass
class PreAnalyzedSong:
    """src/song/decoder.rs:34-67"""
    path: str = ""
    artist: Optional[str] = None
    album_artist: Optional[str] = None
    title: Optional[str] = None
    album: Optional[str] = None
    track_number: Optional[int] = None
    disc_number: Optional[int] = None
    genre: Optional[str] = None
    duration: float = 0.0
    sample_array: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))
    #: Not in the reference.  A decoder may leave the codec's packed frames here ([n_frames, channels] int16 / int32 /
    #: float32, at pcm_rate Hz) instead of filling sample_array: sample-format conversion, down-mix and the conversion
    #: to 22 050 Hz then run on the device behind the copy (bliss_b200_analyze_batch_pcm),
    #: and e.g. 16-bit mono material sends half the bytes over PCIe.
    pcm_frames: Optional[np.ndarray] = None
    pcm_rate: int = SAMPLE_RATE

    def mono(self) -> np.ndarray:
        """sample_array as this backend's decode steps fill it (src/song/decoder.rs:64)"""
        if self.pcm_frames is None:
            return self.sample_array
        m = pcm_to_mono(self.pcm_frames)
        return m if self.pcm_rate == SAMPLE_RATE else native.resample(m, self.pcm_rate)

    def to_song_with_options(self, analysis_options: AnalysisOptions) -> Song:
        """src/song/decoder.rs:85-100"""
        res = analyze_decoded([self], analysis_options)[0]
        if isinstance(res, BlissError):
            raise res
        return self._song(res, analysis_options)

    def _song(self, analysis, analysis_options):
        return Song(path=self.path, artist=self.artist, album_artist=self.album_artist, title=self.title,
                    album=self.album, track_number=self.track_number, disc_number=self.disc_number,
                    genre=self.genre, duration=self.duration, analysis=analysis,
                    features_version=FeaturesVersion(analysis_options.features_version))


class Decoder:
    """trait Decoder, src/song/decoder.rs:115-333.  Subclasses implement `decode`
    (file -> mono f32 22 050 Hz PreAnalyzedSong); the provided methods keep the
    reference's signatures, but analyze_paths batches the decoded buffers into GPU calls
    instead of running Song::analyze on `number_cores` threads."""

    #: how many decoded songs are handed to one bliss_b200_analyze_batch call
    BATCH_SONGS = 64

    @classmethod
    def decode(cls, path: str) -> PreAnalyzedSong:
        raise NotImplementedError

    @classmethod
    def song_from_path(cls, path: str) -> Song:
        return cls.song_from_path_with_options(path, AnalysisOptions())

    @classmethod
    def song_from_path_with_options(cls, path: str, analysis_options: AnalysisOptions) -> Song:
        return cls.decode(path).to_song_with_options(analysis_options)

    @classmethod
    def analyze_paths(cls, paths: Iterable[str]) -> Iterator[Tuple[str, object]]:
        return cls.analyze_paths_with_options(paths, AnalysisOptions())

    @classmethod
    def analyze_paths_with_options(cls, paths: Iterable[str],
                                   analysis_options: AnalysisOptions) -> Iterator[Tuple[str, object]]:
        """Yields (path, Song | BlissError) like the reference's mpsc iterator (src/song/decoder.rs:278-332; order
        of arrival is unspecified there as here).  Same thread layout up to the analysis: min(available cores,
        number_cores) workers, each on a contiguous chunk of the paths (:283-304), run `decode`; decoding errors
        are items (:319-325).  Instead of analysing its own song, a worker hands the decoded buffer to ONE
        batcher thread through a bounded queue (2 x BATCH_SONGS songs: decoders stall rather than pile PCM up), and
        the batcher sends <= BATCH_SONGS buffers per bliss_b200_analyze_batch call while the workers keep decoding
        (the call releases the GIL).  INTEGRATION.md section 3 is the same body in Rust."""
        paths = list(paths)
        results: "queue.Queue" = queue.Queue()  # the channel the reference returns
        if not paths:
            return iter(())
        cores = max(1, min(os.cpu_count() or 1, int(analysis_options.number_cores)))
        chunk_length = max(len(paths) // cores, 1)
        chunks = [paths[i:i + chunk_length] for i in range(0, len(paths), chunk_length)]
        decoded: "queue.Queue" = queue.Queue(maxsize=2 * cls.BATCH_SONGS)
        stop = threading.Event()  # the consumer went away, or a thread failed
        worker_done, end = object(), object()

        def hand_over(item):
            while not stop.is_set():
                try:
                    decoded.put(item, timeout=0.1)
                    return
                except queue.Full:
                    pass

        def worker(chunk):
            try:
                for path in chunk:
                    if stop.is_set():
                        break
                    if os.path.splitext(str(path))[1].lower() == ".cue":  # :305-318: every track of the sheet is an item
                        from .cue import BlissCue
                        try:
                            for song in BlissCue(cls).songs_from_path_with_options(path, analysis_options):
                                results.put((path, song))
                        except BlissError as e:
                            results.put((path, e))
                        continue
                    try:
                        hand_over((path, cls.decode(path)))
                    except BlissError as e:  # decoding errors are items too
                        results.put((path, e))
            except BaseException as e:  # a bug in decode(): surfaces in the consumer, not in a dead thread
                results.put((end, e))  # before the batcher can see `stop` and send its own end marker
                stop.set()
            finally:
                hand_over(worker_done)

        def batcher():
            batch: List[Tuple[str, PreAnalyzedSong]] = []

            def flush():
                analyses = analyze_decoded([p for _, p in batch], analysis_options)
                for (path, pre), res in zip(batch, analyses):
                    results.put((path, res if isinstance(res, BlissError) else pre._song(res, analysis_options)))
                batch.clear()

            try:
                left = len(chunks)
                while left and not stop.is_set():
                    try:
                        item = decoded.get(timeout=0.1)
                    except queue.Empty:
                        continue
                    if item is worker_done:
                        left -= 1
                        continue
                    batch.append(item)
                    if len(batch) >= cls.BATCH_SONGS:
                        flush()
                if batch and not stop.is_set():
                    flush()
                results.put((end, None))
            except BaseException as e:  # e.g. a CUDA failure of the whole call
                stop.set()
                results.put((end, e))

        threads = [threading.Thread(target=worker, args=(c,), daemon=True) for c in chunks]
        threads.append(threading.Thread(target=batcher, daemon=True))
        for t in threads:
            t.start()

        def drain():
            try:
                while True:
                    path, item = results.get()
                    if path is end:
                        if item is not None:
                            raise item
                        return
                    yield (path, item)
            finally:
                stop.set()

        return drain()
