"""Mirror of the distance / ordering functions of src/playlist.rs on top of the C ABI."""
from typing import Callable, Iterable, Iterator, List, Optional, Sequence

import numpy as np

from . import _native as native


class _Metric:
    """A distance function that also remembers how to express itself to the device
    (metric id + optional matrix), so the bulk functions below can run it on the GPU.
    Calling it on two vectors mirrors `Fn(&Array1<f32>, &Array1<f32>) -> f32`
    (src/playlist.rs:41-44)."""

    def __init__(self, metric: int, m: Optional[np.ndarray], name: str):
        self.metric, self.m, self.__name__ = metric, m, name

    def __call__(self, a, b) -> float:
        return float(native.distance(a, b, self.metric, self.m))


#: src/playlist.rs:65-71
euclidean_distance = _Metric(native.METRIC_MAHALANOBIS, None, "euclidean_distance")
#: src/playlist.rs:76-79
cosine_distance = _Metric(native.METRIC_COSINE, None, "cosine_distance")


def mahalanobis_distance(a, b, m) -> float:
    """src/playlist.rs:140-142"""
    return float(native.distance(a, b, native.METRIC_MAHALANOBIS, np.asarray(m, np.float32)))


def mahalanobis_distance_builder(m) -> _Metric:
    """src/playlist.rs:129-131"""
    return _Metric(native.METRIC_MAHALANOBIS, np.ascontiguousarray(m, np.float32), "mahalanobis_distance")


def variance_based_weight_matrix(seeds: Sequence) -> np.ndarray:
    """src/playlist.rs:173-221: diagonal Mahalanobis matrix that weights the dimensions the seeds agree on
    (inverse variance, normalised so that the weights sum to the dimension).  A few f32 operations on
    n_seeds x dim values in the reference's order; feed the result to `mahalanobis_distance_builder`."""
    from .song import ProviderError
    seeds = [np.asarray(getattr(getattr(s, "analysis", s), "internal_analysis", getattr(s, "analysis", s)), np.float32).reshape(-1)
             for s in seeds]
    if len(seeds) < 2:
        raise ProviderError("seeds must contain more than one element")
    n = seeds[0].size
    if n == 0:
        raise ProviderError("seed feature vectors must not be empty")
    if any(s.size != n for s in seeds):
        raise ProviderError("all seed feature vectors must have the same length")
    n_seeds = np.float32(len(seeds))
    mean = np.zeros(n, np.float32)
    for s in seeds:
        mean = mean + s
    mean = mean / n_seeds
    variance = np.zeros(n, np.float32)
    for s in seeds:
        diff = s - mean
        variance = variance + diff * diff
    variance = variance / n_seeds
    weights = np.float32(1.0) / (variance + np.float32(1e-6))
    total = np.float32(0.0)
    for w in weights:  # Array1::sum on a contiguous f32 array: pairwise-unrolled in ndarray, sequential here;
        total = np.float32(total + w)  # the reference's own tests hold the result to 1e-4
    weights = weights * (np.float32(n) / total)
    return np.diag(weights).astype(np.float32)


def closest_album_to_group(group: Sequence, pool: Sequence) -> List:
    """src/playlist.rs:424-485: an "album playlist": `group` first, then the albums of `pool` ordered by the
    euclidean distance of their mean analysis to the group's mean analysis, each album by (disc, track); songs
    of the group and songs without an album tag leave the pool.  Distances run on the device."""
    from .song import ProviderError
    group, pool = list(group), list(pool)
    if not group:
        raise ProviderError("Mean of empty slice")
    song = lambda s: s.as_ref() if hasattr(s, "as_ref") else s
    pool = [s for s in pool if not any(song(g) == song(s) for g in group)]
    albums = {}
    for s in pool:
        album = song(s).album
        if album is not None:
            albums.setdefault(album, []).append(np.asarray(song(s).analysis.internal_analysis, np.float32))
    first = _vectors([song(g) for g in group]).mean(axis=0, dtype=np.float32)
    names = list(albums)
    playlist = list(group)
    if names:
        means = np.stack([np.stack(albums[a]).mean(axis=0, dtype=np.float32) for a in names])
        d = native.distance_matrix(first[None, :], means, native.METRIC_MAHALANOBIS, None)[0]
        for i in np.argsort(d, kind="stable"):
            al = [s for s in pool if song(s).album == names[i]]
            opt = lambda v: (v is not None, v if v is not None else 0)  # Option ordering: None < Some(_)
            al.sort(key=lambda s: (opt(song(s).disc_number), opt(song(s).track_number)))
            playlist.extend(al)
    return playlist


def _song(s):
    """impl AsRef<Song>: a Song, or anything that wraps one (library.LibrarySong), src/playlist.rs:256-262"""
    return s.as_ref() if hasattr(s, "as_ref") else s


def _vectors(songs) -> np.ndarray:
    rows = []
    for s in songs:
        s = _song(s)
        a = getattr(s, "analysis", s)
        rows.append(np.asarray(getattr(a, "internal_analysis", a), np.float32))
    return np.stack(rows) if rows else np.zeros((0, 0), np.float32)


def _as_metric(metric_builder) -> _Metric:
    if isinstance(metric_builder, _Metric):
        return metric_builder
    raise TypeError("this backend runs distances on the GPU: pass euclidean_distance, cosine_distance or a "
                    "mahalanobis_distance_builder(m) metric (a plain callable f(a, b) -> float is accepted by the "
                    "playlist functions and evaluated on the host, as the reference evaluates it)")


def _is_host_callable(metric_builder) -> bool:
    """the reference takes any Fn(&Array1, &Array1) -> f32 (src/playlist.rs:41-44): such a callable cannot run on the
    device and is applied on the host with the reference's own (stable / first-minimum) walk"""
    return callable(metric_builder) and not isinstance(metric_builder, _Metric)


def distance_matrix(rows, cols, metric_builder=euclidean_distance) -> np.ndarray:
    """All-pairs form of FunctionDistanceMetric (one kernel launch)."""
    mt = _as_metric(metric_builder)
    return native.distance_matrix(_vectors(rows), _vectors(cols), mt.metric, mt.m)


def closest_to_songs(initial_songs: Sequence, candidate_songs: Sequence, metric_builder=euclidean_distance) -> List:
    """src/playlist.rs:256-270: candidates sorted (stably) by the summed distance to the seeds."""
    cands = list(candidate_songs)
    if not cands:
        return []
    if _is_host_callable(metric_builder):
        seeds, vc = _vectors(initial_songs), _vectors(cands)
        keys = np.array([np.float32(sum(np.float32(metric_builder(sv, c)) for sv in seeds)) for c in vc], np.float32)
        if np.isnan(keys).any():
            raise ValueError("NaN distance (the reference's n32() panics)")
        return [cands[i] for i in np.argsort(keys, kind="stable")]
    mt = _as_metric(metric_builder)
    order, _ = native.closest_to_songs(_vectors(initial_songs), _vectors(cands), mt.metric, mt.m)
    return [cands[i] for i in order]


def song_to_song(initial_songs: Sequence, candidate_songs: Sequence, metric_builder=euclidean_distance) -> List:
    """src/playlist.rs:272-326: greedy nearest-neighbour chain."""
    cands = list(candidate_songs)
    if not cands:
        return []
    if _is_host_callable(metric_builder):
        cur, vc = list(_vectors(initial_songs)), _vectors(cands)
        alive, order = list(range(len(cands))), []
        while alive:
            keys = [np.float32(sum(np.float32(metric_builder(sv, vc[j])) for sv in cur)) for j in alive]
            k = int(np.argmin(np.asarray(keys, np.float32)))  # first minimum, as ndarray-stats argmin
            order.append(alive.pop(k))
            cur = [vc[order[-1]]]
        return [cands[i] for i in order]
    mt = _as_metric(metric_builder)
    order = native.song_to_song(_vectors(initial_songs), _vectors(cands), mt.metric, mt.m)
    return [cands[i] for i in order]


def dedup_playlist(playlist: Iterable, distance_threshold: Optional[float] = None) -> Iterator:
    """src/playlist.rs:343-349"""
    return dedup_playlist_custom_distance(playlist, distance_threshold, euclidean_distance)


class _BandDistances:
    """d(i, j), j > i, for a walk that only ever looks a little ahead of i: rows are fetched from the device in blocks
    of `rows` songs against the next rows + band candidates (bounded memory: never n x n), and a look-ahead beyond the
    band (a long run of duplicates) falls back to one 1 x rows call.  A plain callable metric is evaluated on the host
    pair by pair, as the reference does."""

    def __init__(self, songs, vec, mt, rows=256, band=32):
        self.songs, self.vec, self.mt, self.rows, self.band = songs, vec, mt, rows, band
        self.block_lo, self.block = -1, None
        self.far_i, self.far_lo, self.far = -1, -1, None

    def __call__(self, i, j):
        if callable(self.mt) and not isinstance(self.mt, _Metric):
            return np.float32(self.mt(self.vec[i], self.vec[j]))
        lo = (i // self.rows) * self.rows
        if j - lo - 1 < self.rows + self.band:
            if self.block_lo != lo:
                hi = min(len(self.vec), lo + self.rows)
                self.block = native.distance_matrix(self.vec[lo:hi], self.vec[lo + 1:lo + 1 + self.rows + self.band],
                                                    self.mt.metric, self.mt.m)
                self.block_lo = lo
            return self.block[i - lo, j - lo - 1]
        if self.far_i != i or not (self.far_lo <= j < self.far_lo + self.rows):
            self.far = native.distance_matrix(self.vec[i:i + 1], self.vec[j:j + self.rows], self.mt.metric, self.mt.m)[0]
            self.far_i, self.far_lo = i, j
        return self.far[j - self.far_lo]


def dedup_playlist_custom_distance(playlist: Iterable, distance_threshold: Optional[float],
                                   metric_builder) -> Iterator:
    """src/playlist.rs:367-402: drops a song when it is closer than the threshold to the
    last kept song, or has the same non-empty title and artist.  Only the distances the walk
    looks at are computed, in bounded blocks on the device (_BandDistances)."""
    songs = list(playlist)
    if not songs:
        return iter(())
    thr = np.float32(0.05 if distance_threshold is None else distance_threshold)
    mt = metric_builder if (callable(metric_builder) and not isinstance(metric_builder, _Metric)) else _as_metric(metric_builder)
    dist = _BandDistances(songs, _vectors(songs), mt)

    def gen():
        i = 0
        n = len(songs)
        while i < n:
            j = i + 1
            while j < n:
                s1, s2 = _song(songs[i]), _song(songs[j])
                same_tags = (getattr(s1, "title", None) is not None and getattr(s2, "title", None) is not None
                             and getattr(s1, "artist", None) is not None and getattr(s2, "artist", None) is not None
                             and s1.title == s2.title and s1.artist == s2.artist)
                if dist(i, j) < thr or same_tags:
                    j += 1
                    continue
                break
            yield songs[i]
            i = j
    return gen()
