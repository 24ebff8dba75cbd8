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


def _vectors(songs) -> np.ndarray:
    rows = []
    for s in songs:
        a = getattr(s, "analysis", s)
        rows.append(np.asarray(getattr(a, "internal_analysis", a), np.float32))
    return np.stack(rows) if rows else np.zeros((0, 0), np.float32)


def _as_metric(metric_builder) -> _Metric:
    if isinstance(metric_builder, _Metric):
        return metric_builder
    raise TypeError("this backend runs distances on the GPU: pass euclidean_distance, cosine_distance or a "
                    "mahalanobis_distance_builder(m) metric")


def distance_matrix(rows, cols, metric_builder=euclidean_distance) -> np.ndarray:
    """All-pairs form of FunctionDistanceMetric (one kernel launch)."""
    mt = _as_metric(metric_builder)
    return native.distance_matrix(_vectors(rows), _vectors(cols), mt.metric, mt.m)


def closest_to_songs(initial_songs: Sequence, candidate_songs: Sequence, metric_builder=euclidean_distance) -> List:
    """src/playlist.rs:256-270: candidates sorted (stably) by the summed distance to the seeds."""
    cands = list(candidate_songs)
    if not cands:
        return []
    mt = _as_metric(metric_builder)
    order, _ = native.closest_to_songs(_vectors(initial_songs), _vectors(cands), mt.metric, mt.m)
    return [cands[i] for i in order]


def song_to_song(initial_songs: Sequence, candidate_songs: Sequence, metric_builder=euclidean_distance) -> List:
    """src/playlist.rs:272-326: greedy nearest-neighbour chain."""
    cands = list(candidate_songs)
    if not cands:
        return []
    mt = _as_metric(metric_builder)
    order = native.song_to_song(_vectors(initial_songs), _vectors(cands), mt.metric, mt.m)
    return [cands[i] for i in order]


def dedup_playlist(playlist: Iterable, distance_threshold: Optional[float] = None) -> Iterator:
    """src/playlist.rs:343-349"""
    return dedup_playlist_custom_distance(playlist, distance_threshold, euclidean_distance)


def dedup_playlist_custom_distance(playlist: Iterable, distance_threshold: Optional[float],
                                   metric_builder) -> Iterator:
    """src/playlist.rs:367-402: drops a song when it is closer than the threshold to the
    last kept song, or has the same non-empty title and artist.  Distances to the
    successor are computed for the whole playlist in one device call."""
    songs = list(playlist)
    if not songs:
        return iter(())
    thr = np.float32(0.05 if distance_threshold is None else distance_threshold)
    mt = _as_metric(metric_builder)
    vec = _vectors(songs)
    dm = native.distance_matrix(vec, vec, mt.metric, mt.m) if len(songs) > 1 else None

    def gen():
        i = 0
        n = len(songs)
        while i < n:
            j = i + 1
            while j < n:
                s1, s2 = songs[i], songs[j]
                same_tags = (getattr(s1, "title", None) is not None and getattr(s2, "title", None) is not None
                             and getattr(s1, "artist", None) is not None and getattr(s2, "artist", None) is not None
                             and s1.title == s2.title and s1.artist == s2.artist)
                if dm[i, j] < thr or same_tags:
                    j += 1
                    continue
                break
            yield songs[i]
            i = j
    return gen()
