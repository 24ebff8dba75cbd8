"""Host-side playlist glue of src/playlist.rs that needs no device arithmetic of its own: the reference's
tests of variance_based_weight_matrix (:1664-1765) and closest_album_to_group (:1113-1262).  The one distance
call of the latter is replaced by its definition (CPU test; on a GPU box it runs through the C ABI)."""
import numpy as np
import pytest

import bliss_rs_b200 as B
from bliss_rs_b200 import playlist as P


def test_variance_based_weight_matrix_errors():
    with pytest.raises(B.ProviderError, match="seeds must contain more than one element"):
        P.variance_based_weight_matrix([np.array([1.0, 2.0, 3.0], np.float32)])
    with pytest.raises(B.ProviderError, match="all seed feature vectors must have the same length"):
        P.variance_based_weight_matrix([np.array([1.0, 2.0, 3.0], np.float32), np.array([1.0, 2.0], np.float32)])
    with pytest.raises(B.ProviderError, match="seed feature vectors must not be empty"):
        P.variance_based_weight_matrix([np.zeros(0, np.float32), np.zeros(0, np.float32)])


def test_variance_based_weight_matrix_values():
    seeds = [np.array(v, np.float32) for v in ([1.0, 0.0, 1.0], [1.0, 100.0, 1.0], [1.0, 200.0, 1.0])]
    m = P.variance_based_weight_matrix(seeds)
    assert m.shape == (3, 3) and m.dtype == np.float32
    assert m[0, 0] > m[1, 1] and m[2, 2] > m[1, 1]
    assert np.count_nonzero(m - np.diag(np.diag(m))) == 0
    assert abs(float(np.trace(m)) - 3.0) < 1e-4                      # weights sum to the dimension
    same = P.variance_based_weight_matrix([np.array([1.0, 2.0, 3.0], np.float32)] * 3)
    assert np.allclose(np.diag(same), 1.0, atol=1e-4)                # identical seeds: every weight 1
    two = P.variance_based_weight_matrix([np.array([0.0, 50.0], np.float32), np.array([0.0, 150.0], np.float32)])
    assert two.shape == (2, 2) and two[0, 0] > two[1, 1]
    # doc example (:166-171) and use as a metric matrix
    ex = P.variance_based_weight_matrix([np.array([0.3, 0.8, 0.5], np.float32), np.array([0.3, 0.2, 0.5], np.float32)])
    assert ex[0, 0] > ex[1, 1]
    assert P.mahalanobis_distance_builder(ex).m.shape == (3, 3)


@pytest.mark.parametrize("version", [B.FeaturesVersion.Version1, B.FeaturesVersion.Version2])
def test_closest_album_to_group_like_the_reference(monkeypatch, version):
    def fake_distance_matrix(rows, cols, metric, m):  # euclidean_distance, src/playlist.rs:65-71
        return np.sqrt(((rows[:, None, :] - cols[None, :, :]) ** 2).sum(-1)).astype(np.float32)
    monkeypatch.setattr(P.native, "distance_matrix", fake_distance_matrix)
    n = version.feature_count()
    mk = lambda path, val, **kw: B.Song(path=path, analysis=B.Analysis([val] * n, version), features_version=version, **kw)
    first = mk("path-to-first", 0.0, album="Album", artist="Artist", track_number=1, disc_number=1)
    second = mk("path-to-third", 10.0, album="Album", artist="Another Artist", track_number=2, disc_number=1)
    other_d1_t1 = mk("path-to-second-2", 0.15, album="Another Album", artist="Artist", track_number=1, disc_number=1)
    other_d1_t2 = mk("path-to-second", 0.1, album="Another Album", artist="Artist", track_number=2, disc_number=1)
    other_d2_t1 = mk("path-to-fourth", 20.0, album="Another Album", artist="Another Artist", track_number=1, disc_number=2)
    other_d2_t4 = mk("path-to-fourth", 20.0, album="Another Album", artist="Another Artist", track_number=4, disc_number=2)
    no_album = mk("path-to-fifth", 40.0, artist="Third Artist")
    pool = [first, other_d1_t2, other_d2_t4, second, other_d2_t1, other_d1_t1, no_album]
    got = P.closest_album_to_group([first, second], pool)
    assert got == [first, second, other_d1_t1, other_d1_t2, other_d2_t1, other_d2_t4]
    # two albums: the closer one comes first, whatever the pool order
    far = [mk("far-%d" % i, 30.0 + i, album="Far", track_number=i) for i in (2, 1)]
    near = [mk("near-%d" % i, 6.0 + i, album="Near", track_number=i) for i in (2, 1)]
    got = P.closest_album_to_group([first, second], far + near + [first])
    assert [s.path for s in got] == ["path-to-first", "path-to-third", "near-1", "near-2", "far-1", "far-2"]
    with pytest.raises(B.ProviderError):
        P.closest_album_to_group([], pool)


def _songs(vals, version=B.FeaturesVersion.Version2, **kw):
    n = version.feature_count()
    return [B.Song(path="p%d" % i, analysis=B.Analysis([v] * n, version), features_version=version, **kw) for i, v in enumerate(vals)]


def test_plain_callable_metrics_run_on_the_host():
    """The reference takes any Fn(&Array1, &Array1) -> f32 (src/playlist.rs:41-44); such a callable cannot run on the
    device: the playlist functions evaluate it on the host with the reference's own walks (stable sort of the summed
    keys, first minimum of the chain).  No device call is made (this test runs without a GPU)."""
    manhattan = lambda a, b: float(np.abs(a - b).sum())
    songs = _songs([0.5, -0.2, 0.9, 0.5, 0.1])
    seed = _songs([0.45])
    got = P.closest_to_songs(seed, songs, manhattan)
    assert [s.path for s in got] == ["p0", "p3", "p4", "p2", "p1"]       # the tie p0 / p3 keeps the input order
    chain = P.song_to_song(seed, songs, manhattan)
    assert [s.path for s in chain] == ["p0", "p3", "p2", "p4", "p1"]     # 0.5, 0.5, then the nearest to the last one
    with pytest.raises(ValueError):
        P.closest_to_songs(seed, songs, lambda a, b: float("nan"))
    kept = list(P.dedup_playlist_custom_distance(_songs([0.0, 0.001, 0.002, 0.5, 0.501, 0.9]), 0.05, manhattan))
    assert [s.path for s in kept] == ["p0", "p3", "p5"]


def test_dedup_walk_never_builds_the_full_matrix(monkeypatch):
    """dedup used to build an n x n matrix.  The walk now fetches bounded blocks (rows x (rows + band))
    and falls back to 1 x rows calls for a run of duplicates longer than the band; the result is the reference's
    (src/playlist.rs:367-402) and no call is larger than the block."""
    calls = []

    def fake_distance_matrix(rows, cols, metric, m):
        calls.append((rows.shape[0], cols.shape[0]))
        return np.sqrt(((rows[:, None, :] - cols[None, :, :]) ** 2).sum(-1)).astype(np.float32)
    monkeypatch.setattr(P.native, "distance_matrix", fake_distance_matrix)
    vals = [0.0] * 50 + [1.0] + [2.0 + 0.1 * i for i in range(700)] + [100.0] * 3
    songs = _songs(vals)
    kept = [s.path for s in P.dedup_playlist_custom_distance(songs, 0.05, P.euclidean_distance)]
    want, i = [], 0
    while i < len(vals):           # the reference's peek loop on the same distances
        j = i + 1
        while j < len(vals) and abs(vals[i] - vals[j]) * np.sqrt(23) < 0.05:
            j += 1
        want.append("p%d" % i)
        i = j
    assert kept == want
    assert max(r * c for r, c in calls) <= 256 * (256 + 32) and len(calls) < 20
    # same title + artist: dropped whatever the distance
    tagged = _songs([0.0, 5.0, 9.0], title="t", artist="a")
    assert [s.path for s in P.dedup_playlist_custom_distance(tagged, 0.05, P.euclidean_distance)] == ["p0"]


def test_vectors_of_different_versions_are_refused():
    from bliss_rs_b200 import _native as N
    v1, v2 = np.zeros((2, 20), np.float32), np.zeros((3, 23), np.float32)
    for fn in (N.closest_to_songs, N.song_to_song, N.distance_matrix):
        with pytest.raises(N.NativeError, match="different lengths"):
            fn(v1, v2)
