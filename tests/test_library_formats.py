"""On-disk formats: the SQLite layout of the reference's Library and the serde-JSON
Vec<Song> cache.  CPU only: the analysis values are given, what is checked is the format and the store / read /
resume semantics, with the reference's own test statements (src/library.rs tests) run against our files."""
import json
import sqlite3

import numpy as np
import pytest

import bliss_rs_b200 as B
from bliss_rs_b200 import library as L


def _song(path, seed, version=B.FeaturesVersion.Version2, **kw):
    rng = np.random.default_rng(seed)
    a = B.Analysis(rng.uniform(-1, 1, version.feature_count()).astype(np.float32), version)
    return B.Song(path=path, analysis=a, features_version=version, duration=250.5, **kw)


def test_schema_matches_reference(tmp_path):
    lib = L.Library(str(tmp_path / "sub" / "songs.db"))
    c = sqlite3.connect(lib.database_path)
    assert c.execute("pragma user_version").fetchone()[0] == 5  # SQLITE_MIGRATIONS.len(), src/library.rs:530-591
    cols = [(r[1], r[2].lower(), r[3], r[4]) for r in c.execute("pragma table_info(song)")]
    # src/library.rs:500-518
    assert cols == [("id", "integer", 0, None), ("path", "text", 1, None), ("duration", "float", 0, None),
                    ("album_artist", "text", 0, None), ("artist", "text", 0, None), ("title", "text", 0, None),
                    ("album", "text", 0, None), ("track_number", "integer", 0, None),
                    ("disc_number", "integer", 0, None), ("genre", "text", 0, None), ("cue_path", "text", 0, None),
                    ("audio_file_path", "text", 0, None), ("stamp", "timestamp", 0, "current_timestamp"),
                    ("version", "integer", 1, None), ("analyzed", "boolean", 0, "false"),
                    ("extra_info", "json", 0, None), ("error", "text", 0, None)]
    fcols = [(r[1], r[2].lower(), r[3]) for r in c.execute("pragma table_info(feature)")]
    assert fcols == [("id", "integer", 0), ("song_id", "integer", 1), ("feature", "real", 1), ("feature_index", "integer", 1)]
    # the insert statements of the reference's test_library_new_create_database (src/library.rs:3903-3933) work
    c.execute("insert into song (id, path, artist, title, album, album_artist, track_number, disc_number, genre, stamp, "
              "version, duration, analyzed, extra_info) values (1, '/random/path', 'Some Artist', 'A Title', 'Some Album', "
              "'Some Album Artist', 1, 1, 'Electronica', '2022-01-01', 1, 250, true, '{\"key\": \"value\"}')")
    c.execute("insert into feature(id, song_id, feature, feature_index) values (2000, 1, 1.1, 1) "
              "on conflict(song_id, feature_index) do update set feature=excluded.feature")
    with pytest.raises(sqlite3.IntegrityError):  # unique(song_id, feature_index)
        c.execute("insert into feature(song_id, feature, feature_index) values (1, 2.2, 1)")


def test_store_and_read_back(tmp_path):
    lib = L.Library(str(tmp_path / "songs.db"))
    s1 = _song("/path/to/song1001", 1, artist="Artist1001", title="Title1001", album="An Album1001",
               album_artist="An Album Artist1001", track_number=3, disc_number=1, genre="Electronica1001")
    s2 = _song("/path/to/cuetrack.cue/CUE_TRACK001", 2)
    s2.cue_info = L.CueInfo("/path/to/cuetrack.cue", "/path/to/cuetrack.flac")
    lib.store_song(L.LibrarySong(s1, {"ignore": True, "metadata_bliss_does_not_have": "x"}))
    lib.store_song(s2)
    # the read statement of the reference (src/library.rs:1440-1447), straight on the file
    c = sqlite3.connect(lib.database_path)
    got = [r[0] for r in c.execute("select feature from feature join song on song.id = feature.song_id "
                                   "where song.path = ? order by feature_index", (s1.path,))]
    assert np.array_equal(np.asarray(got, np.float32), s1.analysis.internal_analysis)  # f32 -> REAL -> f32 is exact
    row = c.execute("select analyzed, version, duration, extra_info, cue_path, error from song where path = ?", (s1.path,)).fetchone()
    assert row[0] == 1 and row[1] == 2 and row[2] == 250.5 and row[4] is None and row[5] is None
    assert json.loads(row[3]) == {"ignore": True, "metadata_bliss_does_not_have": "x"}
    back = lib.songs_from_library()
    assert [b.bliss_song.path for b in back] == [s1.path, s2.path]
    assert back[0].bliss_song.analysis == s1.analysis and back[0].bliss_song.genre == "Electronica1001"
    assert back[1].bliss_song.cue_info == L.CueInfo("/path/to/cuetrack.cue", "/path/to/cuetrack.flac")
    assert back[1].extra_info is None
    assert lib.song_from_path(s1.path).bliss_song.analysis == s1.analysis
    # store again with other values: same row id, features replaced (upsert, src/library.rs:1576-1591)
    sid = c.execute("select id from song where path = ?", (s1.path,)).fetchone()[0]
    s1b = _song(s1.path, 99, artist="Other")
    lib.store_song(s1b)
    c2 = sqlite3.connect(lib.database_path)
    assert c2.execute("select id, artist from song where path = ?", (s1.path,)).fetchone() == (sid, "Other")
    assert c2.execute("select count(*) from feature where song_id = ?", (sid,)).fetchone()[0] == 23
    assert lib.song_from_path(s1.path).bliss_song.analysis == s1b.analysis
    lib.delete_path(s2.path)
    assert c2.execute("select count(*) from feature").fetchone()[0] == 23  # on delete cascade
    with pytest.raises(B.ProviderError):
        lib.delete_path("/not/there")


def test_store_failed_song_like_the_reference(tmp_path):
    """src/library.rs:3717-3758 test_store_failed_song"""
    lib = L.Library(str(tmp_path / "songs.db"))
    lib.store_failed_song("/some/failed/path", B.ProviderError("error with the analysis"), B.FeaturesVersion.Version1)
    c = sqlite3.connect(lib.database_path)
    error, analyzed, version = c.execute("select error, analyzed, version from song where path=?", ("/some/failed/path",)).fetchone()
    assert error == "error happened with the music library provider - error with the analysis"
    assert analyzed == 0 and version == 1
    assert c.execute("select count(*) from feature join song on song.id = feature.song_id where path=?",
                     ("/some/failed/path",)).fetchone()[0] == 0
    failed = lib.get_failed_songs()
    assert failed == [L.ProcessingError("/some/failed/path", error, B.FeaturesVersion.Version1)]
    assert lib.songs_from_library() == []


class _FakeDecoder(B.Decoder):
    """Stands in for decode + GPU analysis (CPU test): counts what it was asked to analyse."""
    seen = []

    @classmethod
    def analyze_paths_with_options(cls, paths, analysis_options):
        for p in paths:
            cls.seen.append(p)
            if "broken" in p:
                yield p, B.DecodingError("while opening format for file '%s'" % p)
            else:
                yield p, _song(p, abs(hash(p)) % 1000, B.FeaturesVersion(analysis_options.features_version))


def test_update_library_resumes_and_drops_other_versions(tmp_path):
    """src/library.rs:1000-1093: analysed paths of this version are skipped, songs of another version go once
    something new is analysed, failures are stored as rows, delete_everything_else prunes."""
    db = str(tmp_path / "songs.db")
    _FakeDecoder.seen = []
    lib = L.Library(db, decoder=_FakeDecoder)
    assert lib.update_library(["/a", "/b", "/broken"]) == (2, 1)
    assert _FakeDecoder.seen == ["/a", "/b", "/broken"]
    lib.close()
    lib = L.Library(db, decoder=_FakeDecoder)  # a second run: the database is the checkpoint
    _FakeDecoder.seen = []
    assert lib.update_library(["/a", "/b", "/c", "/broken"]) == (1, 1)
    assert _FakeDecoder.seen == ["/c", "/broken"]  # failed songs are retried, analysed ones are not
    assert [f.song_path for f in lib.get_failed_songs()] == ["/broken"]
    v1 = B.AnalysisOptions(features_version=B.FeaturesVersion.Version1)
    _FakeDecoder.seen = []
    assert lib.update_library(["/a"], analysis_options=v1) == (1, 0)
    c = sqlite3.connect(db)
    assert c.execute("select path, version from song order by id").fetchall() == [("/a", 1)]  # version-2 rows dropped
    assert c.execute("select count(*) from feature").fetchone()[0] == 20
    lib.analysis_options = v1
    _FakeDecoder.seen = []
    assert lib.update_library(["/d"], delete_everything_else=True) == (1, 0)
    assert [s.bliss_song.path for s in lib.songs_from_library()] == ["/d"]


def test_refuses_newer_and_older_schemas(tmp_path):
    p = str(tmp_path / "new.db")
    c = sqlite3.connect(p)
    c.execute("pragma user_version = 9")
    c.commit()
    with pytest.raises(B.ProviderError):
        L.Library(p)
    p = str(tmp_path / "old.db")
    c = sqlite3.connect(p)
    c.execute("create table song (id integer primary key, path text)")
    c.execute("pragma user_version = 2")
    c.commit()
    with pytest.raises(B.ProviderError):
        L.Library(p)


def test_serde_json_cache_round_trip():
    """examples/playlist.rs:41-46,77-78 caches Vec<Song> with serde_json"""
    s1 = _song("/music/a.flac", 5, artist="A", track_number=2)
    s2 = _song("/music/b.cue/CUE_TRACK002", 6, version=B.FeaturesVersion.Version1)
    s2.cue_info = L.CueInfo("/music/b.cue", "/music/b.wav")
    s1.analysis.internal_analysis[0] = np.float32(0.3846389)  # src/song/mod.rs:556 golden tempo
    text = L.songs_to_json([s1, s2])
    doc = json.loads(text)
    assert list(doc[0].keys()) == ["path", "artist", "title", "album", "album_artist", "track_number", "disc_number",
                                   "genre", "analysis", "duration", "features_version", "cue_info"]  # struct order
    assert doc[0]["duration"] == {"secs": 250, "nanos": 500000000} and doc[0]["features_version"] == 2
    assert doc[0]["analysis"]["features_version"] == 2 and doc[1]["analysis"]["features_version"] == 1
    assert doc[1]["cue_info"] == {"cue_path": "/music/b.cue", "audio_file_path": "/music/b.wav"}
    assert "0.3846389" in text and "0.38463890" not in text  # shortest f32 representation, like serde_json
    back = L.songs_from_json(text)
    assert back[0].analysis == s1.analysis and back[1].analysis == s2.analysis
    assert back[0].artist == "A" and back[0].track_number == 2 and back[1].cue_info == s2.cue_info
    assert abs(back[0].duration - 250.5) < 1e-9
