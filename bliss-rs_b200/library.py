"""On-disk formats of the reference for analysed songs, so that tools built on
bliss-rs (blissify, bliss-mpd ...) can read what the B200 path produced:

* the SQLite database of `Library` (src/library.rs): tables `song` + `feature` (:500-529), schema
  version = number of migrations (`pragma user_version`, :631-679), `store_song` / `store_failed_song`
  semantics (:1544-1670), the read side (`songs_from_library` :1356-1372, `song_from_path` :1414-1463,
  `get_failed_songs` :1671-1690) and the skip-what-is-analysed logic of `update_library`
  (:1000-1093): the database IS the checkpoint / resume mechanism of the reference;
* the serde-JSON `Vec<Song>` cache of examples/playlist.rs:41-46,77-78.

Host-side only (Python `sqlite3` / `json`); the analysis itself goes through `Decoder.analyze_paths`,
i.e. batched `bliss_b200_analyze_batch` calls.  Upgrading databases written by OLD bliss-rs versions
(the SQL migrations :530-591) is not implemented: such a file is refused, never rewritten.
"""
import json
import os
import sqlite3
from dataclasses import dataclass
from typing import Any, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .song import (Analysis, AnalysisOptions, BlissError, Decoder, FeaturesVersion, ProviderError, Song)

#: `Library::SQLITE_MIGRATIONS.len()` (src/library.rs:530-591): what `pragma user_version` holds for a
#: database created by bliss-audio 0.13
SCHEMA_VERSION = 5

# Same tables, columns, types, defaults and constraints as `Library::SQLITE_SCHEMA` (src/library.rs:500-529)
_SONG_COLUMNS = (
    ("id", "integer primary key"),
    ("path", "text not null unique"),
    ("duration", "float"),
    ("album_artist", "text"),
    ("artist", "text"),
    ("title", "text"),
    ("album", "text"),
    ("track_number", "integer"),
    ("disc_number", "integer"),
    ("genre", "text"),
    ("cue_path", "text"),
    ("audio_file_path", "text"),
    ("stamp", "timestamp default current_timestamp"),
    ("version", "integer not null"),
    ("analyzed", "boolean default false"),
    ("extra_info", "json"),
    ("error", "text"),
)
_FEATURE_SQL = (
    "create table feature (id integer primary key, song_id integer not null, feature real not null, "
    "feature_index integer not null, unique(song_id, feature_index), "
    "foreign key(song_id) references song(id) on delete cascade)"
)


@dataclass
class CueInfo:
    """src/cue.rs: CueInfo { cue_path, audio_file_path }"""
    cue_path: str
    audio_file_path: str


@dataclass
class LibrarySong:
    """src/library.rs: LibrarySong<T> { bliss_song, extra_info }"""
    bliss_song: Song
    extra_info: Any = None

    def as_ref(self) -> Song:
        """impl AsRef<Song> for LibrarySong<T>"""
        return self.bliss_song


@dataclass
class ProcessingError:
    """src/library.rs: ProcessingError { song_path, error, features_version }"""
    song_path: str
    error: str
    features_version: FeaturesVersion


class Library:
    """The database half of `Library<Config, D>`; `decoder` is a `Decoder` subclass (src/song/decoder.rs)."""

    def __init__(self, database_path: str, analysis_options: Optional[AnalysisOptions] = None,
                 decoder: Optional[type] = None):
        self.database_path = database_path
        self.analysis_options = analysis_options or AnalysisOptions()
        self.decoder = decoder
        parent = os.path.dirname(os.path.abspath(database_path))
        os.makedirs(parent, exist_ok=True)
        self.conn = sqlite3.connect(database_path)
        self.conn.execute("pragma foreign_keys = on")
        self._upgrade()

    # Library::upgrade, src/library.rs:631-679
    def _upgrade(self):
        version = self.conn.execute("pragma user_version").fetchone()[0]
        if version == SCHEMA_VERSION:
            return
        if version > SCHEMA_VERSION:
            raise ProviderError("bliss-rs version %d is older than the schema version %d" % (version, SCHEMA_VERSION))
        n_tables = self.conn.execute("select count(*) from sqlite_master where type = 'table'").fetchone()[0]
        if version == 0 and n_tables == 0:
            cols = ", ".join("%s %s" % c for c in _SONG_COLUMNS)
            with self.conn:
                self.conn.execute("create table song (%s)" % cols)
                self.conn.execute(_FEATURE_SQL)
                self.conn.execute("pragma user_version = %d" % SCHEMA_VERSION)
            return
        raise ProviderError("database schema version %d predates this writer (no migrations here); "
                            "open it once with bliss-rs to upgrade it" % version)

    # ---- write side -----------------------------------------------------------------------------
    def store_song(self, library_song):
        """Library::store_song (src/library.rs:1544-1633): upsert on `path`, features replaced."""
        if isinstance(library_song, Song):
            library_song = LibrarySong(library_song, None)
        song = library_song.bliss_song
        cue = getattr(song, "cue_info", None)
        with self.conn:
            self.conn.execute(
                "insert into song (path, artist, title, album, album_artist, duration, track_number, disc_number, "
                "genre, analyzed, version, extra_info, cue_path, audio_file_path) "
                "values (?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?, ?) "
                "on conflict(path) do update set artist=excluded.artist, title=excluded.title, album=excluded.album, "
                "track_number=excluded.track_number, disc_number=excluded.disc_number, "
                "album_artist=excluded.album_artist, duration=excluded.duration, genre=excluded.genre, "
                "analyzed=excluded.analyzed, version=excluded.version, extra_info=excluded.extra_info, "
                "cue_path=excluded.cue_path, audio_file_path=excluded.audio_file_path",
                (song.path, song.artist, song.title, song.album, song.album_artist, float(song.duration),
                 song.track_number, song.disc_number, song.genre, True, int(song.features_version),
                 json.dumps(library_song.extra_info), cue.cue_path if cue else None,
                 cue.audio_file_path if cue else None))
            self.conn.execute("delete from feature where song_id in (select id from song where path = ?)", (song.path,))
            self.conn.executemany(
                "insert into feature (song_id, feature, feature_index) values ((select id from song where path = ?), ?, ?) "
                "on conflict(song_id, feature_index) do update set feature=excluded.feature",
                [(song.path, float(v), i) for i, v in enumerate(song.analysis.internal_analysis)])

    def store_failed_song(self, song_path: str, error: BlissError, features_version: FeaturesVersion):
        """src/library.rs:1639-1670: `insert or replace` with the error text, analyzed stays false."""
        with self.conn:
            self.conn.execute("insert or replace into song (path, error, version) values (?, ?, ?)",
                              (song_path, str(error), int(features_version)))

    def delete_path(self, song_path: str) -> int:
        """src/library.rs delete_path: features go with the song (on delete cascade)."""
        with self.conn:
            cur = self.conn.execute("delete from song where path = ?", (song_path,))
        if cur.rowcount == 0:
            raise ProviderError("tried to delete song %s, not existing in the database." % song_path)
        return cur.rowcount

    # ---- read side ------------------------------------------------------------------------------
    _SONG_SELECT = ("select path, artist, title, album, album_artist, track_number, disc_number, genre, duration, "
                    "version, extra_info, cue_path, audio_file_path, id from song ")

    @staticmethod
    def _song_from_row(row) -> LibrarySong:
        (path, artist, title, album, album_artist, track_number, disc_number, genre, duration, version, extra_info,
         cue_path, audio_file_path) = row[:13]
        version = FeaturesVersion.try_from(version)
        song = Song(path=path, artist=artist, title=title, album=album, album_artist=album_artist,
                    track_number=track_number, disc_number=disc_number, genre=genre, duration=float(duration or 0.0),
                    analysis=None, features_version=version)
        song.cue_info = CueInfo(cue_path, audio_file_path) if cue_path is not None else None
        return LibrarySong(song, json.loads(extra_info) if extra_info is not None else None)

    def songs_from_library(self) -> List[LibrarySong]:
        """src/library.rs:1356-1372: analysed songs of the configured features version, by id."""
        ver = int(self.analysis_options.features_version)
        rows = self.conn.execute(self._SONG_SELECT + "where analyzed = true and version = ? order by id", (ver,)).fetchall()
        feats = {}
        for value, sid in self.conn.execute(
                "select feature, song.id from feature join song on song.id = feature.song_id "
                "where song.analyzed = true and song.version = ? order by song_id, feature_index", (ver,)):
            feats.setdefault(sid, []).append(value)
        out = []
        for row in rows:
            ls = self._song_from_row(row)
            ls.bliss_song.analysis = self._analysis(feats.get(row[13], []), ls.bliss_song.features_version)
            out.append(ls)
        return out

    def song_from_path(self, song_path: str) -> LibrarySong:
        """src/library.rs:1414-1463"""
        row = self.conn.execute(self._SONG_SELECT + "where path = ? and analyzed = true", (song_path,)).fetchone()
        if row is None:
            raise ProviderError("Query returned no rows")
        ls = self._song_from_row(row)
        values = [r[0] for r in self.conn.execute(
            "select feature from feature join song on song.id = feature.song_id where song.path = ? "
            "order by feature_index", (song_path,))]
        ls.bliss_song.analysis = self._analysis(values, ls.bliss_song.features_version)
        return ls

    def songs_from_album(self, album_title: str) -> List[LibrarySong]:
        """src/library.rs:1379-1412: the analysed songs of one album, by (disc, track)"""
        ver = int(self.analysis_options.features_version)
        rows = self.conn.execute(self._SONG_SELECT + "where album = ? and analyzed = true and version = ? "
                                 "order by disc_number, track_number", (album_title, ver)).fetchall()
        if not rows:
            raise ProviderError("target album was not found in the database.")
        out = []
        for row in rows:
            ls = self._song_from_row(row)
            values = [r[0] for r in self.conn.execute("select feature from feature where song_id = ? order by feature_index", (row[13],))]
            ls.bliss_song.analysis = self._analysis(values, ls.bliss_song.features_version)
            out.append(ls)
        return out

    def delete_paths(self, paths: Iterable[str]) -> int:
        """src/library.rs:1725-1748: how many rows went"""
        paths = [str(p) for p in paths]
        if not paths:
            return 0
        with self.conn:
            cur = self.conn.execute("delete from song where path in (%s)" % ",".join("?" * len(paths)), paths)
        return cur.rowcount

    # ---- playlists: the distance kernels' consumers (src/library.rs:762-893) ---------------------
    def playlist_from(self, song_paths: Sequence[str]) -> List[LibrarySong]:
        """:762-767: euclidean distance, closest_to_songs, de-duplicated"""
        from . import playlist
        return self.playlist_from_custom(song_paths, playlist.euclidean_distance, playlist.closest_to_songs, True)

    def playlist_from_custom(self, initial_song_paths: Sequence[str], distance, sort_by, deduplicate: bool) -> List[LibrarySong]:
        """:805-848: the initial songs, then the rest of the library as `sort_by(initial, rest, distance)` orders it
        (closest_to_songs / song_to_song: one device call each), optionally through dedup_playlist_custom_distance."""
        from . import playlist
        initial = []
        for p in initial_song_paths:
            try:
                initial.append(self.song_from_path(p))
            except ProviderError:
                raise ProviderError("song '%s' has not been analyzed" % p)
        rest = [s for s in self.songs_from_library() if s.bliss_song.path not in set(initial_song_paths)]
        ordered = initial + list(sort_by(initial, rest, distance))
        if deduplicate:
            ordered = list(playlist.dedup_playlist_custom_distance(ordered, None, distance))
        return ordered

    def album_playlist_from(self, album_title: str, number_albums: int) -> List[LibrarySong]:
        """:850-876: the album, then the `number_albums` closest albums (closest_album_to_group)"""
        from . import playlist
        album = self.songs_from_album(album_title)
        ordered = playlist.closest_album_to_group(album, self.songs_from_library())
        album_count, index, current = 0, 0, album_title
        for s in ordered:
            if s.bliss_song.album != current:
                album_count += 1
                if album_count > number_albums:
                    break
                current = s.bliss_song.album
            index += 1
        return ordered[:index]

    @staticmethod
    def _analysis(values: Sequence[float], version: FeaturesVersion) -> Analysis:
        try:
            return Analysis(np.asarray(values, dtype=np.float32), version)
        except ProviderError:
            raise ProviderError("song has more or less than %d features" % version.feature_count())

    def get_failed_songs(self) -> List[ProcessingError]:
        """src/library.rs:1671-1690"""
        return [ProcessingError(p, e, FeaturesVersion.try_from(v)) for p, e, v in self.conn.execute(
            "select path, error, version from song where error is not null order by id")]

    # ---- update / resume ------------------------------------------------------------------------
    def update_library(self, paths: Iterable[str], delete_everything_else: bool = False,
                       analysis_options: Optional[AnalysisOptions] = None) -> Tuple[int, int]:
        """`update_library_convert_extra_info` (src/library.rs:1000-1093) without the extra-info plumbing:
        paths already analysed with this features version are skipped (resume), songs of another version are
        dropped once anything has to be analysed, the rest goes through the decoder's batched GPU analysis and
        every result -- success or failure -- is stored.  Returns (analysed, failed)."""
        opts = analysis_options or self.analysis_options
        ver = int(opts.features_version)
        paths = list(paths)
        existing = {r[0] for r in self.conn.execute(
            "select path from song where analyzed = true and version = ? order by id", (ver,))}
        if delete_everything_else:
            every = {r[0] for r in self.conn.execute("select path from song where analyzed = true order by id")}
            for p in every - set(paths):
                self.delete_path(p)
        todo = [p for p in paths if p not in existing]
        if todo:
            with self.conn:
                self.conn.execute("delete from song where version != ?", (ver,))
        return self.analyze_paths(todo, opts)

    def analyze_paths(self, paths: Iterable[str], analysis_options: Optional[AnalysisOptions] = None) -> Tuple[int, int]:
        """src/library.rs:1187-1290: analyse and store; a failing song is recorded, it never aborts the run."""
        if self.decoder is None:
            raise ProviderError("this Library was opened without a Decoder")
        opts = analysis_options or self.analysis_options
        ok = failed = 0
        for path, result in self.decoder.analyze_paths_with_options(paths, opts):
            if isinstance(result, BlissError):
                self.store_failed_song(path, result, opts.features_version)
                failed += 1
            else:
                self.store_song(LibrarySong(result, None))
                ok += 1
        return ok, failed

    def close(self):
        self.conn.close()


# ---- serde-JSON cache of Vec<Song> (examples/playlist.rs:41-46, 77-78) -----------------------------------
def song_to_serde(song: Song) -> dict:
    """Field names and shapes of `#[derive(Serialize)] struct Song` (src/song/mod.rs:41-76): PathBuf -> string,
    Duration -> {secs, nanos}, FeaturesVersion -> u16 (src/lib.rs:142-143), Option -> null."""
    secs = int(song.duration)
    nanos = int(round((float(song.duration) - secs) * 1e9))
    if nanos >= 1_000_000_000:
        secs, nanos = secs + 1, nanos - 1_000_000_000
    cue = getattr(song, "cue_info", None)
    return {
        "path": song.path, "artist": song.artist, "title": song.title, "album": song.album,
        "album_artist": song.album_artist, "track_number": song.track_number, "disc_number": song.disc_number,
        "genre": song.genre,
        "analysis": {"internal_analysis": [float(v) for v in song.analysis.internal_analysis],
                     "features_version": int(song.analysis.features_version)},
        "duration": {"secs": secs, "nanos": nanos},
        "features_version": int(song.features_version),
        "cue_info": None if cue is None else {"cue_path": cue.cue_path, "audio_file_path": cue.audio_file_path},
    }


def song_from_serde(d: dict) -> Song:
    ver = FeaturesVersion.try_from(d["features_version"])
    a = d["analysis"]
    song = Song(path=d["path"], artist=d.get("artist"), title=d.get("title"), album=d.get("album"),
                album_artist=d.get("album_artist"), track_number=d.get("track_number"),
                disc_number=d.get("disc_number"), genre=d.get("genre"),
                duration=d["duration"]["secs"] + d["duration"]["nanos"] * 1e-9,
                analysis=Analysis(np.asarray(a["internal_analysis"], np.float32), FeaturesVersion.try_from(a["features_version"])),
                features_version=ver)
    c = d.get("cue_info")
    song.cue_info = CueInfo(c["cue_path"], c["audio_file_path"]) if c else None
    return song


def songs_to_json(songs: Sequence[Song]) -> str:
    """f32 features are written with the shortest decimal that round-trips as f32, like serde_json does for
    `Vec<f32>` (0.3846389, not the f64 expansion 0.38463890552520752)."""
    docs = []
    for s in songs:
        d = song_to_serde(s)
        marker = "@@F32:%d@@" % len(docs)
        vals = ",".join(_f32_repr(v) for v in s.analysis.internal_analysis)
        d["analysis"]["internal_analysis"] = marker
        docs.append(json.dumps(d, separators=(",", ":")).replace('"%s"' % marker, "[" + vals + "]"))
    return "[" + ",".join(docs) + "]"


def _f32_repr(v) -> str:
    v = np.float32(v)
    if not np.isfinite(v):
        return "null"  # serde_json writes non-finite floats as null
    t = np.format_float_positional(v, unique=True, trim="0")  # shortest f32 round-trip
    if len(t) > 24 or abs(float(v)) < 1e-5 and v != 0:
        t = np.format_float_scientific(v, unique=True, trim="0", exp_digits=1).replace("e+", "e")
    return t


def songs_from_json(text: str) -> List[Song]:
    return [song_from_serde(d) for d in json.loads(text)]
