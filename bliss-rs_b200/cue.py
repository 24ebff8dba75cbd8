"""Mirror of src/cue.rs: the songs of a CUE sheet, cut out of ONE decoded buffer per audio file.

The reference decodes each FILE of the sheet once and runs `Song::analyze_with_options` on the sub-slice of every
track (src/cue.rs:208-243).  Here the slices of all tracks of all files go to the device in one batched call
(`analyze_decoded`): the host path copies a buffer whose slices follow each other once, and the kernels address the
tracks by offset (`bliss_b200_analyze_batch_device` takes such offsets directly).

Sheet parsing is the `rcue` crate's job in the reference (rcue 0.1.3, Cargo.toml:94, not vendored): restated here for
the commands src/cue.rs reads -- REM comments, PERFORMER, TITLE, FILE, TRACK, INDEX mm:ss:ff with 75 frames per
second -- in non-strict mode (unknown lines are skipped).  Track boundaries are computed as the reference computes
them, `(index.as_secs_f32() * SAMPLE_RATE as f32) as usize` in f32 (:212-213, :231); the durations the reference's own
test asserts for data/testcue.cue (:311, :356, :402) pin that arithmetic (tests/test_host_abi.py).
"""
import os
import re
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .song import (SAMPLE_RATE, AnalysisOptions, BlissError, DecodingError, PreAnalyzedSong, Song, FeaturesVersion,
                   analyze_decoded)


@dataclass
class Track:
    no: str = ""
    title: Optional[str] = None
    performer: Optional[str] = None
    indices: List[Tuple[str, Tuple[int, int]]] = field(default_factory=list)  # (index number, (seconds, nanoseconds))


@dataclass
class CueFile:
    file: str = ""
    tracks: List[Track] = field(default_factory=list)


@dataclass
class Cue:
    performer: Optional[str] = None
    title: Optional[str] = None
    comments: List[Tuple[str, str]] = field(default_factory=list)
    files: List[CueFile] = field(default_factory=list)


def _unquote(s: str) -> str:
    s = s.strip()
    return s[1:-1] if len(s) >= 2 and s[0] == '"' and s[-1] == '"' else s


def _timestamp(s: str) -> Tuple[int, int]:
    m = re.fullmatch(r"(\d+):(\d+):(\d+)", s.strip())
    if not m:
        raise ValueError("bad timestamp %r" % s)
    minutes, seconds, frames = (int(g) for g in m.groups())
    return minutes * 60 + seconds, frames * 1_000_000_000 // 75


def parse_cue(text: str) -> Cue:
    cue, cur_file, cur_track = Cue(), None, None
    for line in text.splitlines():
        line = line.strip().lstrip("﻿")
        if not line:
            continue
        cmd, _, rest = line.partition(" ")
        cmd = cmd.upper()
        try:
            if cmd == "REM":
                key, _, value = rest.strip().partition(" ")
                cue.comments.append((key, _unquote(value)))
            elif cmd in ("PERFORMER", "TITLE"):
                target = cur_track if cur_track is not None else (cue if cur_file is None else None)
                if target is not None:
                    setattr(target, cmd.lower(), _unquote(rest))
            elif cmd == "FILE":
                m = re.match(r'\s*("[^"]*"|\S+)', rest)
                cur_file, cur_track = CueFile(file=_unquote(m.group(1))), None
                cue.files.append(cur_file)
            elif cmd == "TRACK" and cur_file is not None:
                cur_track = Track(no=rest.split()[0])
                cur_file.tracks.append(cur_track)
            elif cmd == "INDEX" and cur_track is not None:
                no, ts = rest.split()[:2]
                cur_track.indices.append((no, _timestamp(ts)))
        except (ValueError, IndexError, AttributeError):
            continue  # non-strict parsing, as the reference asks of rcue (parse_from_file(.., false), src/cue.rs:110)
    return cue


def sample_index(ts: Tuple[int, int]) -> int:
    """(Duration::as_secs_f32() * SAMPLE_RATE as f32) as usize, src/cue.rs:212-213"""
    secs = np.float32(ts[0]) + np.float32(ts[1]) / np.float32(1_000_000_000)
    return int(np.float32(secs * np.float32(SAMPLE_RATE)))


class BlissCue:
    """BlissCue<D>, src/cue.rs:52-163: `BlissCue(MyDecoder).songs_from_path("album.cue")`"""

    def __init__(self, decoder):
        self.decoder = decoder

    def songs_from_path(self, path: str) -> List[object]:
        return self.songs_from_path_with_options(path, AnalysisOptions())

    def songs_from_path_with_options(self, path: str, analysis_options: AnalysisOptions) -> List[object]:
        """One entry per track, a Song (with cue_info) or the BlissError the reference would have pushed
        (:85-106); a sheet that cannot be read raises DecodingError (:110-116)."""
        from .library import CueInfo
        path = str(path)
        try:
            with open(path, "r", encoding="utf-8", errors="replace") as f:
                cue = parse_cue(f.read())
        except OSError as e:
            raise DecodingError("when opening CUE file '%s': %s" % (path, e))
        genre = next((v for c, v in cue.comments if c.upper() == "GENRE"), None)
        disc = next((v for c, v in cue.comments if c.upper() in ("DISCNUMBER", "DISC")), None)
        try:
            disc_number = int(disc) if disc is not None else None
        except ValueError:
            disc_number = None
        entries: List[object] = []   # BlissError | (slice to analyse, Song without an analysis)
        for cue_file in cue.files:
            parent = os.path.dirname(path)
            audio = os.path.join(parent, cue_file.file) if parent else cue_file.file
            try:
                decoded = self.decoder.decode(audio)
            except BlissError as e:
                entries.append(e)
                continue
            if decoded.pcm_frames is not None and decoded.pcm_rate != SAMPLE_RATE:
                # the sheet's indices count 22 050 Hz samples (:212-213): such a file is cut after its conversion
                decoded = PreAnalyzedSong(path=decoded.path, duration=decoded.duration, sample_array=decoded.mono())
            packed = decoded.pcm_frames is not None
            samples = decoded.pcm_frames if packed else np.asarray(decoded.sample_array, np.float32)
            total = len(samples)
            if total == 0:
                entries.append(DecodingError("empty audio file associated to CUE sheet"))
                continue
            firsts = [(t, sample_index(t.indices[0][1]) if t.indices else None) for t in cue_file.tracks]
            bounds = [(i + 1, t, s, firsts[i + 1][1]) for i, (t, s) in enumerate(firsts[:-1])
                      if s is not None and firsts[i + 1][1] is not None]
            if firsts and firsts[-1][1] is not None:  # the last track runs to the end of the file (:229-241)
                bounds.append((len(firsts), firsts[-1][0], firsts[-1][1], total))
            for index, track, start, end in bounds:
                if not 0 <= start <= end <= total:  # the reference's slice would panic here
                    entries.append(DecodingError("CUE track %s of '%s' lies outside its audio file" % (track.no, path)))
                    continue
                duration = float(np.float32(end - start) / np.float32(SAMPLE_RATE))
                try:
                    track_number = int(track.no)
                except ValueError:
                    track_number = None
                song = Song(path="%s/CUE_TRACK%03d" % (path, index), album=cue.title, artist=track.performer,
                            album_artist=cue.performer, title=track.title, track_number=track_number,
                            disc_number=disc_number, genre=genre, duration=duration,
                            features_version=FeaturesVersion(analysis_options.features_version),
                            cue_info=CueInfo(cue_path=path, audio_file_path=audio))
                piece = samples[start:end]
                entries.append((PreAnalyzedSong(pcm_frames=piece) if packed else PreAnalyzedSong(sample_array=piece), song))
        todo = [e for e in entries if not isinstance(e, BlissError)]
        analyses = analyze_decoded([p for p, _ in todo], analysis_options) if todo else []
        done = iter(analyses)
        out: List[object] = []
        for e in entries:
            if isinstance(e, BlissError):
                out.append(e)
                continue
            analysis = next(done)
            if isinstance(analysis, BlissError):
                out.append(analysis)
            else:
                e[1].analysis = analysis
                out.append(e[1])
        return out
