"""A concrete `Decoder` for the one family of sources this backend can take without the crate's decoders: RIFF/WAVE
PCM files.

The reference's decoders (src/song/decoder/ffmpeg.rs, symphonia.rs) do four things to such a file: unpack the
codec's frames, convert the sample format to f32, down-mix to mono and resample to 22 050 Hz (src/lib.rs:143).
`WavDecoder.decode` does the first on the host (a RIFF chunk walk) and leaves the packed frames in
`PreAnalyzedSong.pcm_frames` (their rate in `pcm_rate`); the others run on the device behind the copy
(`bliss_b200_analyze_batch_pcm`: x * 2^-15 / x * 2^-31, c L + c R with c = (float)sqrt(1/2), mean in channel order for
more channels, then the polyphase resampler of bliss_b200_resample), so `Decoder.analyze_paths` sends the file's own
bytes over PCIe.  Files at 22 050 Hz are bit-identical to the reference's decoders' output; at other rates the
resampler is this backend's own (parity unpinned: DESIGN.md section 7).

Sample widths, as ffmpeg's pcm decoders deliver them (libavcodec/pcm.c behind ffmpeg.rs:190-360):
  8 bit unsigned  -> (x - 128) * 2^-7   (carried as s16: (x - 128) << 8)
  16 bit signed   -> x * 2^-15
  24 bit signed   -> x * 2^-23          (carried as s32: x << 8, what pcm_s24le decodes to)
  32 bit signed   -> x * 2^-31
  32 bit IEEE float as it is
"""
import struct

import numpy as np

from .song import Decoder, DecodingError, PreAnalyzedSong

MAX_CHANNELS = 8  # BLISS_B200_PCM_MAX_CHANNELS, include/bliss_b200.h
MIN_SAMPLE_RATE, MAX_SAMPLE_RATE = 1000, 768000  # BLISS_B200_MIN_SAMPLE_RATE / _MAX_SAMPLE_RATE


def _riff_wave(raw: bytes):
    """(format tag, channels, rate, bits per sample, data bytes) of a RIFF/WAVE file; ValueError otherwise"""
    if len(raw) < 12 or raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    fmt, o = None, 12
    while o + 8 <= len(raw):
        name, size = raw[o:o + 4], struct.unpack_from("<I", raw, o + 4)[0]
        body = o + 8
        if name == b"fmt " and size >= 16 and body + size <= len(raw):
            tag, channels, rate, _, _, bits = struct.unpack_from("<HHIIHH", raw, body)
            if tag == 0xFFFE and size >= 26:  # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
                tag = struct.unpack_from("<H", raw, body + 24)[0]
            fmt = (tag, channels, rate, bits)
        elif name == b"data":
            if fmt is None:
                break
            return fmt + (raw[body:body + size],)  # a truncated file: what is there
        o = body + size + (size & 1)
    raise ValueError("no fmt / data chunk")


class WavDecoder(Decoder):
    @classmethod
    def decode(cls, path: str) -> PreAnalyzedSong:
        try:
            with open(str(path), "rb") as f:
                tag, channels, rate, bits, raw = _riff_wave(f.read())
        except (OSError, ValueError, struct.error) as e:
            raise DecodingError("while opening format for file '%s': %s." % (path, e))
        if not MIN_SAMPLE_RATE <= rate <= MAX_SAMPLE_RATE:
            raise DecodingError("file '%s' runs at %d Hz (%d..%d are taken)." % (path, rate, MIN_SAMPLE_RATE, MAX_SAMPLE_RATE))
        if not 1 <= channels <= MAX_CHANNELS:
            raise DecodingError("file '%s' has %d channels (1..%d are taken)." % (path, channels, MAX_CHANNELS))
        if not ((tag == 1 and bits in (8, 16, 24, 32)) or (tag == 3 and bits == 32)):
            raise DecodingError("file '%s': encoding %d with %d bits per sample." % (path, tag, bits))
        width = bits // 8
        n = len(raw) // (width * channels)  # whole frames only
        raw = raw[:n * width * channels]
        if tag == 3:
            frames = np.frombuffer(raw, "<f4").astype(np.float32, copy=False)
        elif width == 1:
            frames = (np.frombuffer(raw, np.uint8).astype(np.int16) - 128) << 8
        elif width == 2:
            frames = np.frombuffer(raw, "<i2").astype(np.int16, copy=False)
        elif width == 3:
            b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.uint32)
            frames = ((b[:, 0] << 8) | (b[:, 1] << 16) | (b[:, 2] << 24)).view(np.int32)
        else:
            frames = np.frombuffer(raw, "<i4").astype(np.int32, copy=False)
        frames = np.ascontiguousarray(frames.reshape(n, channels))
        return PreAnalyzedSong(path=str(path), duration=n / float(rate), pcm_frames=frames, pcm_rate=int(rate))
