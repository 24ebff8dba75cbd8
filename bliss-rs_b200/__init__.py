"""bliss-rs_b200 -- B200-native implementation of bliss-rs's Song::analyze + distance path.

Host-side mirror (Python) of the reference crate's API for that path, on top of the
C ABI in include/bliss_b200.h (libbliss_b200.so: hand-written sm_100a CUDA kernels).
Names, argument meaning and error behaviour follow the Rust items cited in each
docstring.  No CPU implementation exists in this package.
"""
from . import _native as native
from .song import (AnalysisIndex, AnalysisIndexv1, Analysis, AnalysisOptions, BlissError, AnalysisError,
                   DecodingError, ProviderError, Decoder, FeaturesVersion, PreAnalyzedSong, Song,
                   NUMBER_FEATURES, SAMPLE_RATE, CHANNELS, analyze_batch, analyze_batch_pcm, pcm_to_mono)
from .song import analyze_decoded
from .decoder import WavDecoder
from .cue import BlissCue
from . import cue
from . import playlist
from . import library

__all__ = ["native", "playlist", "library", "AnalysisIndex", "AnalysisIndexv1", "Analysis", "AnalysisOptions", "BlissError",
           "AnalysisError", "DecodingError", "ProviderError", "Decoder", "FeaturesVersion", "PreAnalyzedSong",
           "Song", "NUMBER_FEATURES", "SAMPLE_RATE", "CHANNELS", "analyze_batch", "analyze_batch_pcm", "pcm_to_mono", "analyze_decoded",
           "WavDecoder", "BlissCue", "cue"]
