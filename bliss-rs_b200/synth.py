"""Deterministic synthetic "music-like" PCM (mono f32, 22 050 Hz) for parity tests and
benchmarks: a kick on every beat, a looped chord progression
of harmonic tones with a small random detune, a little broadband noise, random gain.
White noise alone would make the tempo / tuning decisions degenerate.

Pure torch, so the same code generates on the CPU (tests: fed to both the oracle and the
GPU path) and directly in HBM (bench).  Track i of a corpus depends only on (base_seed, i).
"""
import math

import torch

SR = 22050
_SCALE = [0, 2, 4, 5, 7, 9, 11]  # major scale degrees (semitones)
_CHORDS = [[0, 2, 4], [3, 5, 0], [4, 6, 1], [5, 0, 2], [1, 3, 5]]  # scale-degree triads


def track_params(base_seed: int, track_id: int):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(base_seed) * 1000003 + int(track_id)) % (2 ** 62))
    r = torch.rand(16, generator=g, dtype=torch.float64).tolist()
    bpm = 60.0 + 120.0 * r[0]
    root = 48 + int(r[1] * 12)                  # MIDI note of the key
    detune_cents = -30.0 + 60.0 * r[2]
    gain = 0.05 + 0.45 * r[3]
    n_chords = 3 + int(r[4] * 2)
    order = torch.randperm(len(_CHORDS), generator=g)[:n_chords].tolist()
    noise_amp = 10 ** (-30.0 / 20.0) * (0.5 + r[5])
    kick_amp = 0.5 + 0.5 * r[6]
    noise_seed = int(r[7] * (2 ** 31))
    return dict(bpm=bpm, root=root, detune=detune_cents, gain=gain, chords=order, noise_amp=noise_amp,
                kick_amp=kick_amp, noise_seed=noise_seed)


def gen_track(base_seed: int, track_id: int, n_samples: int, device="cpu") -> torch.Tensor:
    p = track_params(base_seed, track_id)
    dev = torch.device(device)
    t = torch.arange(n_samples, device=dev, dtype=torch.float64) / SR
    period = 60.0 / p["bpm"]
    tau = torch.remainder(t, period)
    # kick: 20 ms-ish decaying burst sweeping 120 -> 60 Hz
    kick = torch.sin(2 * math.pi * (60.0 * tau + 0.6 * (1.0 - torch.exp(-tau / 0.01)))) * torch.exp(-tau / 0.04)
    # chord changes every 2 beats
    chord_idx = torch.remainder(torch.floor(t / (2 * period)), len(p["chords"])).to(torch.long)
    f0_table = []
    for ci in p["chords"]:
        notes = []
        for deg in _CHORDS[ci]:
            midi = p["root"] + _SCALE[deg % 7] + 12 * (deg // 7)
            notes.append(440.0 * 2 ** ((midi - 69) / 12.0) * 2 ** (p["detune"] / 1200.0))
        f0_table.append(notes)
    f0_table = torch.tensor(f0_table, device=dev, dtype=torch.float64)  # [n_chords, 3]
    tone = torch.zeros(n_samples, device=dev, dtype=torch.float64)
    for note in range(3):
        f0 = f0_table[:, note][chord_idx]
        ph = torch.remainder(f0 * t, 1.0)
        for h in range(1, 7):
            tone += torch.sin(2 * math.pi * h * ph) / (h * 3.0)
    # slow tremolo so consecutive bars are not bit-identical
    tone *= 0.75 + 0.25 * torch.sin(2 * math.pi * 0.31 * t)
    g = torch.Generator(device=dev)
    g.manual_seed(p["noise_seed"])
    noise = torch.randn(n_samples, device=dev, dtype=torch.float32, generator=g)
    x = p["gain"] * (0.6 * tone.to(torch.float32) + p["kick_amp"] * kick.to(torch.float32)) + p["noise_amp"] * noise
    return torch.clamp(x, -1.0, 1.0).contiguous()


def gen_corpus_flat(base_seed: int, track_ids, lengths, device="cpu", align=4):
    """Concatenates tracks `track_ids` into ONE flat buffer (each start aligned to `align`
    samples).  Returns (pcm [total], offsets list, lengths list)."""
    offsets, total = [], 0
    for n in lengths:
        offsets.append(total)
        total += (int(n) + align - 1) // align * align
    pcm = torch.zeros(max(total, align), device=device, dtype=torch.float32)
    for i, (o, n) in enumerate(zip(offsets, lengths)):
        pcm[o:o + n] = gen_track(base_seed, int(track_ids[i]), int(n), device)
    return pcm, offsets, [int(n) for n in lengths]
