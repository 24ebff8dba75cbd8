// Extracted from INTEGRATION.md by scripts/extract_rust_shim.py -- edit the document, not this file.
// Uncompiled: the build image of this repository has no Rust toolchain.
fn bliss_b200_analyze_batch_pcm(pcm: *const *const c_void, n_frames: *const u64, n_songs: u32,
                                sample_format: c_int /* 1 s16, 2 s32, 3 f32 */, channels: u32, sample_rate: u32,
                                features_version: u16, out: *mut f32, status: *mut i32) -> c_int;
fn bliss_b200_pcm_to_mono(pcm: *const c_void, n_frames: u64, sample_format: c_int, channels: u32,
                          out: *mut f32) -> c_int;
fn bliss_b200_resampled_len(n_samples: u64, sample_rate: u32) -> u64;
fn bliss_b200_resample(pcm: *const f32, n_samples: u64, sample_rate: u32, out: *mut f32, out_capacity: u64,
                       n_out: *mut u64) -> c_int;
