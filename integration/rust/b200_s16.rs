// Extracted from INTEGRATION.md by scripts/extract_rust_shim.py -- edit the document, not this file.
// Uncompiled: the build image of this repository has no Rust toolchain.
fn bliss_b200_analyze_batch_s16(pcm: *const *const i16, n_samples: *const u64, n_songs: u32,
                                features_version: u16, out: *mut f32, status: *mut i32) -> c_int;
