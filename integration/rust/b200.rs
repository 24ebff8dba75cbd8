// Extracted from INTEGRATION.md by scripts/extract_rust_shim.py -- edit the document, not this file.
// Uncompiled: the build image of this repository has no Rust toolchain.
//! Thin FFI over libbliss_b200.so. Every item cites the header entry point it binds.
use crate::{BlissError, BlissResult, FeaturesVersion};
use std::os::raw::{c_char, c_int, c_void};
use std::sync::Once;

#[link(name = "bliss_b200")]
extern "C" {
    fn bliss_b200_init(device: c_int) -> c_int;
    fn bliss_b200_init_devices(n_devices: c_int) -> c_int;
    fn bliss_b200_last_error() -> *const c_char;
    fn bliss_b200_feature_count(features_version: u16) -> u32;
    fn bliss_b200_analyze(pcm: *const f32, n_samples: u64, features_version: u16, out: *mut f32) -> c_int;
    fn bliss_b200_analyze_batch(
        pcm: *const *const f32, n_samples: *const u64, n_songs: u32, features_version: u16,
        out: *mut f32, status: *mut i32,
    ) -> c_int;
    fn bliss_b200_distance(a: *const f32, b: *const f32, dim: u32, metric: c_int, m: *const f32, out: *mut f32) -> c_int;
    fn bliss_b200_distance_matrix(
        rows: *const f32, n_rows: u32, cols: *const f32, n_cols: u32, dim: u32, metric: c_int,
        m: *const f32, out: *mut f32,
    ) -> c_int;
    fn bliss_b200_closest_to_songs(
        seeds: *const f32, n_seeds: u32, cands: *const f32, n_cands: u32, dim: u32, metric: c_int,
        m: *const f32, order: *mut u32, keys: *mut f32,
    ) -> c_int;
}

static INIT: Once = Once::new();
/// BLISS_B200_DEVICE=<n> pins the process to one GPU; otherwise every visible B200 is used: the crate is ONE process
/// (src/song/decoder.rs:282-331), so one `bliss_b200_analyze_batch` call deals its songs over all devices inside the
/// library (longest first, one host thread + copy stream per device) and the batcher below does not change.
fn ensure_init() {
    INIT.call_once(|| unsafe {
        match std::env::var("BLISS_B200_DEVICE").ok().and_then(|s| s.parse().ok()) {
            Some(dev) => assert_eq!(bliss_b200_init(dev), 0, "bliss_b200_init failed (no CPU fallback exists)"),
            None => assert!(bliss_b200_init_devices(0) > 0, "bliss_b200_init_devices failed (no CPU fallback exists)"),
        }
    });
}

fn call_error(code: c_int) -> BlissError {
    let msg = unsafe { std::ffi::CStr::from_ptr(bliss_b200_last_error()) }.to_string_lossy().into_owned();
    BlissError::AnalysisError(format!("b200 backend error {code}: {msg}"))
}

/// status -> BlissResult, same strings as src/song/mod.rs:426-430
fn status_to_result(status: i32, row: &[f32], v: FeaturesVersion) -> BlissResult<crate::Analysis> {
    match status {
        0 => crate::Analysis::new(row.to_vec(), v),
        1 => Err(BlissError::AnalysisError(String::from("empty or too short song."))),
        s => Err(BlissError::AnalysisError(format!("b200 backend: internal error (status {s})"))),
    }
}

pub(crate) fn analyze(sample_array: &[f32], v: FeaturesVersion) -> BlissResult<crate::Analysis> {
    ensure_init();
    let mut out = vec![0f32; v.feature_count()];
    let rc = unsafe { bliss_b200_analyze(sample_array.as_ptr(), sample_array.len() as u64, v as u16, out.as_mut_ptr()) };
    if rc < 0 { return Err(call_error(rc)); }
    status_to_result(rc, &out, v)
}

/// One GPU call for many decoded buffers (the batching seam of analyze_paths_with_options).
pub(crate) fn analyze_batch(buffers: &[&[f32]], v: FeaturesVersion) -> Vec<BlissResult<crate::Analysis>> {
    ensure_init();
    let dim = v.feature_count();
    let ptrs: Vec<*const f32> = buffers.iter().map(|b| b.as_ptr()).collect();
    let lens: Vec<u64> = buffers.iter().map(|b| b.len() as u64).collect();
    let mut out = vec![0f32; dim * buffers.len()];
    let mut status = vec![0i32; buffers.len()];
    let rc = unsafe {
        bliss_b200_analyze_batch(ptrs.as_ptr(), lens.as_ptr(), buffers.len() as u32, v as u16, out.as_mut_ptr(), status.as_mut_ptr())
    };
    if rc < 0 { return buffers.iter().map(|_| Err(call_error(rc))).collect(); }
    status.iter().enumerate().map(|(i, &s)| status_to_result(s, &out[i * dim..(i + 1) * dim], v)).collect()
}

pub(crate) fn mahalanobis(a: &[f32], b: &[f32], m: Option<&[f32]>) -> f32 {
    ensure_init();
    let mut d = 0f32;
    unsafe { bliss_b200_distance(a.as_ptr(), b.as_ptr(), a.len() as u32, 0, m.map_or(std::ptr::null(), |m| m.as_ptr()), &mut d) };
    d
}
