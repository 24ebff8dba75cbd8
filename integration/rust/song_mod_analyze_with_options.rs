// Extracted from INTEGRATION.md by scripts/extract_rust_shim.py -- edit the document, not this file.
// Uncompiled: the build image of this repository has no Rust toolchain.
#[cfg(feature = "b200")]
pub fn analyze_with_options(sample_array: &[f32], analysis_options: &AnalysisOptions) -> BlissResult<Analysis> {
    crate::b200::analyze(sample_array, analysis_options.features_version)
}
