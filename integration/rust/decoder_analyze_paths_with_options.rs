// Extracted from INTEGRATION.md by scripts/extract_rust_shim.py -- edit the document, not this file.
// Uncompiled: the build image of this repository has no Rust toolchain.
#[cfg(feature = "b200")]
fn analyze_paths_with_options<P: Into<PathBuf>, F: IntoIterator<Item = P>>(
    paths: F,
    analysis_options: AnalysisOptions,
) -> mpsc::IntoIter<(PathBuf, BlissResult<Song>)> {
    const BATCH_SONGS: usize = 64; // ~1 GB of PCM for 3-minute tracks; the library chunks and overlaps the copies
    let mut cores = thread::available_parallelism().unwrap_or(NonZeroUsize::new(1).unwrap());
    if cores > analysis_options.number_cores { cores = analysis_options.number_cores; }
    let paths: Vec<PathBuf> = paths.into_iter().map(|p| p.into()).collect();
    let (tx, rx) = mpsc::channel::<(PathBuf, BlissResult<Song>)>();
    if paths.is_empty() { return rx.into_iter(); }
    // decoders -> batcher: decoded songs, still without an analysis
    let (dtx, drx) = mpsc::sync_channel::<PreAnalyzedSong>(2 * BATCH_SONGS);
    let chunk_length = std::cmp::max(paths.len() / cores, 1);
    for chunk in paths.chunks(chunk_length) {
        let (tx_thread, dtx_thread, owned_chunk) = (tx.clone(), dtx.clone(), chunk.to_owned());
        thread::spawn(move || {
            for path in owned_chunk {
                let is_cue = path.extension().map_or(false, |e| e.to_string_lossy().to_lowercase() == "cue");
                if is_cue {
                    match BlissCue::<Self>::songs_from_path(&path) {
                        Ok(songs) => songs.into_iter().for_each(|s| tx_thread.send((path.to_owned(), s)).unwrap()),
                        Err(e) => tx_thread.send((path.to_owned(), Err(e))).unwrap(),
                    }
                    continue;
                }
                match Self::decode(&path) {
                    Ok(pre) => dtx_thread.send(pre).unwrap(),                  // decoding stays on the CPU threads
                    Err(e) => tx_thread.send((path.to_owned(), Err(e))).unwrap(), // errors are items (:319-325)
                }
            }
        });
    }
    drop(dtx);
    // the batcher: one GPU call per <= BATCH_SONGS decoded songs
    thread::spawn(move || {
        let mut batch: Vec<PreAnalyzedSong> = Vec::with_capacity(BATCH_SONGS);
        let flush = |batch: &mut Vec<PreAnalyzedSong>| {
            let buffers: Vec<&[f32]> = batch.iter().map(|p| p.sample_array.as_slice()).collect();
            let results = crate::b200::analyze_batch(&buffers, analysis_options.features_version);
            for (pre, analysis) in batch.iter().zip(results) {
                let song = analysis.map(|analysis| pre.to_song_with_analysis(analysis, analysis_options));
                tx.send((pre.path.to_owned(), song)).unwrap();
            }
            batch.clear();
        };
        for pre in drx {
            batch.push(pre);
            if batch.len() == BATCH_SONGS { flush(&mut batch); }
        }
        if !batch.is_empty() { flush(&mut batch); }
    });
    rx.into_iter()
}
