#!/usr/bin/env python3
"""bench.py -- songs/sec of the Song::analyze hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port)

A "step" = one pass of the hot path over one batch: every rank analyses its shard of
synthetic 3-min 22 050 Hz f32 mono tracks (BASELINE.json configs[1]: 1024 tracks per GPU, full
descriptor set, PCM resident in HBM), the 23-float rows reach every rank (stored straight into the
peers' buffers by the analysis' last kernel + a one-warp epoch barrier; --gather nccl = the
all_gather comparator) and each rank computes its row block of the all-pairs distance matrix
(configs[3] shape).  Weak scaling: per-GPU work is fixed, value = all songs of all ranks /
max-over-ranks device time.

Prints ONE JSON line on rank 0 (see the task contract): value, e2e (same metric through the
C-ABI call with pinned HOST f32 buffers, H2D + D2H inside the timed region), e2e_s16 (the same
through bliss_b200_analyze_batch_s16: 16-bit host buffers converted on the device; an extra, the
metric itself is quoted on f32 PCM), roofline of the dominant kernel (CUDA-event timed inside this
run; roofline.traffic = its ncu DRAM bytes per launch, roofline.traffic_detail = the same in GB beside the
algorithmic bytes and the percentages of the resources that bind it, from profiles/ncu_traffic.json), cpu_baseline (oracle on the host cores), clocks, gpu_launches.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TRACK_SAMPLES = 3 * 60 * 22050          # 3 969 000 samples = 15 876 000 B
SONGS_PER_GPU = 1024                    # BASELINE.json configs[1]
E2E_SONGS = 256                         # songs per e2e step (pinned host -> device inside the timed region)
METRIC = "songs/sec (3-min 22050Hz f32 PCM) at 1/2/4/8 B200 vs ref CPU; STFT HBM GB/s"
BASE_SEED = 20260925


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.  ONE nvidia-smi process
    (started by rank 0) polls every local GPU of the job: eight pollers at 100 ms would contend on
    the driver with the ranks' own launches."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        self.indices, self.rows, self.proc = list(indices), [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", ",".join(str(i) for i in self.indices), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        per = {}
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                g = per.setdefault(int(f[0]), {"sm": [], "mx": [], "pw": []})
                g["sm"].append(float(f[1])); g["mx"].append(float(f[2])); g["pw"].append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not per:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons), "samples": 0}
        med = {k: float(np.median(v["sm"])) for k, v in per.items()}
        return {"sm_mhz": min(med.values()), "sm_max_mhz": max(max(v["mx"]) for v in per.values()),
                "power_w_max": max(max(v["pw"]) for v in per.values()),
                "samples": sum(len(v["sm"]) for v in per.values()), "reasons": sorted(reasons),
                "per_gpu_sm_mhz_median": {str(k): v for k, v in sorted(med.items())}}


def algorithmic_bytes_per_song():
    """DESIGN.md section 'roofline': bytes each kernel must move per 3-min song."""
    n = TRACK_SAMPLES
    n_s, n_t = (n - 512) // 128 + 1, (n - 512) // 256 + 1
    n_c, n_l = -(-n // 2205), -(-n // 1024)
    return {
        "pvoc512_kernel": 4 * n + 12 * n_s + 4 * n_t,          # read every sample once, 3 floats/frame + flux
        "timedomain_kernel": 4 * n + 4 * n_l + 4 * (n // 256),
        "stft8192_kernel": 4 * n + 4 * 4097 * n_c,             # read samples once, spill f32 magnitudes
        "chroma_kernel": 4 * 4097 * n_c + 80 * ((n_c + 127) // 128),
        "tuning_kernel": 0, "peakpick_kernel": 8 * n_t, "beattrack_kernel": 4 * n_t, "finalize_kernel": 12 * n_s,
    }


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust crate cannot be built
    here) on all host threads.  One step = 4 x `cores` songs, contiguous chunks per worker thread, scheduled
    like Decoder::analyze_paths_with_options (src/song/decoder.rs:278-332)."""
    if rank != 0:
        return
    import torch
    from bliss_rs_b200 import synth
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    n_songs = 4 * cores  # four songs per worker thread and step: ~2 s of wall per step on any box
    distinct = min(n_songs, 16)
    base = [synth.gen_track(BASE_SEED, i, TRACK_SAMPLES).numpy() for i in range(distinct)]
    songs = [base[i % distinct] for i in range(n_songs)]
    for _ in range(max(args.warmup, 0)):
        O.analyze_batch(songs[:cores], 2, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, _ = O.analyze_batch(songs, 2, n_threads=cores)
    dt = time.perf_counter() - t0
    val = n_songs * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "songs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic 3-min 22050 Hz f32 mono tracks, full 23-feature analysis "
                               "(CPU arm: %d songs per step, four per host thread)" % n_songs,
                   "track_samples": TRACK_SAMPLES},
        "cpu_baseline": {"value": val, "unit": "songs/s", "cores": cores, "kind": "port",
                         "sample": "%d x 3-min synthetic tracks per step, %d steps, %d threads (C oracle, "
                                   "-O3 -march=native, full complex FFT per frame like the reference)"
                                   % (n_songs, args.steps, cores),
                         "note": "C port of the reference's algorithm, not the Rust crate (no Rust toolchain here): rustfft's "
                                 "AVX kernels are expected to be 2-3x faster per FFT (BASELINE.md section 2)"},
        "e2e": {"value": val, "unit": "songs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


def _emit(line: dict):
    """The one JSON line goes to the ORIGINAL stdout; everything else any library prints (NCCL's
    version banner, warnings) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def stft_microbench(nat, torch, pcm, offs, S, stream, peak, peak_src):
    """BASELINE.json configs[2]: 512-point hanningz STFT, hop 256 (= PVocTempo, src/aubio.rs:338-425), 257 magnitudes
    per frame materialised in HBM, over >= 10 000 tracks: the step's resident tracks are looped (stated), each pass
    re-reads every sample from HBM (the input is far larger than L2)."""
    n_t = (TRACK_SAMPLES - 512) // 256 + 1
    R = min(S, 1024)
    try:
        mags = torch.empty((R * n_t, 257), dtype=torch.float32, device=pcm.device)
    except Exception as e:  # no room next to the analysis scratch: say so rather than fail the bench line
        return {"skipped": "no HBM for the magnitude buffer: %s" % str(e)[:80]}
    lens = [TRACK_SAMPLES] * R
    passes = max(1, -(-10000 // R))
    nat.stft512_mag_device(pcm.data_ptr(), offs[:R], lens, mags.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        nat.stft512_mag_device(pcm.data_ptr(), offs[:R], lens, mags.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tracks = passes * R
    rd = tracks * TRACK_SAMPLES * 4 / 1e9 / (ms / 1e3)
    wr = tracks * n_t * 257 * 4 / 1e9 / (ms / 1e3)
    del mags
    torch.cuda.empty_cache()
    return {"workload": "configs[2]: 512-pt hanningz STFT, hop 256, 257 magnitudes per frame written to HBM",
            "tracks": tracks, "resident_tracks": R, "passes": passes,
            "looped_resident": "the %d resident tracks (%.1f GB, far larger than L2) are re-read from HBM %d times"
                               % (R, R * TRACK_SAMPLES * 4 / 1e9, passes),
            "ms_total": ms, "tracks_per_s": tracks / (ms / 1e3), "read_gbs": rd, "write_gbs": wr,
            "frac_read": rd / peak, "frac_read_write": (rd + wr) / peak, "peak_gbs": peak, "peak_source": peak_src}


def _gpu_numa_cpus(torch, d):
    """CPUs of the NUMA node GPU d hangs off (sysfs; None when the container hides it)."""
    try:
        pr = torch.cuda.get_device_properties(d)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return node, (cpus & os.sched_getaffinity(0)) or None
    except Exception:
        return None, None


def measure_e2e(nat, torch, world, ES, host_songs, feats_ref, args, dim):
    """songs/s through bliss_b200_analyze_batch[_s16] from pinned host memory, one call for all `world` GPUs, plus the
    box's plain-copy H2D ceiling measured the same way (all GPUs copying at once from the same buffers)."""
    n_dev = nat.init_devices(world) if world > 1 else 1
    # one pinned copy of the ES songs per NUMA node that hosts a GPU, allocated and touched from that node's CPUs
    all_cpus = os.sched_getaffinity(0)
    node_of, bufs = {}, {}
    for d in range(n_dev):
        node, cpus = _gpu_numa_cpus(torch, d)
        node_of[d] = node if node is not None else -1
        if node_of[d] not in bufs:
            try:
                if cpus:
                    os.sched_setaffinity(0, cpus)
                b = torch.empty(ES * TRACK_SAMPLES, dtype=torch.float32, pin_memory=True)
                b.copy_(host_songs)
            finally:
                os.sched_setaffinity(0, all_cpus)
            bufs[node_of[d]] = b
    # song k of a call goes to device k % n_dev (equal lengths: the library's longest-first deal is round-robin)
    n_call = ES * n_dev
    ptrs = (ctypes.c_void_p * n_call)(*[bufs[node_of[k % n_dev]].data_ptr() + 4 * (k // n_dev) * TRACK_SAMPLES
                                        for k in range(n_call)])
    hlens = (ctypes.c_uint64 * n_call)(*([TRACK_SAMPLES] * n_call))
    out = np.zeros((n_call, dim), np.float32)
    status = np.zeros(n_call, np.int32)
    nat.analyze_batch_ptrs(ptrs, hlens, 2, out, status)  # warm-up (allocations on every device)
    steps = max(2, min(args.steps, 4))
    t0 = time.perf_counter()
    for _ in range(steps):
        nat.analyze_batch_ptrs(ptrs, hlens, 2, out, status)
    dt = time.perf_counter() - t0
    want = np.repeat(feats_ref, n_dev, axis=0)  # call song k = resident song k // n_dev
    # the plain-copy ceiling of the same transfer pattern: every GPU copies its ES songs at once, nothing else runs
    ceil_gbs = None
    try:
        dsts = [torch.empty(ES * TRACK_SAMPLES, dtype=torch.float32, device="cuda:%d" % d) for d in range(n_dev)]
        strs = [torch.cuda.Stream(device=d) for d in range(n_dev)]

        def copy_all():
            for d in range(n_dev):
                with torch.cuda.stream(strs[d]):
                    dsts[d].copy_(bufs[node_of[d]], non_blocking=True)
            for d in range(n_dev):
                strs[d].synchronize()
        copy_all()
        tc = time.perf_counter()
        for _ in range(3):
            copy_all()
        ceil_gbs = 3 * n_dev * ES * TRACK_SAMPLES * 4 / 1e9 / (time.perf_counter() - tc)
        del dsts
        torch.cuda.empty_cache()
    except Exception as e:
        print("[bench] H2D ceiling not measured: %s" % e, file=sys.stderr)
    gbs = n_call * steps * TRACK_SAMPLES * 4 / 1e9 / dt
    e2e = {"value": n_call * steps / dt, "unit": "songs/s", "h2d_bytes_per_step": n_call * TRACK_SAMPLES * 4,
           "d2h_bytes_per_step": n_call * dim * 4, "songs_per_step": n_call, "steps": steps, "devices": n_dev,
           "bitwise_equal_to_device_path": bool(np.array_equal(out, want)) and bool((status == 0).all()),
           "h2d_gbs": gbs, "h2d_ceiling_gbs": ceil_gbs, "frac_of_h2d_ceiling": (gbs / ceil_gbs) if ceil_gbs else None,
           "numa_nodes_of_gpus": [node_of[d] for d in range(n_dev)],
           "note": "ONE bliss_b200_analyze_batch call from one process over %d GPU(s) (bliss_b200_init_devices), pinned host "
                   "buffers local to each GPU's NUMA node; PCIe-bound (15.9 MB per song): see frac_of_h2d_ceiling" % n_dev}
    # 16-bit sources: bliss_b200_analyze_batch_s16 (s16 -> f32 on the device, half the PCIe bytes).  Reported next to
    # `e2e`, never instead of it: BASELINE.json's metric is quoted on f32 PCM.
    bufs16 = {k: torch.empty(ES * TRACK_SAMPLES, dtype=torch.int16, pin_memory=True) for k in bufs}
    for k in bufs:
        bufs16[k].copy_((bufs[k] * 32767.0).round().to(torch.int16))
    ptrs16 = (ctypes.c_void_p * n_call)(*[bufs16[node_of[k % n_dev]].data_ptr() + 2 * (k // n_dev) * TRACK_SAMPLES
                                          for k in range(n_call)])
    out16 = np.zeros((n_call, dim), np.float32)
    nat.analyze_batch_s16_ptrs(ptrs16, hlens, 2, out16, status)  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        nat.analyze_batch_s16_ptrs(ptrs16, hlens, 2, out16, status)
    dt16 = time.perf_counter() - t0
    # the same samples converted on the host and sent as f32 must give the same bits
    for k in bufs:
        bufs[k].copy_(bufs16[k].to(torch.float32) / 32768.0)
    nat.analyze_batch_ptrs(ptrs, hlens, 2, out, status)
    e2e_s16 = {"value": n_call * steps / dt16, "unit": "songs/s", "h2d_bytes_per_step": n_call * TRACK_SAMPLES * 2,
               "d2h_bytes_per_step": n_call * dim * 4, "songs_per_step": n_call, "steps": steps, "devices": n_dev,
               "bitwise_equal_to_f32_path": bool(np.array_equal(out16, out)),
               "note": "bliss_b200_analyze_batch_s16: signed 16-bit mono 22 050 Hz samples, x/32768 on the device"}
    del bufs16
    # CD-format sources (interleaved s16 stereo at 44 100 Hz, what most files decode to): down-mix, the conversion to
    # 22 050 Hz and the analysis behind ONE bliss_b200_analyze_batch_pcm call.  An extra next to `e2e`, twice its
    # bytes per song; the resampler's parity against the reference's swresample / rubato is unpinned.
    e2e_cd = None
    try:
        per = min(ES, 32)                                 # songs per device in this leg (31.8 MB each)
        n_cd = per * n_dev
        cd = {}
        for k in bufs:
            t = bufs[k][:per * TRACK_SAMPLES].view(per, TRACK_SAMPLES)
            nxt = torch.cat([t[:, 1:], t[:, -1:]], 1)
            up2 = torch.stack([t, 0.5 * (t + nxt)], 2).reshape(per, 2 * TRACK_SAMPLES)          # linear 2 x
            q = (up2 * 32767.0).round().to(torch.int16)
            st = torch.stack([q, q], 2)                                                           # L = R
            cd[k] = torch.empty(st.shape, dtype=torch.int16, pin_memory=True)
            cd[k].copy_(st)
        fb = 4
        ptrs_cd = (ctypes.c_void_p * n_cd)(*[cd[node_of[k % n_dev]].data_ptr() + fb * (k // n_dev) * 2 * TRACK_SAMPLES
                                             for k in range(n_cd)])
        lens_cd = (ctypes.c_uint64 * n_cd)(*([2 * TRACK_SAMPLES] * n_cd))
        out_cd = np.zeros((n_cd, dim), np.float32)
        st_cd = np.zeros(n_cd, np.int32)
        nat.analyze_batch_pcm_ptrs(ptrs_cd, lens_cd, nat.PCM_S16, 2, 44100, 2, out_cd, st_cd)  # warm-up
        t0 = time.perf_counter()
        for _ in range(steps):
            nat.analyze_batch_pcm_ptrs(ptrs_cd, lens_cd, nat.PCM_S16, 2, 44100, 2, out_cd, st_cd)
        dtcd = time.perf_counter() - t0
        # the three steps apart, through the C ABI, on song 0: the fused call must give the same bits
        mono0 = nat.resample(nat.pcm_to_mono(cd[node_of[0]][0].numpy()), 44100)
        st0, f0 = nat.analyze_batch([mono0], 2)
        e2e_cd = {"value": n_cd * steps / dtcd, "unit": "songs/s", "h2d_bytes_per_step": n_cd * 2 * TRACK_SAMPLES * fb,
                  "d2h_bytes_per_step": n_cd * dim * 4, "songs_per_step": n_cd, "steps": steps, "devices": n_dev,
                  "h2d_gbs": n_cd * steps * 2 * TRACK_SAMPLES * fb / 1e9 / dtcd,
                  "frac_of_h2d_ceiling": (n_cd * steps * 2 * TRACK_SAMPLES * fb / 1e9 / dtcd / ceil_gbs) if ceil_gbs else None,
                  "all_ok": bool((st_cd == 0).all()),
                  "bitwise_equal_to_separate_steps": bool(st0[0] == 0 and np.array_equal(f0[0], out_cd[0])),
                  "max_abs_diff_to_features_of_the_22k05_version": float(np.abs(out_cd - np.repeat(feats_ref[:per], n_dev, axis=0)[:n_cd]).max()),
                  "note": "bliss_b200_analyze_batch_pcm: interleaved s16 stereo at 44 100 Hz (3-min songs, 31.8 MB each); "
                          "down-mix + polyphase resampler + analysis on the device; resampler parity unpinned"}
        del cd
    except Exception as e:  # an extra: never takes the bench line down
        print("[bench] e2e_cd not measured: %s" % e, file=sys.stderr)
    return e2e, e2e_s16, e2e_cd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--songs-per-gpu", type=int, default=SONGS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5],
                    help="2 = BASELINE.json configs[1] (the bench line the driver reads); 4 = configs[3] (100 k tracks, "
                         "waves, fused gather, all-pairs row blocks: bench_config4.py); 5 = configs[4] (Zipf mixed-duration "
                         "corpus end to end through ONE multi-device call + playlist order vs the oracle: bench_config5.py)")
    ap.add_argument("--kernels-only", action="store_true",
                    help="A/B runs: device-timed value + per-kernel times only (no e2e legs, no CPU baseline)")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "nccl"],
                    help="N>1 feature-row exchange: p2p = rows stored into every rank's buffer by the analysis "
                         "itself + one-warp epoch barrier; nccl = all_gather_into_tensor; auto = p2p, else nccl")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.config == 4:
        import bench_config4
        sys.argv = [sys.argv[0]]
        return bench_config4.main(emit=_emit)  # (stdout stays on stderr: NCCL prints its banner there)
    if args.config == 5:
        import bench_config5
        return bench_config5.main(emit=_emit, argv=["--gpus", str(max(args.gpus, world))] + (["--songs", os.environ["BLISS_CFG5_SONGS"]] if os.environ.get("BLISS_CFG5_SONGS") else []))

    import torch
    import torch.distributed as dist
    import bliss_rs_b200 as B
    from bliss_rs_b200 import synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nat = B.native
    nat.init(local_rank)
    S = args.songs_per_gpu
    dim = 23

    # ---- synthetic corpus, generated in HBM (not timed) --------------------------------
    from bliss_rs_b200 import multigpu as M
    n_total = world * S
    my_ids = M.shard_round_robin(n_total, world, rank)       # BASELINE config 4: round-robin shards
    lengths = [TRACK_SAMPLES] * S
    pcm, offs, lens = synth.gen_corpus_flat(BASE_SEED, my_ids, lengths, device=dev)
    feats = torch.zeros((S, dim), dtype=torch.float32, device=dev)
    all_feats = torch.zeros((n_total, dim), dtype=torch.float32, device=dev)
    row_lo, row_hi = M.row_block(n_total, world, rank)       # this rank's rows of the all-pairs matrix
    dmat = torch.zeros((row_hi - row_lo, n_total), dtype=torch.float32, device=dev)
    weights = nat.feature_weights(2)
    stream = torch.cuda.current_stream().cuda_stream

    # N>1: the feature rows of all ranks must meet before the distance row block.  Fused form: the
    # analysis' last kernel stores each row into every rank's buffer over NVLink (global song order, so
    # no permutation either) and a one-warp epoch barrier replaces the collective.
    gather, gather_info = None, None
    if world > 1 and args.gather in ("auto", "p2p"):
        ok = torch.ones(1, device=dev)
        try:
            gather = M.PeerGather(n_total, dev)
        except Exception as e:  # no peer access / IPC refused by the container
            if args.gather == "p2p":
                raise
            ok.zero_()
            print("[bench] peer gather unavailable on rank %d: %s" % (rank, e), file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            gather = None
    if world > 1:
        gather_info = {"kind": "p2p-fused" if gather else "nccl"}

    def step_nccl(out):
        nat.analyze_batch_device(pcm.data_ptr(), offs, lens, 2, feats.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(all_feats, feats)        # the ONE collective of the path (NCCL)
            cols = M.round_robin_to_global(all_feats, world)      # rank-major -> global song order
        else:
            cols = feats
        rows = cols[row_lo:row_hi]
        nat.distance_matrix_device(rows.data_ptr(), row_hi - row_lo, cols.data_ptr(), n_total, dim,
                                   out.data_ptr(), nat.METRIC_MAHALANOBIS, weights, stream)
        return cols

    def step_p2p(out):
        gather.scatter(pcm.data_ptr(), offs, lens, 2, rank, world, feats.data_ptr(), stream)
        cols = gather.commit(n_total, dim, stream)
        rows = cols[row_lo:row_hi]
        nat.distance_matrix_device(rows.data_ptr(), row_hi - row_lo, cols.data_ptr(), n_total, dim,
                                   out.data_ptr(), nat.METRIC_MAHALANOBIS, weights, stream)
        return cols

    def step():
        return step_p2p(dmat) if gather else step_nccl(dmat)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    if gather and args.gather == "auto":
        # a peer mapping that opened but does not carry the stores shows up as a barrier time-out in the
        # warm-up: every rank then drops to the NCCL exchange together instead of failing the run
        ok = torch.ones(1, device=dev)
        try:
            gather.check()
        except Exception as e:
            ok.zero_()
            print("[bench] peer gather failed its warm-up on rank %d: %s" % (rank, e), file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            gather = None
            gather_info = {"kind": "nccl", "note": "fused peer-store exchange failed its warm-up check"}
            for _ in range(args.warmup):
                step()
    feats_warm = feats.clone()   # every later step must reproduce these bits (no timing-dependent result)
    # The sampler is started BEFORE the barrier: forking nvidia-smi from a process this size takes ~75 ms,
    # which used to delay rank 0's first launch -- every other rank then sat in its first exchange waiting
    # for rank 0 inside the timed region (seen as 471 vs 396 ms per 5 steps on ranks 1-7 at N=8).
    sampler = ClockSampler(range(world)) if rank == 0 else None  # one node: local GPUs 0..world-1
    if sampler:
        sampler.start()
    barrier()
    launches0 = nat.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = nat.launch_count() - launches0
    deterministic = torch.tensor([float(torch.equal(feats, feats_warm))], device=dev)
    clocks = sampler.stop() if sampler else {}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    lz = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        # per-rank diagnostics (device ms of the timed region, median SM clock, max power) for the scaling analysis
        mine = torch.tensor([ms], dtype=torch.float64, device=dev)
        allr = torch.zeros((world, 1), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = {"ms": [round(v, 2) for v in allr[:, 0].tolist()]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lz, op=dist.ReduceOp.SUM)
        dist.all_reduce(deterministic, op=dist.ReduceOp.MIN)
    ms = float(t.item())
    value = world * S * args.steps / (ms / 1e3)
    if gather:
        # the fused exchange against the NCCL one: same rows, same distance block, bit for bit
        gather.check()
        cols_p2p = step_p2p(dmat).clone()
        dmat2 = torch.empty_like(dmat)
        cols_nccl = step_nccl(dmat2).clone()
        torch.cuda.synchronize()
        same = torch.tensor([float(torch.equal(cols_p2p, cols_nccl) and torch.equal(dmat, dmat2))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        gather_info["bitwise_equal_to_nccl_path"] = bool(same.item())
        del dmat2
        tn0, tn1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        tn0.record()
        for _ in range(args.steps):
            step_nccl(dmat)
        tn1.record()
        barrier()
        tn = torch.tensor([tn0.elapsed_time(tn1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        gather_info["nccl_path_ms_per_step"] = float(tn.item()) / args.steps
        gather_info["fused_path_ms_per_step"] = ms / args.steps

    # ---- per-kernel device times (CUDA events on the launching stream) -> roofline -------
    nat.set_profiling(True)
    nat.get_profile()
    prof_steps = 2
    for _ in range(prof_steps):
        step()
    kms, klaunch = nat.get_profile()
    nat.set_profiling(False)
    names = nat.kernel_names()
    peak, peak_src = _peaks()
    alg = algorithmic_bytes_per_song()
    total_kms = sum(kms) or 1.0
    kernels = []
    for nm, m_, l_ in zip(names, kms, klaunch):
        if l_ == 0:
            continue
        avg_ms = m_ / l_
        b = alg.get(nm, 0) * S
        kernels.append({"kernel": nm, "avg_ms": avg_ms, "share": m_ / total_kms,
                        "algorithmic_gbs": (b / 1e9) / (avg_ms / 1e3) if avg_ms > 0 else None})
    kernels.sort(key=lambda k: -k["share"])
    dom = kernels[0]
    traffic = traffic_detail = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            ent = json.load(open(tp)).get(dom["kernel"])
            # measured DRAM bytes per song (ncu --set full, profiles/) x songs per launch: bytes per launch, a plain
            # number like the algorithmic bytes behind `achieved`
            traffic = float(ent["bytes_per_song"]) * S
            traffic_detail = {"unit": "bytes per launch", "gb_per_launch": traffic / 1e9,
                              "algorithmic_gb_per_launch": alg.get(dom["kernel"], 0) * S / 1e9,
                              "source": "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum)"}
            # what the same ncu capture says the kernel is actually bound by (percent of peak)
            for k in ("l1tex_data_pipe_pct", "issue_active_pct", "fma_pipe_pct", "dram_pct", "lts_pct"):
                if k in ent:
                    traffic_detail[k] = ent[k]
        except Exception:
            traffic = traffic_detail = None
    step_read_gbs = S * TRACK_SAMPLES * 4 / 1e9 / (ms / args.steps / 1e3)  # every PCM byte once per step / step time
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["algorithmic_gbs"], "peak": peak,
                "unit": "GB/s", "frac": dom["algorithmic_gbs"] / peak, "traffic": traffic,
                "traffic_source": "static: ncu capture committed under profiles/ (not re-measured in this run)",
                "traffic_detail": traffic_detail, "peak_source": peak_src,
                "step_read_gbs": step_read_gbs, "step_read_frac": step_read_gbs / peak,
                "step_read_note": "BASELINE.md section 3: songs x 4N bytes / step time / HBM peak (per GPU); the FFT kernels "
                                  "are FP32-pipe bound, so this sits far below 1 by construction",
                "avg_launch_ms": dom["avg_ms"], "share_of_step": dom["share"],
                "note": "FFT work: at algorithmic-minimum traffic the kernel sits on the SM (FP32 pipe ~55 %, L1 / shared-memory "
                        "data pipe ~47 %, issue ~44 % in the ncu capture: no pipe saturated, latency between barriers; "
                        "DESIGN.md section 4); the HBM roofline is reported because BASELINE.json fixes it",
                "kernels": kernels}

    # ---- e2e: the C-ABI call with pinned HOST buffers (H2D + D2H inside the timed region) --
    # ---- STFT-only micro-benchmark (BASELINE.json configs[2]; the "STFT HBM GB/s" half of the metric) -------------
    stft_micro = None
    if world == 1 and not args.kernels_only:
        stft_micro = stft_microbench(nat, torch, pcm, offs, S, stream, peak, peak_src)

    if args.kernels_only:
        if rank == 0:
            _emit({"metric": METRIC, "value": value, "unit": "songs/s", "n_gpus": world, "steps": args.steps,
                   "warmup": args.warmup, "ms_per_step": ms / args.steps, "kernels_only": True,
                   "config": {"kernel_variant_mask": int(os.environ.get("BLISS_B200_VARIANT", "0") or 0),
                              "songs_per_gpu": S},
                   "roofline": roofline, "clocks": clocks, "gpu_launches": int(lz.item()),
                   "bitwise_reproducible_across_steps": bool(deterministic.item())})
        if world > 1:
            dist.barrier()
            if gather:
                gather.destroy()
            dist.destroy_process_group()
        return

    # ---- e2e: the C-ABI call with pinned HOST buffers (H2D + D2H inside the timed region) -------------------------
    # ONE process drives every GPU (bliss_b200_init_devices): the reference is one process (worker threads + a channel,
    # src/song/decoder.rs:282-331), so that is what its drop-in has to be measured through.  Under torchrun rank 0 owns
    # the call; the other ranks free their GPUs and wait on a CPU (gloo) barrier -- an NCCL barrier would spin on the
    # very SMs being measured.
    ES = min(E2E_SONGS, S)
    host_songs = torch.empty(ES * TRACK_SAMPLES, dtype=torch.float32)
    if rank == 0:
        for i in range(ES):
            host_songs[i * TRACK_SAMPLES:(i + 1) * TRACK_SAMPLES].copy_(pcm[offs[i]:offs[i] + TRACK_SAMPLES])
    feats_ref = feats[:ES].cpu().numpy()
    cpu_songs = cpu_feats = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = min(max(4 * (os.cpu_count() or 1), 64), 256, S)  # a few seconds of wall on all cores (tens of core-seconds)
        cpu_songs = [pcm[offs[i]:offs[i] + TRACK_SAMPLES].cpu().numpy() for i in range(n_cpu)]
        cpu_feats = feats[:n_cpu].cpu().numpy()
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None
    if gather:
        gather.destroy()
        gather = None
    del pcm, dmat, all_feats
    torch.cuda.synchronize()
    if rank != 0:
        nat.shutdown()      # this rank's context and scratch leave its GPU: rank 0 opens its own context there
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier(group=cpu_group)
    e2e = e2e_s16 = e2e_cd = None
    if rank == 0:
        e2e, e2e_s16, e2e_cd = measure_e2e(nat, torch, world, ES, host_songs, feats_ref, args, dim)
    if world > 1:
        dist.barrier(group=cpu_group)

    # ---- CPU baseline (oracle = port of the reference algorithm) on rank 0, bounded sample ----
    cpu = None
    if cpu_songs is not None:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        t0 = time.perf_counter()
        ost, ofe = O.analyze_batch(cpu_songs, 2, n_threads=cores)
        dtc = time.perf_counter() - t0
        err = np.abs(cpu_feats - ofe)
        tol = 1e-4 * np.maximum(1.0, np.abs(ofe))
        cpu = {"value": len(cpu_songs) / dtc, "unit": "songs/s", "cores": cores, "kind": "port",
               "sample": "%d of the same synthetic 3-min tracks (D2H-copied), %d threads, %.1f s wall"
                         % (len(cpu_songs), cores, dtc),
               "note": "C port of the reference's algorithm (plain radix FFT, -O3 -march=native), not the Rust crate: "
                       "rustfft's AVX kernels are expected to be 2-3x faster per FFT (BASELINE.md section 2), so the "
                       "speed-up over the crate itself is that much lower than value / cpu_baseline.value",
               "parity_max_abs_err": float(err.max()), "parity_within_1e-4": bool((err <= tol).all()),
               "parity_per_feature_max_abs_err": [float(v) for v in err.max(0)],
               "parity_tempo_mismatches": int((err[:, 0] > 1e-3).sum())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "songs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: %d synthetic 3-min 22050 Hz f32 mono tracks per GPU, full 23-feature "
                                   "analysis + feature-row exchange + all-pairs distance row block" % S,
                       "songs_per_gpu": S, "track_samples": TRACK_SAMPLES,
                       "kernel_variant_mask": int(os.environ.get("BLISS_B200_VARIANT", "0") or 0), "parallelism": "songs sharded %d-way" % world,
                       "l2": "inputs (%.1f GB PCM per GPU) are far larger than the 126 MB L2; no flush needed"
                             % (S * TRACK_SAMPLES * 4 / 1e9)},
            "e2e": e2e, "e2e_s16": e2e_s16, "e2e_cd": e2e_cd, "stft_microbench": stft_micro, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "per_rank": per_rank,
            "gather": gather_info,
            "gpu_launches": int(lz.item()), "bitwise_reproducible_across_steps": bool(deterministic.item()),
        }
        _emit(line)
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
