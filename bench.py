#!/usr/bin/env python3
"""Benchmark of the AXCTD decode path (BASELINE.json metric: audio-seconds decoded
per second, x real-time, per GPU and at 2/4/8 B200; HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W              # this engine
    python bench.py --impl reference --steps K --warmup W      # reference CPU path (oracle port) on host cores

Workload (config.workload): the per-GPU share of BASELINE config 4 -- a batch of
independent 12-minute synthetic drops, alternating 44.1 / 48 kHz, generated
directly in HBM by the device twin of synth.py (distinct seed per drop).  One
step = one pass of the whole decode path over the batch.  Weak scaling: every
rank decodes its own --drops drops, no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import synth  # noqa: E402

METRIC = "audio_seconds_decoded_per_second"
UNIT = "x_realtime"


def workload_config(drops, duration, total_samples, shards=1):
    return {"workload": f"BASELINE config 4 share: {drops} drops/GPU x {duration:.0f} s, 44.1/48 kHz alternating, "
                        f"SNR 40/25/10 dB, device-generated (synth.py twin), default 1200 Hz lowpass",
            "drops_per_gpu": drops, "samples_per_gpu": total_samples, "sharding": "by drop, no collective", "sub_batches_in_flight_per_gpu": shards,
            "cache": "inputs (%.1f GB per GPU) larger than L2" % (2e-9 * total_samples)}


def drop_specs(n_drops, duration_s, rank):
    return [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=duration_s, seed=100000 * (rank + 1) + i,
                           snr_db=(40.0, 25.0, 10.0)[i % 3]) for i in range(n_drops)]


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md), sampled through NVML every few
    milliseconds (the timed region of a default run is under 0.1 s: `nvidia-smi -lms` does not even start in that
    time); falls back to one-shot nvidia-smi queries when the NVML bindings are missing."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period_s=0.004):
        super().__init__(daemon=True)
        self.index, self.period, self.rows, self.stop_flag = index, period_s, [], False
        self.nv, self.h, self.max_mhz = None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv, self.h = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _sample(self):
        if self.nv is not None:
            nv = self.nv
            mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            return mhz, mask
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        self.max_mhz = float(out[1])
        mask = 0
        for (name, bit), v in zip(self.REASONS, out[2:6]):
            if v.strip().lower().startswith("active"):
                mask |= bit
        return float(out[0]), mask

    def run(self):
        while not self.stop_flag:
            try:
                self.rows.append(self._sample())
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=15)
        sm = [r[0] for r in self.rows]
        mask = 0
        for r in self.rows:
            mask |= r[1]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for name, bit in self.REASONS if mask & bit), "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------- CPU baseline
def _oracle_one(args):
    pcm, fs = args
    sys.path.insert(0, ROOT)
    from oracle import axctd_oracle as ao
    t0 = time.perf_counter()
    op = ao.process_pcm(pcm, fs)
    return time.perf_counter() - t0, len(op.time)


def _reference_one(args):
    """The UNMODIFIED reference CLI (processAXCTD.main(), reference processAXCTD.py:47-183: WAV read, normalise,
    run(), output file) under oracle/ref_shim.py, on a WAV file written beforehand."""
    wav, out = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    t0 = time.perf_counter()
    ref_shim.run_cli(["-i", wav, "-o", out])
    dt = time.perf_counter() - t0
    with open(out) as f:
        rows = sum(1 for _ in f)
    return dt, rows


def reference_available():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    return ref_shim.available()


def cpu_baseline_run(pcms, fss, durations, cores, kind="port", workdir=None):
    """One process per host core, one drop each; aggregate audio-seconds per wall-second.  kind "reference": the
    unmodified reference CLI on WAV files in `workdir`; kind "port": oracle/axctd_oracle.py (numpy restatement)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    if kind == "reference":
        jobs = [(os.path.join(workdir, f"drop{i}.wav"), os.path.join(workdir, f"drop{i}.txt")) for i in range(len(pcms))]
        fn = _reference_one
    else:
        jobs, fn = list(zip(pcms, fss)), _oracle_one
    t0 = time.perf_counter()
    with ctx.Pool(processes=cores) as pool:
        out = pool.map(fn, jobs)
    wall = time.perf_counter() - t0
    return sum(durations) / wall, wall, out


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def write_wavs(pcms, fss, workdir):
    for i, (p, fs) in enumerate(zip(pcms, fss)):
        synth.write_wav(os.path.join(workdir, f"drop{i}.wav"), p, fs)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores -- the unmodified
    reference (baseline/_ref, installed by __graft_entry__.build()) when present, else the oracle port.  Nothing of
    this repository's engine is loaded: the PCM comes from the numpy generator (synth.generate_drop)."""
    import tempfile
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(host_cores(), args.cpu_procs) if args.cpu_procs else host_cores()
    kind = "reference" if reference_available() and not args.cpu_port else "port"
    if kind == "reference":                   # import the reference's modules once; the forked workers inherit them
        import ref_shim
        ref_shim.load_modules()
    dur = args.cpu_duration if args.cpu_duration > 0 else (120.0 if kind == "reference" else 720.0)
    specs = drop_specs(cores, dur, 0)
    with tempfile.TemporaryDirectory() as td:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(processes=cores) as pool:
            pcms = pool.map(synth.generate_drop, specs)
        fss = [s.fs for s in specs]
        if kind == "reference":
            write_wavs(pcms, fss, td)
        for _ in range(args.warmup):
            cpu_baseline_run(pcms, fss, [dur] * cores, cores, kind, td)
        t0 = time.perf_counter()
        rows = 0
        for _ in range(args.steps):
            _, _, out = cpu_baseline_run(pcms, fss, [dur] * cores, cores, kind, td)
            rows = sum(o[1] for o in out)
        wall = time.perf_counter() - t0
        port = None
        if kind == "reference":          # the port beside it, one step, for continuity with round 1
            v, w, _ = cpu_baseline_run(pcms, fss, [dur] * cores, cores, "port")
            port = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"the same {cores} drops, one step, {w:.1f} s wall"}
    value = args.steps * cores * dur / wall
    what = ("UNMODIFIED reference CLI (processAXCTD.main() under oracle/ref_shim.py: WAV read + run() + output file)"
            if kind == "reference" else "oracle/axctd_oracle.py (numpy port of the reference's path; baseline/_ref not installed)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.drops, args.duration,
                                      sum(int(round(sp.duration_s * sp.fs)) for sp in drop_specs(args.drops, args.duration, 0)), args.shards),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"bounded sample of the workload per step: {cores} drops of the batch's kind "
                                       f"(44.1/48 kHz alternating, SNR 40/25/10 dB) cut to {dur:.0f} s each, one process per "
                                       f"host core, {what}; rows written per step: {rows}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if port:
        line["cpu_baseline_port"] = port
    print(json.dumps(line))


# ---------------------------------------------------------------- the other BASELINE configs
def measure_configs(device, opts):
    """BASELINE configs 1, 2, 3 and 5 next to the headline (config 4), each through the public API with the recording
    in pinned HOST memory (upload, decode and result download inside the timed region), plus the device time of the
    decode alone.  Synthetic recordings are generated on the device and moved to the host before anything is timed."""
    import torch
    from axctdprocessor_b200 import engine, segment
    eng = engine.Engine(device)
    for k, v in opts.items():
        eng.set_option(k, v)
    out = {}

    def pinned_drop(spec):
        n = int(round(spec.duration_s * spec.fs))
        g = eng.batch([n], [eng.config(spec.fs if spec.fs <= 50000 else spec.fs / 2, decimate=2 if spec.fs > 50000 else 1)])
        g.synth_fill(0, spec)
        t = torch.empty(n, dtype=torch.int16).pin_memory()
        t.numpy()[:] = g.download(0)
        g.close()
        return t

    def med(xs):
        return float(np.median(xs))

    # configs 1 and 2: one 12-minute 44.1 kHz drop (40 dB, default low-pass / 10 dB, band-pass as -u documents it)
    for name, spec, st in (("config1_single_drop", synth.config_spec("config1"), None),
                           ("config2_single_drop_bandpass_10db", synth.config_spec("config2"), {"usebandpass": True})):
        host = pinned_drop(spec)
        cfg = eng.config(spec.fs, settings=st)
        b = eng.batch([len(host)], [cfg])
        e2e, dev, rows = [], [], 0
        for rep in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            b.upload_ptr(0, host.data_ptr(), len(host))
            b.run()
            r = b.result(0, full=False)
            dt = time.perf_counter() - t0
            if rep >= 3:
                e2e.append(1e3 * dt); dev.append(b.timing())
            rows = int(r.summary.n_rows)
        ph = {k: med([d[k] for d in dev]) for k in dev[0]}
        out[name] = {"e2e_ms": med(e2e), "device_ms": ph["total_ms"], "x_realtime_e2e": spec.duration_s / (1e-3 * med(e2e)),
                     "x_realtime_device": spec.duration_s / (1e-3 * ph["total_ms"]), "phases_ms": {k: round(v, 4) for k, v in ph.items()},
                     "status": int(r.summary.status), "frames": int(r.summary.n_frames), "rows": rows,
                     "h2d_bytes": 2 * len(host)}
        b.close()
        del host

    # config 3: 96 kHz, one hour, five drops back to back; cut by the segmentation driver, halved on the device
    specs = [synth.DropSpec(fs=96000, duration_s=720.0, seed=3300 + i, snr_db=(40.0, 25.0, 10.0)[i % 3]) for i in range(5)]
    parts = [pinned_drop(s) for s in specs]
    rec = torch.empty(sum(len(p) for p in parts), dtype=torch.int16).pin_memory()
    torch.cat(parts, out=rec)
    del parts
    t = []
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        segs = segment.process_recording(eng, rec.numpy(), 48000.0, decimate=2)
        t.append(time.perf_counter() - t0)
    out["config3_96k_1h_recording"] = {"e2e_ms": 1e3 * med(t[1:]), "x_realtime_e2e": 3600.0 / med(t[1:]), "segments": len(segs),
                                       "frames": [int(r.summary.n_frames) for _, _, r in segs],
                                       "status": [int(r.summary.status) for _, _, r in segs], "h2d_bytes": 2 * len(rec)}
    del rec

    # config 5: parameter sweep over a 4-hour low-SNR archive (20 drops x 720 s at 44.1 kHz, 8 dB): the archive is
    # uploaded and cut once, every parameter point decodes every segment (segments x points drops, filled on the device)
    specs = [synth.DropSpec(fs=44100, duration_s=720.0, seed=5500 + i, snr_db=8.0) for i in range(20)]
    parts = [pinned_drop(s) for s in specs]
    rec = torch.empty(sum(len(p) for p in parts), dtype=torch.int16).pin_memory()
    torch.cat(parts, out=rec)
    del parts
    points = [{"refreshrate": rr, "mark_space_freqs": ms, "deadfreq": df}
              for rr in (0.5, 1.0, 2.0, 4.0, 8.0) for ms in ([400.0, 800.0], [405.0, 795.0]) for df in (3000.0, 2500.0)]
    t, frames = [], 0
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scan, (p, r400, _) = segment.scan_batch(eng, rec.numpy(), 44100.0)
        cuts = segment.find_drops(p, r400, 44100.0, len(rec))
        frames = 0
        wave = 5                                         # parameter points per batch
        for w0 in range(0, len(points), wave):
            pts = points[w0:w0 + wave]
            cfgs = [eng.config(44100.0, settings=pt) for pt in pts for _ in cuts]
            b = eng.batch([hi - lo for _ in pts for lo, hi in cuts], cfgs)
            for q in range(len(pts)):
                for i, (lo, hi) in enumerate(cuts):
                    b.copy_from(q * len(cuts) + i, scan, 0, lo, hi - lo)
            b.run()
            frames += sum(int(b.summary(i).n_frames) for i in range(b.n))
            b.close()
        scan.close()
        t.append(time.perf_counter() - t0)
    audio = len(points) * len(rec) / 44100.0
    out["config5_sweep_4h_archive"] = {"e2e_ms": 1e3 * t[-1], "x_realtime_e2e": audio / t[-1], "points": len(points),
                                       "segments": len(cuts), "drop_decodes": len(points) * len(cuts), "frames": frames,
                                       "audio_seconds_decoded": audio, "h2d_bytes": 2 * len(rec),
                                       "note": "chunk 0.5/1/2/4/8 x fs, mark/space (400,800)/(405,795), dead 3000/2500 Hz"}
    # live receiver: one 12-minute drop arriving a second at a time through stream.StreamingDecoder (axctd_batch_stream_*):
    # every poll uploads the new second only and decodes the iterations that became complete
    from axctdprocessor_b200.stream import StreamingDecoder
    spec = synth.config_spec("config1")
    host = pinned_drop(spec).numpy()
    sd = StreamingDecoder(spec.fs, engine=eng, max_seconds=spec.duration_s + 10.0, norm_seconds=2.0)
    wall = []
    for a in range(0, len(host), spec.fs):
        t0 = time.perf_counter()
        sd.push(host[a:a + spec.fs])
        sd.poll()
        wall.append(1e3 * (time.perf_counter() - t0))
    r = sd.finish()
    runs = sd.runs
    out["streaming_720s_drop_1s_polls"] = {
        "polls": len(runs), "status": int(r.summary.status), "frames": int(r.summary.n_frames),
        "device_ms_per_poll_first_minutes": med([x["device_ms"] for x in runs[60:120]]),
        "device_ms_per_poll_last_minute": med([x["device_ms"] for x in runs[-61:-1]]),
        "filter_ms_per_poll_first_minutes": med([x["filter_ms"] for x in runs[60:120]]),
        "filter_ms_per_poll_last_minute": med([x["filter_ms"] for x in runs[-61:-1]]),
        "wall_ms_per_push_and_poll": med(wall[60:]), "h2d_bytes_per_poll": 2 * spec.fs,
        "note": "per poll: H2D of the new second, tone sums / filter / windows over the new samples, edges and bits of the new iterations; "
                "the rows handed out are final (tests hold every poll to the oracle's per-iteration lists)"}
    sd._own = False
    sd.close()
    eng.close()
    return out


_REAL_STDOUT = None


def emit_line(line):
    """The one JSON line, on the process's real stdout (see run_native: under torchrun descriptor 1 is handed to stderr)."""
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, text.encode())


# ---------------------------------------------------------------- this engine
def run_native(args):
    import torch
    import torch.distributed as dist
    from axctdprocessor_b200 import engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL prints its banner ("NCCL version ...") on file descriptor 1 from C code (NCCL_DEBUG_FILE does not move it): from
        # here on descriptor 1 is stderr, and the JSON line goes to the saved real stdout -- stdout carries that line only
        sys.stdout.flush()
        global _REAL_STDOUT
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from axctdprocessor_b200 import batch as axbatch
    numa_bound = axbatch.bind_host_thread_to_gpu(local) if world > 1 else False      # (one rank per GPU: keep its pinned buffers local)
    opts = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in args.opt}
    specs = drop_specs(args.drops, args.duration, rank)
    n = [int(round(s.duration_s * s.fs)) for s in specs]
    # The GPU's share of the drops is decoded as --shards sub-batches in flight at once (batch.ConcurrentDecoder:
    # one engine, CUDA stream and host thread each); they are timed together on a master stream.
    master = torch.cuda.Stream()
    streams = [torch.cuda.Stream() for _ in range(max(1, min(args.shards, len(specs))))]
    b = axbatch.ConcurrentDecoder(local, n, [s.fs for s in specs], shards=len(streams), engine_options=opts,
                                  streams=[st.cuda_stream for st in streams])
    for i, s in enumerate(specs):
        b.synth_fill(i, s)
    audio_s = sum(s.duration_s for s in specs)
    total_samples = sum(n)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ("value")
    b.run(steps=args.warmup)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = b.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    done = [torch.cuda.Event() for _ in streams]
    ev0.record(master)
    for st in streams:
        st.wait_event(ev0)
    timings = b.run(steps=args.steps, after=lambda k: done[k].record(streams[k]))
    for d in done:
        master.wait_event(d)
    ev1.record(master)
    barrier()
    # per step: the sub-batches' demodulation passes run one after the other (heavy_chain), so their durations add
    filt_ms = [sum(t[q]["filter_ms"] for t in timings) for q in range(args.steps)]
    tone_ms = [sum(t[q]["tone_ms"] for t in timings) for q in range(args.steps)]
    launches = b.launch_count - l0
    ms = reduce_max(ev0.elapsed_time(ev1))
    clocks = sampler.finish()
    ms_per_step = ms / args.steps
    value = world * audio_s / (ms_per_step / 1e3)

    # parity guard on the timed data: every drop decoded, nothing flagged
    stats = [b.summary(i) for i in range(len(specs))]
    bad = [int(s.status) for s in stats if s.status != 0]
    rows = int(sum(s.n_rows for s in stats)); frames = int(sum(s.n_frames for s in stats))
    crossings = int(sum(s.n_crossings for s in stats))

    # ---- roofline of the dominant kernel (continuous filter / crossing / bit-DFT pass)
    peak, peak_src = measured_peak_gbs()
    # 2 B per input sample + outputs (one byte per bit, 48 B per frame).  The kernel's own output
    # is larger (12 B per crossing: index i32, |S1| f32, |S2| f32 -- intermediate records); that figure is kept beside it.
    bits_total = int(sum(s.n_bits for s in stats))
    alg_bytes = 2.0 * total_samples + 1.0 * bits_total + 48.0 * frames
    alg_bytes_kernel = 2.0 * total_samples + 12.0 * crossings
    f_ms = float(np.mean(filt_ms))
    achieved = alg_bytes / (f_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "filter_traffic.json")
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_sample") * total_samples if tj.get("dram_bytes_per_sample") else None
        except Exception:
            traffic = None

    # ---- end to end through the public batch API with HOST buffers ("e2e")
    pool_n = min(args.host_pool, len(specs))
    pinned = []
    for i in range(pool_n):
        t = torch.empty(n[i], dtype=torch.int16).pin_memory()
        t.numpy()[:] = b.download(i)
        pinned.append(t)
    # drops i and i+2k share a length (same rate): cycle the pinned pool over matching lengths
    src_idx = [(i % pool_n) if n[i % pool_n] == n[i] else (i % 2) % pool_n for i in range(len(specs))]
    src = [pinned[j] for j in src_idx]
    h2d_bytes = int(sum(2 * x for x in n))

    # The batch goes through the package's ingest pipeline (batch.PipelinedDecoder: two engines / streams) in
    # --e2e-parts parts, so that the host->device copy of one part overlaps the decode of the previous one.  The
    # timed region starts with nothing in flight and ends when the last part's results are on the host.
    nparts = max(1, min(args.e2e_parts, len(specs)))
    parts = [list(range(k, len(specs), nparts)) for k in range(nparts)]
    pipe = axbatch.PipelinedDecoder(local, slots=2, engine_options=opts)

    def submit(h):
        pipe.submit([src[i].data_ptr() for i in h], [n[i] for i in h], [specs[i].fs for i in h])

    def e2e_run(steps):
        out, nfr = 0, 0
        order = [p for _ in range(steps) for p in parts]
        submit(order[0])
        for q in range(len(order)):
            if q + 1 < len(order):
                submit(order[q + 1])
            for r in pipe.collect(full=False):
                out += r.rows.nbytes + r.chunks.nbytes + 2048
                nfr += int(r.summary.n_frames)
        return out // steps, nfr // steps

    d2h_bytes, e2e_frames = e2e_run(1)
    barrier()
    # what the box can move host -> device with every rank copying at once (same pinned buffers, copies only):
    # the ceiling of the end-to-end figure, which carries 2 bytes per sample over these links
    probe_dev = torch.empty(max(n), dtype=torch.int16, device="cuda")
    def h2d_only():
        for i in range(len(specs)):
            probe_dev[:n[i]].copy_(src[i], non_blocking=True)
        torch.cuda.synchronize()
    h2d_only()
    barrier()
    t0 = time.perf_counter()
    h2d_only()
    h2d_s = reduce_max(time.perf_counter() - t0)
    h2d_ceiling_gbs = world * h2d_bytes / h2d_s / 1e9
    del probe_dev
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.e2e_steps)
    torch.cuda.synchronize()
    barrier()
    e2e_s = reduce_max((time.perf_counter() - t0) / args.e2e_steps)
    e2e_value = world * audio_s / e2e_s
    e2e_gbs = world * h2d_bytes / e2e_s / 1e9
    # parity guard on the end-to-end leg: every pooled recording decodes to the frames it gave device-resident
    assert e2e_frames == sum(int(stats[j].n_frames) for j in src_idx), (e2e_frames, frames)
    pipe.close()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.drops, args.duration, total_samples, len(streams)),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                    "steps": args.e2e_steps, "parts": nparts, "host_thread_bound_to_gpu_numa_node": bool(numa_bound),
                    "h2d_gbs": e2e_gbs, "h2d_ceiling_gbs": h2d_ceiling_gbs, "fraction_of_h2d_ceiling": e2e_gbs / h2d_ceiling_gbs,
                    "h2d_ceiling_note": "aggregate host->device rate of this box with all ranks copying the same pinned buffers and nothing else, measured in this run (tools/h2d_probe.py, profiles/r2_run2_h2d_probe_n*.json: 55 / 115 / 186 GB/s at 1 / 4 / 8 GPUs on the 32-vCPU single-NUMA guest, whatever the memory type, pool size, streams in flight or CPU binding)",
                    "host_pool_drops": pool_n, "note": "pinned host PCM -> axctd_batch_upload -> run -> compact rows to host, wall clock from an idle pipeline to the last result; the batch goes through batch.PipelinedDecoder in parts so that H2D overlaps decode"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_demod_fused (int16 -> SOS IIR f64 -> zero crossings -> mark/space windows f32)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": "profiles/filter_traffic.json: dram__bytes_read + dram__bytes_write of k_demod_fused per sample from one ncu --set full capture, scaled to this run's sample count (static, not re-measured here)",
                         "peak_source": peak_src, "kernel_ms": f_ms, "tone_kernels_ms": float(np.mean(tone_ms)),
                         "algorithmic_bytes": alg_bytes,
                         "algorithmic_bytes_definition": "SURVEY 8(d): 2 B x samples + 1 B x bits + 48 B x frames",
                         "achieved_incl_crossing_records": alg_bytes_kernel / (f_ms * 1e-3) / 1e9,
                         "frac_incl_crossing_records": alg_bytes_kernel / (f_ms * 1e-3) / 1e9 / peak, "note": "issue-bound: 7 FP64-pipe ops (2.2 issue cycles each on B200) + 4 IDP.2A + ~20 other instructions per sample; HBM is not the binding unit"},
            "decoded": {"frames": frames, "rows": rows, "drops_not_ok": bad}}
    if rank == 0 and world == 1 and not args.no_cpu:
        import tempfile
        cores = host_cores()
        k = min(cores, len(specs), args.cpu_procs or cores)
        kind = "reference" if reference_available() and not args.cpu_port else "port"
        dur = min(args.duration, args.cpu_duration if args.cpu_duration > 0 else (180.0 if kind == "reference" else 720.0))
        pcms = [b.download(i)[:int(round(dur * specs[i].fs))] for i in range(k)]
        fss = [specs[i].fs for i in range(k)]
        with tempfile.TemporaryDirectory() as td:
            if kind == "reference":
                import ref_shim
                ref_shim.load_modules()
                write_wavs(pcms, fss, td)
            v, wall, _ = cpu_baseline_run(pcms, fss, [dur] * k, k, kind, td)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": k, "kind": kind,
                                "sample": f"the first {dur:.0f} s of {k} of the batch's drops, one process per core, "
                                          + ("UNMODIFIED reference CLI (baseline/_ref under oracle/ref_shim.py)" if kind == "reference"
                                             else "oracle/axctd_oracle.py (numpy port)") + f", {wall:.1f} s wall"}
    b.close()
    if rank == 0 and world == 1 and not args.no_configs:
        torch.cuda.empty_cache()
        line["configs"] = measure_configs(local, opts)
    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--drops", type=int, default=128, help="drops per GPU (config 4: 1024 drops over 8 GPUs)")
    ap.add_argument("--duration", type=float, default=720.0)
    ap.add_argument("--host-pool", type=int, default=32, help="distinct pinned host drops cycled by the e2e leg (32 = 2.1 GB per rank, well beyond the host's 60 MB L3)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-parts", type=int, default=4, help="parts the batch is cut into for the ingest pipeline")
    ap.add_argument("--cpu-procs", type=int, default=0)
    ap.add_argument("--cpu-duration", type=float, default=0.0, help="seconds per drop in the CPU legs (0: 120 for the real reference, 720 for the port)")
    ap.add_argument("--cpu-port", action="store_true", help="time the oracle port even when baseline/_ref is installed")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the side measurements of BASELINE configs 1, 2, 3 and 5")
    ap.add_argument("--shards", type=int, default=4, help="sub-batches in flight per GPU (batch.ConcurrentDecoder)")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (A/B experiments)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
