#!/usr/bin/env python3
"""bench.py -- STFT frames/s of the fused STFT->dB->palette path (BASELINE.json metric) on N GPUs of one node.

Workload (config.workload = "cfg2-batch"): BASELINE configs[1] geometry -- stereo 48 kHz, FFT 2048, hop 512, Hann,
AbsMean mix, dB, Jade/256 palette over -50..+50 dB, one ARGB32 pixel per bin (1025 rows) -- applied to a batch of
independent synthetic streams (noise*0.1 + sine sweep).  A "step" is `passes_per_step` passes of the hot path over that batch
(>= 100 ms of kernels per step, so that the timed region lasts seconds and the clock record covers it).

  value   : frames/s with the inputs already resident in HBM (one main kernel launch + one for the boundary columns per
            pass, CUDA events, max over ranks)
  e2e     : frames/s through the reference-facing C ABI call jade_render_batch with pinned HOST buffers
            (H2D of the samples and D2H of the pixel columns inside the timed region)
  latency : per-block time of the real-time path (512-sample blocks, push + fetch through the C ABI, from a native C++
            caller -- tools/native/latency_probe -- with the Python loop's numbers beside it)
  roofline: algorithmic bytes (4*hop*C + 4*R per frame) / kernel time against the measured HBM copy bandwidth; `secondary`
            holds the on-chip limits beside it (FP32-pipe cycles and shared-memory wavefronts per frame from the committed
            ncu profile x the measured frame rate / (SMs x SM clock)) -- SURVEY 8d: "report both"
  cpu_baseline: the CPU oracle port timed on this box's host cores on a bounded sample

`--impl reference` times the reference's CPU implementation of the same path (oracle port / compiled reference).
"""
import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FS = 48000.0
N_FFT = 2048
HOP = 512
CHANNELS = 2
ROWS = N_FFT // 2 + 1
BYTES_ALG = 4 * HOP * CHANNELS + 4 * ROWS  # 8196 B per frame (SURVEY 8d / BASELINE.md section 3)
STREAM_SECONDS = 20.0
STREAMS_PER_GPU = 256
WORKLOAD = dict(workload="cfg2-batch", sample_rate=48000, fft_size=N_FFT, hop=HOP, channels=CHANNELS, window="hann",
                mix="absmean", palette="jade256", range_db=[-50, 50], rows=ROWS, pixel="ARGB32")


def workload_config(world):
    """`config` of the JSON line: the workload only, identical for both arms (--impl ours / reference)."""
    return dict(WORKLOAD, streams_per_gpu=STREAMS_PER_GPU, seconds_per_stream=STREAM_SECONDS,
                l2_policy="inputs+outputs of one pass (2 x 1.97 GB per GPU) exceed L2 (126 MB) many times; no explicit flush",
                parallelism=f"streams sharded over {world} GPU(s), no collective")


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_runner():
    """(callable(nstreams, threads) -> (frames/s, frames), kind, description).  Prefers the reference's own classes
    compiled in place (oracle/_ref/libjade_ref.so: Spectrogram.cpp + CColorpalette.cpp against the stub JUCE/TGM headers,
    FFT = the oracle's stand-in for the absent TGM `spectrum`); falls back to the restated oracle port."""
    import ctypes as C

    import oracle_lib as O
    import signals
    x = signals.streams(1, CHANNELS, int(FS * 2.0), FS, kind="mix")[0]
    flat = x.reshape(-1)
    if O.have_ref_spec():
        R = O.ref()

        def run(nstreams, threads):
            frames = C.c_long(0)
            fps = R.jr_bench_batch(FS, N_FFT, O.FEED["p25"], O.WIN["hann"], CHANNELS, O.PAL["jade"], 256, -50.0, 50.0, flat,
                                   x.shape[1], nstreams, threads, C.byref(frames))
            return fps, frames.value
        return run, "reference", ("the reference's own Spectrogram::processSynchronBlock + getMem + CColorPalette::getRGBColor "
                                  "(Spectrogram.cpp / CColorpalette.cpp compiled in place; FFT = oracle stand-in for the "
                                  "absent TGM spectrum class), one instance per host thread, perc25 feed = hop 512")
    cfg = O.BatchCfg(FS, N_FFT, HOP, O.WIN["hann"], CHANNELS, O.MIX["absmean"], O.PAL["jade"], 256, 0, -50.0, 50.0, 0)

    def run(nstreams, threads):
        frames = C.c_long(0)
        fps = O.lib().jo_bench_batch(C.byref(cfg), flat, x.shape[1], nstreams, threads, C.byref(frames))
        return fps, frames.value
    return run, "port", "oracle port (restated Spectrogram.cpp loop, radix-2 FFT stand-in, restated palette)"


def cpu_baseline(sample_seconds=12.0, threads=None):
    """The reference's CPU path on the host cores, on a bounded sample of the bench workload."""
    threads = threads or os.cpu_count() or 1
    run, kind, what = _cpu_runner()
    t0 = time.time()
    fps1, per_stream = run(1, 1)  # calibrate on one thread, then size the sample for ~sample_seconds on all threads
    cal = time.time() - t0
    nstreams = max(threads, int(sample_seconds * fps1 * threads / max(per_stream, 1)))
    nstreams = (nstreams + threads - 1) // threads * threads
    fps, frames = run(nstreams, threads)
    return dict(value=fps, unit="frames/s", cores=threads, kind=kind,
                sample=f"{nstreams} streams x 2.0 s of the cfg2-batch workload ({frames} frames); {what}; "
                       f"1-thread rate {fps1:.0f} frames/s (calibration {cal:.1f} s)",
                single_thread_value=fps1)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    threads = os.cpu_count() or 1
    run, kind, what = _cpu_runner()
    # one step = 4 streams per host thread x 2 s (bounded sample of the workload)
    per_step_streams = 4 * threads
    frames = 0
    for _ in range(warm):
        run(per_step_streams, threads)
    t0 = time.time()
    total = 0
    for _ in range(steps):
        _, frames = run(per_step_streams, threads)
        total += frames
    dt = time.time() - t0
    fps = total / dt
    sample = (f"bounded sample of the workload: each step = {per_step_streams} streams x 2.0 s of cfg2-batch ({frames} frames) on "
              f"{threads} host threads (FFT plan cached per thread); {what}")
    line = dict(impl="reference", metric="stft_frames_per_sec", value=fps, unit="frames/s", n_gpus=args.gpus, steps=steps,
                warmup=warm, ms_per_step=dt / steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", config=workload_config(args.gpus),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=threads, kind=kind, sample=sample),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import signals
    from jadespectrogram_b200 import Engine, host_alloc, host_free, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # pinned host buffers of the e2e leg on the GPU's own NUMA node (no-op on single-node boxes)
    orig_affinity = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    props = torch.cuda.get_device_properties(local)
    numa_node = sharding.bind_host_to_gpu_node(getattr(props, "pci_domain_id", 0), getattr(props, "pci_bus_id", 0),
                                               getattr(props, "pci_device_id", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    steps, warm = args.steps, max(args.warmup, 3)
    S = args.streams
    nsamp = int(FS * STREAM_SECONDS)
    eng = Engine(local, sample_rate=FS, fft_size=N_FFT, hop=HOP, channels=CHANNELS, window="hann", mix_mode="absmean",
                 max_push=512)
    eng.set_palette_scheme("jade", 256)
    eng.set_value_range(-50.0, 50.0)
    ncols = eng.columns_for(nsamp)
    frames_step = S * ncols

    # ---- inputs resident in HBM (generated on the device), outputs in HBM
    d_in = torch.empty((S, CHANNELS, nsamp), dtype=torch.float32, device=dev)
    d_pix = torch.empty((S, ncols, ROWS), dtype=torch.int32, device=dev)
    # a dedicated (non-default) stream: kernels, timing events and the engine all use this one stream
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0
    eng.synth_device(d_in.data_ptr(), S, CHANNELS, nsamp, CHANNELS * nsamp, nsamp, kind="mix", seed=20240601 + rank,
                     cuda_stream=stream)
    torch.cuda.synchronize()

    def one_pass():
        eng.render_device(d_in.data_ptr(), S, nsamp, CHANNELS * nsamp, nsamp, 0, ncols, d_pix.data_ptr(), None, stream)

    # passes per step: >= 100 ms of kernels per step (calibrated once, the same on every rank)
    if args.passes > 0:
        passes = args.passes
    else:
        for _ in range(3):
            one_pass()
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(4):
            one_pass()
        c1.record()
        torch.cuda.synchronize()
        passes = max(1, int(-(-100.0 // (c0.elapsed_time(c1) / 4))))
        passes = int(round(sharding.reduce_max(float(passes), dev)))
    frames_pass = frames_step
    frames_step = frames_pass * passes

    def step():
        for _ in range(passes):
            one_pass()

    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = eng.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    launches = eng.kernel_launches - l0
    total_ms = ev[0].elapsed_time(ev[steps])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    clocks = sampler.stop() if rank == 0 else None  # sampled during the timed region only
    max_ms = sharding.reduce_max(total_ms, dev)  # the slowest rank defines the job time
    value = frames_step * steps * world / (max_ms * 1e-3)
    kernel_ms = statistics.mean(per_launch_ms) / passes  # one pass = the main kernel + the launch for the boundary columns

    # ---- 1-GPU == N-GPU evidence on separate devices: rank 0 renders stream 0 of every other rank (same seed, same engine
    # configuration) and compares a checksum of the pixel columns with the one that rank computed itself
    shard_parity = None
    if world > 1:
        own = d_pix[0].to(torch.int64).sum().reshape(1)
        sums = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(sums, own)
        if rank == 0:
            shard_parity = True
            d_one = torch.empty((1, CHANNELS, nsamp), dtype=torch.float32, device=dev)
            d_onepix = torch.empty((1, ncols, ROWS), dtype=torch.int32, device=dev)
            for r in range(1, world):
                eng.synth_device(d_one.data_ptr(), 1, CHANNELS, nsamp, CHANNELS * nsamp, nsamp, kind="mix", seed=20240601 + r,
                                 cuda_stream=stream)
                eng.render_device(d_one.data_ptr(), 1, nsamp, CHANNELS * nsamp, nsamp, 0, ncols, d_onepix.data_ptr(), None, stream)
                torch.cuda.synchronize()
                shard_parity = shard_parity and int(d_onepix[0].to(torch.int64).sum().item()) == int(sums[r].item())

    if args.only_kernel:  # short command for ncu captures: no e2e / latency / CPU legs
        if rank == 0:
            emit(dict(metric="stft_frames_per_sec", value=value, unit="frames/s", n_gpus=world, steps=steps,
                      kernel_ms=kernel_ms, gpu_launches=int(launches), note="--only-kernel", passes_per_step=passes,
                      config=dict(kernel=eng.kernel_name)))
        eng.close()
        return

    # ---- e2e: host buffers through jade_render_batch (H2D + kernel + D2H inside the timed region)
    Se = min(S, args.e2e_streams)
    while True:  # pinned host memory is a shared resource of the box (one rank per GPU): halve the batch if it is short
        h_in = h_pix = None
        try:
            h_in = host_alloc((Se, CHANNELS, nsamp), np.float32)
            h_pix = host_alloc((Se, ncols, ROWS), np.uint32)
            break
        except Exception:
            if h_in is not None:
                host_free(h_in)
            if Se <= 8:
                raise
            Se //= 2
    h_in[:] = d_in[:Se].cpu().numpy()
    e2e_steps = max(2, min(steps, 6))
    for _ in range(2):
        eng.render_batch(h_in, out_pix=h_pix)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.render_batch(h_in, out_pix=h_pix)
    e2e_s = time.perf_counter() - t0
    e2e_value = Se * ncols * e2e_steps * world / sharding.reduce_max(e2e_s, dev)
    if numa_node is not None and orig_affinity is not None:
        os.sched_setaffinity(0, orig_affinity)  # the CPU baseline leg below uses every host core
    check = int(h_pix[0, ncols // 2].astype(np.uint64).sum())  # the step's result is read on the host
    e2e = dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=int(h_in.nbytes), d2h_bytes_per_step=int(h_pix.nbytes),
               steps=e2e_steps, streams_per_step=Se, api="jade_render_batch (pinned host buffers)", host_numa_node=numa_node, checksum=check)
    host_free(h_in)
    host_free(h_pix)

    # ---- real-time path latency: 512-sample stereo blocks, push + fetch through the C ABI
    lat = None
    if rank == 0:
        eng.reset()
        blk = signals.streams(1, CHANNELS, 512 * 64, FS, kind="mix")[0]
        ts = []
        nblk = args.latency_blocks
        for b in range(nblk + 200):
            piece = blk[:, (b % 64) * 512:(b % 64 + 1) * 512]
            t0 = time.perf_counter()
            eng.push(piece)
            p, _, _ = eng.fetch(max_cols=4, want_db=False)
            t1 = time.perf_counter()
            if b >= 200:
                ts.append((t1 - t0) * 1e6)
        ts.sort()
        lat = dict(p50_us=ts[len(ts) // 2], p99_us=ts[int(len(ts) * 0.99)], blocks=nblk, block_samples=512,
                   path="jade_push_samples + jade_fetch_columns (1 column per block), host timer around both calls",
                   caller="python (ctypes wrappers)")
        # The plugin's host is C++: the same loop from a native caller of the C ABI (tools/native/latency_probe.cpp) is the
        # headline; the Python numbers above stay as `python_*` (ctypes / numpy add a few microseconds per block).
        probe = ROOT / "tools" / "native" / "latency_probe"
        if probe.exists():
            try:
                out = subprocess.run([str(probe), str(nblk), str(local)], capture_output=True, text=True, timeout=120)
                nat = json.loads(out.stdout.strip().splitlines()[-1])
                lat = dict(p50_us=nat["p50_us"], p99_us=nat["p99_us"], push_p50_us=nat["push_p50_us"],
                           fetch_p50_us=nat["fetch_p50_us"], blocks=nat["blocks"], block_samples=512, path=lat["path"],
                           caller="native C++ (tools/native/latency_probe.cpp)", python_p50_us=lat["p50_us"],
                           python_p99_us=lat["p99_us"])
            except Exception as ex:  # keep the Python measurement
                lat["native_probe_error"] = str(ex)[:200]

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = BYTES_ALG * frames_pass / (kernel_ms * 1e-3) / 1e9
        traffic, secondary = None, None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                prof = json.loads(tp.read_text())
                traffic = prof.get("cfg2-batch")
                per = prof.get("cfg2-batch-per-frame")
                if per:
                    # on-chip limits beside the HBM one: cycles per frame from the committed ncu profile of this kernel, turned
                    # into busy fractions with the frame rate and SM clock measured in THIS run
                    clk = (clocks or {}).get("sm_mhz") or 1965.0
                    fps_gpu = frames_pass / (kernel_ms * 1e-3)
                    sm_cycles = props.multi_processor_count * clk * 1e6
                    secondary = dict(fp32_pipe_frac=per["fp32_pipe_cycles"] * fps_gpu / sm_cycles,
                                     smem_pipe_frac=per["smem_wavefronts"] * fps_gpu / sm_cycles,
                                     issue_frac=per["warp_instructions"] / 4.0 * fps_gpu / sm_cycles,
                                     per_frame=per, sm_mhz_used=clk,
                                     note="per-frame counts from profiles/ (ncu), rate and clock from this run")
            except Exception:
                traffic = None
        cpu = cpu_baseline(args.cpu_seconds) if not args.no_cpu else None
        line = dict(
            metric="stft_frames_per_sec", value=value, unit="frames/s", n_gpus=world, steps=steps, warmup=warm,
            ms_per_step=max_ms / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
            data="synthetic (0.1*white noise + linear sine sweep per stream, generated on the device)",
            config=workload_config(world), kernel=eng.kernel_name,
            run=dict(passes_per_step=passes, frames_per_pass_per_gpu=frames_pass, frames_per_step_per_gpu=frames_step,
                     input_bytes_per_pass_per_gpu=int(d_in.numel() * 4), output_bytes_per_pass_per_gpu=int(d_pix.numel() * 4),
                     timed_region_s=max_ms * 1e-3),
            e2e=e2e, gpu_launches=int(launches), clocks=clocks, latency=lat, shard_parity=shard_parity,
            roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic,
                          traffic_unit="DRAM bytes per launch, from the committed ncu capture of this kernel (profiles/traffic.json), "
                                       "not measured in this run",
                          algorithmic_bytes_per_launch=BYTES_ALG * frames_pass,
                          peak_source=peak_src, bytes_per_frame=BYTES_ALG, frames_per_launch=frames_pass,
                          kernel_ms=kernel_ms, secondary=secondary),
            cpu_baseline=cpu)
        emit(line)
    if world > 1:
        # The other ranks are done after the e2e reduction.  They wait for rank 0's latency and CPU legs on the HOST (a key in
        # the rendezvous store), not in an NCCL barrier, which would keep their GPUs spinning for ~20 s.
        try:
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                store.set("jade_bench_done", "1")
            else:
                store.wait(["jade_bench_done"])
        except Exception:
            pass
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU per pass")
    ap.add_argument("--passes", type=int, default=0, help="passes over the batch per step (0: calibrate to >= 100 ms per step)")
    ap.add_argument("--e2e-streams", type=int, default=128)
    ap.add_argument("--latency-blocks", type=int, default=2000)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only-kernel", action="store_true", help="device-resident loop only (for ncu)")
    ap.add_argument("--geometry", default=None, help="experiments only: FFT,HOP,CHANNELS instead of the cfg2-batch workload")
    args = ap.parse_args()
    if args.geometry:
        global N_FFT, HOP, CHANNELS, ROWS, BYTES_ALG, WORKLOAD
        N_FFT, HOP, CHANNELS = (int(v) for v in args.geometry.split(","))
        ROWS = N_FFT // 2 + 1
        BYTES_ALG = 4 * HOP * CHANNELS + 4 * ROWS
        WORKLOAD = dict(WORKLOAD, workload=f"custom-{N_FFT}-{HOP}-{CHANNELS}", fft_size=N_FFT, hop=HOP, channels=CHANNELS, rows=ROWS)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
