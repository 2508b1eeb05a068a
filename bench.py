#!/usr/bin/env python
"""bench.py — decoded Mpixel/s of a synthetic 1080p 4:2:0 baseline-JPEG batch (BASELINE.json metric).

One process per GPU (torchrun for N > 1); the batch is sharded by image (weak scaling: every
GPU decodes `--images` images), no collective on the data path.  Prints ONE JSON line.

  value      whole-job Mpixel/s with the JPEG bitstreams already resident in HBM
             (timed: unstuff/RST pre-pass + Huffman decode + IDCT/colour, CUDA events)
  e2e        same metric through the C ABI with HOST buffers: H2D of every bitstream from
             pinned memory, decode, D2H of every RGB image into pinned memory
  roofline   the IDCT+colour kernel: algorithmic bytes (blocks*128 B in + W*H*3 B out) / its
             CUDA-event time, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline  the oracle (restatement of the reference decoder, O(N^4) cosf IDCT) on the host cores

--impl reference times the reference's own CPU algorithm (the oracle port; the Rust crate
cannot be built here) on the same workload, bounded to a sample that finishes in minutes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "decoded_mpixel_per_s_1080p_420_batch"
UNIT = "Mpixel/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--images", type=int, default=1024, help="images per GPU (BASELINE configs[2]: 1024 on 1 B200)")
    ap.add_argument("--distinct", type=int, default=256, help="distinct synthetic images cycled to fill the batch")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--subsampling", default="420")
    ap.add_argument("--quality", type=int, default=85)
    ap.add_argument("--restart-interval", type=int, default=0, help="MCUs per restart interval (0 = none; needs the DRI "
                    "extension the reference panics on: BASELINE configs[3], e.g. 3840x2160 444 with 480/16/1)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the CPU baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--waves", type=int, default=0, help="also time a job of WAVES x --images images decoded wave after wave "
                    "through the one batch object, every wave's scans and RGB output resident in HBM (BASELINE configs[4] "
                    "on one GPU: 16 waves of 1024)")
    ap.add_argument("--e2e-chunk", type=int, default=128, help="images per chunk of the pipelined end-to-end run")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_throughput(files, width, height, nthreads, sample):
    """The reference's algorithm (oracle port, cosf per term like transform.rs:79-81) on `nthreads` host threads,
    one image per thread at a time. Returns (Mpixel/s, seconds, images)."""
    import oracle_ffi as O
    from concurrent.futures import ThreadPoolExecutor
    O.lib()
    todo = [files[i % len(files)] for i in range(sample)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(nthreads) as ex:
        secs = list(ex.map(lambda f: O.time_decode(f, O.LAYOUT_REF, O.EXT_NONE, O.COS_CALL, 1), todo))
    dt = time.perf_counter() - t0
    if any(s < 0 for s in secs):
        raise RuntimeError("oracle failed on a benchmark image")
    return sample * width * height / dt / 1e6, dt, sample


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from jpeg_rust_b200 import synth
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or max(cores, 8)
    files = synth.synth_corpus(min(sample, args.distinct), args.width, args.height, args.subsampling, args.quality)
    vals = []
    for _ in range(args.warmup):
        cpu_reference_throughput(files[:cores], args.width, args.height, cores, min(cores, len(files)))
    t_total = 0.0
    for _ in range(args.steps):
        v, dt, n = cpu_reference_throughput(files, args.width, args.height, cores, sample)
        vals.append(v)
        t_total += dt
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic {args.width}x{args.height} {args.subsampling} q{args.quality} baseline JPEG; "
                               f"each step decodes a bounded sample of {sample} images on the host CPU",
                   "images_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images per step, one image per thread, oracle (C restatement of the "
                                   "reference: linear-search Huffman, O(N^4) f32 IDCT with cosf per term); the Rust "
                                   "crate itself cannot be built in this image (no rustc/cargo)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers of the
    end-to-end leg (first touch) and the threads feeding the copies are local to the GPU's PCIe root.  Best
    effort: returns the node, or None when the topology cannot be read (then nothing is changed)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from jpeg_rust_b200 import LAYOUT_SPEC, Batch, _ffi, context, parse_descriptor, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)   # before any pinned allocation: first touch places the pages
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- inputs: `distinct` synthetic images (different per rank), cycled to `images`, each copy with its own
    # host region and its own device buffers
    files = synth.synth_corpus(args.distinct, args.width, args.height, args.subsampling, args.quality,
                               restart_interval=args.restart_interval, first_index=rank * args.distinct)
    ext = _ffi.EXT_DRI if args.restart_interval else _ffi.EXT_NONE
    n = args.images
    sizes = [len(files[i % args.distinct]) for i in range(n)]
    offs = np.zeros(n + 1, np.int64)
    offs[1:] = np.cumsum([(s + 63) // 64 * 64 for s in sizes])
    host_in = torch.empty(int(offs[-1]), dtype=torch.uint8).pin_memory()
    hin = host_in.numpy()
    for i in range(n):
        hin[offs[i]:offs[i] + sizes[i]] = np.frombuffer(files[i % args.distinct], np.uint8)
    descs = (_ffi.ImageDesc * n)()
    for i in range(n):
        st, d, _ = parse_descriptor(hin[offs[i]:offs[i] + sizes[i]], ext, LAYOUT_SPEC)
        assert st == 0, st
        descs[i] = d
    ctx = context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    batch = Batch(descs=descs, device=local_rank, keepalive=host_in)
    stats = batch.stats()
    pixels = stats["pixels"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: bitstreams resident in HBM -> RGB resident in HBM
    batch.upload()
    for _ in range(args.warmup):
        batch.decode()
    statuses, _ = batch.results()
    bad = [s for s in statuses if s != 0]
    if bad:
        raise SystemExit(f"decode failed for {len(bad)} images, first status {bad[0]}")
    launches0 = batch.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)          # let nvidia-smi start sampling before the timed region
    barrier()
    t_wall0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        batch.decode()            # the call a user makes: image groups pipelined over two streams inside the library
    e1.record(stream)
    barrier()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = batch.launch_count() - launches0
    ms_per_step = total_ms / args.steps
    value = world * pixels / (ms_per_step * 1e-3) / 1e6
    # stage breakdown and the IDCT/colour kernel's own duration: the same work, stage after stage on one stream
    ent_ev, idct_ev = [], []
    for k in range(args.steps):
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record(stream)
        batch.entropy()
        b.record(stream)
        batch.idct()
        c.record(stream)
        ent_ev.append((a, b))
        idct_ev.append((b, c))
    barrier()
    idct_ms = sum(a.elapsed_time(b) for a, b in idct_ev) / args.steps
    ent_ms = sum(a.elapsed_time(b) for a, b in ent_ev) / args.steps
    # per-kernel device times (average of a few profiled decodes) and what each achieves against the HBM roofline
    prof = None
    for _ in range(3):
        pr = batch.profile()
        prof = pr if prof is None else {k: prof[k] + pr[k] for k in pr}
    prof = {k: v / 3 for k, v in prof.items()}
    peak, peak_src = measured_peak()
    idct_bytes = stats["coef_bytes"] + stats["rgb_bytes"]
    achieved = idct_bytes / (idct_ms * 1e-3) / 1e9
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel, per image
        t = json.load(open(os.path.join(ROOT, "profiles", "idct_traffic.json")))
        if (t["width"], t["height"], t["subsampling"]) == (args.width, args.height, args.subsampling):
            traffic = t["dram_bytes_per_image"] * n
    except Exception:
        pass
    kname = {"420": "idct_colour_kernel<2,2,false>", "422": "idct_colour_kernel<2,1,false>", "444": "idct_colour_kernel<1,1,false>",
             "440": "idct_colour_kernel<1,2,false>", "gray": "idct_colour_kernel<1,1,true>"}[args.subsampling]
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "ms_per_launch": idct_ms,
                "algorithmic_bytes_per_launch": idct_bytes,
                "kernels": {
                    "prepass_count+scan+write": {"ms": prof["prepass_count"] + prof["prepass_scan"] + prof["prepass_write"],
                                                 "algorithmic_GBps": 2 * stats["scan_bytes"] / ((prof["prepass_count"] + prof["prepass_scan"] + prof["prepass_write"]) * 1e-3) / 1e9,
                                                 "split_ms": [prof["prepass_count"], prof["prepass_scan"], prof["prepass_write"]],
                                                 "bound": "instruction issue and latency of the classify / scan / byte-scatter chain (two reads of the raw bytes, whole-sector stores)"},
                    "sync": {"ms": prof["sync"], "bitstream_GBps": stats["scan_bytes"] / (prof["sync"] * 1e-3) / 1e9,
                             "bound": "instruction issue + load latency of a serial bit-dependent decode (~45 instructions per symbol step, ~20 of 32 lanes active)"},
                    "verify_scan": {"ms": prof["verify_scan"], "bound": "latency of the longest repair walk"},
                    "decode_write": {"ms": prof["decode_write"],
                                     "algorithmic_GBps": (stats["scan_bytes"] + stats["coef_bytes"]) / (prof["decode_write"] * 1e-3) / 1e9,
                                     "frac_of_hbm_peak": (stats["scan_bytes"] + stats["coef_bytes"]) / (prof["decode_write"] * 1e-3) / 1e9 / peak,
                                     "bound": "instruction issue (~55 instructions per symbol step plus the cooperative block flush, ~19 of 32 lanes active), not HBM"},
                    "idct_colour": {"ms": prof["idct_colour"]}},
                "entropy_stage": {"ms": ent_ms, "bitstream_GBps": stats["scan_bytes"] / (ent_ms * 1e-3) / 1e9,
                                  "algorithmic_GBps": (stats["scan_bytes"] + stats["coef_bytes"]) / (ent_ms * 1e-3) / 1e9}}

    # ---- BASELINE configs[4] on this GPU: a job larger than one set of arenas, wave after wave through the same batch
    # object.  Every wave has its own scan bytes (one resident device buffer per job) and its
    # own slice of one resident output arena; bitstream and coefficient arenas are the batch's, reused.
    waves = None
    if args.waves > 1:
        import ctypes as C
        K = args.waves
        wave_out = batch.output_bytes()
        out_arena = torch.empty(K * wave_out + 256, dtype=torch.uint8, device="cuda")
        out_base = (out_arena.data_ptr() + 255) // 256 * 256
        scan_stride = (int(offs[-1]) + 255) // 256 * 256
        scans = torch.empty(K * scan_stride, dtype=torch.uint8, device="cuda")
        for w in range(K):
            scans[w * scan_stride:w * scan_stride + int(offs[-1])].copy_(host_in, non_blocking=True)
        scan_offs = [int(descs[i].scan) - hin.ctypes.data for i in range(n)]   # where image i's entropy-coded bytes start
        torch.cuda.synchronize()

        def run_job():
            for w in range(K):
                batch.set_device_scans(scans.data_ptr() + w * scan_stride, scan_offs)
                batch.set_device_output(out_base + w * wave_out, wave_out)
                batch.decode()
        run_job()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record(stream)
        run_job()
        w1.record(stream)
        barrier()
        job_ms = max_over_ranks(w0.elapsed_time(w1))
        first = torch.empty(wave_out, dtype=torch.uint8, device="cuda")
        first.copy_(out_arena[out_base - out_arena.data_ptr():out_base - out_arena.data_ptr() + wave_out])
        last = out_arena[out_base - out_arena.data_ptr() + (K - 1) * wave_out:out_base - out_arena.data_ptr() + K * wave_out]
        assert torch.equal(first, last), "waves of identical input differ"
        batch.set_device_output(None, 0)
        batch.upload()   # the batch's own scans again
        batch.decode()
        waves = {"images": K * n, "waves": K, "value": world * K * pixels / (job_ms * 1e-3) / 1e6, "unit": UNIT, "ms": job_ms,
                 "resident_GB": {"scans": K * scan_stride / 1e9, "rgb": K * wave_out / 1e9},
                 "note": "per wave: device-to-device hand-over of the wave's scan bytes, decode into the wave's slice of one "
                         "resident output arena; arenas of one batch object reused"}
        del out_arena, scans

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region.  The batch is cut into chunks that
    # alternate between two contexts (= two streams) so that the H2D of one chunk, the kernels of another and the D2H
    # of a third overlap; the plans and device arenas of the chunks are created once, outside the timed region.
    e2e = None
    if not args.no_e2e:
        from jpeg_rust_b200 import Context
        import ctypes as C
        per = args.width * args.height * 3
        host_out = torch.empty(n * per, dtype=torch.uint8).pin_memory()
        chunk = max(1, min(args.e2e_chunk, n))
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        ctxs = [Context(local_rank), Context(local_rank)]
        for cx, st in zip(ctxs, streams):
            cx.set_stream(st.cuda_stream)
        chunks = []
        for k, i0 in enumerate(range(0, n, chunk)):
            m = min(chunk, n - i0)
            sub = (_ffi.ImageDesc * m).from_address(C.addressof(descs) + i0 * C.sizeof(_ffi.ImageDesc))
            cb = Batch(descs=sub, device=local_rank, keepalive=(host_in, descs), ctx=ctxs[k % 2])
            chunks.append((cb, [host_out.data_ptr() + (i0 + i) * per for i in range(m)]))

        def e2e_step():
            for cb, ptrs in chunks:
                cb.upload()
                cb.decode()
                cb.download_ptrs(ptrs)

        e2e_step()
        torch.cuda.synchronize()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ksteps = max(1, min(args.steps, 3))
        streams[1].wait_stream(streams[0])
        s0.record(streams[0])
        streams[1].wait_event(s0)
        for _ in range(ksteps):
            e2e_step()
        streams[0].wait_stream(streams[1])
        s1.record(streams[0])
        barrier()
        e2e_ms = max_over_ranks(s0.elapsed_time(s1)) / ksteps
        for cb, _ in chunks:
            statuses, _ = cb.results()
            assert all(s == 0 for s in statuses)
        # keep the result honest: the host copy of the last image equals what the resident-input run left on the device
        chk = np.empty((args.height, args.width, 3), np.uint8)
        batch.download_ptrs([0] * (n - 1) + [chk.ctypes.data])
        batch.ctx.sync()
        assert np.array_equal(chk.reshape(-1), host_out[(n - 1) * per:n * per].numpy()), "e2e output differs from the device result"
        e2e = {"value": world * pixels / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(sum(sizes)),
               "d2h_bytes_per_step": int(n * per), "ms_per_step": e2e_ms,
               "note": f"jpgpu_batch_upload + decode + download of {len(chunks)} chunks of {chunk} images alternating "
                       "between two contexts/streams (copies overlap kernels); pinned host memory; plans reused",
               "numa_node": numa}
        for cb, _ in chunks:
            cb.close()
        for cx in ctxs:
            cx.close()

    if not sampler.lines:   # very short runs: keep the GPU busy with the same work until nvidia-smi has reported
        t_end = time.perf_counter() + 1.0
        while not sampler.lines and time.perf_counter() < t_end:
            batch.decode()
            torch.cuda.synchronize()
    clocks = sampler.stop()

    # ---- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or 2 * cores
        v, dt, cnt = cpu_reference_throughput(files, args.width, args.height, cores, sample)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cnt} of the batch's images, one per thread, {dt:.1f} s wall; oracle = C restatement of the "
                         "reference decoder (not the Rust binary: no rustc in the image)"}

    # ---- BASELINE configs[0]: one lena.jpeg through the single-image entry point (JPEGImage.parse = parse + H2D +
    # decode + D2H, synchronous), next to the reference's algorithm on one host core
    config0 = None
    lena_path = os.path.join(ROOT, "tests", "golden", "fixtures", "lena.jpeg")
    if rank == 0 and os.path.exists(lena_path):
        from jpeg_rust_b200 import JPEGImage, LAYOUT_REF
        lena = open(lena_path, "rb").read()
        for _ in range(3):
            JPEGImage.parse(lena, layout=LAYOUT_REF, device=local_rank)
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            JPEGImage.parse(lena, layout=LAYOUT_REF, device=local_rank)
        gpu_ms = (time.perf_counter() - t0) / reps * 1e3
        config0 = {"workload": "lena.jpeg 512x512 4:2:2, one image per call, host bytes in, host RGB out, REF layout",
                   "gpu_ms_per_image": gpu_ms, "gpu_mpixel_per_s": 512 * 512 / gpu_ms / 1e3}
        if world == 1 and not args.no_cpu_baseline:
            import oracle_ffi as O
            cpu_s = O.time_decode(lena, O.LAYOUT_REF, O.EXT_NONE, O.COS_CALL, 1)
            config0["cpu_ms_per_image_1_core"] = cpu_s * 1e3

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (IDCT/colour), i16/u8 (entropy)", "data": "synthetic",
            "config": {"workload": f"{n} synthetic {args.width}x{args.height} {args.subsampling} q{args.quality} baseline "
                                   f"JPEGs per GPU ({args.distinct} distinct, own buffers per copy), SPEC layout"
                                   + (f", restart interval {args.restart_interval} MCUs" if args.restart_interval else ""),
                       "images_per_gpu": n, "l2": "inputs larger than L2 (no flush needed): "
                       f"{stats['scan_bytes'] / 1e6:.0f} MB bitstream, {stats['coef_bytes'] / 1e9:.2f} GB coefficients",
                       "parallelism": f"images sharded over {world} GPU(s), no collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "config0_single_image": config0, "config4_waves": waves,
            "stage_ms": {"entropy": ent_ms, "idct_colour": idct_ms, "note": "stages run one after the other on one stream; "
                         "`value` is timed over jpgpu_batch_decode, which overlaps the image groups of a batch"},
        }
        print(json.dumps(line), flush=True)
    batch.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
