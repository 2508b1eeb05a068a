"""N > 1 path on CPU: two gloo ranks shard a batch by image range (no data-path collective), each decodes its
shard, and the gathered result equals the single-process result image by image (determinism 1 vs N).
The per-rank decode uses the CPU simulation here; on the GPU box the same host logic drives the CUDA library
(tests/test_gpu_parity.py, bench.py --gpus N)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jpeg_rust_b200 import shard_range, synth


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 16, 1024, 16384):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, files, q):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sim_ffi as S
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(len(files), rank, world)
    rs, _ = S.decode_batch(files[lo:hi])
    digest = torch.tensor([int(np.sum(r.rgb.astype(np.int64) * 31 + 7)) for r in rs], dtype=torch.int64)
    pixels = torch.tensor([sum(r.width * r.height for r in rs)], dtype=torch.int64)
    # only bookkeeping is exchanged: pixel counts (for the throughput figure) and digests (for this test)
    dist.all_reduce(pixels)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, digest.tolist()))
    dist.barrier()
    if rank == 0:
        q.put((int(pixels.item()), gathered))
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_process_result():
    files = [synth.synth_jpeg(700 + i, 96 + 8 * i, 64, "420") for i in range(7)]
    import sim_ffi as S
    single, _ = S.decode_batch(files)
    want = [int(np.sum(r.rgb.astype(np.int64) * 31 + 7)) for r in single]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, files, q)) for r in range(2)]
    for p in procs:
        p.start()
    pixels, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = [None] * len(files)
    for lo, hi, dig in gathered:
        got[lo:hi] = dig
    assert got == want
    assert pixels == sum(r.width * r.height for r in single)


def test_partition_by_scan_bytes():
    """jpgpu_partition (host only): contiguous ranges that cover all images, balanced by scan bytes - the partition
    jpgpu_multi_plan gives every device."""
    import ctypes as C
    from jpeg_rust_b200 import _ffi
    L = _ffi.lib()
    rng = np.random.default_rng(3)
    for n in (0, 1, 5, 64, 1000):
        descs = (_ffi.ImageDesc * max(n, 1))()
        sizes = rng.integers(100, 100000, n)
        for i in range(n):
            descs[i].scan_len = int(sizes[i])
        for parts in (1, 2, 3, 8):
            first = (C.c_size_t * (parts + 1))()
            assert L.jpgpu_partition(descs, n, parts, first) == 0
            f = list(first)
            assert f[0] == 0 and f[-1] == n and all(a <= b for a, b in zip(f, f[1:]))
            if n >= 64:
                loads = [int(sizes[a:b].sum()) for a, b in zip(f, f[1:])]
                assert max(loads) - min(loads) <= 2 * int(sizes.max())


def test_multi_create_without_a_device_fails_cleanly():
    """No CPU fallback, no crash: without a usable sm_100 device the multi-device handle reports JPGPU_ERR_NO_DEVICE."""
    import ctypes as C
    from jpeg_rust_b200 import _ffi
    if torch.cuda.is_available():
        return
    h = C.c_void_p()
    devs = (C.c_int * 2)(0, 1)
    assert _ffi.lib().jpgpu_multi_create(devs, 2, C.byref(h)) == _ffi.ERR_NO_DEVICE and not h


def test_partition_keeps_the_scans_of_a_frame_together():
    """A frame of non-interleaved scans is three consecutive descriptors (frame_part 1, 2, 2): no range may start inside it."""
    import ctypes as C
    from jpeg_rust_b200 import _ffi
    n = 30
    descs = (_ffi.ImageDesc * n)()
    for i in range(n):
        descs[i].scan_len = 1000 + 37 * i
        descs[i].frame_part = (1, 2, 2)[i % 3]
    for parts in (2, 3, 4, 7):
        first = (C.c_size_t * (parts + 1))()
        assert _ffi.lib().jpgpu_partition(descs, n, parts, first) == 0
        assert first[0] == 0 and first[parts] == n
        assert all(f == n or descs[f].frame_part != 2 for f in first)
