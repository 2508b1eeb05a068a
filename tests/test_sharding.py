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
