"""N > 1 host logic on CPU: the batch of independent textures is partitioned over ranks with no data-path
collective; a gloo all_gather of per-texture checksums proves every texture is produced exactly once and that the
sharded result equals the single-process result.  (The per-texture generator here is the CPU oracle -- the test
exercises the sharding/rendezvous logic of vk_compute_mipmaps_b200.batch, not the CUDA kernels.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_partition():
    from vk_compute_mipmaps_b200 import batch
    for n in (0, 1, 7, 512):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += batch.shard_indices(n, r, world)
            assert sorted(seen) == list(range(n))
            counts = batch.shard_counts(n, world)
            assert sum(counts) == n and max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        batch.shard_indices(4, 2, 2)


def test_fnv_checksum_is_order_sensitive():
    from vk_compute_mipmaps_b200 import batch
    a = np.arange(100000, dtype=np.uint32).view(np.uint8)
    b = a.copy()
    b[[5, 70000]] = b[[70000, 5]]
    assert batch.fnv1a64(a) != batch.fnv1a64(b)
    assert batch.fnv1a64(a) == batch.fnv1a64(a.copy())
    # block boundary handling: same bytes, different lengths
    assert batch.fnv1a64(a[:65536]) != batch.fnv1a64(a[:65537])


def _worker(rank, world, port, n_tex, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import _oracle
    from vk_compute_mipmaps_b200 import batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = _oracle.load_oracle()
    w = h = 32
    mine = batch.shard_indices(n_tex, rank, world)
    sums = torch.zeros(n_tex, dtype=torch.int64)
    for k in mine:
        chain, _ = o.shader_chain(_oracle.random_level0(w, h, 1000 + k), w, h)
        sums[k] = np.int64(np.uint64(batch.fnv1a64(chain)).view(np.int64))
    owner = torch.full((n_tex,), -1, dtype=torch.int64)
    owner[mine] = rank
    dist.barrier()
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)    # every texture has exactly one non-zero contribution
    dist.all_reduce(owner, op=dist.ReduceOp.MAX)
    if rank == 0:
        out_q.put((sums.tolist(), owner.tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_batch_matches_single_process():
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    from vk_compute_mipmaps_b200 import batch
    _oracle.build_oracle()
    n_tex, world = 9, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_tex, q)) for r in range(world)]
    [p.start() for p in procs]
    sums, owner = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert owner == [k % world for k in range(n_tex)]
    o = _oracle.load_oracle()
    for k in range(n_tex):
        chain, _ = o.shader_chain(_oracle.random_level0(32, 32, 1000 + k), 32, 32)
        assert sums[k] == int(np.uint64(batch.fnv1a64(chain)).view(np.int64))


def test_device_checksum_equals_numpy_evaluation_and_is_order_sensitive():
    """The checksum bench.py uses to compare sharded batch runs with the single-GPU run (torch int64 arithmetic,
    wraps mod 2^64) against an independent numpy uint64 evaluation."""
    import torch
    from vk_compute_mipmaps_b200 import batch
    rng = np.random.default_rng(3)
    for n in (4, 4 * 1000, 4 * ((1 << 22) + 17)):
        a = rng.integers(0, 256, n, dtype=np.uint8)
        w = a.view("<u4").astype(np.uint64)
        idx = np.arange(w.size, dtype=np.uint64)
        with np.errstate(over="ignore"):
            mult = (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x2545F4914F6CDD1D)) | np.uint64(1)
            want = int(((w + np.uint64(1)) * mult).sum(dtype=np.uint64))
        assert batch.device_checksum(torch.from_numpy(a)) == want
    b = a.copy()
    b[[5, 70000]] = b[[70000, 5]]
    assert batch.device_checksum(torch.from_numpy(b)) != want
    assert batch.fold_checksums([1, 2, 3]) != batch.fold_checksums([3, 2, 1])
    with pytest.raises(ValueError):
        batch.device_checksum(torch.zeros(5, dtype=torch.uint8))
