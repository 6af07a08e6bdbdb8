"""Sharding of a batch of independent textures over ranks (one process per GPU).

A single mip chain is never sharded (levels are serially dependent, SURVEY.md
section 8e); a batch is: texture k belongs to rank ``k % world``.  No data-path
collective exists -- torch.distributed is used by callers only to agree on a
start barrier and to gather timings / checksums.
"""


def shard_indices(num_textures, rank, world_size):
    """Indices of the textures rank ``rank`` owns (round robin)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return list(range(rank, num_textures, world_size))


def shard_counts(num_textures, world_size):
    return [len(range(r, num_textures, world_size)) for r in range(world_size)]


def fnv1a64(data):
    """Order-sensitive checksum of a bytes-like object (numpy uint8 array ok).

    Computed with numpy in blocks; used to compare shards' outputs across ranks
    ("checksum of checksums") without moving the textures.
    """
    import numpy as np
    a = np.frombuffer(memoryview(data), dtype=np.uint8)
    # 64-bit polynomial hash, vectorised: sum(a[i] * P^(n-1-i)) mod 2^64
    P = np.uint64(1099511628211)
    h = np.uint64(14695981039346656037)
    block = 1 << 16
    pw = np.empty(block, dtype=np.uint64)
    acc = np.uint64(1)
    with np.errstate(over="ignore"):
        for i in range(block - 1, -1, -1):
            pw[i] = acc
            acc = acc * P
        pblock = acc  # P^block
        for s in range(0, a.size, block):
            c = a[s:s + block].astype(np.uint64)
            k = c.size
            if k == block:
                h = h * pblock + np.sum(c * pw, dtype=np.uint64)
            else:
                h = h * (P ** np.uint64(k)) + np.sum(c * pw[block - k:], dtype=np.uint64)
    return int(h)


_MIX_A = 0x9E3779B97F4A7C15 - (1 << 64)   # as signed 64-bit constants (torch int64 arithmetic wraps mod 2^64)
_MIX_B = 0x2545F4914F6CDD1D


def device_checksum(chain):
    """Order-sensitive 64-bit checksum of a packed chain that stays on the device it lives on.

    ``chain``: 1-D torch uint8 tensor whose length is a multiple of 4 (CPU or CUDA).  Word i (little-endian
    uint32) contributes (word + 1) * (i * A + B | 1) mod 2^64.  Used by the batch bench to compare the outputs of
    sharded runs (texture k on rank k % world) with the single-GPU run without moving 45 GB through the host.
    Returns a Python int in [0, 2^64).
    """
    import torch
    if chain.dtype != torch.uint8 or chain.dim() != 1 or chain.numel() % 4:
        raise ValueError("device_checksum expects a flat uint8 tensor of whole texels")
    words = chain.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    total = 0
    block = 1 << 22
    for s in range(0, words.numel(), block):  # blocks bound the temporaries (32 MB each)
        w = words[s:s + block]
        idx = torch.arange(s, s + w.numel(), dtype=torch.int64, device=chain.device)
        mult = (idx * _MIX_A + _MIX_B) | 1
        total = (total + int(((w + 1) * mult).sum().item())) & 0xFFFFFFFFFFFFFFFF
    return total


def fold_checksums(values):
    """Checksum of checksums, order-sensitive (values: iterable of ints in texture-index order)."""
    h = 14695981039346656037
    for v in values:
        h = ((h ^ (int(v) & 0xFFFFFFFFFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h
