"""Read-sharded multi-GPU basecalling driver (BASELINE config 4).

Reads are independent: no LSTM state crosses chunks (bonito/util.py:152-166) and a read's chunks only meet
again in stitch (util.py:169-188).  So G GPUs run G independent pipelines, one process per GPU: rank g takes
reads {r : r mod G == g}, keeps all chunks of a read on its GPU, and the only collective is one all_gather of
a few counters at the end (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Host side here: the sharding plan, the chunk table of a read set (same windows as util.chunk) and the counter
gather.  Device side (ReadSetBasecaller): whole reads are uploaded once, chunk batches are cut on the GPU by
index (xb_gather_chunks), the fused encoder + decode runs per batch, rows are stitched on the GPU (xb_stitch) and
only the stitched base strings return to the host.
"""
import numpy as np
import torch

from .util import chunk_starts


def shard_reads(n_reads, rank, world):
    """Indices of the reads rank `rank` of `world` basecalls (round robin: balances read-length drift)."""
    return np.arange(rank, n_reads, world, dtype=np.int64)


def plan_chunks(read_lengths, chunksize, overlap):
    """Chunk table of a read set, in read order.

    Returns dict of int64 arrays: chunk_read (owning read), chunk_start (first sample within the read; negative
    for a short read = number of left-pad zeros), chunk_first / chunk_count (per read: its rows in the table).
    Windows are exactly those of util.chunk (bonito/util.py:152-166)."""
    chunk_read, chunk_start, first, count = [], [], [], []
    for r, n in enumerate(read_lengths):
        n = int(n)
        first.append(len(chunk_read))
        if n < chunksize:
            starts = [n - chunksize]                    # left pad with zeros (util.py:160)
        else:
            starts = chunk_starts(n, chunksize, overlap)
        chunk_read += [r] * len(starts)
        chunk_start += starts
        count.append(len(starts))
    as64 = lambda v: np.asarray(v, dtype=np.int64)
    return {'chunk_read': as64(chunk_read), 'chunk_start': as64(chunk_start), 'chunk_first': as64(first),
            'chunk_count': as64(count)}


def gather_counters(counters, device=None):
    """all_gather a small dict of numbers over the default process group; returns {name: [v_rank0, ...]}.
    With no process group initialised (single GPU) every list has one entry."""
    import torch.distributed as dist
    names = sorted(counters)
    mine = torch.tensor([float(counters[k]) for k in names], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(out, mine)
    else:
        out = [mine]
    table = torch.stack(out).cpu().numpy()
    return {k: table[:, i].tolist() for i, k in enumerate(names)}
