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
    n = np.asarray(read_lengths, dtype=np.int64)
    step = chunksize - overlap
    short = n < chunksize
    stub = np.where(short, 0, (n - overlap) % step)
    regular = np.where(short, 0, (n - stub - chunksize) // step + 1)      # windows of signal[stub:].unfold(...)
    count = np.where(short, 1, regular + (stub > 0))
    first = np.concatenate([[0], np.cumsum(count)[:-1]])
    total = int(count.sum())
    chunk_read = np.repeat(np.arange(len(n), dtype=np.int64), count)
    k = np.arange(total, dtype=np.int64) - first[chunk_read]              # index of the chunk within its read
    has_stub = (stub > 0)[chunk_read]
    start = stub[chunk_read] + (k - has_stub) * step                      # k-th regular window (after the stub chunk)
    start = np.where(has_stub & (k == 0), 0, start)                       # leading stub chunk = signal[:chunksize]
    start = np.where(short[chunk_read], n[chunk_read] - chunksize, start) # short read: left pad (negative start)
    return {'chunk_read': chunk_read, 'chunk_start': start.astype(np.int64), 'chunk_first': first.astype(np.int64),
            'chunk_count': count.astype(np.int64)}


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


class ReadSetBasecaller:
    """Basecall a whole read set on one GPU without per-chunk host work (BASELINE config 4, one rank of it).

    model: xna_basecaller_b200.crf.Model on a CUDA device.  basecall(signals) takes a list of 1-D float32 (or int16)
    numpy arrays and returns (list of base strings in input order, counters).  Reads are uploaded once (pinned
    staging, one H2D), chunk batches are cut on the device (xb_gather_chunks: the windows of util.chunk), every batch
    runs the fused encoder + CRF decode into one (n_chunks, T) int8 row buffer, the rows are stitched on the device
    (xb_stitch: util.stitch + to_str semantics) and only the stitched letters and lengths come back."""

    def __init__(self, model, chunksize=4000, overlap=500, batchsize=512):
        self.model, self.chunksize, self.overlap, self.batchsize = model, chunksize, overlap, batchsize
        self.stride = model.stride
        self.device = next(model.parameters()).device
        self._pinned = {}                               # grow-only pinned staging buffers, reused across calls

    def _staging(self, n, dtype):
        buf = self._pinned.get(dtype)
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pinned[dtype] = buf
        return buf[:n]

    def basecall(self, signals, scaling=None, offset=None):
        """signals: normalised reads (float32, or int16 already in model units) -- or, with scaling / offset given (one per
        read: channel range / digitisation and channel offset, fast5.py:66,64), RAW int16 DAC reads, which are scaled,
        trimmed and med/MAD-normalised on the GPU first (xb_preprocess_reads = the reference's Read.__init__)."""
        import time
        dev, cs, ov, T = self.device, self.chunksize, self.overlap, self.chunksize // self.stride
        n_reads = len(signals)
        if n_reads == 0:
            return [], {'reads': 0, 'samples': 0, 'chunks': 0, 'seconds': 0.0, 'seconds_stage_h2d': 0.0,
                        'seconds_gpu_batches': 0.0, 'seconds_stitch_d2h': 0.0, 'seconds_strings': 0.0}
        lengths = np.fromiter((len(s) for s in signals), dtype=np.int64, count=n_reads)
        raw_lengths = lengths
        offsets = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
        dtype = torch.int16 if signals[0].dtype == np.int16 else torch.float32
        if scaling is not None and dtype != torch.int16:
            raise ValueError('raw reads (scaling / offset given) must be int16 DAC samples')
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        host = self._staging(int(lengths.sum()), dtype)
        np.concatenate(signals, out=host.numpy())          # one pass into pinned memory, then a single H2D
        sig = host.to(dev, non_blocking=True)
        eng = self.model.seqdist.engine
        if scaling is not None:
            h0 = eng.get(dev, self.batchsize, T, bf16=next(self.model.parameters()).dtype == torch.bfloat16)
            sig, out_len, _ = h0.preprocess(sig, offsets, lengths, np.asarray(scaling, dtype=np.float64),
                                            np.asarray(offset, dtype=np.int32))
            lengths = out_len.cpu().numpy().astype(np.int64)       # trimmed lengths decide the chunk table
            if int(lengths.min()) <= 0:
                raise ValueError('read %d has no samples left after trimming' % int(lengths.argmin()))
        plan = plan_chunks(lengths, cs, ov)
        n_chunks = len(plan['chunk_read'])
        as_dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
        read_offset, read_len = as_dev(offsets, torch.int64), as_dev(lengths, torch.int32)
        chunk_read, chunk_start = as_dev(plan['chunk_read'], torch.int32), as_dev(plan['chunk_start'], torch.int32)

        t1 = time.perf_counter()
        h = eng.get(dev, min(self.batchsize, n_chunks), T, bf16=next(self.model.parameters()).dtype == torch.bfloat16)
        self.model.encoder.sync_weights(h)
        rows = torch.empty(n_chunks, T, dtype=torch.int8, device=dev)
        batch = torch.empty(min(self.batchsize, n_chunks), cs, dtype=torch.float32, device=dev)
        for lo in range(0, n_chunks, self.batchsize):
            hi = min(lo + self.batchsize, n_chunks)
            x = h.gather_chunks(sig, read_offset, read_len, chunk_read[lo:hi], chunk_start[lo:hi], cs, out=batch[:hi - lo])
            scores = h.encoder(x)
            seq, _, _ = h.decode(scores, want_qstring=False)
            rows[lo:hi] = seq
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        out_stride = int(plan['chunk_count'].max()) * T
        out, out_len = h.stitch(rows, plan['chunk_first'], plan['chunk_count'], lengths, cs, ov, self.stride, out_stride)
        out_host, len_host = out.cpu().numpy(), out_len.cpu().numpy()
        seconds = time.perf_counter() - t0
        flat = out_host.view('u1')
        strings = [flat[i, :len_host[i]].tobytes().decode('ascii') for i in range(n_reads)]
        return strings, {'reads': n_reads, 'samples': int(raw_lengths.sum()), 'chunks': n_chunks, 'seconds': seconds,
                         'seconds_stage_h2d': t1 - t0, 'seconds_gpu_batches': t2 - t1, 'seconds_stitch_d2h': seconds - (t2 - t0),
                         'seconds_strings': time.perf_counter() - t0 - seconds}


def basecall_reads(model, reads, chunksize=4000, overlap=500, batchsize=512, raw=False):
    """The reference's basecall() contract (crf/basecall.py:96-119) on top of the device-side read-set pipeline: consumes an
    iterable of read objects (`.read_id`, `.signal`; with raw=True `.signal` holds int16 DAC samples and `.scaling` /
    `.offset` the channel calibration, fast5.py:64-66) and yields (read, {'sequence', 'qstring', 'sig_move'}) in input order
    -- the pairs xna_basecaller_b200.io.Writer turns into FASTQ records and summary rows."""
    reads = list(reads)
    caller = ReadSetBasecaller(model, chunksize, overlap, batchsize)
    if raw:
        strings, _ = caller.basecall([np.asarray(r.signal) for r in reads], scaling=[r.scaling for r in reads],
                                     offset=[r.offset for r in reads])
    else:
        strings, _ = caller.basecall([np.asarray(r.signal) for r in reads])
    for read, seq in zip(reads, strings):
        yield read, {'sequence': seq, 'qstring': 'O' * len(seq),
                     'sig_move': np.zeros(len(read.signal) // caller.stride * caller.stride, dtype=bool)}


def basecall_sharded(model, signals, chunksize=4000, overlap=500, batchsize=512, rank=0, world=1):
    """One rank's share of a read set (reads r with r mod world == rank) + the gathered per-rank counters."""
    mine = shard_reads(len(signals), rank, world)
    strings, counters = ReadSetBasecaller(model, chunksize, overlap, batchsize).basecall([signals[i] for i in mine])
    table = gather_counters(counters, device=next(model.parameters()).device if world > 1 else None)
    return dict(zip(mine.tolist(), strings)), table
