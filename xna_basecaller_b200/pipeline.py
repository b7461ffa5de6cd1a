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
    """Basecall read sets on one GPU without per-chunk host work (BASELINE config 4, one rank of it).

    model: xna_basecaller_b200.crf.Model on a CUDA device.  A BLOCK of reads (a list of 1-D float32 or int16 numpy arrays)
    goes through three phases:
      stage   the reads are packed back to back into a pinned staging buffer and uploaded with one H2D copy on the copy
              stream (host work + PCIe; touches no kernel)
      launch  chunk batches are cut on the device (xb_gather_chunks: the windows of util.chunk), every batch runs the fused
              encoder + CRF decode (xb_basecall_chunks) into one (n_chunks, T) int8 row buffer, the rows are stitched on the
              device (xb_stitch: util.stitch + to_str semantics) and the stitched letters + lengths start their D2H copy
      finish  waits for that copy and cuts the Python strings
    basecall(signals) runs the three phases for one block.  basecall_stream(blocks) pipelines them over a stream of
    blocks: block i+1 is staged (helper thread, copy stream) while block i computes, and block i's strings are cut while
    block i+1 computes -- copies and host work hide under the kernels, results come out lazily and in order."""

    def __init__(self, model, chunksize=4000, overlap=500, batchsize=512, reverse=False):
        self.model, self.chunksize, self.overlap, self.batchsize = model, chunksize, overlap, batchsize
        self.reverse = reverse                          # `bonito basecaller --reverse`: scores reverse-complemented, reverse stitch
        self.stride = model.stride
        self.device = next(model.parameters()).device
        self._pinned = {}                               # grow-only pinned staging buffers, reused across calls
        self._copy_stream = None
        self._h2d_done = {}

    def _staging(self, key, n, dtype):
        buf = self._pinned.get((key, dtype))
        if buf is None or buf.numel() < n:
            buf = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pinned[(key, dtype)] = buf
        return buf[:n]

    def _handle(self, N):
        eng = self.model.seqdist.engine
        h = eng.get(self.device, N, self.chunksize // self.stride,
                    bf16=next(self.model.parameters()).dtype == torch.bfloat16)
        self.model.encoder.sync_weights(h)
        return h

    # ------------------------------------------------------------------ phase 1: host staging + H2D (copy stream)
    def stage(self, signals, scaling=None, offset=None, slot=0):
        import time
        t0 = time.perf_counter()
        # stage() runs on a helper thread in basecall_stream, and the CUDA current device is per-thread state that starts at
        # device 0: without this every rank > 0 opens a context on GPU 0 and routes its pinned-memory and event calls
        # through it (measured on 8 GPUs: 106 ms per block instead of 8 ms, serialised with rank 0's kernels)
        torch.cuda.set_device(self.device)
        n_reads = len(signals)
        blk = {'n_reads': n_reads, 'scaling': scaling, 'offset': offset, 'slot': slot, 't0': t0}
        if n_reads == 0:
            return blk
        lengths = np.fromiter((len(s) for s in signals), dtype=np.int64, count=n_reads)
        dtype = torch.int16 if signals[0].dtype == np.int16 else torch.float32
        if scaling is not None and dtype != torch.int16:
            raise ValueError('raw reads (scaling / offset given) must be int16 DAC samples')
        prev = self._h2d_done.get(slot)
        if prev is not None:
            prev.synchronize()                              # the slot's previous upload has left the pinned buffer
        host = self._staging('in%d' % slot, int(lengths.sum()), dtype)
        np.concatenate(signals, out=host.numpy())          # one pass into pinned memory, then a single H2D
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            sig = host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._h2d_done[slot] = ev
        blk.update(lengths=lengths, offsets=np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64), sig=sig,
                   event=ev, seconds_stage=time.perf_counter() - t0)
        return blk

    # ------------------------------------------------------------------ phase 2: kernels + async D2H (current stream)
    def launch(self, blk):
        import time
        if blk['n_reads'] == 0:
            return blk
        t1 = time.perf_counter()
        dev, cs, ov, T = self.device, self.chunksize, self.overlap, self.chunksize // self.stride
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(blk['event'])
        sig, lengths, offsets = blk['sig'], blk['lengths'], blk['offsets']
        sig.record_stream(cur)
        if blk['scaling'] is not None:
            h0 = self._handle(self.batchsize)
            sig, out_len, _ = h0.preprocess(sig, offsets, lengths, np.asarray(blk['scaling'], dtype=np.float64),
                                            np.asarray(blk['offset'], dtype=np.int32))
            lengths = out_len.cpu().numpy().astype(np.int64)       # trimmed lengths decide the chunk table (one sync)
            if int(lengths.min()) <= 0:
                raise ValueError('read %d has no samples left after trimming' % int(lengths.argmin()))
        plan = plan_chunks(lengths, cs, ov)
        n_chunks = len(plan['chunk_read'])
        as_dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev, non_blocking=True)
        read_offset, read_len = as_dev(offsets, torch.int64), as_dev(lengths, torch.int32)
        chunk_read, chunk_start = as_dev(plan['chunk_read'], torch.int32), as_dev(plan['chunk_start'], torch.int32)
        h = self._handle(min(self.batchsize, n_chunks))
        rows = torch.empty(n_chunks, T, dtype=torch.int8, device=dev)
        batch = torch.empty(min(self.batchsize, n_chunks), cs, dtype=torch.float32, device=dev)
        for lo in range(0, n_chunks, self.batchsize):
            hi = min(lo + self.batchsize, n_chunks)
            x = h.gather_chunks(sig, read_offset, read_len, chunk_read[lo:hi], chunk_start[lo:hi], cs, out=batch[:hi - lo])
            if self.reverse:      # the score tensor is permuted between encoder and decode (CTC_CRF.reverse_complement)
                seq, _, _ = self.model.seqdist.decode_packed(self.model.seqdist.reverse_complement(h.encoder(x)))
                rows[lo:hi] = seq
            else:
                h.basecall_chunks(x, out=rows[lo:hi])
        out_stride = int(plan['chunk_count'].max()) * T
        out, out_len = h.stitch(rows, plan['chunk_first'], plan['chunk_count'], lengths, cs, ov, self.stride, out_stride,
                                reverse=self.reverse)
        out_host = self._staging('out%d' % blk['slot'], out.numel(), torch.int8).view(out.shape)
        len_host = self._staging('len%d' % blk['slot'], out_len.numel(), torch.int32)
        out_host.copy_(out, non_blocking=True)
        len_host.copy_(out_len, non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        blk.update(out_host=out_host, len_host=len_host, done=done, n_chunks=n_chunks, keep=(out, out_len, rows, batch),
                   seconds_launch=time.perf_counter() - t1)
        return blk

    # ------------------------------------------------------------------ phase 3: wait for the letters, cut the strings
    def finish(self, blk):
        import time
        n_reads = blk['n_reads']
        if n_reads == 0:
            return [], {'reads': 0, 'samples': 0, 'chunks': 0, 'seconds': 0.0, 'seconds_stage_h2d': 0.0,
                        'seconds_gpu_batches': 0.0, 'seconds_strings': 0.0}
        t2 = time.perf_counter()
        blk['done'].synchronize()
        t3 = time.perf_counter()
        flat, lens = blk['out_host'].numpy().view('u1'), blk['len_host'].numpy()
        strings = [flat[i, :lens[i]].tobytes().decode('ascii') for i in range(n_reads)]
        t4 = time.perf_counter()
        counters = {'reads': n_reads, 'samples': int(blk['lengths'].sum()), 'chunks': blk['n_chunks'],
                    'seconds': t4 - blk['t0'], 'seconds_stage_h2d': blk['seconds_stage'],
                    'seconds_gpu_batches': blk['seconds_launch'] + (t3 - t2), 'seconds_strings': t4 - t3}
        blk.pop('keep', None)
        return strings, counters

    def basecall(self, signals, scaling=None, offset=None):
        """One block.  signals: normalised reads (float32, or int16 already in model units) -- or, with scaling / offset given
        (one per read: channel range / digitisation and channel offset, fast5.py:66,64), RAW int16 DAC reads, which are scaled,
        trimmed and med/MAD-normalised on the GPU first (xb_preprocess_reads = the reference's Read.__init__).
        Returns (list of base strings in input order, counters)."""
        return self.finish(self.launch(self.stage(signals, scaling, offset)))

    def basecall_stream(self, blocks):
        """blocks: iterable of lists of reads, or of (signals, scaling, offset) tuples.  Yields (strings, counters) per block,
        in order, with the three phases of consecutive blocks overlapped (see the class docstring)."""
        from concurrent.futures import ThreadPoolExecutor

        def as_args(b):
            return b if isinstance(b, tuple) else (b, None, None)

        it = iter(blocks)
        with ThreadPoolExecutor(1) as pool:
            def stage_next(slot):
                try:
                    b = next(it)
                except StopIteration:
                    return None
                return pool.submit(self.stage, *as_args(b), slot=slot)

            i = 0
            staged = stage_next(0)
            in_flight = None
            while staged is not None:
                blk = staged.result()
                # the staging slot about to be refilled belongs to the block in flight two rounds ago: already finished
                nxt = stage_next((i + 1) & 1)
                launched = self.launch(blk)
                if in_flight is not None:
                    yield self.finish(in_flight)
                in_flight = launched
                staged = nxt
                i += 1
            if in_flight is not None:
                yield self.finish(in_flight)


def basecall_reads(model, reads, chunksize=4000, overlap=500, batchsize=512, raw=False, block_reads=2048, reverse=False):
    """The reference's basecall() contract (crf/basecall.py:96-119) on top of the device-side read-set pipeline: consumes an
    ITERATOR of read objects lazily, `block_reads` at a time (`.read_id`, `.signal`; with raw=True `.signal` holds int16 DAC
    samples and `.scaling` / `.offset` the channel calibration, fast5.py:64-66) and yields (read, {'sequence', 'qstring',
    'sig_move'}) in input order -- the pairs xna_basecaller_b200.io.Writer turns into FASTQ records and summary rows.
    While one block computes, the next one is pulled from the iterator and staged."""
    from itertools import islice
    caller = ReadSetBasecaller(model, chunksize, overlap, batchsize, reverse=reverse)
    it = iter(reads)
    pending = []                                        # blocks of read objects, in the order their results will arrive

    def blocks():
        while True:
            blk = list(islice(it, block_reads))
            if not blk:
                return
            pending.append(blk)
            sigs = [np.asarray(r.signal) for r in blk]
            yield (sigs, [r.scaling for r in blk], [r.offset for r in blk]) if raw else sigs

    for strings, _ in caller.basecall_stream(blocks()):
        for read, seq in zip(pending.pop(0), strings):
            yield read, {'sequence': seq, 'qstring': 'O' * len(seq),
                         'sig_move': np.zeros(len(read.signal) // caller.stride * caller.stride, dtype=bool)}


class WorkQueue:
    """Dynamic pull of work items 0..n-1 from ONE shared counter ("dynamic pull from one host work queue"):
    every rank takes the next unclaimed index when it is ready for more, so a rank that drew long reads or a slower GPU
    simply takes fewer items.  The counter lives in the process group's key-value store (the TCPStore torchrun creates;
    `store.add` is atomic); without a store (single process) it is a local counter.  No data-path collective."""

    def __init__(self, n_items, store=None, key='xb_work_queue'):
        self.n_items, self.store, self.key, self._local = int(n_items), store, key, 0

    def __iter__(self):
        while True:
            if self.store is not None:
                k = int(self.store.add(self.key, 1)) - 1
            else:
                k, self._local = self._local, self._local + 1
            if k >= self.n_items:
                return
            yield k


def basecall_sharded(model, signals, chunksize=4000, overlap=500, batchsize=512, rank=0, world=1):
    """One rank's share of a read set (reads r with r mod world == rank) + the gathered per-rank counters."""
    mine = shard_reads(len(signals), rank, world)
    strings, counters = ReadSetBasecaller(model, chunksize, overlap, batchsize).basecall([signals[i] for i in mine])
    table = gather_counters(counters, device=next(model.parameters()).device if world > 1 else None)
    return dict(zip(mine.tolist(), strings)), table
