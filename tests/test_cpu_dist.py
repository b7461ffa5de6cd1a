"""world_size-2 gloo test of the multi-GPU host logic: reads shard by index with no data-path
collective; the only exchange is one all_gather of per-rank counters."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from xna_basecaller_b200 import pipeline
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lengths = np.random.RandomState(11).randint(4000, 20000, size=101)
    mine = pipeline.shard_reads(len(lengths), rank, world)
    plan = pipeline.plan_chunks(lengths[mine], 4000, 500)
    counters = {'reads': len(mine), 'samples': int(lengths[mine].sum()), 'chunks': len(plan['chunk_read']),
                'seconds': 1.0 + rank}
    table = pipeline.gather_counters(counters)
    # dynamic work queue over the group's store: the two ranks together take every block exactly once
    store = dist.distributed_c10d._get_default_store()
    took = list(pipeline.WorkQueue(37, store=store, key='wq_test'))
    table['took'] = pipeline.gather_counters({'n': len(took), 'sum': sum(took), 'sumsq': sum(k * k for k in took)})
    if rank == 0:
        torch.save(table, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_counter_gather(tmp_path):
    out = str(tmp_path / 'table.pt')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    table = torch.load(out)
    lengths = np.random.RandomState(11).randint(4000, 20000, size=101)
    assert table['reads'] == [51.0, 50.0]
    assert sum(table['samples']) == float(lengths.sum())
    assert table['seconds'] == [1.0, 2.0]
    assert sum(table['chunks']) > 101
    took = table['took']
    assert sum(took['n']) == 37 and sum(took['sum']) == sum(range(37)) and sum(took['sumsq']) == sum(k * k for k in range(37))


def test_single_process_counter_gather():
    from xna_basecaller_b200 import pipeline
    assert pipeline.gather_counters({'a': 3, 'b': 0.5}) == {'a': [3.0], 'b': [0.5]}
    assert list(pipeline.WorkQueue(5)) == [0, 1, 2, 3, 4]
