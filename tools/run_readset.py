"""BASELINE config 4 (scaled): synthetic read set sharded by read across the ranks of one box.

    python tools/run_readset.py [n_reads] [--raw]               # 1 GPU; --raw: int16 DAC reads, pre-processed on the GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_readset.py [n_reads]

Reads: length ~ N(10000, 1000) clipped to [4000, 20000] (seed 11), x ~ N(0,1); chunksize 4000, overlap 500, batch 512.
Each rank basecalls reads r mod G == rank; one all_gather of counters at the end; samples/s = sum of samples / max
seconds over ranks (upload, chunking, encoder, decode, stitch, D2H of the strings all inside the timed region).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bonito_oracle as bo                 # weight generator only
from xna_basecaller_b200 import pipeline
from xna_basecaller_b200.crf import Model


def sup_config(alphabet):
    return {'global_norm': {'state_len': 3}, 'input': {'features': 1}, 'labels': {'labels': list(alphabet)},
            'model': {'package': 'xna_basecaller_b200.crf'},
            'encoder': {'stride': 5, 'activation': 'swish', 'features': 768, 'winlen': 19, 'scale': 5.0,
                        'rnn_type': 'lstm', 'blank_score': 2.0}}


RAW = '--raw' in sys.argv
args = [a for a in sys.argv[1:] if not a.startswith('--')]
n_reads = int(args[0]) if args else 4000
rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rs = np.random.RandomState(11)
lengths = np.clip(rs.normal(10000, 1000, size=n_reads), 4000, 20000).astype(np.int64)
model = Model(sup_config('NACGTX'))
model.load_state_dict(bo.reference_state_dict(n_base=5, seed=25))
model = model.half().eval().to('cuda:%d' % local)
mine = pipeline.shard_reads(n_reads, rank, world)
g = np.random.RandomState(1000 + rank)
caller = pipeline.ReadSetBasecaller(model, 4000, 500, 512)
if RAW:        # int16 DAC samples (2 B/sample over PCIe) + per-read calibration; scaling, trim, med/MAD on the GPU
    signals = [np.clip(np.round(400 + 40 * g.standard_normal(int(lengths[i])) +
                                np.repeat(25 * g.standard_normal(int(lengths[i]) // 8 + 1), 8)[:int(lengths[i])]),
                       -2000, 2000).astype(np.int16) for i in mine]
    kw = {'scaling': np.full(len(mine), 1437.976 / 8192), 'offset': g.randint(-300, 300, len(mine))}
else:
    signals = [g.standard_normal(int(lengths[i])).astype(np.float32) for i in mine]
    kw = {}
caller.basecall(signals, **kw)                          # warm-up (handle, weights, pinned staging buffers)
strings, counters = caller.basecall(signals, **kw)
table = pipeline.gather_counters(counters, device=torch.device('cuda', local) if world > 1 else None)
if rank == 0:
    total = sum(table['samples'])
    sec = max(table['seconds'])
    print(json.dumps({'config': 'configs[3] scaled: %d reads%s, cs 4000, ov 500, batch 512, sharded r mod G' % (n_reads, ' (raw int16 + GPU pre-processing)' if RAW else ''),
                      'n_gpus': world, 'samples': total, 'seconds_max_rank': sec, 'samples_per_s': total / sec,
                      'chunks': sum(table['chunks']), 'reads_per_rank': table['reads'],
                      'rank0_seconds': {k: v[0] for k, v in table.items() if k.startswith('seconds_')},
                      'mean_bases_per_read': float(np.mean([len(s) for s in strings]))}))
if world > 1:
    dist.destroy_process_group()
