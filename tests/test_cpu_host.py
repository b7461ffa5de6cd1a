"""CPU tests of the host-side mirror of the reference interface (no GPU): util chunk / stitch / batchify against
the reference golden vectors, the layer registry and state_dict surface, the plugin selector, the sharding plan,
and that the product path fails loudly without a CUDA device (no fallback)."""
import os

import numpy as np
import pytest
import torch

from make_golden import ALPHABETS
from oracle import bonito_oracle as bo
from xna_basecaller_b200 import nn, util, pipeline
from xna_basecaller_b200.crf import Model, basecall
from xna_basecaller_b200.crf import model as crf_model
from xna_basecaller_b200.crf.basecall import apply_stride_to_moves, stitch_results, to_str

from test_cpu_oracle import _stitch_cases


def sup_config(alphabet):
    return {'global_norm': {'state_len': 3}, 'input': {'features': 1}, 'labels': {'labels': list(alphabet)},
            'model': {'package': 'xna_basecaller_b200.crf'},
            'encoder': {'stride': 5, 'activation': 'swish', 'features': 768, 'winlen': 19, 'scale': 5.0,
                        'rnn_type': 'lstm', 'blank_score': 2.0}}


@pytest.mark.parametrize('cs_,ov,L', list(_stitch_cases()))
def test_util_chunk_stitch_match_reference_golden(golden, cs_, ov, L):
    g = golden['stitch']
    key = 'c%d_o%d_L%d_' % (cs_, ov, L)
    sig = torch.arange(L, dtype=torch.float32)
    ch = util.chunk(sig, cs_, ov)
    assert ch.shape[1:] == (1, cs_)
    assert np.array_equal(ch[:, 0, 0].numpy().astype(np.int64), g[key + 'first'])
    assert np.array_equal(ch[-1, 0, -3:].numpy().astype(np.int64), g[key + 'lastrow'])
    T = cs_ // 5
    lab = torch.arange(ch.shape[0] * T, dtype=torch.int32).reshape(ch.shape[0], T)
    assert np.array_equal(util.stitch(lab, cs_, ov, L, 5).numpy(), g[key + 'stitched'])
    assert np.array_equal(util.stitch(lab, cs_, ov, L, 5, reverse=True).numpy(), g[key + 'stitched_rev'])
    d = stitch_results({'a': lab, 'b': lab.numpy()}, L, cs_, ov, 5)
    assert np.array_equal(d['a'].numpy(), g[key + 'stitched']) and np.array_equal(d['b'], g[key + 'stitched'])
    # the chunk table of the read-set pipeline names the same windows
    plan = pipeline.plan_chunks([L], cs_, ov)
    starts = np.maximum(plan['chunk_start'], 0) if L >= cs_ else np.zeros(1, dtype=np.int64)
    assert np.array_equal(starts, g[key + 'first'])
    assert plan['chunk_count'][0] == ch.shape[0]


def test_chunk_short_read_left_pads_and_chunksize_zero():
    sig = torch.arange(1, 8, dtype=torch.float32)
    ch = util.chunk(sig, 10, 2)
    assert ch.shape == (1, 1, 10) and ch[0, 0].tolist() == [0, 0, 0, 1, 2, 3, 4, 5, 6, 7]
    assert util.chunk(sig, 0, 0).shape == (1, 1, 7)
    plan = pipeline.plan_chunks([7], 10, 2)
    assert plan['chunk_start'].tolist() == [-3]


def test_batchify_unbatchify_round_trip_ragged():
    rs = np.random.RandomState(0)
    items = [(('r%d' % i, 0, n), torch.from_numpy(rs.randn(n, 1, 6).astype(np.float32))) for i, n in enumerate([3, 1, 7, 2, 5])]
    batches = list(util.batchify(iter(items), batchsize=4))
    assert [v.shape[0] for _, v in batches] == [4, 4, 4, 4, 2]
    want = list(bo.batchify(iter(items), 4))
    for (k1, v1), (k2, v2) in zip(batches, want):
        assert k1 == k2 and torch.equal(v1, v2)
    # dict-valued results regroup by key in input order
    scored = [(k, {'sequence': v[:, 0, :].numpy(), 'moves': v[:, 0, :] > 0}) for k, v in batches]
    back = list(util.unbatchify(iter(scored)))
    assert [k for k, _ in back] == [k for k, _ in items]
    for (k, d), (_, v) in zip(back, items):
        assert np.array_equal(d['sequence'], v[:, 0, :].numpy())
    assert list(util.batchify(iter([]), 4)) == []


def test_concat_select_size_types():
    assert util.concat(['ab', 'c']) == 'abc'
    assert util.concat([[1], [2, 3]]) == [1, 2, 3]
    assert util.size([1, 2, 3]) == 3 and util.size(np.zeros((2, 5)), 1) == 5
    x = np.arange(12).reshape(3, 4)
    assert np.array_equal(util.select_range(x, 1, 3, dim=1), x[:, 1:3])
    with pytest.raises(TypeError):
        util.concat([1, 2])


def test_layer_registry_and_state_dict_surface():
    for name in ('serial', 'convolution', 'lstm', 'linearcrfencoder', 'permute', 'reverse', 'swish', 'relu', 'tanh'):
        assert name in nn.layers
    m = Model(sup_config(ALPHABETS[5]))
    want = bo.reference_state_dict(n_base=5, seed=1)
    assert sorted(m.state_dict().keys()) == sorted(want.keys())
    assert all(m.state_dict()[k].shape == want[k].shape for k in want)
    m.load_state_dict(want)
    assert m.stride == 5 and m.alphabet == ALPHABETS[5]
    assert m.encoder[-1].expand_blanks is True and m.encoder[-1].blank_score == 2.0
    assert [type(l).__name__ for l in m.encoder] == ['Convolution'] * 3 + ['Permute'] + ['LSTM'] * 5 + ['LinearCRFEncoder']
    assert [l.reverse for l in m.encoder[4:9]] == [True, False, True, False, True]
    assert (m.encoder[4].rnn.bias_hh_l0 == 0).all() and not m.encoder[4].rnn.bias_hh_l0.requires_grad
    assert torch.equal(m.seqdist.idx, bo.crf_idx(5, 3))
    assert m.seqdist.n_score() == 750
    # to_dict / from_dict round trip keeps the architecture and the key set
    d = nn.to_dict(m.encoder)
    rebuilt = nn.from_dict(d)
    assert sorted(rebuilt.state_dict().keys()) == sorted(m.encoder.state_dict().keys())
    assert nn.to_dict(rebuilt) == d
    cfg = sup_config(ALPHABETS[6])
    cfg['encoder'] = d | {'type': 'serial'}
    cfg['encoder']['sublayers'][-1]['n_base'] = 6
    m6 = Model(cfg)
    assert m6.encoder[-1].linear.out_features == 6 ** 4


def test_load_symbol_and_load_model(tmp_path):
    cfg = sup_config(ALPHABETS[5])
    assert util.load_symbol(cfg, 'Model') is Model and util.load_symbol(cfg, 'basecall') is basecall
    d = tmp_path / 'model'
    d.mkdir()
    lines = ['[global_norm]', 'state_len = 3', '[input]', 'features = 1', '[model]', 'package = "xna_basecaller_b200.crf"',
             '[labels]', 'labels = [ "N", "A", "C", "G", "T", "X",]', '[encoder]', 'stride = 5', 'activation = "swish"',
             'features = 768', 'winlen = 19', 'scale = 5.0', 'rnn_type = "lstm"', 'blank_score = 2.0',
             '[basecaller]', 'batchsize = 384', 'chunksize = 3600', 'overlap = 500']
    (d / 'config.toml').write_text('\n'.join(lines) + '\n')
    with pytest.raises(FileNotFoundError):
        util.load_model(str(d), 'cpu', half=False)
    sd = bo.reference_state_dict(n_base=5, seed=2)
    saved = {'module.' + k: v for k, v in sd.items()}        # DataParallel-style prefix is tolerated
    torch.save(saved, str(d / 'weights_3.tar'))
    torch.save({k: v * 0 for k, v in saved.items()}, str(d / 'weights_1.tar'))
    m = util.load_model(str(d), 'cpu', half=False, chunksize=4000)
    assert m.config['basecaller'] == {'batchsize': 384, 'chunksize': 4000, 'overlap': 500}
    assert torch.equal(m.state_dict()['encoder.9.linear.weight'], sd['encoder.9.linear.weight'])   # latest = weights_3
    assert not m.training
    assert util.load_symbol(str(d), 'Model') is Model


def test_no_cpu_fallback():
    m = Model(sup_config(ALPHABETS[5])).eval()
    x = torch.zeros(2, 1, 100)
    with pytest.raises(RuntimeError, match='CUDA device'):
        m(x)
    with pytest.raises(RuntimeError, match='CUDA device'):
        m.seqdist.logZ(torch.zeros(4, 2, 750))
    with pytest.raises(RuntimeError, match='fused stem'):
        m.encoder[0](x)
    if not torch.cuda.is_available():
        from xna_basecaller_b200._lib import Handle
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            Handle('NACGTX', 3)


def test_apply_stride_to_moves_and_to_str():
    class M:
        stride = 5
    seq = np.array([65, 0, 67, 88, 0], dtype=np.int8)
    out = apply_stride_to_moves(M, {'sequence': seq, 'qstring': np.where(seq != 0, ord('O'), 0), 'moves': np.zeros(5)})
    assert out['sequence'] == 'ACX' and out['qstring'] == 'OOO'
    assert out['sig_move'].shape == (25,) and not out['sig_move'].any()
    assert to_str(np.zeros(4, dtype=np.int8)) == ''


def test_shard_plan_covers_every_read_once():
    lengths = np.random.RandomState(11).randint(4000, 20000, size=1001)
    for world in (1, 2, 4, 8):
        shards = [pipeline.shard_reads(len(lengths), r, world) for r in range(world)]
        allr = np.sort(np.concatenate(shards))
        assert np.array_equal(allr, np.arange(len(lengths)))
        sizes = [lengths[s].sum() for s in shards]
        assert max(sizes) / min(sizes) < 1.1
    plan = pipeline.plan_chunks(lengths, 4000, 500)
    assert plan['chunk_count'].sum() == len(plan['chunk_read'])
    assert np.array_equal(np.repeat(np.arange(len(lengths)), plan['chunk_count']), plan['chunk_read'])
    for r in (0, 17, 1000):
        ch = util.chunk(torch.arange(int(lengths[r]), dtype=torch.float32), 4000, 500)
        f, c = plan['chunk_first'][r], plan['chunk_count'][r]
        assert np.array_equal(plan['chunk_start'][f:f + c], ch[:, 0, 0].numpy().astype(np.int64))


def test_lr_schedules_match_reference_golden():
    """schedule.linear_warmup_cosine_decay / linear_cooldown (bonito/schedule.py:7-17,56-67) through torch's LambdaLR:
    learning rates at selected steps equal the values the reference module produced (tests/golden/schedule.json was
    written from /root/reference/ub-bonito/bonito/schedule.py; the full 750-step traces were compared with == then)."""
    import json
    import os
    from xna_basecaller_b200 import schedule
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'schedule.json')))

    class Loader:
        def __len__(self):
            return 250

    for key, want in gold.items():
        name, kw = key.split(' ', 1)
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=2e-3)
        sch = getattr(schedule, name)(**json.loads(kw))(opt, Loader(), 4, 1)
        got = []
        for _ in range(750):
            got.append(opt.param_groups[0]['lr'])
            opt.step()
            sch.step()
        assert [got[i] for i in (0, 1, 50, 99, 100, 249, 250, 499, 500, 749)] == want


# ----------------------------------------------------------------------------- training data path
def _write_ctc_dir(path, n, L=200, Lmax=30, seed=0, indices=None):
    import os
    rs = np.random.RandomState(seed)
    os.makedirs(path, exist_ok=True)
    chunks = rs.randn(n, L).astype(np.float16)
    lengths = rs.randint(5, Lmax + 1, n).astype(np.uint16)
    targets = np.zeros((n, Lmax), dtype=np.uint8)
    for i, k in enumerate(lengths):
        targets[i, :k] = rs.randint(1, 7, k)
    np.save(os.path.join(path, 'chunks.npy'), chunks)
    np.save(os.path.join(path, 'references.npy'), targets)
    np.save(os.path.join(path, 'reference_lengths.npy'), lengths)
    if indices is not None:
        np.save(os.path.join(path, 'indices.npy'), np.asarray(indices))
    return chunks, targets, lengths


def test_load_numpy_datasets_limit_and_indices(tmp_path):
    """bonito/data.py:129-163: plain directory with `limit`; indices.npy sub-sampling (out-of-range indices dropped first,
    then `limit`)."""
    from xna_basecaller_b200 import data
    d = str(tmp_path / 'plain')
    chunks, targets, lengths = _write_ctc_dir(d, 50)
    c, t, l = data.load_numpy_datasets(limit=20, directory=d)
    assert c.shape == (20, 200) and np.array_equal(c, chunks[:20]) and np.array_equal(t, targets[:20]) and np.array_equal(l, lengths[:20])
    c, t, l = data.load_numpy_datasets(directory=d)
    assert len(l) == 50
    d2 = str(tmp_path / 'indexed')
    chunks, targets, lengths = _write_ctc_dir(d2, 40, indices=[7, 3, 99, 12, 5])
    c, t, l = data.load_numpy_datasets(limit=3, directory=d2)
    assert np.array_equal(c, chunks[[7, 3, 12]]) and np.array_equal(l, lengths[[7, 3, 12]])


def test_load_numpy_split_and_items(tmp_path, capsys):
    """data.py:100-126: 97 % / 3 % split without a validation directory, validation/ used when present; item dtypes of
    ChunkDataSet (data.py:53-57); augmentation arguments are refused, not ignored."""
    from xna_basecaller_b200 import data
    d = str(tmp_path / 'set')
    chunks, targets, lengths = _write_ctc_dir(d, 100)
    tr, va = data.load_numpy(None, d)
    assert 'splitting training set' in capsys.readouterr().out
    assert len(tr['dataset']) == 97 and len(va['dataset']) == 3 and tr['shuffle'] and not va['shuffle']
    chunk, target, length = va['dataset'][1]
    assert chunk.shape == (1, 200) and chunk.dtype == np.float32 and target.dtype == np.int64 and length.dtype == np.int64
    assert np.array_equal(chunk[0], chunks[98].astype(np.float32)) and int(length) == int(lengths[98])
    _write_ctc_dir(str(tmp_path / 'set' / 'validation'), 11, seed=5)
    tr, va = data.load_numpy(60, d)
    assert len(tr['dataset']) == 60 and len(va['dataset']) == 11
    with pytest.raises(NotImplementedError):
        data.load_numpy(None, d, spike_kwargs={'prop_ubs': 0.09})


def test_device_chunk_loader_covers_every_item_once(tmp_path):
    """DeviceChunkLoader (logic checked on the CPU device): every item exactly once per pass, short last batch, reproducible
    shuffles that differ between passes, the Y -> X replacement of data.py:82-83."""
    from xna_basecaller_b200 import data
    d = str(tmp_path / 'set')
    chunks, targets, lengths = _write_ctc_dir(d, 37)
    ds = data.ChunkDataSet(*data.load_numpy_datasets(directory=d))
    ds.replace_6_letter = True
    ld = data.DeviceChunkLoader(ds, batch_size=8, shuffle=True, device='cpu', seed=3)
    assert len(ld) == 5 and len(ld.sampler) == 37
    passes = []
    for _ in range(2):
        seen = []
        for x, t, l in ld:
            assert x.dtype == torch.float32 and x.shape[1:] == (1, 200) and t.dtype == torch.int64 and l.dtype == torch.int64
            assert not (t == 6).any()
            for row, ln in zip(x[:, 0], l):          # identify the item by its signal
                i = int(np.flatnonzero((chunks.astype(np.float32) == row.numpy()).all(1))[0])
                assert int(ln) == int(lengths[i])
                seen.append(i)
        assert sorted(seen) == list(range(37))
        passes.append(seen)
    assert passes[0] != passes[1] and passes[0] != list(range(37))
    again = data.DeviceChunkLoader(ds, batch_size=8, shuffle=True, device='cpu', seed=3)
    assert [int(l[0]) for _, _, l in again] == [int(lengths[passes[0][k]]) for k in range(0, 37, 8)]
    plain = data.DeviceChunkLoader(ds, batch_size=16, shuffle=False, device='cpu')
    assert torch.equal(torch.cat([l for _, _, l in plain]), torch.from_numpy(lengths.astype(np.int64)))
    # batch_multiple=8 (what the training step of this package needs): 37 items -> batches of 16, 16 and nothing left over
    m8 = data.DeviceChunkLoader(ds, batch_size=16, shuffle=True, device='cpu', seed=4, batch_multiple=8)
    assert len(m8) == 2 and [x.shape[0] for x, _, _ in m8] == [16, 16]
    m8b = data.DeviceChunkLoader(ds, batch_size=24, shuffle=False, device='cpu', batch_multiple=8)
    assert [x.shape[0] for x, _, _ in m8b] == [24, 8]
    with pytest.raises(ValueError):
        data.DeviceChunkLoader(ds, batch_size=12, device='cpu', batch_multiple=8)
