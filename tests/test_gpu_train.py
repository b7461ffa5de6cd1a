"""GPU parity of the training step (BASELINE configs[4], fwd_bwd): the encoder backward (xb_encoder_fwd_train /
xb_encoder_bwd) against torch autograd through the fp32 oracle restatement of the reference modules
(bonito/training.py:91-117 runs exactly that autograd).

Tolerance: gradients travel as bf16 and the forward runs on fp16 operands, so every gradient tensor is compared by its
relative error in the 2-norm (<= 3e-2) and by the cosine of the angle to the reference (>= 0.999)."""
import numpy as np
import pytest
import torch

from make_golden import ALPHABETS, REF_SCALE, synthetic_signal
from oracle import bonito_oracle as bo

pytestmark = pytest.mark.gpu

REL_TOL, COS_TOL = 3e-2, 0.999


def _reference_grads(sd, x, cotangent, n_base):
    params = {k: v.clone().requires_grad_() for k, v in sd.items()}
    scores = bo.encoder_forward(params, x, n_base)
    scores.backward(cotangent)
    return scores.detach(), {k: p.grad for k, p in params.items()}


def _compare(got, ref, skip=()):
    worst = {}
    for k, g in ref.items():
        if k in skip or g is None:
            continue
        a, b = got[k].double().cpu().flatten(), g.double().flatten()
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        cos = (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30)).item()
        worst[k] = (rel, cos)
    return worst


# N = 136: two row tiles of the BPTT step kernel, the second clipped at the batch edge by its tensor maps
@pytest.mark.parametrize('n_base,N,L', [(5, 8, 500), (6, 16, 300), (5, 136, 100)])
def test_encoder_backward_matches_autograd(n_base, N, L):
    from xna_basecaller_b200._lib import Handle
    sd = bo.reference_state_dict(n_base=n_base, seed=12, **REF_SCALE)
    x = synthetic_signal(41, N, L)
    T = L // 5
    g = torch.Generator().manual_seed(9)
    C, NZ = n_base ** 3, n_base + 1
    cot = torch.randn(T, N, C * NZ, generator=g) * 1e-2
    ref_scores, ref = _reference_grads(sd, x, cot, n_base)
    h = Handle(ALPHABETS[n_base], 3, max_N=N, max_T=T, train=True)
    h.load_weights(sd)
    scores = h.encoder_train(x.cuda())
    assert (scores.cpu() - ref_scores).abs().max().item() < 1e-2
    # the training forward is the inference forward plus stores: same bits
    assert torch.equal(scores, h.encoder(x.cuda()))
    got = h.encoder_backward(cot.cuda())
    torch.cuda.synchronize()
    worst = _compare(got, ref, skip=[k for k in ref if k.endswith('bias_hh_l0')])
    for k, (rel, cos) in sorted(worst.items()):
        print('%-32s rel %.4f  cos %.6f' % (k, rel, cos))
    for k, (rel, cos) in worst.items():
        assert rel <= REL_TOL and cos >= COS_TOL, (k, rel, cos)
    # bias_hh receives the gradient of bias_ih (they enter the gates as a sum)
    for layer in range(4, 9):
        assert torch.equal(got['encoder.%d.rnn.bias_hh_l0' % layer], got['encoder.%d.rnn.bias_ih_l0' % layer])
    h.close()


def test_adamw_step_matches_torch():
    """xb_adamw_step (clip_grad_norm_(2.0) + AdamW) against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW over five
    steps on tensors of the model's shapes (one larger than a 65536-element block, one tiny)."""
    from xna_basecaller_b200.training import AdamW
    g = torch.Generator(device='cuda').manual_seed(3)
    shapes = [(3072, 768), (625, 768), (3072,), (4, 1, 5), (16,)]
    ours = [torch.nn.Parameter(torch.randn(*s, device='cuda', generator=g) * 0.1) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    kw = dict(lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    opt, opt_ref = AdamW(ours, **kw), torch.optim.AdamW(ref, **kw)
    for step in range(5):
        scale = 10.0 if step % 2 == 0 else 1e-3               # one step clipped, one not
        for p, q in zip(ours, ref):
            p.grad = torch.randn(p.shape, device='cuda', generator=g) * scale
            q.grad = p.grad.clone()
        norm = opt.step(max_norm=2.0).item()
        norm_ref = torch.nn.utils.clip_grad_norm_(ref, max_norm=2.0).item()
        opt_ref.step()
        assert abs(norm - norm_ref) <= 1e-5 * norm_ref
        for p, q in zip(ours, ref):
            assert (p - q).abs().max().item() <= 2e-6
    assert opt.state[ours[0]]['step'] == 5


def test_trainer_train_one_step_through_the_plugin():
    """Trainer.train_one_step (training.py:91-117) on the plugin Model: loss.backward() reaches every parameter through
    xb_ctc_crf_loss_bwd + xb_encoder_bwd, gradients equal the direct C-ABI calls, and a few clipped AdamW steps on one
    batch lower the loss."""
    from make_golden import synthetic_targets
    from xna_basecaller_b200 import util
    from xna_basecaller_b200.training import Trainer
    cfg = {'global_norm': {'state_len': 3}, 'input': {'features': 1}, 'labels': {'labels': ALPHABETS[5]},
           'model': {'package': 'xna_basecaller_b200.crf'},
           'encoder': {'stride': 5, 'activation': 'swish', 'features': 768, 'winlen': 19, 'scale': 5.0,
                       'rnn_type': 'lstm', 'blank_score': 2.0}}
    model = util.load_symbol(cfg, 'Model')(cfg)
    sd = bo.reference_state_dict(n_base=5, seed=12, **REF_SCALE)
    model.load_state_dict(sd)
    N, L = 8, 600
    x = synthetic_signal(51, N, L)
    tg, tl = synthetic_targets(7, N, 5, 40, 60)
    trainer = Trainer(model, 'cuda')
    trainer.init_optimizer(1e-3)
    model.train()
    # gradients of one forward / backward through autograd == the oracle's autograd on the same loss
    scores = model(x.cuda())
    loss = model.seqdist.ctc_loss(scores.float(), tg.cuda(), tl.cuda())
    loss.backward()
    params = {k: v.clone().requires_grad_() for k, v in sd.items()}
    ref_loss = bo.CRF(3, ALPHABETS[5]).ctc_loss(bo.encoder_forward(params, x, 5), tg, tl)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 2e-3 * abs(ref_loss.item())
    named = dict(model.named_parameters())
    for k, q in params.items():
        if k.endswith('bias_hh_l0'):
            continue                                   # frozen at zero in the reference (nn.py:209-213): no gradient
        a, b = named[k].grad.double().cpu().flatten(), q.grad.double().flatten()
        rel = ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
        assert rel <= 5e-2, (k, rel)
    first = None
    for step in range(6):
        losses, grad_norm = trainer.train_one_step((x, tg, tl))
        assert np.isfinite(losses['loss']) and np.isfinite(grad_norm) and grad_norm > 0
        first = losses['loss'] if first is None else first
    assert losses['loss'] < first
    print('loss %.4f -> %.4f in 6 steps, last grad norm %.3f' % (first, losses['loss'], grad_norm))


def test_device_chunk_loader_feeds_the_trainer(tmp_path):
    """Training data path (bonito/data.py): a ctc-data directory -> load_numpy -> DeviceChunkLoader (data set resident in
    HBM, batches gathered on the device) -> Trainer.train_one_step.  Passes over the set lower the loss."""
    from make_golden import synthetic_targets
    from xna_basecaller_b200 import data, util
    from xna_basecaller_b200.training import Trainer
    N, L = 16, 600
    x = synthetic_signal(61, N, L)
    tg, tl = synthetic_targets(9, N, 5, 40, 60)
    d = str(tmp_path)
    np.save(d + '/chunks.npy', x[:, 0].numpy().astype(np.float16))
    np.save(d + '/references.npy', tg.numpy().astype(np.uint8))
    np.save(d + '/reference_lengths.npy', tl.numpy().astype(np.uint16))
    np.save(d + '/indices.npy', np.arange(N))
    train_kwargs, _ = data.load_numpy(None, d)                     # 97 % / 3 % split: 15 + 1 chunks
    # unshuffled here, so that every pass trains on the same 8 chunks and the loss must fall (the 7 left-over chunks of the
    # 15 are cut off by batch_multiple=8; shuffled, others would be left over in the next pass)
    loader = data.DeviceChunkLoader(batch_size=8, device='cuda', batch_multiple=8, **dict(train_kwargs, shuffle=False))
    assert len(loader.sampler) == 15 and len(loader) == 1
    cfg = {'global_norm': {'state_len': 3}, 'input': {'features': 1}, 'labels': {'labels': ALPHABETS[5]},
           'model': {'package': 'xna_basecaller_b200.crf'},
           'encoder': {'stride': 5, 'activation': 'swish', 'features': 768, 'winlen': 19, 'scale': 5.0,
                       'rnn_type': 'lstm', 'blank_score': 2.0}}
    model = util.load_symbol(cfg, 'Model')(cfg)
    model.load_state_dict(bo.reference_state_dict(n_base=5, seed=12, **REF_SCALE))
    trainer = Trainer(model, 'cuda', train_loader=loader)
    trainer.init_optimizer(1e-3)
    per_pass = []
    for _ in range(6):
        seen, total = 0, 0.0
        for batch in loader:
            assert batch[0].is_cuda and batch[0].shape[1:] == (1, L)
            losses, grad_norm = trainer.train_one_step(batch)
            assert np.isfinite(losses['loss']) and np.isfinite(grad_norm)
            seen += batch[0].shape[0]
            total += losses['loss'] * batch[0].shape[0]
        assert seen == 8
        per_pass.append(total / seen)
    assert per_pass[-1] < per_pass[0], per_pass
