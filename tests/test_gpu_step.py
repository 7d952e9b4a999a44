"""End-to-end parity of `single_step` (tf2/run.py:557-622) with the CPU oracle:
config 1 of BASELINE.json (ResNet-18, batch 32, 64x64, one replica).

fp32 storage + CUDA-core conv engine is the 1e-3 parity mode (north_star
tolerance); the bf16 tcgen05 path is checked against the same oracle with the
looser tolerance bf16 activations allow.  Both the reference initialisation
(zero last-BN gamma => exactly-zero in-block conv gradients) and a
"warm" state are covered.
"""
import collections

import pytest
import torch

from util import rel_err, cfg_from_flags

pytestmark = pytest.mark.gpu


def _setup(flags, precision, conv_engine, warm, B=32, S=64, depth=18, use_blur=True):
    from simclr_b200 import engine, run, flags_def
    from oracle import model as OM
    flags_def.set_flags(resnet_depth=depth, image_size=S, train_batch_size=B, use_blur=use_blur,
                        b200_precision=precision, b200_conv_engine=conv_engine, weight_decay=1e-4)
    eng = engine.set_engine(engine.Engine(precision=precision, conv_engine=conv_engine))
    trainer = run.Trainer(num_classes=1000, num_examples=50000, seed=0)
    cfg = cfg_from_flags(flags_def.FLAGS)
    om = OM.Model(cfg, 1000)
    assert [(k, s) for k, (s, _) in om.vs.trainable.items()] == \
        [(v.name, v.shape) for v in trainer.model.trainable_variables], 'variable names/shapes must match the oracle'
    P, S_ = om.init(0)
    if warm:
        g = torch.Generator().manual_seed(5)
        for k in P:
            if k.endswith('gamma:0'):
                P[k] = torch.rand(P[k].shape, generator=g) + 0.5
            elif k.endswith('beta:0'):
                P[k] = torch.randn(P[k].shape, generator=g) * 0.1
        for k in S_:
            S_[k] = torch.rand(S_[k].shape, generator=g) + 0.5 if 'variance' in k else torch.randn(S_[k].shape, generator=g) * 0.1
    trainer.model.vs.load(P)
    trainer.model.vs.load(S_)
    return trainer, om, P, S_


def _data(B, S, seed=1):
    g = torch.Generator().manual_seed(seed)
    f = torch.rand(B, S, S, 6, generator=g)
    lab = torch.nn.functional.one_hot(torch.randint(0, 1000, (B,), generator=g), 1000).float()
    sigma = [0.9, 1.6]
    sel = (torch.rand(2, B, generator=g) < 0.5).to(torch.uint8)
    return f, lab, sigma, sel


@pytest.mark.parametrize('warm', [False, True])
def test_step_parity_fp32(flags, warm):
    from oracle import step as OS
    B, S = 32, 64
    trainer, om, P, S_ = _setup(flags, 'fp32', 'simt', warm, B, S)
    f, lab, sigma, sel = _data(B, S)
    lr = 0.3
    trainer.optimizer.learning_rate = lr
    V = collections.OrderedDict((k, torch.zeros_like(v)) for k, v in P.items())
    Pn, Sn, Vn, info = OS.single_step(om, P, S_, V, [f], [lab], lr,
                                      blur_draws=[[(sigma[0], sel[0]), (sigma[1], sel[1])]])
    trainer.model.set_blur_draws(torch.tensor(sigma), sel)
    loss = trainer.single_step(f.cuda(), lab.cuda())
    torch.cuda.synchronize()
    assert abs(loss.item() - info['loss'].item()) < 1e-4 * abs(info['loss'].item())
    assert rel_err(trainer.metrics['logits_con'], info['logits_con'][0]) < 1e-4
    worst = 0.0
    for v in trainer.model.trainable_variables:
        ref = info['grads'][v.name]
        err = rel_err(v.grad, ref)
        if ref.norm() == 0:
            assert err < 1e-6, (v.name, err)          # exactly-zero grads at the reference init (Q3)
        else:
            assert err < 1e-3, (v.name, err)
            worst = max(worst, err)
    for v in trainer.model.trainable_variables:
        assert rel_err(v.value, Pn[v.name]) < 1e-3, v.name
    for v in trainer.model.vs.moving:
        assert rel_err(v.value, Sn[v.name]) < 1e-4, v.name
    print('worst grad rel err', worst)


@pytest.mark.parametrize('precision,conv_engine,tol', [('bf16', 'tc', 0.08), ('fp32', 'tc', 0.02), ('bf16', 'simt', 0.08)])
def test_step_tensor_core_path(flags, precision, conv_engine, tol):
    """Same step through the tcgen05 engine.  bf16 activations cannot meet 1e-3
    through 18 layers; the loss must agree to 1% and every non-zero gradient
    tensor to `tol` relative (typical observed error is printed)."""
    from oracle import step as OS
    B, S = 32, 64
    trainer, om, P, S_ = _setup(flags, precision, conv_engine, True, B, S)
    f, lab, sigma, sel = _data(B, S)
    info = OS.forward_backward(om, P, S_, [f], [lab], blur_draws=[[(sigma[0], sel[0]), (sigma[1], sel[1])]])
    trainer.model.set_blur_draws(torch.tensor(sigma), sel)
    trainer.optimizer.learning_rate = 0.0
    loss = trainer.single_step(f.cuda(), lab.cuda())
    torch.cuda.synchronize()
    assert abs(loss.item() - info['loss'].item()) < 1e-2 * abs(info['loss'].item())
    errs = {v.name: rel_err(v.grad, info['grads'][v.name]) for v in trainer.model.trainable_variables
            if info['grads'][v.name].norm() > 0}
    worst = max(errs, key=errs.get)
    print('median grad rel err', sorted(errs.values())[len(errs) // 2], 'worst', worst, errs[worst])
    assert errs[worst] < tol, (worst, errs[worst])


def test_two_steps_graph_replay(flags):
    """CUDA-graph capture/replay of the step gives the same weights as eager steps."""
    B, S = 16, 32
    results = []
    for use_graph in (False, True):
        trainer, om, P, S_ = _setup(flags, 'bf16', 'tc', True, B, S, use_blur=False)
        f, lab, _, _ = _data(B, S)
        f, lab = f.cuda(), lab.cuda()
        trainer.optimizer.learning_rate = 0.05
        if use_graph:
            vs = trainer.model.vs
            w0, m0 = vs.flat_value.clone(), vs.flat_moving.clone()
            trainer.capture(f, lab, warmup=1)
            # the warm-up step moved the state: restore the initial one before replaying
            vs.flat_value.copy_(w0); vs.flat_moving.copy_(m0)
            trainer.optimizer._flat_v.zero_(); trainer.optimizer.iterations = 0
            trainer.replay(); trainer.replay()
        else:
            trainer.single_step(f, lab); trainer.single_step(f, lab)
        torch.cuda.synchronize()
        results.append(trainer.model.vs.flat_value.clone())
    assert rel_err(results[1], results[0]) < 2e-3     # atomics in wgrad/BN make it non-bitwise
