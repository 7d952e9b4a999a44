"""GPU diagnostics (scratch): per-variable gradient error tables and tf32 probes."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from util import rel_err, cfg_from_flags
from simclr_b200 import engine, run, flags_def
from simclr_b200._lib import lib, stream_ptr
from oracle import model as OM, step as OS

flags_def.FLAGS(['diag'])


def table(precision, conv_engine, warm=True, B=32, S=64):
    flags_def.set_flags(resnet_depth=18, image_size=S, train_batch_size=B, use_blur=True,
                        b200_precision=precision, b200_conv_engine=conv_engine, weight_decay=1e-4)
    engine.set_engine(engine.Engine(precision=precision, conv_engine=conv_engine))
    trainer = run.Trainer(num_classes=1000, num_examples=50000, seed=0)
    om = OM.Model(cfg_from_flags(flags_def.FLAGS), 1000)
    P, S_ = om.init(0)
    g = torch.Generator().manual_seed(5)
    if warm:
        for k in P:
            if k.endswith('gamma:0'):
                P[k] = torch.rand(P[k].shape, generator=g) + 0.5
            elif k.endswith('beta:0'):
                P[k] = torch.randn(P[k].shape, generator=g) * 0.1
    trainer.model.vs.load(P)
    g = torch.Generator().manual_seed(1)
    f = torch.rand(B, S, S, 6, generator=g)
    lab = torch.nn.functional.one_hot(torch.randint(0, 1000, (B,), generator=g), 1000).float()
    sigma = [0.9, 1.6]; sel = (torch.rand(2, B, generator=g) < 0.5).to(torch.uint8)
    info = OS.forward_backward(om, P, S_, [f], [lab], blur_draws=[[(sigma[0], sel[0]), (sigma[1], sel[1])]])
    trainer.model.set_blur_draws(torch.tensor(sigma), sel)
    trainer.optimizer.learning_rate = 0.0
    loss = trainer.single_step(f.cuda(), lab.cuda())
    torch.cuda.synchronize()
    print('==== %s %s warm=%s loss %.6f oracle %.6f' % (precision, conv_engine, warm, loss.item(), info['loss'].item()))
    print('proj_out err', rel_err(trainer.metrics['logits_con'], info['logits_con'][0]))
    for v in trainer.model.trainable_variables:
        ref = info['grads'][v.name]
        print('%-95s %10.3e  |ref| %.3e' % (v.name[-95:], rel_err(v.grad, ref), ref.norm().item()))


def tf32_probe():
    from test_gpu_tc import _pack, conv_reference
    for (N, H, W, Cin, Cout, k, s) in [(2, 16, 16, 64, 256, 1, 1), (2, 12, 12, 64, 64, 3, 1)]:
        for dtype in (torch.float32, torch.bfloat16):
            x = torch.ones(N, H, W, Cin).to(dtype); w = (torch.ones(k, k, Cin, Cout) * 0.5).to(dtype)
            wf, wd = _pack(w, dtype, k, Cin, Cin, Cout)
            y = torch.full((N, H, W, Cout), float('nan'), device='cuda')
            code = 0 if dtype == torch.float32 else 1
            lib.conv2d_fprop_tc(x.cuda(), wf, y, code, 0, N, H, W, Cin, Cout, k, k, s, None, stream_ptr())
            torch.cuda.synchronize()
            print('tf32 probe', dtype, (N, H, W, Cin, Cout, k, s), 'y[0,5,5,:4]', y[0, 5, 5, :4].tolist(), 'expect', 0.5 * Cin * k * k,
                  'wf[0,:4]', wf[0, :4].tolist(), 'nan count', int(torch.isnan(y).sum()))
            xr = torch.randn(N, H, W, Cin).to(dtype); wr = (torch.randn(k, k, Cin, Cout) * 0.1).to(dtype)
            wf, wd = _pack(wr, dtype, k, Cin, Cin, Cout)
            lib.conv2d_fprop_tc(xr.cuda(), wf, y, code, 0, N, H, W, Cin, Cout, k, k, s, None, stream_ptr())
            torch.cuda.synchronize()
            print('   random rel err', rel_err(y, conv_reference(xr.double(), wr.double(), k, s)))


if __name__ == '__main__':
    what = sys.argv[1:] or ['tf32', 'fp32simt', 'bf16simt']
    if 'tf32' in what:
        tf32_probe()
    if 'fp32simt' in what:
        table('fp32', 'simt')
    if 'bf16simt' in what:
        table('bf16', 'simt')
    if 'fp32tc' in what:
        table('fp32', 'tc')
    if 'bf16tc' in what:
        table('bf16', 'tc')
