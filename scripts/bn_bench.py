"""Bandwidth of the BatchNorm-family kernels on ResNet-50 activation shapes (1024 views).
Usage: python scripts/bn_bench.py [--views 1024]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simclr_b200._lib import lib, stream_ptr

ap = argparse.ArgumentParser()
ap.add_argument('--views', type=int, default=1024)
ap.add_argument('--reps', type=int, default=5)
args = ap.parse_args()
st = stream_ptr()
SHAPES = [(112, 64), (56, 64), (56, 256), (28, 128), (28, 512), (14, 256), (14, 1024), (7, 512), (7, 2048)]
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device='cuda')
bf = torch.bfloat16

def timeit(fn):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(args.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

tot = {}
for hw, C in SHAPES:
    rows = args.views * hw * hw
    y = torch.randn(rows, C, device='cuda').to(bf)
    res = torch.randn(rows, C, device='cuda').to(bf)
    z = torch.empty_like(y); dz = torch.randn(rows, C, device='cuda').to(bf); dz2 = torch.randn(rows, C, device='cuda').to(bf)
    dy = torch.empty_like(y)
    sums = torch.zeros(2 * C, dtype=torch.float64, device='cuda')
    stats = torch.rand(4, C, device='cuda') + 0.5
    mean, rstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
    gamma = torch.ones(C, device='cuda'); coef = torch.empty(3 * C, device='cuda')
    nb = rows * C * 2
    ops = [
        ('stats', 1, lambda: lib.bn_stats(y, 1, rows, C, sums, st)),
        ('apply', 2, lambda: lib.bn_apply(y, 1, None, z, 1, rows, C, scale, shift, 1, st)),
        ('apply+res', 3, lambda: lib.bn_apply(y, 1, res, z, 1, rows, C, scale, shift, 1, st)),
        ('bwd_reduce', 2, lambda: lib.bn_bwd_reduce(dz, None, None, 1, y, 1, rows, C, mean, rstd, sums, st)),
        ('bwd_reduce+res', 5, lambda: lib.bn_bwd_reduce(dz, dz2, z, 1, y, 1, rows, C, mean, rstd, sums, st)),
        ('bwd_relu_reduce', 2, lambda: lib.bn_bwd_relu_reduce(dz, 1, y, 1, rows, C, mean, rstd, scale, shift, sums, st)),
        ('bwd_apply', 3, lambda: lib.bn_bwd_apply(dz, 1, y, 1, dy, 1, rows, C, mean, rstd, gamma, sums, sums, float(rows), None, None, coef, None, None, st)),
        ('bwd_apply+mask', 3, lambda: lib.bn_bwd_apply(dz, 1, y, 1, dy, 1, rows, C, mean, rstd, gamma, sums, sums, float(rows), None, None, coef, scale, shift, st)),
    ]
    line = '%3dx%-3d C=%-4d %6.0f MB |' % (hw, hw, C, nb / 1e6)
    for name, passes, fn in ops:
        ms = timeit(fn)
        gbs = passes * nb / ms / 1e6
        line += ' %s %.0f' % (name, gbs)
        t = tot.setdefault(name, [0.0, 0.0]); t[0] += passes * nb; t[1] += ms
    print(line, flush=True)
    del y, res, z, dz, dz2, dy
print('TOTAL GB/s: ' + '  '.join('%s %.0f' % (k, v[0] / v[1] / 1e6) for k, v in tot.items()))
