"""A/B of two builds of the BN reduction kernels on small shapes (debug aid): compares sums bit-for-bit-ish."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simclr_b200 import _lib

def mk(path):
    l = _lib._Lib(); 
    _lib_path = _lib.LIB_PATH
    _lib.LIB_PATH = path
    l.load()
    _lib.LIB_PATH = _lib_path
    return l

new = mk(os.path.join(ROOT, 'simclr_b200', 'libsimclr_b200.so'))
old = mk(os.path.join(ROOT, 'build_tmp', 'libsimclr_bnold.so'))
st = _lib.stream_ptr()
torch.manual_seed(0)
shapes = [(16384, 32), (16384, 64), (4096, 64), (4096, 128), (4096, 256), (1024, 128), (1024, 256), (1024, 512),
          (256, 256), (256, 512), (256, 1024), (64, 512), (64, 1024), (64, 2048), (16, 32), (16, 64), (16, 128), (16, 2048),
          (600, 64), (75, 2048), (64, 4096), (50, 128)]
for dt, code in ((torch.float32, 0), (torch.bfloat16, 1)):
    for rows, C in shapes:
        y = (torch.randn(rows, C, device='cuda') * 1.5 + 0.3).to(dt)
        dz = torch.randn(rows, C, device='cuda').to(dt)
        dz2 = torch.randn(rows, C, device='cuda').to(dt)
        z = torch.relu(torch.randn(rows, C, device='cuda')).to(dt)
        stats = torch.rand(4, C, device='cuda') + 0.5
        mean, rstd, scale, shift = stats[0], stats[1], stats[2], stats[3] - 1.0
        out = {}
        for name, lib in (('new', new), ('old', old)):
            r = []
            s = torch.empty(2 * C, dtype=torch.float64, device='cuda')
            lib.bn_stats(y, code, rows, C, s, st); r.append(s.clone())
            lib.bn_bwd_relu_reduce(dz, code, y, code, rows, C, mean, rstd, scale, shift, s, st); r.append(s.clone())
            d = dz.clone(); lib.bn_bwd_reduce(d, None, None, code, y, code, rows, C, mean, rstd, s, st); r.append(s.clone())
            d = dz.clone(); lib.bn_bwd_reduce(d, dz2, z, code, y, code, rows, C, mean, rstd, s, st); r.append(s.clone()); r.append(d.double().flatten())
            d = dz.clone(); lib.bn_bwd_reduce(d, None, z, code, y, code, rows, C, mean, rstd, s, st); r.append(s.clone())
            out[name] = r
        torch.cuda.synchronize()
        errs = []
        for a, b in zip(out['new'], out['old']):
            errs.append(float((a - b).norm() / (b.norm() + 1e-30)))
        flag = 'BAD' if max(errs) > 1e-5 else 'ok'
        print('%-8s rows=%6d C=%5d  stats %.1e relu_reduce %.1e reduce %.1e reduce+res %.1e (dz %.1e) reduce+mask %.1e  %s'
              % (str(dt)[6:], rows, C, errs[0], errs[1], errs[2], errs[3], errs[4], errs[5], flag), flush=True)
