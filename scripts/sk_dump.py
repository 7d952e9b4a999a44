"""Debug aid: run the SK/ResNet-D fp32 verification step and dump every gradient (A/B of two library builds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from simclr_b200 import flags_def
import test_gpu_step as T

tag = sys.argv[1]
if tag == 'cmp':
    a = torch.load(os.path.join(ROOT, 'gpurun_out', 'grads_%s.pt' % sys.argv[2]))
    b = torch.load(os.path.join(ROOT, 'gpurun_out', 'grads_%s.pt' % sys.argv[3]))
    for i, k in enumerate(a):
        d = float((a[k] - b[k]).norm() / (b[k].norm() + 1e-30))
        sc = float((a[k] * b[k]).sum() / ((b[k] * b[k]).sum() + 1e-300))
        if d > 1e-5 and (i < 3 or i > 180):
            print('%3d %-70s diff %.2e  scale-1 %+.2e  norm %.3e' % (i, k[-70:], d, sc - 1.0, float(b[k].norm())))
    sys.exit(0)
flags_def.FLAGS(['x'])
flags_def.set_flags(sk_ratio=0.0625)
trainer, om, P, S_ = T._setup(flags_def.FLAGS, 'fp32', 'simt', False, 8, 64, depth=50, use_blur=False)
f, lab, _, _ = T._data(8, 64)
trainer.optimizer.learning_rate = 0.0
loss = trainer.single_step(f.cuda(), lab.cuda())
torch.cuda.synchronize()
out = {v.name: v.grad.detach().cpu().double().clone() for v in trainer.model.trainable_variables}
torch.save(out, os.path.join(ROOT, 'gpurun_out', 'grads_%s.pt' % tag))
print(tag, 'loss', float(loss), 'vars', len(out))
