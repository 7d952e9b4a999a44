"""Generates the golden fixtures in this directory from the oracle.

The reference (TensorFlow) cannot be imported in this environment (SURVEY.md
8c), so these vectors are outputs of the CPU restatement, not of the reference:
they pin the oracle against drift and give the GPU tests fixed inputs.
Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.config import default_cfg          # noqa: E402
from oracle import model as M, objective as O, step as St, data_util as D   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

torch.manual_seed(1234)
h = torch.randn(16, 16)
loss, logits, labels = O.add_contrastive_loss(h, True, 0.1)
np.savez(os.path.join(HERE, 'ntxent_b8_d16.npz'), hidden=h.numpy(), temperature=0.1, loss=loss.item(),
         logits_ab=logits.numpy(), labels=labels.numpy())

img = torch.rand(1, 64, 64, 3)
np.savez(os.path.join(HERE, 'blur_64.npz'), image=img.numpy(), sigma=1.3,
         blurred=D.gaussian_blur(img, 6, 1.3).numpy())

cfg = default_cfg(resnet_depth=18, image_size=32, train_batch_size=4, use_blur=False)
m = M.Model(cfg, 10)
P, S = m.init(0)
f = torch.rand(4, 32, 32, 6)
lab = torch.nn.functional.one_hot(torch.tensor([1, 3, 5, 7]), 10).float()
info = St.forward_backward(m, P, S, [f], [lab])
np.savez(os.path.join(HERE, 'r18_step.npz'), features=f.numpy(), labels=lab.numpy(), loss=info['loss'].item(),
         sup_grad_norm=info['grads']['head_supervised/linear_layer/dense_3/kernel:0'].norm().item())
print('golden fixtures written to', HERE)
