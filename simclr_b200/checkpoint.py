"""Checkpoint / resume of the training driver (`tf2/run.py:263-337`): `CheckpointManager` keeps
`ckpt-<step>.npz` files in `model_dir` (variables under the reference's Keras names, optimizer
slots, `global_step`), prunes to `keep_checkpoint_max`, and `try_restore_from_checkpoint` follows
the reference's order: latest checkpoint in `model_dir` (weights + step + optimizer state), else
`--checkpoint` (weights only, optional `--zero_init_logits_layer`).

The container format is NumPy's (TensorFlow is not available to write or read tensor-bundle
checkpoints); the variable NAMES are the reference's, so a converter from a released SimCLR
checkpoint is a dictionary copy (HWIO kernels, [in, out] dense kernels: no transposes).
"""
import glob
import os
import re

import numpy as np
import torch
from absl import logging

from .flags_def import FLAGS


class CheckpointManager:
    def __init__(self, model, optimizer, directory, max_to_keep=5):
        self.model, self.optimizer = model, optimizer
        self.directory, self.max_to_keep = directory, max_to_keep

    def _paths(self):
        if not self.directory:
            return []
        ps = glob.glob(os.path.join(self.directory, 'ckpt-*.npz'))
        return sorted(ps, key=lambda p: int(re.search(r'ckpt-(\d+)\.npz$', p).group(1)))

    @property
    def latest_checkpoint(self):
        ps = self._paths()
        return ps[-1] if ps else None

    def save(self, step):
        os.makedirs(self.directory, exist_ok=True)
        out = {'global_step': np.asarray(int(step), dtype=np.int64)}
        for v in self.model.variables:
            out['model/' + v.name] = v.value.detach().float().cpu().numpy()
        opt = self.optimizer
        if opt is not None:
            for name, t in optimizer_state(opt).items():
                out['optimizer/' + name] = t.detach().float().cpu().numpy()
        path = os.path.join(self.directory, 'ckpt-%d.npz' % int(step))
        tmp = path + '.tmp.npz'
        np.savez(tmp, **out)
        os.replace(tmp, path)
        if self.max_to_keep:
            for old in self._paths()[:-self.max_to_keep]:
                os.remove(old)
        return path

    def restore(self, path, weights_only=False):
        """Returns the restored global step (0 with `weights_only`).  Unknown / missing entries are skipped
        like `expect_partial()`."""
        data = np.load(path)
        byname = {v.name: v for v in self.model.variables}
        for k in data.files:
            if k.startswith('model/') and k[6:] in byname:
                v = byname[k[6:]]
                if tuple(data[k].shape) == v.shape:
                    v.value.copy_(torch.from_numpy(data[k]))
        if weights_only:
            return 0
        step = int(data['global_step']) if 'global_step' in data.files else 0
        if self.optimizer is not None:
            self.optimizer.ensure_built(self.model.trainable_variables)
            for name, t in optimizer_state(self.optimizer).items():
                if 'optimizer/' + name in data.files and tuple(data['optimizer/' + name].shape) == tuple(t.shape):
                    t.copy_(torch.from_numpy(data['optimizer/' + name]))
            self.optimizer.iterations = step
        return step


def optimizer_state(opt):
    """name -> device tensor of every optimizer slot buffer."""
    if hasattr(opt, '_flat_v') and getattr(opt, '_flat_v', None) is not None:          # LARS momentum slots
        return {'Momentum': opt._flat_v}
    slots = getattr(opt, '_slots', None)
    if isinstance(slots, list):
        return {'slot_%d' % i: t for i, t in enumerate(slots)}
    return {}


def try_restore_from_checkpoint(model, optimizer):
    """tf2/run.py:308-337.  Returns the CheckpointManager; `optimizer.iterations` is the restored step."""
    manager = CheckpointManager(model, optimizer, FLAGS.model_dir, FLAGS.keep_checkpoint_max)
    latest = manager.latest_checkpoint
    if latest:
        logging.info('Restoring from latest checkpoint: %s', latest)
        manager.restore(latest)
    elif FLAGS.checkpoint:
        logging.info('Restoring from given checkpoint: %s', FLAGS.checkpoint)
        manager.restore(FLAGS.checkpoint, weights_only=True)
        if FLAGS.zero_init_logits_layer:
            for v in model.trainable_variables:
                if 'head_supervised' in v.name:
                    logging.info('Initializing output layer parameter %s to zero', v.name)
                    v.value.zero_()
    return manager
