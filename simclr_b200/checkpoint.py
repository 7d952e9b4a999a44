"""Checkpoint / resume of the training driver (`tf2/run.py:263-337`): `CheckpointManager` keeps
`ckpt-<step>.npz` files in `model_dir` (variables under the reference's Keras names, optimizer
slots, `global_step`), prunes to `keep_checkpoint_max`, and `try_restore_from_checkpoint` follows
the reference's order: latest checkpoint in `model_dir` (weights + step + optimizer state), else
`--checkpoint` (weights only, optional `--zero_init_logits_layer`).

Own checkpoints use NumPy's container format; the variable NAMES are the reference's.  TensorFlow checkpoints
(`<prefix>.index` + `.data-*`: what the reference writes and what the released SimCLR weights are stored in) are READ
through `tf_checkpoint.TensorBundleReader` -- `--checkpoint=<prefix>`, or a `model_dir` holding `ckpt-N.index` files:
variables are joined on the object graph's `full_name` (HWIO kernels, [in, out] dense kernels: no transposes),
optimizer slots are not imported.  That reader is unpinned against TensorFlow-written files (tf_checkpoint.py).
"""
import glob
import os
import re

import numpy as np
import torch
from absl import logging

from .flags_def import FLAGS
from . import tf_checkpoint


class CheckpointManager:
    def __init__(self, model, optimizer, directory, max_to_keep=5):
        self.model, self.optimizer = model, optimizer
        self.directory, self.max_to_keep = directory, max_to_keep

    def _paths(self):
        if not self.directory:
            return []
        ps = glob.glob(os.path.join(self.directory, 'ckpt-*.npz'))
        return sorted(ps, key=lambda p: int(re.search(r'ckpt-(\d+)\.npz$', p).group(1)))

    def _tf_prefixes(self):
        """`ckpt-N` prefixes of TensorFlow checkpoints in the directory (a model_dir written by the reference)."""
        if not self.directory:
            return []
        ps = [p[:-6] for p in glob.glob(os.path.join(self.directory, 'ckpt-*.index'))]
        return sorted((p for p in ps if re.search(r'ckpt-(\d+)$', p)), key=lambda p: int(re.search(r'ckpt-(\d+)$', p).group(1)))

    @property
    def latest_checkpoint(self):
        ps = self._paths()
        if ps:
            return ps[-1]
        tf = self._tf_prefixes()
        return tf[-1] if tf else None

    def _restore_tf(self, prefix, weights_only):
        reader = tf_checkpoint.TensorBundleReader(prefix)
        arrays = reader.variables_by_name()
        byname = {v.name: v for v in self.model.variables}
        hit = 0
        for full, a in arrays.items():
            v = byname.get(full + ':0') or byname.get(full)
            if v is not None and tuple(a.shape) == tuple(v.shape):
                v.value.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
                hit += 1
        logging.info('TensorFlow checkpoint %s: %d of %d model variables restored (%d tensors in the file)',
                     prefix, hit, len(byname), len(arrays))
        if hit == 0:
            raise ValueError('no variable of %s matches the model (names are joined on the object graph full_name)' % prefix)
        if weights_only:
            return 0
        step = 0
        for key in ('global_step' + tf_checkpoint.VARIABLE_SUFFIX, 'global_step'):
            if key in reader.entries:
                step = int(reader.get_tensor(key))
                break
        if self.optimizer is not None:
            logging.warning('optimizer slots of TensorFlow checkpoints are not imported: momentum restarts from zero')
            self.optimizer.iterations = step
        return step

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
        if tf_checkpoint.is_tf_checkpoint(path):
            return self._restore_tf(path, weights_only)
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
