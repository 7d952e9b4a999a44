"""Input pipeline of the reference (`tf2/data.py:29-115`) over decoded images held in host memory.

`build_input_fn(builder, global_batch_size, topology, is_training)` keeps the reference's signature and
returns `_input_fn(input_context)`, an iterator factory yielding `(features, labels)` batches of the model's
input contract (`tf2/data.py:52-62`): pretraining -> two augmented views concatenated on channels
`[B,H,W,6]`, otherwise one view `[B,H,W,3]`, fp32 in [0,1] on the device, labels one-hot.  The per-image
map (`get_preprocess_fn` -> `data_util.preprocess_image`) runs as ONE augmentation kernel launch per view
over the whole batch (`simclr_augment`: crop + bicubic resize [+ flip + colour ops], draws on the host).

What stands in for TFDS: `ArrayBuilder` (uint8 images + integer labels per split, e.g. loaded from an
`.npz` under `--data_dir`); reading TFRecords / decoding JPEG is not part of this repo.  Shuffling follows
the reference's recipe -- a buffer of `batch_size * (50 if image_size <= 32 else 10)` elements drawn
uniformly, `repeat(-1)`, `drop_remainder` when training -- and each input pipeline reads its own
contiguous shard of the split (`input_context`).
"""
import collections
import functools
import os
import random

import numpy as np
import torch
from absl import logging

from .flags_def import FLAGS
from . import data_util
from .engine import get_engine

InputContext = collections.namedtuple('InputContext', 'num_input_pipelines input_pipeline_id num_replicas_in_sync')


def _per_replica_batch_size(ctx, global_batch_size):
    if global_batch_size % ctx.num_replicas_in_sync:
        raise ValueError('global batch %d is not divisible by %d replicas' % (global_batch_size, ctx.num_replicas_in_sync))
    return global_batch_size // ctx.num_replicas_in_sync


class _Split:
    def __init__(self, images, labels):
        assert len(images) == len(labels)
        self.images, self.labels = images, np.asarray(labels, dtype=np.int64)
        self.num_examples = len(labels)


class _Info:
    def __init__(self, splits, num_classes):
        self.splits = splits
        self.features = {'label': collections.namedtuple('L', 'num_classes')(num_classes)}


class ArrayBuilder:
    """Stand-in for the `tfds.builder(...)` object `tf2/run.py:469-475` uses: `.info.splits[name].num_examples`,
    `.info.features['label'].num_classes`, and the decoded examples of a split."""

    def __init__(self, splits, num_classes):
        """splits: {name: (images, labels)}; images: list / object array of uint8 [Hs,Ws,3] arrays (any sizes)
        or one uint8 array [N,Hs,Ws,3]."""
        self.info = _Info({k: _Split(*v) for k, v in splits.items()}, num_classes)

    @classmethod
    def from_npz(cls, path):
        """`<split>_images`, `<split>_labels` arrays (+ optional scalar `num_classes`)."""
        z = np.load(path, allow_pickle=True)
        names = sorted(k[:-7] for k in z.files if k.endswith('_images'))
        splits = {n: (z[n + '_images'], z[n + '_labels']) for n in names}
        nc = int(z['num_classes']) if 'num_classes' in z.files else int(max(int(np.max(v[1])) for v in splits.values()) + 1)
        return cls(splits, nc)

    def download_and_prepare(self):
        pass

    def examples(self, split):
        return self.info.splits[split]


def get_preprocess_fn(is_training, is_pretrain):
    """Get function that accepts an image and returns a preprocessed image (tf2/data.py:101-115)."""
    # Disable test cropping for small images (e.g. CIFAR)
    test_crop = not (FLAGS.image_size <= 32)
    color_jitter_strength = FLAGS.color_jitter_strength if is_pretrain else 0.
    return functools.partial(data_util.preprocess_image, height=FLAGS.image_size, width=FLAGS.image_size,
                             is_training=is_training, color_jitter_strength=color_jitter_strength, test_crop=test_crop)


def build_input_fn(builder, global_batch_size, topology, is_training):
    """tf2/data.py:29-92.  `topology` is accepted for signature parity (TPU only)."""
    del topology

    def _input_fn(input_context, seed=0, make_batch=None):
        batch_size = _per_replica_batch_size(input_context, global_batch_size)
        logging.info('Global batch size: %d', global_batch_size)
        logging.info('Per-replica batch size: %d', batch_size)
        num_classes = builder.info.features['label'].num_classes
        split = builder.examples(FLAGS.train_split if is_training else FLAGS.eval_split)
        P, p = input_context.num_input_pipelines, input_context.input_pipeline_id
        logging.info('num_input_pipelines: %d', P)
        lo, hi = split.num_examples * p // P, split.num_examples * (p + 1) // P        # this pipeline's shard
        pretrain = is_training and FLAGS.train_mode == 'pretrain'
        H = FLAGS.image_size
        jitter = FLAGS.color_jitter_strength if pretrain else 0.
        test_crop = not (FLAGS.image_size <= 32)
        rng = random.Random(seed * 1000003 + p)

        def order():
            """dataset.shuffle(batch * buffer_multiplier).repeat(-1) when training; one ordered pass otherwise."""
            if not is_training:
                yield from range(lo, hi)
                return
            buffer_size = batch_size * (50 if FLAGS.image_size <= 32 else 10)
            buf = []
            while True:
                for i in range(lo, hi):
                    buf.append(i)
                    if len(buf) >= buffer_size:
                        yield buf.pop(rng.randrange(len(buf)))
                # (repeat: the buffer keeps its remainder across epochs, as tf.data does)

        def batches():
            e = get_engine() if make_batch is None else None
            idx = []
            for i in order():
                idx.append(i)
                if len(idx) == batch_size:
                    yield (make_batch or make)(e, idx)
                    idx = []
            if idx and not is_training:            # drop_remainder=is_training
                yield (make_batch or make)(e, idx)

        def make(e, idx):
            images = [torch.from_numpy(np.ascontiguousarray(split.images[i])) for i in idx]
            n = len(idx)
            if pretrain:
                feats = torch.empty((n, H, H, 6), dtype=torch.float32, device=e.device)
                for v in range(2):                       # two transformations (tf2/data.py:55-58)
                    draws = [data_util.draw_train_augmentation(im.shape[0], im.shape[1], jitter, rng) for im in images]
                    if jitter <= 0:
                        draws = [dict(d, color=dict(d['color'], apply_jitter=False, apply_gray=False)) for d in draws]
                    data_util.preprocess_for_train_batch(images, draws, H, H, out=feats, channel_offset=3 * v)
            elif is_training:
                draws = [data_util.draw_train_augmentation(im.shape[0], im.shape[1], 0., rng) for im in images]
                draws = [dict(d, color=dict(d['color'], apply_jitter=False, apply_gray=False)) for d in draws]
                feats = data_util.preprocess_for_train_batch(images, draws, H, H)
            else:
                feats = data_util.preprocess_for_eval_batch(images, H, H, crop=test_crop)
            lab = torch.nn.functional.one_hot(torch.from_numpy(split.labels[idx]), num_classes).float().to(e.device)
            return feats, lab

        return batches()

    return _input_fn


def build_distributed_dataset(builder, batch_size, is_training, strategy, topology):
    """tf2/data.py:95-98: one input pipeline per replica (process)."""
    input_fn = build_input_fn(builder, batch_size, topology, is_training)
    R = strategy.num_replicas_in_sync
    return input_fn(InputContext(R, strategy.replica_id, R))
