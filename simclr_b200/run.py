"""Pretraining driver: the flag surface and `single_step` of the reference's
`tf2/run.py` on the B200 engine.

`python -m simclr_b200.run --train_batch_size=512 ...` (under torchrun for more
than one GPU: one process per GPU, NCCL over NVLink) trains on synthetic
tensors of the reference's input contract (`tf2/data.py:52-62`: [B,H,W,6] fp32 in
[0,1] + one-hot labels); dataset reading, checkpoints, eval and export of the
reference driver are outside this path.
"""
import os

import torch
import torch.distributed as dist
from absl import app
from absl import logging

from .flags_def import FLAGS
from . import engine as engine_lib
from . import model as model_lib
from . import objective as obj_lib
from ._lib import lib, stream_ptr


class Trainer:
    """Owns model + optimizer and runs `single_step` (tf2/run.py:557-622)."""

    def __init__(self, num_classes=None, num_examples=None, seed=0, engine=None):
        self.engine = engine or engine_lib.get_engine()
        self.strategy = self.engine.ctx
        self.num_classes = num_classes or FLAGS.b200_num_classes
        self.num_examples = num_examples or FLAGS.b200_num_examples
        self.model = model_lib.Model(self.num_classes, seed=seed)
        self.learning_rate = model_lib.WarmUpAndCosineDecay(FLAGS.learning_rate, self.num_examples)
        self.optimizer = model_lib.build_optimizer(self.learning_rate)
        self.metrics = {}
        self._graph = None
        self._static = None

    # ------------------------------------------------------------------
    def single_step(self, features, labels):
        """One synchronous data-parallel step on this replica's shard.

        features [B,H,W,6] fp32, labels one-hot [B,classes] fp32 (or None).
        Loss = contrastive + supervised + weight decay, divided by the number of
        replicas (tf2/run.py:587-617); gradients are summed across replicas
        (Keras `apply_gradients`, C3) and LARS is applied."""
        loss = self.forward_backward(features, labels)
        self.reduce_gradients()
        self.optimizer.apply_gradients([(v.grad, v) for v in self.model.trainable_variables])
        return loss

    def forward_backward(self, features, labels):
        """Forward, losses and the explicit backward pass: fills `.grad` of every trainable
        variable with THIS replica's contribution.  The collectives inside (SyncBN statistics,
        embedding / log-sum-exp all-gathers) are our own peer-memory kernels when
        `strategy.comm` is set, so this part is CUDA-graph capturable at any replica count."""
        e, model, R = self.engine, self.model, self.strategy.num_replicas_in_sync
        e.begin_step(model.vs.flat_grad)       # BN-sum pool and the flat gradient buffer zeroed once
        try:
            return self._forward_backward(features, labels)
        finally:
            e.end_step()

    def _forward_backward(self, features, labels):
        e, model, R = self.engine, self.model, self.strategy.num_replicas_in_sync
        st = stream_ptr()
        projection_head_outputs, supervised_head_outputs = model(features, training=True)
        B = features.shape[0]
        loss = None
        d_proj = d_sup = None
        if projection_head_outputs is not None:
            con_loss, logits_con, labels_con, ctx = obj_lib.contrastive_forward(
                projection_head_outputs, hidden_norm=FLAGS.hidden_norm, temperature=FLAGS.temperature,
                strategy=self.strategy if R > 1 else None, want_labels=False)
            loss = con_loss
            d_proj = obj_lib.contrastive_backward(ctx, 1.0 / (B * R))
            self.metrics['contrast_loss'] = con_loss
            self.metrics['logits_con'] = logits_con
        if supervised_head_outputs is not None:
            # l = concat([l, l], 0) (tf2/run.py:600-602): row r uses labels[r % B]
            sup_loss, d_sup = obj_lib.add_supervised_loss(labels, supervised_head_outputs,
                                                          grad_scale=1.0 / (supervised_head_outputs.shape[0] * R))
            loss = sup_loss if loss is None else loss + sup_loss
            self.metrics['supervised_loss'] = sup_loss
        weight_decay = model_lib.add_weight_decay(model, adjust_per_optimizer=True)
        self.metrics['weight_decay'] = weight_decay
        loss = loss + weight_decay
        self.metrics['total_loss'] = loss
        model.backward(d_proj, d_sup)
        # d(weight_decay)/dW = wd * W on the supervised-head kernel; loss / R per replica
        if 'lars' in FLAGS.optimizer:
            for v in model.trainable_variables:
                if 'head_supervised' in v.name and 'bias' not in v.name:
                    lib.axpy(float(FLAGS.weight_decay) / R, v.value, v.grad, v.numel, st)
        return loss

    def reduce_gradients(self):
        """C3: cross-replica SUM of the gradients, one NCCL all-reduce over the flat buffer."""
        if self.strategy.num_replicas_in_sync > 1:
            self.strategy.all_reduce_sum(self.model.vs.flat_grad)

    # ------------------------------------------------------------------
    def capture(self, features, labels, warmup=2, restore=False):
        """Captures `single_step` in a CUDA graph (the reference runs its steps inside
        one `tf.while_loop`, tf2/run.py:626-638).  `features`/`labels` become the
        static input buffers: copy new data into them before `replay()`.

        The `warmup` eager steps run on a side stream first (lazy initialisation, allocator
        warm-up) and are REAL steps: they update weights, momentum, moving statistics and
        advance `optimizer.iterations`.  `restore=True` snapshots that state before the
        warm-up and puts it back after the capture, so the first `replay()` is step
        `optimizer.iterations` of the schedule on the pre-capture weights.  The capture
        itself executes nothing and does not advance the schedule."""
        self._static = (features, labels)
        opt, vs = self.optimizer, self.model.vs
        opt.ensure_built(self.model.trainable_variables)      # H2D table copies must not land in the capture
        snap = None
        if restore:
            snap = (vs.flat_value.clone(), vs.flat_moving.clone(), opt._flat_v.clone(), opt.iterations)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.single_step(features, labels)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.optimizer.stage_learning_rate()
        R = self.strategy.num_replicas_in_sync
        if R == 1:
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._graph_loss = self.single_step(features, labels)
            self._graph_apply = None
        else:
            # More than one replica: the in-step collectives are peer-memory kernels (capturable); the
            # gradient all-reduce is NCCL and stays between two graphs.
            if self.strategy.comm is None:
                raise RuntimeError('CUDA-graph capture with %d replicas needs the peer-memory collectives '
                                   '(strategy.comm); run eagerly instead' % R)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._graph_loss = self.forward_backward(features, labels)
            self._graph_apply = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph_apply):
                self.optimizer.apply_gradients([(v.grad, v) for v in self.model.trainable_variables])
        if snap is not None:
            vs.flat_value.copy_(snap[0]); vs.flat_moving.copy_(snap[1]); opt._flat_v.copy_(snap[2])
            opt.iterations = snap[3]
        return self._graph

    def replay(self):
        self.optimizer.prepare_replay()
        self._graph.replay()
        if self._graph_apply is not None:
            self.reduce_gradients()
            self._graph_apply.replay()
        return self._graph_loss


def init_distributed():
    """One process per GPU under torchrun; no-op for a single process."""
    if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized():
        local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    return dist.get_rank() if dist.is_initialized() else 0


def synthetic_batch(batch, image_size, num_classes, device, seed):
    """Synthetic inputs of the reference's input contract."""
    g = torch.Generator(device=device).manual_seed(seed)
    features = torch.rand(batch, image_size, image_size, 6, device=device, generator=g)
    idx = torch.randint(0, num_classes, (batch,), device=device, generator=g)
    labels = torch.nn.functional.one_hot(idx, num_classes).float()
    return features, labels


def main(argv):
    if len(argv) > 1:
        raise app.UsageError('Too many command-line arguments.')
    if FLAGS.mode != 'train' or FLAGS.train_mode != 'pretrain':
        raise NotImplementedError('only --mode=train --train_mode=pretrain is on the B200 path')
    rank = init_distributed()
    engine_lib.set_engine(engine_lib.Engine())
    trainer = Trainer()
    R = trainer.strategy.num_replicas_in_sync
    assert FLAGS.train_batch_size % R == 0
    B = FLAGS.train_batch_size // R            # per-replica batch size (tf2/data.py:45)
    features, labels = synthetic_batch(B, FLAGS.image_size, trainer.num_classes, trainer.engine.device, 1234 + rank)
    train_steps = model_lib.get_train_steps(trainer.num_examples)
    for step in range(train_steps):
        loss = trainer.single_step(features, labels)
        if rank == 0 and (step % 10 == 0 or step == train_steps - 1):
            logging.info('Step: [%d] total_loss = %f', step, float(loss))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    app.run(main)
