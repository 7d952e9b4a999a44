"""Oracle restatement of `tf2/model.py` (heads, schedule, weight decay, Model).

Test infrastructure only -- see `oracle/__init__.py`.
"""
import math

import torch
import torch.nn.functional as F

from . import resnet as resnet_lib
from . import data_util
from .bf16_emul import q, qb, qw      # identities unless `emulate_bf16()` is active


def get_train_steps(cfg, num_examples):
    """tf2/model.py:72-75."""
    return cfg.train_steps or (num_examples * cfg.train_epochs // cfg.train_batch_size + 1)


def warmup_and_cosine_decay(cfg, base_learning_rate, num_examples, step):
    """tf2/model.py:78-110 evaluated at integer `step`."""
    warmup_steps = int(round(cfg.warmup_epochs * num_examples // cfg.train_batch_size))
    if cfg.learning_rate_scaling == 'linear':
        scaled_lr = base_learning_rate * cfg.train_batch_size / 256.
    elif cfg.learning_rate_scaling == 'sqrt':
        scaled_lr = base_learning_rate * math.sqrt(cfg.train_batch_size)
    else:
        raise ValueError('Unknown learning rate scaling {}'.format(cfg.learning_rate_scaling))
    if step < warmup_steps:
        return step / float(warmup_steps) * scaled_lr
    total_steps = get_train_steps(cfg, num_examples)
    decay_steps = total_steps - warmup_steps
    s = min(step - warmup_steps, decay_steps)
    return scaled_lr * 0.5 * (1.0 + math.cos(math.pi * s / decay_steps))


class LinearLayer:
    """tf2/model.py:119-154."""

    def __init__(self, vs, cfg, scope, cin, num_classes, use_bias=True, use_bn=False,
                 name='linear_layer'):
        scope = scope + '/' + name
        dname = vs.namer('dense')
        self.kernel = vs.add('%s/%s/kernel:0' % (scope, dname), (cin, num_classes), 'normal_0.01')
        self.bias = None
        if use_bias and not use_bn:
            self.bias = vs.add('%s/%s/bias:0' % (scope, dname), (num_classes,), 'zeros')
        self.bn = (resnet_lib.BatchNormRelu(vs, cfg, scope, num_classes, relu=False,
                                            center=use_bias) if use_bn else None)
        self.cout = num_classes

    def __call__(self, P, S, x, training, fp32_out=False):
        # `fp32_out` only matters under bf16-storage emulation: the CUDA path keeps the last projection
        # layer and the supervised logits in fp32 (their gradient is still handed to wgrad in bf16).
        assert x.dim() == 2
        x = qb(x) @ qw(P[self.kernel])
        x = qb(x) if fp32_out else q(x)
        if self.bias is not None:
            x = x + P[self.bias]
        if self.bn is not None:
            x = self.bn(P, S, x, training, store=False)    # BN (+ReLU) output is rounded by the caller
        return x


class ProjectionHead:
    """tf2/model.py:157-213 (only 'nonlinear' works in the reference)."""

    def __init__(self, vs, cfg, cin):
        self.cfg = cfg
        self.linear_layers = []
        scope = 'projection_head'
        if cfg.proj_head_mode == 'none':
            pass
        elif cfg.proj_head_mode == 'linear':
            self.linear_layers = [LinearLayer(vs, cfg, scope, cin, cfg.proj_out_dim,
                                              use_bias=False, use_bn=True, name='l_0')]
        elif cfg.proj_head_mode == 'nonlinear':
            for j in range(cfg.num_proj_layers):
                if j != cfg.num_proj_layers - 1:
                    self.linear_layers.append(LinearLayer(
                        vs, cfg, scope, cin, cin, use_bias=True, use_bn=True, name='nl_%d' % j))
                else:
                    self.linear_layers.append(LinearLayer(
                        vs, cfg, scope, cin, cfg.proj_out_dim, use_bias=False, use_bn=True,
                        name='nl_%d' % j))
        else:
            raise ValueError('Unknown head projection mode {}'.format(cfg.proj_head_mode))

    def __call__(self, P, S, x, training):
        cfg = self.cfg
        if cfg.proj_head_mode != 'nonlinear':
            raise ValueError("only proj_head_mode='nonlinear' is callable (reference quirk Q1)")
        hiddens_list = [x]
        for j in range(cfg.num_proj_layers):
            last = j == cfg.num_proj_layers - 1
            h = self.linear_layers[j](P, S, hiddens_list[-1], training, fp32_out=last)
            if not last:
                h = q(F.relu(h))
            hiddens_list.append(h)
        return hiddens_list[-1], hiddens_list[cfg.ft_proj_selector]


class SupervisedHead:
    """tf2/model.py:216-225."""

    def __init__(self, vs, cfg, cin, num_classes):
        self.linear_layer = LinearLayer(vs, cfg, 'head_supervised', cin, num_classes)

    def __call__(self, P, S, x, training):
        return self.linear_layer(P, S, x, training, fp32_out=True)


class Model:
    """tf2/model.py:228-280."""

    def __init__(self, cfg, num_classes):
        self.cfg = cfg
        self.vs = resnet_lib.VarStore()
        self.resnet_model = resnet_lib.resnet(
            self.vs, cfg, cfg.resnet_depth, cfg.width_multiplier,
            cifar_stem=cfg.image_size <= 32)
        self._projection_head = ProjectionHead(self.vs, cfg, self.resnet_model.cout)
        self.supervised_head = None
        if cfg.train_mode == 'finetune' or cfg.lineareval_while_pretraining:
            sel_dim = self.resnet_model.cout  # every hidden layer keeps the input width
            self.supervised_head = SupervisedHead(self.vs, cfg, sel_dim, num_classes)

    def init(self, seed=0, dtype=torch.float32):
        return self.vs.init(seed, dtype)

    def __call__(self, P, S, inputs, training, blur_draws=None, endpoints=None):
        """inputs: [B,H,W,3*T].  `blur_draws`: list per view of (sigma, selector[B])
        standing in for the reference's `tf.random` draws (tf2/data_util.py:407,425)."""
        cfg = self.cfg
        if training and cfg.train_mode == 'pretrain' and cfg.fine_tune_after_block > -1:
            raise ValueError('Does not support layer freezing during pretraining,'
                             'should set fine_tune_after_block<=-1 for safety.')
        num_transforms = inputs.shape[3] // 3
        features_list = list(torch.split(inputs, 3, dim=-1))
        assert len(features_list) == num_transforms
        if cfg.use_blur and training and cfg.train_mode == 'pretrain':
            features_list = data_util.batch_random_blur(
                features_list, cfg.image_size, cfg.image_size, draws=blur_draws)
        features = torch.cat(features_list, 0)          # view-major (tf2/model.py:259)
        hiddens = self.resnet_model(P, S, features, training, endpoints=endpoints)
        proj_out, sup_in = self._projection_head(P, S, hiddens, training)
        if cfg.train_mode == 'finetune':
            return None, self.supervised_head(P, S, sup_in, training)
        elif cfg.train_mode == 'pretrain' and cfg.lineareval_while_pretraining:
            return proj_out, self.supervised_head(P, S, sup_in.detach(), training)
        return proj_out, None


def add_weight_decay(cfg, P, adjust_per_optimizer=True):
    """tf2/model.py:47-69.  `tf.nn.l2_loss(v) = sum(v**2)/2`."""
    if adjust_per_optimizer and 'lars' in cfg.optimizer:
        l2 = [0.5 * (v ** 2).sum() for n, v in P.items()
              if 'head_supervised' in n and 'bias' not in n]
        return cfg.weight_decay * sum(l2) if l2 else 0
    # `model.trainable_weights` (tf2/model.py:64-66): variables of frozen (trainable=False) layers are left out
    frozen = resnet_lib.frozen_variable_names(cfg, P)
    l2 = [0.5 * (v ** 2).sum() for n, v in P.items() if 'batch_normalization' not in n and n not in frozen]
    return cfg.weight_decay * sum(l2)
