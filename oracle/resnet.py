"""Oracle restatement of `tf2/resnet.py` (PyTorch-CPU, autograd gives the grads).

Test infrastructure only -- see `oracle/__init__.py`.  Each class follows the
reference class of the same name; `file:line` citations are relative to
`/root/reference/`.  TensorFlow semantics encoded here:
NHWC activations / HWIO kernels at the API, explicit FixedPadding before
strided VALID convs, TF-'SAME' pooling (pad after), unfused BatchNorm with
biased variance, eps 1e-5, zero-gamma last BN in every block.

Variables live in a flat `OrderedDict` name -> tensor (`P` trainable, `S`
moving statistics).  Names imitate Keras so the LARS name filters
(`tf2/model.py:40-42`) select the same tensors.
"""
from collections import OrderedDict
import math

import torch
import torch.nn.functional as F

from .bf16_emul import q, qb, qw, enabled as _bf16_emulated   # identities unless `emulate_bf16()` is active

BATCH_NORM_EPSILON = 1e-5  # tf2/resnet.py:28


class _Namer:
    """Keras-style unique layer names: conv2d, conv2d_1, ..."""

    def __init__(self):
        self.counts = {}

    def __call__(self, base):
        n = self.counts.get(base, 0)
        self.counts[base] = n + 1
        return base if n == 0 else '%s_%d' % (base, n)


class VarStore:
    """Holds variable specs in creation order (== `model.trainable_variables`)."""

    def __init__(self):
        self.namer = _Namer()
        self.trainable = OrderedDict()   # name -> (shape, init)
        self.moving = OrderedDict()      # name -> (shape, init)

    def add(self, name, shape, init, trainable=True):
        d = self.trainable if trainable else self.moving
        assert name not in d, name
        d[name] = (tuple(shape), init)
        return name

    def init(self, seed=0, dtype=torch.float32):
        """Reference initialisers."""
        g = torch.Generator().manual_seed(seed)
        P, S = OrderedDict(), OrderedDict()
        for name, (shape, init) in self.trainable.items():
            P[name] = _init_tensor(shape, init, g).to(dtype)
        for name, (shape, init) in self.moving.items():
            S[name] = _init_tensor(shape, init, g).to(dtype)
        return P, S


def _init_tensor(shape, init, g):
    if init == 'zeros':
        return torch.zeros(shape, dtype=torch.float64)
    if init == 'ones':
        return torch.ones(shape, dtype=torch.float64)
    if init == 'variance_scaling':
        # tf.keras.initializers.VarianceScaling(): scale 1, fan_in, truncated
        # normal; tf2/resnet.py:202.  fan_in = kh*kw*Cin for HWIO.
        fan_in = 1
        for s in shape[:-1]:
            fan_in *= s
        std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
        t = torch.empty(shape, dtype=torch.float64)
        torch.nn.init.trunc_normal_(t, 0.0, 1.0, -2.0, 2.0, generator=g)
        return t * std
    if init == 'normal_0.01':
        # RandomNormal(stddev=0.01), tf2/model.py:145
        return torch.randn(shape, dtype=torch.float64, generator=g) * 0.01
    raise ValueError(init)


# ----------------------------------------------------------------------------
# Layers
# ----------------------------------------------------------------------------

class BatchNormRelu:
    """tf2/resnet.py:31-78.  (Sync)BatchNormalization + optional ReLU."""

    def __init__(self, vs, cfg, scope, channels, relu=True, init_zero=False,
                 center=True, scale=True):
        self.relu = relu
        self.cfg = cfg
        lname = vs.namer('batch_norm_relu')
        bn = vs.namer('sync_batch_normalization' if cfg.global_bn
                      else 'batch_normalization')
        pre = '%s/%s/%s' % (scope, lname, bn)
        self.gamma = vs.add(pre + '/gamma:0', (channels,),
                            'zeros' if init_zero else 'ones') if scale else None
        self.beta = vs.add(pre + '/beta:0', (channels,), 'zeros') if center else None
        self.mm = vs.add(pre + '/moving_mean:0', (channels,), 'zeros', False)
        self.mv = vs.add(pre + '/moving_variance:0', (channels,), 'ones', False)

    def __call__(self, P, S, x, training, store=True):
        # x: [N, C, H, W] or [N, C].  `store` only matters under bf16-storage emulation: False when
        # the CUDA path fuses this BN with a following add / ReLU and never materialises its output.
        axes = [0] + list(range(2, x.dim()))
        shape = [1, -1] + [1] * (x.dim() - 2)
        if training:
            mean = x.mean(dim=axes)
            var = ((x - mean.view(shape)) ** 2).mean(dim=axes)  # biased (A4)
            with torch.no_grad():
                d = self.cfg.batch_norm_decay
                S[self.mm] = S[self.mm] - (S[self.mm] - mean.detach()) * (1 - d)
                S[self.mv] = S[self.mv] - (S[self.mv] - var.detach()) * (1 - d)
        else:
            mean, var = S[self.mm], S[self.mv]
        inv = torch.rsqrt(var + BATCH_NORM_EPSILON)
        if self.gamma is not None:
            inv = inv * P[self.gamma]
        y = (x - mean.view(shape)) * inv.view(shape)
        if self.beta is not None:
            y = y + P[self.beta].view(shape)
        if self.relu:
            y = F.relu(y)
        return q(y) if store else y


def fixed_padding(x, kernel_size):
    """tf2/resnet.py:160-180 on NCHW tensors."""
    pad_total = kernel_size - 1
    pad_beg = pad_total // 2
    pad_end = pad_total - pad_beg
    return F.pad(x, (pad_beg, pad_end, pad_beg, pad_end))


class Conv2dFixedPadding:
    """tf2/resnet.py:183-208.  stride 1 -> 'SAME'; stride > 1 -> pad + 'VALID'."""

    def __init__(self, vs, scope, cin, filters, kernel_size, strides):
        self.k, self.s = kernel_size, strides
        lname = vs.namer('conv2d_fixed_padding')
        cname = vs.namer('conv2d')
        self.kernel = vs.add('%s/%s/%s/kernel:0' % (scope, lname, cname),
                             (kernel_size, kernel_size, cin, filters),
                             'variance_scaling')
        self.cout = filters

    def __call__(self, P, S, x, training):
        w = qw(P[self.kernel]).permute(3, 2, 0, 1)  # HWIO -> OIHW
        x = qb(x)
        if self.s > 1:
            x = fixed_padding(x, self.k)
            return q(F.conv2d(x, w, stride=self.s))
        return q(F.conv2d(x, w, padding=(self.k - 1) // 2))


class PlainConv1x1:
    """`tf.keras.layers.Conv2D(k=1)` inside SK/SE (tf2/resnet.py:238-253,285-301)."""

    def __init__(self, vs, scope, cin, filters, use_bias=False):
        cname = vs.namer('conv2d')
        self.kernel = vs.add('%s/%s/kernel:0' % (scope, cname),
                             (1, 1, cin, filters), 'variance_scaling')
        self.bias = vs.add('%s/%s/bias:0' % (scope, cname), (filters,),
                           'zeros') if use_bias else None

    def __call__(self, P, S, x, training):
        y = F.conv2d(x, P[self.kernel].permute(3, 2, 0, 1))
        if self.bias is not None:
            y = y + P[self.bias].view(1, -1, 1, 1)
        return y


class SK_Conv2D:
    """tf2/resnet.py:217-277."""

    def __init__(self, vs, cfg, scope, cin, filters, strides, sk_ratio, min_dim=32):
        scope = scope + '/' + vs.namer('sk_conv2d')
        self.filters = filters
        self.conv = Conv2dFixedPadding(vs, scope, cin, 2 * filters, 3, strides)
        self.bn = BatchNormRelu(vs, cfg, scope, 2 * filters)
        mid_dim = max(int(filters * sk_ratio), min_dim)
        self.conv0 = PlainConv1x1(vs, scope, filters, mid_dim)
        self.bn1 = BatchNormRelu(vs, cfg, scope, mid_dim)
        self.conv1 = PlainConv1x1(vs, scope, mid_dim, 2 * filters)

    def __call__(self, P, S, x, training):
        assert not _bf16_emulated(), 'bf16-storage emulation does not cover SK blocks'
        x = self.conv(P, S, x, training)
        x = self.bn(P, S, x, training)
        streams = torch.stack(torch.split(x, self.filters, dim=1))    # [2,N,f,H,W]
        g = streams.sum(0).mean(dim=(2, 3), keepdim=True)              # [N,f,1,1]
        g = self.bn1(P, S, self.conv0(P, S, g, training), training)
        mixing = self.conv1(P, S, g, training)
        mixing = torch.stack(torch.split(mixing, self.filters, dim=1))
        mixing = torch.softmax(mixing, dim=0)
        return (streams * mixing).sum(0)


class SE_Layer:
    """tf2/resnet.py:280-311 (expand width taken from the input, Q9)."""

    def __init__(self, vs, scope, cin, filters, se_ratio):
        scope = scope + '/' + vs.namer('se_layer')
        self.reduce = PlainConv1x1(vs, scope, cin, max(1, int(filters * se_ratio)), True)
        self.expand = PlainConv1x1(vs, scope, max(1, int(filters * se_ratio)), cin, True)

    def __call__(self, P, S, x, training):
        assert not _bf16_emulated(), 'bf16-storage emulation does not cover SE blocks'
        t = x.mean(dim=(2, 3), keepdim=True)
        t = self.expand(P, S, F.relu(self.reduce(P, S, t, training)), training)
        return torch.sigmoid(t) * x


def avg_pool_2x2(x, strides):
    """ResNet-D shortcut pooling, tf2/resnet.py:333-340,401-408."""
    if strides > 1:
        x = fixed_padding(x, 2)                      # pad (0,1)
        return F.avg_pool2d(x, 2, strides)           # 'VALID'
    # 'SAME' with stride 1: pad after by 1, divisor counts valid elements only
    ones = torch.ones_like(x[:1, :1])
    num = F.avg_pool2d(F.pad(x, (0, 1, 0, 1)), 2, 1) * 4
    den = F.avg_pool2d(F.pad(ones, (0, 1, 0, 1)), 2, 1) * 4
    return num / den


class _Shortcut:
    def __init__(self, vs, cfg, scope, cin, filters_out, strides):
        self.cfg, self.strides = cfg, strides
        if cfg.sk_ratio > 0:
            self.conv = Conv2dFixedPadding(vs, scope, cin, filters_out, 1, 1)
        else:
            self.conv = Conv2dFixedPadding(vs, scope, cin, filters_out, 1, strides)
        self.bn = BatchNormRelu(vs, cfg, scope, filters_out, relu=False)

    def __call__(self, P, S, x, training):
        if self.cfg.sk_ratio > 0:
            x = avg_pool_2x2(x, self.strides)
        return self.bn(P, S, self.conv(P, S, x, training), training)


class ResidualBlock:
    """tf2/resnet.py:314-382."""

    def __init__(self, vs, cfg, scope, cin, filters, strides, use_projection=False):
        scope = scope + '/' + vs.namer('residual_block')
        self.cfg = cfg
        self.shortcut = (_Shortcut(vs, cfg, scope, cin, filters, strides)
                         if use_projection else None)
        self.c1 = Conv2dFixedPadding(vs, scope, cin, filters, 3, strides)
        self.b1 = BatchNormRelu(vs, cfg, scope, filters)
        self.c2 = Conv2dFixedPadding(vs, scope, filters, filters, 3, 1)
        self.b2 = BatchNormRelu(vs, cfg, scope, filters, relu=False, init_zero=True)
        self.se = SE_Layer(vs, scope, filters, filters, cfg.se_ratio) if cfg.se_ratio > 0 else None
        self.cout = filters

    def __call__(self, P, S, x, training):
        shortcut = x if self.shortcut is None else self.shortcut(P, S, x, training)
        x = self.b1(P, S, self.c1(P, S, x, training), training)
        x = self.b2(P, S, self.c2(P, S, x, training), training, store=self.se is not None)
        if self.se is not None:
            x = self.se(P, S, x, training)
        return q(F.relu(x + shortcut))


class BottleneckBlock:
    """tf2/resnet.py:385-487 (DropBlock is dead code)."""

    def __init__(self, vs, cfg, scope, cin, filters, strides, use_projection=False):
        scope = scope + '/' + vs.namer('bottleneck_block')
        self.cfg = cfg
        self.shortcut = (_Shortcut(vs, cfg, scope, cin, 4 * filters, strides)
                         if use_projection else None)
        self.c1 = Conv2dFixedPadding(vs, scope, cin, filters, 1, 1)
        self.b1 = BatchNormRelu(vs, cfg, scope, filters)
        if cfg.sk_ratio > 0:
            self.sk = SK_Conv2D(vs, cfg, scope, filters, filters, strides, cfg.sk_ratio)
        else:
            self.sk = None
            self.c2 = Conv2dFixedPadding(vs, scope, filters, filters, 3, strides)
            self.b2 = BatchNormRelu(vs, cfg, scope, filters)
        self.c3 = Conv2dFixedPadding(vs, scope, filters, 4 * filters, 1, 1)
        self.b3 = BatchNormRelu(vs, cfg, scope, 4 * filters, relu=False, init_zero=True)
        # tf2/resnet.py:474-476 builds SE with `filters`; expand width follows the input.
        self.se = SE_Layer(vs, scope, 4 * filters, filters, cfg.se_ratio) if cfg.se_ratio > 0 else None
        self.cout = 4 * filters

    def __call__(self, P, S, x, training):
        shortcut = x if self.shortcut is None else self.shortcut(P, S, x, training)
        x = self.b1(P, S, self.c1(P, S, x, training), training)
        if self.sk is not None:
            x = self.sk(P, S, x, training)
        else:
            x = self.b2(P, S, self.c2(P, S, x, training), training)
        x = self.b3(P, S, self.c3(P, S, x, training), training, store=self.se is not None)
        if self.se is not None:
            x = self.se(P, S, x, training)
        return q(F.relu(x + shortcut))


class BlockGroup:
    """tf2/resnet.py:490-526: first block always uses a projection shortcut."""

    def __init__(self, vs, cfg, scope, cin, filters, block_fn, blocks, strides, name):
        scope = scope + '/' + name
        self.layers = [block_fn(vs, cfg, scope, cin, filters, strides, use_projection=True)]
        for _ in range(1, blocks):
            self.layers.append(block_fn(vs, cfg, scope, self.layers[-1].cout, filters, 1))
        self.cout = self.layers[-1].cout

    def __call__(self, P, S, x, training):
        for layer in self.layers:
            x = layer(P, S, x, training)
        return x


def max_pool_3x3_s2_same(x):
    """MaxPooling2D(3, 2, 'SAME') (tf2/resnet.py:605-611): pad after only (A3)."""
    h, w = x.shape[2], x.shape[3]
    oh, ow = -(-h // 2), -(-w // 2)
    ph = max((oh - 1) * 2 + 3 - h, 0)
    pw = max((ow - 1) * 2 + 3 - w, 0)
    x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2), value=float('-inf'))
    return F.max_pool2d(x, 3, 2)


MODEL_PARAMS = {  # tf2/resnet.py:709-734
    18: (ResidualBlock, [2, 2, 2, 2]), 34: (ResidualBlock, [3, 4, 6, 3]),
    50: (BottleneckBlock, [3, 4, 6, 3]), 101: (BottleneckBlock, [3, 4, 23, 3]),
    152: (BottleneckBlock, [3, 8, 36, 3]), 200: (BottleneckBlock, [3, 24, 36, 3]),
}


class Resnet:
    """tf2/resnet.py:529-699."""

    def __init__(self, vs, cfg, block_fn, layers, width_multiplier, cifar_stem=False):
        scope = 'resnet'
        self.cfg = cfg
        wm = width_multiplier
        self.stem = []
        self.cifar_stem = cifar_stem
        if cifar_stem:                                         # :551-564
            self.stem.append(Conv2dFixedPadding(vs, scope, 3, 64 * wm, 3, 1))
            self.stem.append(BatchNormRelu(vs, cfg, scope, 64 * wm))
        elif cfg.sk_ratio > 0:                                 # ResNet-D stem :566-591
            self.stem.append(Conv2dFixedPadding(vs, scope, 3, 64 * wm // 2, 3, 2))
            self.stem.append(BatchNormRelu(vs, cfg, scope, 64 * wm // 2))
            self.stem.append(Conv2dFixedPadding(vs, scope, 64 * wm // 2, 64 * wm // 2, 3, 1))
            self.stem.append(BatchNormRelu(vs, cfg, scope, 64 * wm // 2))
            self.stem.append(Conv2dFixedPadding(vs, scope, 64 * wm // 2, 64 * wm, 3, 1))
            self.stem.append(BatchNormRelu(vs, cfg, scope, 64 * wm))
        else:                                                  # :593-604
            self.stem.append(Conv2dFixedPadding(vs, scope, 3, 64 * wm, 7, 2))
            self.stem.append(BatchNormRelu(vs, cfg, scope, 64 * wm))
        self.groups = []
        cin = 64 * wm
        for i, (f, s) in enumerate(zip([64, 128, 256, 512], [1, 2, 2, 2])):
            g = BlockGroup(vs, cfg, scope, cin, f * wm, block_fn, layers[i], s,
                           'block_group%d' % (i + 1))
            self.groups.append(g)
            cin = g.cout
        self.cout = cin

    def __call__(self, P, S, x_nhwc, training, endpoints=None):
        cfg = self.cfg
        # Finetuning with fine_tune_after_block = k >= 0 (tf2/resnet.py:548-549,619-692): the stem and block groups
        # 1..k are `trainable=False` Keras layers -- their BatchNorm runs in inference mode [TF semantics] -- and a
        # tf.stop_gradient sits in front of group k+1 (:675-681).  The caller leaves the frozen variables out of
        # the optimizer (`frozen_variable_names`).
        ft = cfg.fine_tune_after_block if cfg.train_mode == 'finetune' else -1
        train_all = training
        training = train_all and ft < 0
        x = q(x_nhwc.permute(0, 3, 1, 2))      # (emulation: the network input is cast to bf16)
        for layer in self.stem:
            x = layer(P, S, x, training)
            if endpoints is not None and isinstance(layer, Conv2dFixedPadding):
                endpoints['initial_conv'] = x
        if not self.cifar_stem:
            x = qb(max_pool_3x3_s2_same(x))
        if endpoints is not None:
            endpoints['initial_max_pool'] = x
        for i, g in enumerate(self.groups):
            if ft == i:
                x = x.detach()                                  # tf.stop_gradient, :675-677
            x = g(P, S, x, train_all and not (ft >= 0 and i < ft))
            if endpoints is not None:
                endpoints['block_group%d' % (i + 1)] = x
        if ft == 4:
            x = x.detach()                                      # :680-681
        x = q(x.mean(dim=(2, 3)))                               # :693-696
        if endpoints is not None:
            endpoints['final_avg_pool'] = x
        return x


def frozen_variable_names(cfg, names):
    """Variables of `trainable=False` layers in a finetuning run (tf2/resnet.py:548-549,619-692)."""
    ft = cfg.fine_tune_after_block if cfg.train_mode == 'finetune' else -1
    if ft < 0:
        return set()
    out = set()
    for n in names:
        if not n.startswith('resnet/'):
            continue
        grp = [i for i in range(1, 5) if '/block_group%d/' % i in n]
        if not grp or grp[0] <= ft:
            out.add(n)
    return out


def resnet(vs, cfg, resnet_depth, width_multiplier, cifar_stem=False):
    """tf2/resnet.py:702-747."""
    if resnet_depth not in MODEL_PARAMS:
        raise ValueError('Not a valid resnet_depth:', resnet_depth)
    block_fn, layers = MODEL_PARAMS[resnet_depth]
    return Resnet(vs, cfg, block_fn, layers, width_multiplier, cifar_stem=cifar_stem)
