"""Flag defaults of the reference (`tf2/run.py:37-238`) as a plain namespace.

The oracle never touches absl; every function takes a `cfg` object carrying the
flag values it reads (the reference reads the process-global `FLAGS`).
"""
from types import SimpleNamespace

_DEFAULTS = dict(
    learning_rate=0.3,              # tf2/run.py:37
    learning_rate_scaling='linear', # :41
    warmup_epochs=10,               # :45
    weight_decay=1e-6,              # :49
    batch_norm_decay=0.9,           # :51
    train_batch_size=512,           # :55
    train_epochs=100,               # :63
    train_steps=0,                  # :67
    train_mode='pretrain',          # :106
    lineareval_while_pretraining=True,  # :110
    fine_tune_after_block=-1,       # :122
    optimizer='lars',               # :163
    momentum=0.9,                   # :167
    temperature=0.1,                # :183
    hidden_norm=True,               # :187
    proj_head_mode='nonlinear',     # :191
    proj_out_dim=128,               # :195
    num_proj_layers=3,              # :199
    ft_proj_selector=0,             # :203
    global_bn=True,                 # :208
    width_multiplier=1,             # :212
    resnet_depth=50,                # :216
    sk_ratio=0.,                    # :220
    se_ratio=0.,                    # :224
    image_size=224,                 # :228
    color_jitter_strength=1.0,      # :232
    use_blur=True,                  # :236
)


def default_cfg(**overrides):
    d = dict(_DEFAULTS)
    for k in overrides:
        if k not in d:
            raise KeyError('unknown flag %r' % k)
    d.update(overrides)
    return SimpleNamespace(**d)
