"""Device-side data augmentation of the reference (`tf2/data_util.py`).

`prepare_views` is the fused input stage of `Model.__call__`
(tf2/model.py:250-259): channel split into views, `batch_random_blur`
(tf2/data_util.py:413-440), batch concat (view-major), cast to the activation
dtype and pad 3 -> 4 channels, in two passes over the images.
"""
import torch

from ._lib import lib, stream_ptr
from .engine import get_engine


def draw_blur(num_views, batch, device, blur_probability=0.5, generator=None):
    """The `tf.random` draws of `batch_random_blur`: one sigma ~ U[0.1, 2.0] per view
    shared by the batch (tf2/data_util.py:407) and a Bernoulli(p)
    selector per sample (tf2/data_util.py:425-430)."""
    sigma = torch.empty(num_views, dtype=torch.float32, device=device).uniform_(0.1, 2.0, generator=generator)
    selector = (torch.rand(num_views, batch, device=device, generator=generator) < blur_probability).to(torch.uint8)
    return sigma, selector


def prepare_views(inputs, num_transforms, use_blur, image_size, draws=None):
    """inputs [B,H,W,3*T] fp32 in [0,1] -> [T*B,H,W,4] activation dtype."""
    e = get_engine()
    assert inputs.dtype == torch.float32
    inputs = inputs.contiguous()
    B, H, W, C = inputs.shape
    T = num_transforms
    assert C == 3 * T
    out = e.empty((T * B, H, W, 4))
    sigma = selector = tmp = None
    if use_blur:
        if draws is None:
            sigma, selector = draw_blur(T, B, e.device)
        else:
            sigma, selector = draws
            sigma = torch.as_tensor(sigma, dtype=torch.float32, device=e.device).contiguous()
            selector = torch.as_tensor(selector, device=e.device).to(torch.uint8).contiguous()
        tmp = e.empty((T * B, H, W, 3), torch.float32)
    lib.input_prep(inputs, out, e.code(out.dtype), B, H, W, T, int(use_blur), image_size // 10,
                   sigma, selector, tmp, stream_ptr())
    return out


def batch_random_blur(images_list, height, width, blur_probability=0.5, draws=None):
    """tf2/data_util.py:413-440 on a list of [B,H,W,3] fp32 views; returns fp32 views."""
    e = get_engine()
    T = len(images_list)
    B, H, W, _ = images_list[0].shape
    x = torch.cat(images_list, dim=-1).contiguous()
    if draws is None:
        draws = draw_blur(T, B, e.device, blur_probability)
    sigma = torch.as_tensor(draws[0], dtype=torch.float32, device=e.device).contiguous()
    selector = torch.as_tensor(draws[1], device=e.device).to(torch.uint8).contiguous()
    out = e.empty((T * B, H, W, 4), torch.float32)
    tmp = e.empty((T * B, H, W, 3), torch.float32)
    lib.input_prep(x, out, 0, B, H, W, T, 1, height // 10, sigma, selector, tmp, stream_ptr())
    return [out[t * B:(t + 1) * B, :, :, :3] for t in range(T)]


def preprocess_for_train_batch(images, draws, height, width, out=None, channel_offset=0):
    """`preprocess_for_train` (tf2/data_util.py:443-475) for a batch of uint8 images on device.

    images: list of uint8 tensors [Hs,Ws,3] (any device); draws: list of dicts with keys
    box=(y,x,h,w), flip, color=dict(apply_jitter, perm, brightness, contrast, saturation, hue,
    apply_gray) -- the reference's tf.random draws.  Returns fp32 [n,height,width,3], or writes
    into `out` [n,height,width,C] at `channel_offset` (two views -> [B,H,W,6], tf2/data.py:55-58)."""
    import numpy as np
    e = get_engine()
    n = len(images)
    offs, hw, total = [], [], 0
    for im in images:
        assert im.dtype == torch.uint8 and im.dim() == 3 and im.shape[2] == 3
        offs.append(total); hw.append([im.shape[0], im.shape[1]]); total += im.numel()
    src = torch.cat([im.reshape(-1) for im in images]).to(e.device)
    box = np.asarray([d['box'] for d in draws], dtype=np.int32)
    for (y, x, h, w), (Hs, Ws) in zip(box, hw):
        if not (0 <= y and 0 <= x and h > 0 and w > 0 and y + h <= Hs and x + w <= Ws):
            raise ValueError('crop box out of the image')
    flip = np.asarray([1 if d['flip'] else 0 for d in draws], dtype=np.uint8)
    col = np.zeros((n, 8), dtype=np.float32)
    for i, d in enumerate(draws):
        c = d['color']
        code = sum(int(op) << (2 * t) for t, op in enumerate(c['perm']))
        col[i] = [1.0 if c['apply_jitter'] else 0.0, code, c['brightness'], c['contrast'], c['saturation'], c['hue'],
                  1.0 if c['apply_gray'] else 0.0, 0.0]
    t = lambda a: torch.from_numpy(a).to(e.device)
    if out is None:
        out = e.empty((n, height, width, 3), torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[:3] == (n, height, width)
    lib.augment(src, t(np.asarray(offs, dtype=np.int64)), t(np.asarray(hw, dtype=np.int32)), t(box), t(flip), t(col),
                out, n, height, width, out.shape[3], channel_offset, stream_ptr())
    return out
