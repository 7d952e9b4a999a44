import torch


def rel_err(a, b):
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    den = b.norm().item()
    if den == 0:
        return (a - b).norm().item()
    return (a - b).norm().item() / den


def cfg_from_flags(F):
    """Namespace for the oracle carrying the current FLAGS values."""
    from oracle.config import default_cfg, _DEFAULTS
    return default_cfg(**{k: getattr(F, k) for k in _DEFAULTS})


def structured_batch(B, S, num_classes=1000, seed=0):
    """A well-conditioned synthetic batch of the reference's input contract (tf2/data.py:52-62):
    [B,S,S,6] fp32 in [0,1] -- two views per sample -- plus one-hot labels.

    Each sample is its own smooth image (a few random 2-D sinusoids per colour channel over a
    per-sample base colour); its two views are different crops (scale / shift of the coordinate
    grid, optional flip) with a brightness / contrast jitter, i.e. what `preprocess_for_train`
    hands the model on real data.  Unlike i.i.d.-noise pixels (which every view averages to the
    same feature vector, so that the batch statistics of the deeper layers divide by almost
    nothing) the samples stay distinct through the network and positives stay similar: the
    step is then well conditioned and rounding is not amplified."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, S), torch.linspace(0, 1, S), indexing='ij')
    views = [[], []]
    for _ in range(B):
        K = 4
        freq = torch.rand(3, K, 2, generator=g) * 5.0 + 0.5
        phase = torch.rand(3, K, generator=g) * 6.2831853
        amp = torch.rand(3, K, generator=g) * 0.25
        base = torch.rand(3, generator=g) * 0.6 + 0.2
        for v in range(2):
            scale = 0.5 + 0.5 * torch.rand(1, generator=g).item()
            ox = torch.rand(1, generator=g).item() * (1 - scale)
            oy = torch.rand(1, generator=g).item() * (1 - scale)
            flip = torch.rand(1, generator=g).item() < 0.5
            x = ox + scale * (1 - xx if flip else xx)
            y = oy + scale * yy
            img = torch.empty(S, S, 3)
            for c in range(3):
                t = torch.zeros(S, S)
                for k in range(K):
                    t = t + amp[c, k] * torch.sin(6.2831853 * (freq[c, k, 0] * x + freq[c, k, 1] * y) + phase[c, k])
                img[..., c] = base[c] + t
            bright = 0.8 + 0.4 * torch.rand(1, generator=g).item()
            contrast = 0.8 + 0.4 * torch.rand(1, generator=g).item()
            m = img.mean(dim=(0, 1), keepdim=True)
            img = ((img - m) * contrast + m) * bright
            views[v].append(img.clamp(0, 1))
    f = torch.cat([torch.stack(views[0]), torch.stack(views[1])], dim=-1).contiguous()
    lab = torch.nn.functional.one_hot(torch.randint(0, num_classes, (B,), generator=g), num_classes).float()
    return f, lab
