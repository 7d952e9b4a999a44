"""Oracle restatement of `single_step` (`tf2/run.py:557-622`) over R simulated
replicas.  Test infrastructure only -- never imported by `simclr_b200/`.
"""
from collections import OrderedDict

import torch

from . import model as model_lib
from . import objective as obj_lib
from . import lars as lars_lib


def forward_backward(model, P, S, features, labels, blur_draws=None):
    """Loss and gradients of one synchronous step.

    features: list over replicas of [B,H,W,6]; labels: list of one-hot [B,classes]
    (or None).  Returns dict(loss, con_loss, sup_loss, wd, grads, logits_con,
    labels_con, S_new, proj_out, sup_out).

    With `global_bn` the backbone runs once over the concatenation of all
    replicas' view-major batches (SyncBN == BN over the global batch);
    otherwise once per replica.  Per-replica loss is divided by R and summed
    (tf2/run.py:617 + cross-replica gradient SUM).
    """
    cfg = model.cfg
    R = len(features)
    Pg = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in P.items())
    Sn = OrderedDict((k, v.clone()) for k, v in S.items())
    proj, sup = [], []
    if cfg.global_bn and R > 1:
        # view-major inside each replica; order across replicas is irrelevant to BN
        B = features[0].shape[0]
        views = []
        for r in range(R):
            fl = list(torch.split(features[r], 3, dim=-1))
            if cfg.use_blur:
                from . import data_util
                fl = data_util.batch_random_blur(fl, cfg.image_size, cfg.image_size,
                                                 draws=None if blur_draws is None else blur_draws[r])
            views.append(torch.cat(fl, 0))
        x = torch.cat(views, 0)                                    # [R*2B,H,W,3]
        hid = model.resnet_model(Pg, Sn, x, True)
        po, si = model._projection_head(Pg, Sn, hid, True)
        so = model.supervised_head(Pg, Sn, si.detach(), True) if model.supervised_head and \
            cfg.lineareval_while_pretraining else None
        for r in range(R):
            proj.append(po[r * 2 * B:(r + 1) * 2 * B])
            sup.append(None if so is None else so[r * 2 * B:(r + 1) * 2 * B])
    else:
        for r in range(R):
            po, so = model(Pg, Sn, features[r], True,
                           blur_draws=None if blur_draws is None else blur_draws[r])
            proj.append(po)
            sup.append(so)
    if cfg.train_mode == 'finetune':
        con = [None] * R                                          # supervised loss only (tf2/model.py:267-270)
    else:
        con = obj_lib.contrastive_loss_replicas(proj, cfg.hidden_norm, cfg.temperature)
    total = 0.
    con_losses, sup_losses = [], []
    wd = model_lib.add_weight_decay(cfg, Pg, adjust_per_optimizer=True)
    for r in range(R):
        loss = 0.
        if con[r] is not None:
            loss = con[r][0]
            con_losses.append(con[r][0].detach())
        if sup[r] is not None:
            # tf2/run.py:600-602: the labels are doubled only when pretraining with the linear-eval head
            l = labels[r]
            if cfg.train_mode == 'pretrain' and cfg.lineareval_while_pretraining:
                l = torch.cat([labels[r], labels[r]], 0)
            sl = obj_lib.add_supervised_loss(l, sup[r])
            sup_losses.append(sl.detach())
            loss = loss + sl
        loss = loss + wd
        total = total + loss / R
    grads = torch.autograd.grad(total, list(Pg.values()), allow_unused=True)
    G = OrderedDict()
    for (k, v), g in zip(Pg.items(), grads):
        G[k] = torch.zeros_like(v) if g is None else g
    return dict(loss=total.detach(), con_loss=con_losses, sup_loss=sup_losses,
                wd=wd.detach() if torch.is_tensor(wd) else wd, grads=G,
                logits_con=[None if c is None else c[1].detach() for c in con],
                labels_con=[None if c is None else c[2] for c in con],
                S_new=Sn, proj_out=[None if p is None else p.detach() for p in proj],
                sup_out=[None if s is None else s.detach() for s in sup])


def single_step(model, P, S, V, features, labels, learning_rate, blur_draws=None):
    """forward/backward + LARS apply.  Returns (P_new, S_new, V_new, info)."""
    cfg = model.cfg
    info = forward_backward(model, P, S, features, labels, blur_draws)
    P_new, V_new = lars_lib.lars_apply(
        P, info['grads'], V, learning_rate, momentum=cfg.momentum,
        weight_decay=cfg.weight_decay, exclude_from_weight_decay=lars_lib.LARS_EXCLUDE)
    return OrderedDict(P_new), info['S_new'], OrderedDict(V_new), info
