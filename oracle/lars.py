"""Oracle restatement of `tf2/lars_optimizer.py:83-157`.  Test infrastructure only."""
import re

import torch

EETA_DEFAULT = 0.001  # tf2/lars_optimizer.py:22


def _use_weight_decay(name, weight_decay, exclude):
    """tf2/lars_optimizer.py:139-148."""
    if not weight_decay:
        return False
    for r in exclude or []:
        if re.search(r, name) is not None:
            return False
    return True


def _do_layer_adaptation(name, exclude):
    """tf2/lars_optimizer.py:150-157."""
    for r in exclude or []:
        if re.search(r, name) is not None:
            return False
    return True


def lars_apply(P, G, V, learning_rate, momentum=0.9, use_nesterov=False, weight_decay=0.0,
               exclude_from_weight_decay=None, exclude_from_layer_adaptation=None,
               classic_momentum=True, eeta=EETA_DEFAULT):
    """One `apply_gradients` over dicts name->tensor.  Returns (new_P, new_V).

    Follows `_resource_apply_dense` line by line; `learning_rate` is `lr_t`, the
    schedule evaluated at the pre-increment iteration."""
    if not exclude_from_layer_adaptation:
        exclude_from_layer_adaptation = exclude_from_weight_decay
    new_P, new_V = {}, {}
    for name, param in P.items():
        grad = G[name]
        v = V[name]
        if _use_weight_decay(name, weight_decay, exclude_from_weight_decay):
            grad = grad + weight_decay * param
        if classic_momentum:
            trust_ratio = 1.0
            if _do_layer_adaptation(name, exclude_from_layer_adaptation):
                w_norm = torch.linalg.vector_norm(param)
                g_norm = torch.linalg.vector_norm(grad)
                if w_norm > 0 and g_norm > 0:
                    trust_ratio = eeta * w_norm / g_norm
            scaled_lr = learning_rate * trust_ratio
            next_v = momentum * v + scaled_lr * grad
            update = momentum * next_v + scaled_lr * grad if use_nesterov else next_v
            next_param = param - update
        else:
            next_v = momentum * v + grad
            update = momentum * next_v + grad if use_nesterov else next_v
            trust_ratio = 1.0
            if _do_layer_adaptation(name, exclude_from_layer_adaptation):
                w_norm = torch.linalg.vector_norm(param)
                v_norm = torch.linalg.vector_norm(update)
                if w_norm > 0 and v_norm > 0:
                    trust_ratio = eeta * w_norm / v_norm
            scaled_lr = trust_ratio * learning_rate
            next_param = param - scaled_lr * update
        new_P[name], new_V[name] = next_param, next_v
    return new_P, new_V


LARS_EXCLUDE = ['batch_normalization', 'bias', 'head_supervised']  # tf2/model.py:40-42
