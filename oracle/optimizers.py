"""Oracle restatement of the two other optimizers `build_optimizer` can return (tf2/model.py:31-34):
`tf.keras.optimizers.SGD(lr, momentum, nesterov=True)` and `tf.keras.optimizers.Adam(lr)` [TF semantics:
Keras update rules, Adam epsilon 1e-7, bias correction folded into the step size].  Test infrastructure only."""
import math
from collections import OrderedDict

import torch


def sgd_nesterov_apply(P, G, V, learning_rate, momentum=0.9, nesterov=True):
    """v <- m v - lr g;  w <- w + m v - lr g  (nesterov)  |  w <- w + v.  Returns (new_P, new_V)."""
    Pn, Vn = OrderedDict(), OrderedDict()
    for k in P:
        v = momentum * V[k] - learning_rate * G[k]
        Vn[k] = v
        Pn[k] = P[k] + (momentum * v - learning_rate * G[k] if nesterov else v)
    return Pn, Vn


def adam_apply(P, G, M, V, learning_rate, step, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
    """`step` is 1-based.  Returns (new_P, new_M, new_V)."""
    lr_t = learning_rate * math.sqrt(1.0 - beta_2 ** step) / (1.0 - beta_1 ** step)
    Pn, Mn, Vn = OrderedDict(), OrderedDict(), OrderedDict()
    for k in P:
        m = M[k] + (G[k] - M[k]) * (1.0 - beta_1)
        v = V[k] + (G[k] * G[k] - V[k]) * (1.0 - beta_2)
        Mn[k], Vn[k] = m, v
        Pn[k] = P[k] - lr_t * m / (torch.sqrt(v) + epsilon)
    return Pn, Mn, Vn
