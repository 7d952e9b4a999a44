"""Oracle restatement of `tf2/objective.py`.  Test infrastructure only.

`strategy` stands in for `tf.distribute.Strategy`: a `SimStrategy` carries the
replica id and the hidden vectors of *all* simulated replicas, so that
`tpu_cross_replica_concat` (scatter into zeros + all-reduce SUM,
tf2/objective.py:114-122) is reproduced as a differentiable concat -- which is
exactly what the all-reduce computes, including its backward.
"""
import torch

LARGE_NUM = 1e9  # tf2/objective.py:24


class SimStrategy:
    def __init__(self, num_replicas, replica_id, all_hidden1, all_hidden2):
        self.num_replicas_in_sync = num_replicas
        self.replica_id = replica_id
        self.all_hidden1 = all_hidden1   # list over replicas of [B, D]
        self.all_hidden2 = all_hidden2


def l2_normalize(x, axis=-1, epsilon=1e-12):
    """tf.math.l2_normalize: x * rsqrt(max(sum(x^2), eps))  (SURVEY A6)."""
    sq = (x * x).sum(dim=axis, keepdim=True)
    return x * torch.rsqrt(torch.clamp(sq, min=epsilon))


def add_supervised_loss(labels, logits):
    """tf2/objective.py:27-32: mean categorical CE from logits."""
    lse = torch.logsumexp(logits, dim=-1)
    return (lse - (labels * logits).sum(-1)).mean()


def softmax_cross_entropy_with_logits(labels, logits):
    lse = torch.logsumexp(logits, dim=-1)
    return lse - (labels * logits).sum(-1)


def labels_and_masks(batch_size, replica_id, enlarged_batch_size, dtype):
    """tf2/objective.py:64-69: integer index construction (bit-exact contract)."""
    labels_idx = torch.arange(batch_size, dtype=torch.int64) + replica_id * batch_size
    labels = torch.nn.functional.one_hot(labels_idx, enlarged_batch_size * 2).to(dtype)
    masks = torch.nn.functional.one_hot(labels_idx, enlarged_batch_size).to(dtype)
    return labels_idx, labels, masks


def add_contrastive_loss(hidden, hidden_norm=True, temperature=1.0, strategy=None):
    """tf2/objective.py:35-89.  Returns (loss, logits_ab, labels)."""
    if hidden_norm:
        hidden = l2_normalize(hidden, -1)
    hidden1, hidden2 = torch.split(hidden, hidden.shape[0] // 2, 0)
    batch_size = hidden1.shape[0]
    if strategy is not None:
        hidden1_large = tpu_cross_replica_concat(hidden1, strategy, which=1)
        hidden2_large = tpu_cross_replica_concat(hidden2, strategy, which=2)
        enlarged_batch_size = hidden1_large.shape[0]
        _, labels, masks = labels_and_masks(batch_size, strategy.replica_id,
                                            enlarged_batch_size, hidden.dtype)
    else:
        hidden1_large, hidden2_large = hidden1, hidden2
        _, labels, masks = labels_and_masks(batch_size, 0, batch_size, hidden.dtype)

    logits_aa = hidden1 @ hidden1_large.t() / temperature
    logits_aa = logits_aa - masks * LARGE_NUM
    logits_bb = hidden2 @ hidden2_large.t() / temperature
    logits_bb = logits_bb - masks * LARGE_NUM
    logits_ab = hidden1 @ hidden2_large.t() / temperature
    logits_ba = hidden2 @ hidden1_large.t() / temperature

    loss_a = softmax_cross_entropy_with_logits(labels, torch.cat([logits_ab, logits_aa], 1))
    loss_b = softmax_cross_entropy_with_logits(labels, torch.cat([logits_ba, logits_bb], 1))
    loss = (loss_a + loss_b).mean()
    return loss, logits_ab, labels


def tpu_cross_replica_concat(tensor, strategy=None, which=1):
    """tf2/objective.py:92-127.  In the simulation the other replicas' tensors
    are the live autograd tensors of those replicas, so gradients flow back to
    them exactly as they do through TF's all-reduce."""
    if strategy is None or strategy.num_replicas_in_sync <= 1:
        return tensor
    src = strategy.all_hidden1 if which == 1 else strategy.all_hidden2
    parts = [tensor if r == strategy.replica_id else src[r]
             for r in range(strategy.num_replicas_in_sync)]
    return torch.cat(parts, 0)


def contrastive_loss_replicas(hiddens, hidden_norm=True, temperature=1.0):
    """Runs `add_contrastive_loss` on R simulated replicas.

    hiddens: list over replicas of [2B, D].  Returns the list of per-replica
    (loss, logits_ab, labels).  The training loss of the job is
    sum_r loss_r / R (tf2/run.py:617 + gradient SUM across replicas)."""
    R = len(hiddens)
    if R == 1:
        return [add_contrastive_loss(hiddens[0], hidden_norm, temperature, None)]
    normed = [l2_normalize(h, -1) if hidden_norm else h for h in hiddens]
    h1 = [torch.split(h, h.shape[0] // 2, 0)[0] for h in normed]
    h2 = [torch.split(h, h.shape[0] // 2, 0)[1] for h in normed]
    out = []
    for r in range(R):
        st = SimStrategy(R, r, h1, h2)
        # hidden_norm already applied above on the shared tensors; l2_normalize is
        # idempotent only up to rounding, so pass the pre-normalised rows through
        # with hidden_norm=False to keep one normalisation per row as in TF.
        out.append(add_contrastive_loss(normed[r], False, temperature, st))
    return out


def contrast_metrics(logits_ab, labels):
    """tf2/metrics.py:23-36: contrastive accuracy and entropy."""
    acc = (labels.argmax(1) == logits_ab.argmax(1)).to(logits_ab.dtype).mean()
    prob = torch.softmax(logits_ab, -1)
    entropy = -(prob * torch.log(prob + 1e-8)).sum(-1).mean()
    return acc, entropy
