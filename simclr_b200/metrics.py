"""Training / evaluation metrics of the reference (`tf2/metrics.py`) and their TensorBoard flush.

`Mean`, `Accuracy`, `TopKCategoricalAccuracy` stand in for the `tf.keras.metrics` classes the
reference instantiates (`tf2/run.py:370-379,529-545`): same names, `update_state` / `result` /
`reset_states`.  Values are accumulated as device tensors (no host sync inside the step) and read
once per summary interval.  `SummaryWriter` writes real TensorBoard event files (TFRecord framing,
`Event{wall_time, step, summary{value{tag, simple_value}}}` hand-encoded: TensorFlow is not a
dependency), so runs show up in the same dashboards as the reference's.
"""
import os
import struct
import time

import torch
from absl import logging

from . import objective as obj_lib


class Mean:
    def __init__(self, name):
        self.name = name
        self.reset_states()

    def update_state(self, value):
        v = value.detach().reshape(()).float() if torch.is_tensor(value) else float(value)
        self._sum = v if self._sum is None else self._sum + v
        self._n += 1

    def result(self):
        if self._n == 0:
            return 0.0
        return float(self._sum) / self._n

    def reset_states(self):
        self._sum, self._n = None, 0


class Accuracy(Mean):
    """tf.keras.metrics.Accuracy: update_state(y_true_idx, y_pred_idx)."""

    def update_state(self, y_true, y_pred):
        super().update_state((y_true == y_pred).float().mean())


class TopKCategoricalAccuracy(Mean):
    """tf.keras.metrics.TopKCategoricalAccuracy(k): update_state(one_hot_labels, logits)."""

    def __init__(self, k, name):
        super().__init__(name)
        self.k = k

    def update_state(self, y_true, y_pred):
        topk = y_pred.topk(min(self.k, y_pred.shape[1]), dim=1).indices
        hit = (topk == y_true.argmax(1, keepdim=True)).any(1)
        super().update_state(hit.float().mean())


def update_pretrain_metrics_train(contrast_loss, contrast_acc, contrast_entropy, loss, logits_con, labels_con,
                                  replica_id=0):
    """tf2/metrics.py:23-36.  The accuracy / entropy pair comes from one kernel over logits_ab
    (`simclr_contrast_metrics`); `labels_con` is accepted for signature parity (the positive of row i is
    column replica_id*B + i by construction, tf2/objective.py:64-67)."""
    contrast_loss.update_state(loss)
    m = obj_lib.contrast_metrics(logits_con, replica_id)
    contrast_acc.update_state(m[0])
    contrast_entropy.update_state(m[1])


def update_pretrain_metrics_eval(contrast_loss_metric, contrastive_top_1_accuracy_metric,
                                 contrastive_top_5_accuracy_metric, contrast_loss, logits_con, labels_con):
    """tf2/metrics.py:39-46."""
    contrast_loss_metric.update_state(contrast_loss)
    contrastive_top_1_accuracy_metric.update_state(labels_con.argmax(1), logits_con.argmax(1))
    contrastive_top_5_accuracy_metric.update_state(labels_con, logits_con)


def update_finetune_metrics_train(supervised_loss_metric, supervised_acc_metric, loss, labels, logits):
    """tf2/metrics.py:49-55.  `labels` may hold half as many rows as `logits` (l = concat([l, l], 0))."""
    supervised_loss_metric.update_state(loss)
    reps = logits.shape[0] // labels.shape[0]
    lab = labels.argmax(1).repeat(reps)
    supervised_acc_metric.update_state((lab == logits.argmax(1)).float().mean())


def update_finetune_metrics_eval(label_top_1_accuracy_metrics, label_top_5_accuracy_metrics, outputs, labels):
    """tf2/metrics.py:58-62."""
    label_top_1_accuracy_metrics.update_state(labels.argmax(1), outputs.argmax(1))
    label_top_5_accuracy_metrics.update_state(labels, outputs)


def _float_metric_value(metric):
    return float(metric.result())


def log_and_write_metrics_to_summary(all_metrics, global_step, writer=None):
    """tf2/metrics.py:70-74."""
    for metric in all_metrics:
        metric_value = _float_metric_value(metric)
        logging.info('Step: [%d] %s = %f', global_step, metric.name, metric_value)
        if writer is not None:
            writer.scalar(metric.name, metric_value, step=global_step)


# ---------------------------------------------------------------------------------------------
# TensorBoard event files without TensorFlow
# ---------------------------------------------------------------------------------------------
def _crc32c_table():
    poly, table = 0x82F63B78, []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        table.append(c)
    return table


_CRC_TABLE = _crc32c_table()


def crc32c(data):
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked_crc(data):
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n):
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num, wire, payload):
    return _varint((num << 3) | wire) + payload


def encode_scalar_event(tag, value, step, wall_time):
    """Event{1: wall_time double, 2: step int64, 5: Summary{1: Value{1: tag, 2: simple_value float}}}."""
    val = _field(1, 2, _varint(len(tag.encode())) + tag.encode()) + _field(2, 5, struct.pack('<f', value))
    summ = _field(1, 2, _varint(len(val)) + val)
    return (_field(1, 1, struct.pack('<d', wall_time)) + _field(2, 0, _varint(int(step))) +
            _field(5, 2, _varint(len(summ)) + summ))


def encode_file_version_event(wall_time):
    v = b'brain.Event:2'
    return _field(1, 1, struct.pack('<d', wall_time)) + _field(3, 2, _varint(len(v)) + v)


class SummaryWriter:
    """`tf.summary.create_file_writer(model_dir)` + `tf.summary.scalar` + `flush` (tf2/run.py:499,646-657)."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, 'events.out.tfevents.%010d.simclr_b200' % int(time.time()))
        self._f = open(self.path, 'ab')
        self._record(encode_file_version_event(time.time()))

    def _record(self, data):
        header = struct.pack('<Q', len(data))
        self._f.write(header + struct.pack('<I', _masked_crc(header)) + data + struct.pack('<I', _masked_crc(data)))

    def scalar(self, tag, value, step):
        self._record(encode_scalar_event(tag, float(value), step, time.time()))

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()
