"""CPU oracle for the SimCLR pretrain step -- TEST INFRASTRUCTURE ONLY.

This package is a PyTorch-CPU restatement of the arithmetic of the reference's
TF2 tree (`tf2/objective.py`, `tf2/lars_optimizer.py`, `tf2/resnet.py`,
`tf2/model.py`, `tf2/data_util.py`, `tf2/run.py:557-622`).  It exists so the
CUDA path can be checked against something; it is never the product path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it.  Nothing under `simclr_b200/` does.

PARITY UNPINNED: TensorFlow is not installable in this environment and the
reference ships no tests and no golden vectors, so the
restatement cannot be checked against outputs of the reference itself.  What
pins it instead: the known-answer anchors that do exist in the reference
(parameter counts, endpoint shapes, LR-scaling identity -- SURVEY.md 4.2),
closed-form NT-Xent values, fp64-vs-fp32 self-consistency and the
sharded == global identity.  The TensorFlow semantics it encodes are listed in
SURVEY.md Appendix A.
"""
