"""Run under torchrun with N ranks: one sharded pretrain step (NCCL all-gather of
embeddings + lse, SyncBN all-reduce, flat gradient all-reduce) checked on rank 0
against the oracle's N-replica simulation of the same global batch."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist
from util import rel_err, cfg_from_flags
from simclr_b200 import engine, run, flags_def
from oracle import model as OM, step as OS

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
conv_engine = sys.argv[2] if len(sys.argv) > 2 else 'simt'
global_bn = (sys.argv[3] != 'local') if len(sys.argv) > 3 else True
check_graph = len(sys.argv) > 4 and sys.argv[4] == 'graph'     # also: 2 graph replays == 2 eager steps
flags_def.FLAGS(['multi'])
rank = run.init_distributed()
R = dist.get_world_size()
B, S = 16, 64
flags_def.set_flags(resnet_depth=18, image_size=S, train_batch_size=B * R, use_blur=True, weight_decay=1e-4,
                    global_bn=global_bn, b200_precision=precision, b200_conv_engine=conv_engine)
eng = engine.set_engine(engine.Engine(precision=precision, conv_engine=conv_engine))
trainer = run.Trainer(num_classes=1000, num_examples=50000, seed=0)
g = torch.Generator().manual_seed(1)
feats = [torch.rand(B, S, S, 6, generator=g) for _ in range(R)]
labs = [torch.nn.functional.one_hot(torch.randint(0, 1000, (B,), generator=g), 1000).float() for _ in range(R)]
sigma = [0.9, 1.6]
sels = [(torch.rand(2, B, generator=g) < 0.5).to(torch.uint8) for _ in range(R)]
trainer.model.set_blur_draws(torch.tensor(sigma), sels[rank])
trainer.optimizer.learning_rate = 0.3
loss = trainer.single_step(feats[rank].cuda(), labs[rank].cuda())
torch.cuda.synchronize()
# the per-replica losses differ; the job loss is their mean
lt = loss.detach().clone().reshape(1)
dist.all_reduce(lt)
ok = True
if rank == 0:
    om = OM.Model(cfg_from_flags(flags_def.FLAGS), 1000)
    P, S_ = om.init(0)
    draws = [[(sigma[0], sels[r][0]), (sigma[1], sels[r][1])] for r in range(R)]
    info = OS.forward_backward(om, P, S_, feats, labs, blur_draws=draws)
    job_loss = float(lt.item()) / R
    res = {'world': R, 'precision': precision, 'engine': conv_engine, 'global_bn': global_bn,
           'collectives': 'peer' if trainer.strategy.comm is not None else 'nccl',
           'loss': job_loss, 'oracle_loss': float(info['loss'])}
    tol = 1e-3 if precision == 'fp32' and conv_engine == 'simt' else 0.5
    worst, wname = 0.0, ''
    for v in trainer.model.trainable_variables:
        ref = info['grads'][v.name]
        err = rel_err(v.grad, ref)
        if ref.norm() == 0:
            ok = ok and err < 1e-6
        else:
            if err > worst:
                worst, wname = err, v.name
    res['worst_grad_rel_err'] = worst; res['worst_name'] = wname
    ok = ok and worst < tol and abs(job_loss - res['oracle_loss']) < (1e-4 if tol == 1e-3 else 1e-2) * abs(res['oracle_loss'])
    res['ok'] = bool(ok)
    print('MULTI_GPU_CHECK ' + json.dumps(res), flush=True)
# all ranks must hold identical weights after the LARS step
w = trainer.model.vs.flat_value.clone()
w0 = w.clone()
dist.broadcast(w0, 0)
same = bool(torch.equal(w, w0))
t = torch.tensor([1.0 if same else 0.0], device='cuda')
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print('MULTI_GPU_WEIGHTS_IDENTICAL %d' % int(t.item()), flush=True)
    ok = ok and t.item() == 1.0
if check_graph:
    # segmented CUDA-graph replay at R > 1 (forward+backward graph | NCCL all-reduce | LARS graph) against eager steps
    vs, opt = trainer.model.vs, trainer.optimizer
    snap = (vs.flat_value.clone(), vs.flat_moving.clone(), opt._flat_v.clone(), opt.iterations)
    f_dev, l_dev = feats[rank].cuda(), labs[rank].cuda()
    for _ in range(2):
        trainer.single_step(f_dev, l_dev)
    torch.cuda.synchronize()
    w_eager = vs.flat_value.clone()
    vs.flat_value.copy_(snap[0]); vs.flat_moving.copy_(snap[1]); opt._flat_v.copy_(snap[2]); opt.iterations = snap[3]
    trainer.capture(f_dev, l_dev, warmup=1, restore=True)
    for _ in range(2):
        trainer.replay()
    torch.cuda.synchronize()
    err = rel_err(vs.flat_value, w_eager)
    w0 = vs.flat_value.clone(); dist.broadcast(w0, 0)
    t = torch.tensor([1.0 if torch.equal(vs.flat_value, w0) else 0.0], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        g_ok = err < 2e-3 and t.item() == 1.0
        print('MULTI_GPU_GRAPH ' + json.dumps({'replay_vs_eager_rel_err': err, 'ranks_identical': bool(t.item() == 1.0), 'ok': bool(g_ok)}), flush=True)
        ok = ok and g_ok
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
