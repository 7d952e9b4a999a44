"""CPU tests of the host side: the C-ABI library loads and exports every symbol
the header declares, the flag surface equals the reference's, the variable
inventory equals the oracle's, and the multi-replica plumbing (gloo, 2 ranks)."""
import math
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from simclr_b200 import _lib
    decls = _lib.parse_header()
    assert len(decls) >= 30
    _lib.lib.load()                 # getattr on every declared symbol: raises if one is missing
    assert _lib.lib.version() >= 100
    import ctypes
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in decls:
        assert hasattr(dll, name), name


def test_argument_errors_are_reported_without_a_gpu():
    from simclr_b200._lib import lib, SimclrError
    with pytest.raises(SimclrError) as ei:
        lib.ntxent_normalize(None, 4, 8, 1, None, None, None)
    assert 'null pointer' in str(ei.value)
    with pytest.raises(SimclrError):
        lib.conv2d_fprop_tc(1, 1, 1, 1, 1, 1, 8, 8, 8, 8, 2, 2, 1, None, None)     # even kernel size


def test_flag_surface(flags):
    F = flags.FLAGS
    assert len(flags.REFERENCE_FLAG_NAMES) == 48
    expected = dict(learning_rate=0.3, learning_rate_scaling='linear', warmup_epochs=10, weight_decay=1e-6,
                    batch_norm_decay=0.9, train_batch_size=512, train_epochs=100, train_steps=0,
                    optimizer='lars', momentum=0.9, temperature=0.1, hidden_norm=True, proj_head_mode='nonlinear',
                    proj_out_dim=128, num_proj_layers=3, ft_proj_selector=0, global_bn=True, width_multiplier=1,
                    resnet_depth=50, sk_ratio=0., se_ratio=0., image_size=224, color_jitter_strength=1.0,
                    use_blur=True, use_tpu=True, lineareval_while_pretraining=True, fine_tune_after_block=-1,
                    mode='train', train_mode='pretrain', dataset='imagenet2012', eval_batch_size=256)
    for k in flags.REFERENCE_FLAG_NAMES:
        assert k in F, k
    fresh = {k: F[k].default for k in expected}
    assert fresh == expected
    with pytest.raises(Exception):
        F['learning_rate_scaling'].parse('cubic')        # enum is enforced like in the reference


def test_flag_names_match_reference_source():
    """If the reference tree is mounted (build container only), compare the flag names with it."""
    ref = '/root/reference/tf2/run.py'
    if not os.path.exists(ref):
        pytest.skip('reference not mounted')
    import re
    from simclr_b200 import flags_def
    names = re.findall(r"flags\.DEFINE_\w+\(\s*'(\w+)'", open(ref).read())
    assert names == flags_def.REFERENCE_FLAG_NAMES


@pytest.mark.parametrize('kw', [dict(resnet_depth=18, image_size=64), dict(resnet_depth=50), dict(resnet_depth=50, width_multiplier=2),
                                dict(resnet_depth=34, image_size=32, global_bn=False),
                                dict(resnet_depth=50, sk_ratio=0.0625, image_size=64), dict(resnet_depth=152, width_multiplier=2, sk_ratio=0.0625),
                                dict(resnet_depth=18, se_ratio=0.0625, image_size=64), dict(resnet_depth=50, se_ratio=0.25, sk_ratio=0.0625)])
def test_variable_inventory_matches_oracle(flags, kw):
    from simclr_b200 import resnet, model, engine
    from oracle import model as OM
    from util import cfg_from_flags
    flags.set_flags(**kw)
    vs = engine.VarStore()
    net = resnet.resnet(flags.FLAGS.resnet_depth, flags.FLAGS.width_multiplier, cifar_stem=flags.FLAGS.image_size <= 32, vs=vs)
    model.ProjectionHead(vs, net.cout)
    model.SupervisedHead(1000, vs, net.cout)
    om = OM.Model(cfg_from_flags(flags.FLAGS), 1000)
    assert [(v.name, v.shape, v.init) for v in vs.trainable] == [(k, s, i) for k, (s, i) in om.vs.trainable.items()]
    assert [(v.name, v.shape) for v in vs.moving] == [(k, s) for k, (s, i) in om.vs.moving.items()]


def test_lr_schedule_matches_oracle(flags):
    from simclr_b200 import model
    from oracle import model as OM
    from util import cfg_from_flags
    flags.set_flags(train_batch_size=4096, warmup_epochs=10, train_epochs=100)
    sched = model.WarmUpAndCosineDecay(0.3, 1281167)
    cfg = cfg_from_flags(flags.FLAGS)
    for step in [0, 1, 100, 3127, 3128, 20000, 31279, 40000]:
        assert sched(step) == OM.warmup_and_cosine_decay(cfg, 0.3, 1281167, step)
    assert model.get_train_steps(1281167) == 1281167 * 100 // 4096 + 1


def test_lars_name_filters(flags):
    from simclr_b200 import lars_optimizer
    opt = lars_optimizer.LARSOptimizer(0.1, weight_decay=1e-4,
                                       exclude_from_weight_decay=['batch_normalization', 'bias', 'head_supervised'])
    assert opt._use_weight_decay('resnet/conv2d/kernel:0') and opt._do_layer_adaptation('resnet/conv2d/kernel:0')
    for n in ['resnet/batch_norm_relu/sync_batch_normalization/gamma:0', 'head_supervised/linear_layer/dense_3/kernel:0',
              'head_supervised/linear_layer/dense_3/bias:0']:
        assert not opt._use_weight_decay(n) and not opt._do_layer_adaptation(n)
    with pytest.raises(NotImplementedError):
        lars_optimizer.LARSOptimizer(0.1, use_nesterov=True)


def test_model_errors(flags):
    from simclr_b200 import resnet, engine
    with pytest.raises(ValueError):
        resnet.resnet(51, 1)


# ---- multi-replica plumbing over gloo (2 ranks on CPU) ------------------------
def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from simclr_b200.engine import ReplicaContext
    from oracle import objective as O
    ctx = ReplicaContext()
    assert ctx.num_replicas_in_sync == world and ctx.replica_id == rank
    B, D = 6, 16
    g = torch.Generator().manual_seed(100)
    hs = [torch.randn(2 * B, D, generator=g, dtype=torch.float64) for _ in range(world)]
    z = O.l2_normalize(hs[rank])
    z_all = ctx.all_gather(z)                                   # [R][2B][D] rank-major
    assert z_all.shape == (world, 2 * B, D)
    for r in range(world):
        assert torch.equal(z_all[r], O.l2_normalize(hs[r]))
    # sharded loss from the gathered tensor == the oracle's replica simulation
    h1 = z_all[:, :B].reshape(world * B, D); h2 = z_all[:, B:].reshape(world * B, D)
    st = O.SimStrategy(world, rank, list(z_all[:, :B]), list(z_all[:, B:]))
    loss, _, labels = O.add_contrastive_loss(z, False, 0.1, st)
    ref = O.contrastive_loss_replicas(hs, True, 0.1)[rank]
    assert abs(loss.item() - ref[0].item()) < 1e-12
    assert torch.equal(labels, ref[2])
    t = torch.full((4,), float(rank + 1), dtype=torch.float64)
    ctx.all_reduce_sum(t)
    assert torch.equal(t, torch.full((4,), float(sum(range(1, world + 1))), dtype=torch.float64))
    dist.barrier()
    dist.destroy_process_group()
    out.put((rank, 'ok'))


def test_replica_context_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(2))
    assert got == [(0, 'ok'), (1, 'ok')]


def test_augmentation_draws_follow_the_reference_ranges():
    """`draw_train_augmentation` reproduces the distributions of tf2/data_util.py:53-75,246-320,382-390."""
    import random
    from simclr_b200 import data_util
    rng = random.Random(1)
    n_flip = n_jit = n_gray = 0
    for _ in range(400):
        d = data_util.draw_train_augmentation(240, 320, 1.0, rng)
        y, x, h, w = d['box']
        assert 0 <= y and 0 <= x and y + h <= 240 and x + w <= 320
        assert h * w >= 0.1 * 240 * 320 - 1e-9                      # min_object_covered (A9)
        assert 0.70 <= w / h <= 1.40                                  # aspect in [3/4, 4/3] up to rounding
        c = d['color']
        assert sorted(c['perm']) == [0, 1, 2, 3]
        assert 0.2 <= c['brightness'] <= 1.8 and 0.2 <= c['contrast'] <= 1.8 and 0.2 <= c['saturation'] <= 1.8
        assert -0.2 <= c['hue'] <= 0.2
        n_flip += d['flip']; n_jit += c['apply_jitter']; n_gray += c['apply_gray']
    assert 150 < n_flip < 250 and 280 < n_jit < 360 and 40 < n_gray < 120


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract (CPU arm): exactly one JSON line on stdout, with the reference-arm keys, even when a
    library writes to fd 1 behind Python's back."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                        '--cpu_batch', '2', '--image_size', '32', '--resnet_depth', '18'],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS='4'))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'images/sec pretrain step' and d['unit'] == 'images/s'
    assert d['higher_is_better'] is True and d['value'] > 0
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_profiles_are_consistent_with_their_sources(tmp_path):
    """profiles/: the committed launch list regenerates the committed share table and traffic.json, and the
    traffic bench.py reports is within 10 % of the algorithmic bytes (no wasted DRAM re-reads)."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out_txt, out_json = tmp_path / 'shares.txt', tmp_path / 'traffic.json'
    r = subprocess.run([sys.executable, os.path.join(root, 'scripts', 'summarize_launches.py'),
                        os.path.join(root, 'profiles', 'r02_launches_step.csv'), str(out_txt), str(out_json)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    new = json.load(open(out_json)); old = json.load(open(os.path.join(root, 'profiles', 'traffic.json')))
    assert new['launches'] == old['launches'] and abs(new['dram_bytes_per_launch'] - old['dram_bytes_per_launch']) < 1.0
    assert open(out_txt).read() == open(os.path.join(root, 'profiles', 'r02_launch_shares.txt')).read()
    line = json.load(open(os.path.join(root, 'profiles', 'r02_bench_n1_final.json')))
    roof = line['roofline']
    assert roof['traffic'] is not None and 0.9 < roof['traffic'] / roof['algorithmic_bytes'] < 1.1
    assert line['gpu_launches'] > 0 and line['e2e']['h2d_bytes_per_step'] > 0


def test_metrics_and_event_file(tmp_path):
    """tf2/metrics.py stand-ins and the hand-encoded TensorBoard event records."""
    import struct
    import torch
    from simclr_b200 import metrics as M
    assert M.crc32c(b'123456789') == 0xE3069283                     # CRC-32C check value
    m = M.Mean('train/x'); m.update_state(1.0); m.update_state(torch.tensor(3.0)); assert m.result() == 2.0
    m.reset_states(); assert m.result() == 0.0
    logits = torch.tensor([[0.1, 0.9, 0.0], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5]])
    labels = torch.nn.functional.one_hot(torch.tensor([1, 2, 2]), 3).float()
    a1 = M.Accuracy('a1'); a5 = M.TopKCategoricalAccuracy(2, 'a2')
    M.update_finetune_metrics_eval(a1, a5, logits, labels)
    assert abs(a1.result() - 2 / 3) < 1e-6 and abs(a5.result() - 2 / 3) < 1e-6
    sl, sa = M.Mean('l'), M.Mean('a')
    M.update_finetune_metrics_train(sl, sa, 0.5, labels[:1], torch.cat([logits[:1], logits[:1]]))   # labels doubled
    assert sa.result() == 1.0
    w = M.SummaryWriter(str(tmp_path))
    w.scalar('train/total_loss', 1.5, step=7); w.flush(); w.close()
    data = open(w.path, 'rb').read()
    recs, off = [], 0
    while off < len(data):
        (n,) = struct.unpack('<Q', data[off:off + 8])
        assert struct.unpack('<I', data[off + 8:off + 12])[0] == M._masked_crc(data[off:off + 8])
        body = data[off + 12:off + 12 + n]
        assert struct.unpack('<I', data[off + 12 + n:off + 16 + n])[0] == M._masked_crc(body)
        recs.append(body); off += 16 + n
    assert len(recs) == 2 and b'brain.Event:2' in recs[0]
    assert b'train/total_loss' in recs[1] and struct.pack('<f', 1.5) in recs[1] and recs[1][9:11] == bytes([0x10, 0x07])   # step = 7


def test_checkpoint_manager_roundtrip(tmp_path):
    """CheckpointManager: reference variable names, pruning to max_to_keep, latest, partial restore."""
    import torch
    from simclr_b200 import checkpoint as C

    class V:
        def __init__(self, name, t): self.name, self.value, self.shape = name, t, tuple(t.shape)

    class FakeModel:
        def __init__(self):
            self.variables = [V('resnet/conv2d/kernel:0', torch.arange(6.).view(1, 1, 2, 3)), V('head/bias:0', torch.ones(4))]
            self.trainable_variables = self.variables

    class FakeOpt:
        def __init__(self): self._flat_v = torch.full((8,), 2.0); self.iterations = 0
        def ensure_built(self, vs): pass

    m, o = FakeModel(), FakeOpt()
    mgr = C.CheckpointManager(m, o, str(tmp_path), max_to_keep=2)
    assert mgr.latest_checkpoint is None
    for step in (10, 20, 30):
        m.variables[1].value.fill_(float(step)); mgr.save(step)
    assert [os.path.basename(p) for p in mgr._paths()] == ['ckpt-20.npz', 'ckpt-30.npz']
    m2, o2 = FakeModel(), FakeOpt(); o2._flat_v.zero_()
    m2.variables.append(V('new/extra:0', torch.zeros(2)))               # not in the file: skipped (expect_partial)
    step = C.CheckpointManager(m2, o2, str(tmp_path)).restore(mgr.latest_checkpoint)
    assert step == 30 and o2.iterations == 30 and float(m2.variables[1].value[0]) == 30.0 and float(o2._flat_v[0]) == 2.0
    m3 = FakeModel(); m3.variables[1].value.zero_()
    assert C.CheckpointManager(m3, None, str(tmp_path)).restore(mgr.latest_checkpoint, weights_only=True) == 0
    assert float(m3.variables[1].value[0]) == 30.0


def test_input_pipeline_host_logic(tmp_path):
    """tf2/data.py:29-98 over an ArrayBuilder: per-replica batch size, per-pipeline shards, shuffle buffer + repeat +
    drop_remainder when training, one ordered pass with a ragged last batch otherwise, npz round trip."""
    import numpy as np
    from simclr_b200 import data as D, flags_def
    F = flags_def.FLAGS
    if not F.is_parsed():
        F(['test'])
    saved = {k: getattr(F, k) for k in ('image_size', 'train_mode', 'train_split', 'eval_split')}
    try:
        flags_def.set_flags(image_size=64, train_mode='pretrain', train_split='train', eval_split='validation')
        imgs = np.zeros((50, 8, 8, 3), dtype=np.uint8)
        path = str(tmp_path / 'toy.npz')
        np.savez(path, train_images=imgs, train_labels=np.arange(50) % 5, validation_images=imgs[:11],
                 validation_labels=np.arange(11) % 5, num_classes=5)
        b = D.ArrayBuilder.from_npz(path)
        assert b.info.splits['train'].num_examples == 50 and b.info.features['label'].num_classes == 5
        fake = lambda e, idx: list(idx)
        ctx0, ctx1 = D.InputContext(2, 0, 2), D.InputContext(2, 1, 2)
        it0 = D.build_input_fn(b, 8, None, True)(ctx0, seed=1, make_batch=fake)
        it1 = D.build_input_fn(b, 8, None, True)(ctx1, seed=1, make_batch=fake)
        a = [next(it0) for _ in range(20)]; c = [next(it1) for _ in range(20)]
        assert all(len(x) == 4 for x in a + c)                                  # 8 global / 2 replicas
        assert set(sum(a, [])) <= set(range(0, 25)) and set(sum(c, [])) <= set(range(25, 50))   # disjoint shards
        assert set(sum(a, [])) == set(range(0, 25))                             # repeat(-1): every example comes around
        assert sum(a, [])[:25] != list(range(25))                               # shuffled (buffer of 40 > shard)
        ev = list(D.build_input_fn(b, 8, None, False)(D.InputContext(1, 0, 1), make_batch=fake))
        assert [len(x) for x in ev] == [8, 3] and sum(ev, []) == list(range(11))   # ordered, ragged last batch kept
        with pytest.raises(ValueError):
            D.build_input_fn(b, 9, None, True)(ctx0, make_batch=fake)
    finally:
        flags_def.set_flags(**saved)
