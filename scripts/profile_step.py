"""Runs eager pretrain steps with an NVTX range around the last one (for ncu --nvtx-include "profiled/")."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from simclr_b200 import engine, run, flags_def

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=512)
ap.add_argument('--steps', type=int, default=2)
ap.add_argument('--resnet_depth', type=int, default=50)
ap.add_argument('--image_size', type=int, default=224)
args = ap.parse_args()
flags_def.FLAGS(['profile'])
flags_def.set_flags(resnet_depth=args.resnet_depth, image_size=args.image_size, train_batch_size=args.batch,
                    b200_precision='bf16', b200_conv_engine='tc')
eng = engine.set_engine(engine.Engine(precision='bf16', conv_engine='tc'))
trainer = run.Trainer(num_classes=1000, num_examples=1281167, seed=0)
f, l = run.synthetic_batch(args.batch, args.image_size, 1000, eng.device, 1234)
for i in range(args.steps - 1):
    trainer.single_step(f, l)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push('profiled')
loss = trainer.single_step(f, l)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print('loss', float(loss))
