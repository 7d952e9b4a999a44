"""NVLink peer-memory collectives of the step (`csrc/comm.cu`): host side.

One symmetric allocation per rank (torch.distributed._symmetric_memory: cuMem + handle
exchange; PyTorch is the allocator / rendezvous plumbing, the data path is our kernels),
carved into

  bn      NSLOT x world x SLOT_BYTES   one-shot SyncBN exchanges ([2C] doubles per rank)
  gather  per channel: world x capacity  all-gather regions ('z': embeddings, 'lse')
  flags   u64 words

`PeerComm.create(group)` returns None when peer memory cannot be set up (single rank, no
P2P, SIMCLR_COMM=nccl): the callers then use torch.distributed / NCCL for the same
collectives.  Either path computes the same values; the peer path is what the scaling
numbers are measured with.
"""
import os

import torch
import torch.distributed as dist

from ._lib import lib, stream_ptr

NSLOT = 4
MAX_CHANNELS = 8192                      # widest BatchNorm (ResNet 2x / 4x heads): 2*C doubles per slot
SLOT_BYTES = 2 * MAX_CHANNELS * 8
GATHER_CHANNELS = {'z': 4 << 20, 'lse': 256 << 10}     # bytes per rank


class PeerComm:
    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.device = device
        self.slot_bytes = SLOT_BYTES
        W = self.world
        off = 0
        self.bn_off = off; off += NSLOT * W * SLOT_BYTES
        self.gather_off = {}
        for name, cap in GATHER_CHANNELS.items():
            self.gather_off[name] = off; off += W * cap
        self.bn_flag_off = off; off += NSLOT * W * 8
        self.gather_flag_off = {}
        for name in GATHER_CHANNELS:
            self.gather_flag_off[name] = off; off += W * 8
        self.total = (off + 4095) // 4096 * 4096
        self.buf = symm.empty(self.total, dtype=torch.uint8, device=device)
        self.handle = symm.rendezvous(self.buf, self.group.group_name)
        self.peers_dev = int(self.handle.buffer_ptrs_dev)
        self.buf.zero_()
        # sequence numbers (start at 1: flag words start at 0) and CTA arrival counters, local memory
        # [sequence, nanoseconds spent waiting for peers] per channel
        self.bn_seq = torch.tensor([1, 0], dtype=torch.int64, device=device)
        self.gather_seq = {n: torch.tensor([1, 0], dtype=torch.int64, device=device) for n in GATHER_CHANNELS}
        self.gather_arrive = {n: torch.zeros(1, dtype=torch.int32, device=device) for n in GATHER_CHANNELS}
        torch.cuda.synchronize(device)
        dist.barrier(self.group)             # nobody pushes before every buffer is zeroed
        torch.cuda.synchronize(device)

    @staticmethod
    def create(group=None, device=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
            return None
        if os.environ.get('SIMCLR_COMM', 'peer') == 'nccl':
            return None
        device = device or torch.device('cuda', torch.cuda.current_device())
        try:
            return PeerComm(group, device)
        except Exception as exc:      # no P2P / symmetric memory on this system: NCCL carries the collectives
            if os.environ.get('SIMCLR_COMM') == 'peer_required':
                raise
            import warnings
            warnings.warn('simclr_b200: peer-memory collectives unavailable (%r); using NCCL' % (exc,))
            return None

    def wait_ns(self):
        """Nanoseconds this rank has spent inside the exchange kernels waiting for its peers, since creation
        (device counters: a synchronising read).  The slowest rank of a step waits ~0; the others wait for it."""
        return int(self.bn_seq[1].item()) + sum(int(t[1].item()) for t in self.gather_seq.values())

    # -- fused SyncBN exchanges ----------------------------------------------------
    def _bn_args(self):
        return (self.peers_dev, self.rank, self.world, self.bn_off, self.bn_flag_off, NSLOT, SLOT_BYTES, self.bn_seq)

    def bn_finalize(self, sums, count_local, gamma, beta, eps, momentum, mm, mv, mean, rstd, scale, shift, C):
        lib.comm_bn_finalize(sums, float(count_local), gamma, beta, eps, momentum, mm, mv, mean, rstd, scale, shift, C,
                             *self._bn_args(), stream_ptr())

    def bn_bwd_coef(self, sums, count_local, mean, rstd, gamma, coef, dgamma, dbeta, C):
        lib.comm_bn_bwd_coef(sums, float(count_local), mean, rstd, gamma, coef, dgamma, dbeta, C,
                             *self._bn_args(), stream_ptr())

    # -- all-gather ------------------------------------------------------------------
    def all_gather(self, t, channel):
        """[...] -> [world, ...] (rank-major), a view of the local gather region of `channel`."""
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        if nbytes % 16 or nbytes > GATHER_CHANNELS[channel]:
            raise ValueError('all_gather(%s): %d bytes per rank (need a 16-byte multiple <= %d)'
                             % (channel, nbytes, GATHER_CHANNELS[channel]))
        off = self.gather_off[channel]
        lib.comm_all_gather(t, nbytes, self.peers_dev, self.rank, self.world, off, self.gather_flag_off[channel],
                            nbytes, self.gather_seq[channel], self.gather_arrive[channel], stream_ptr())
        return self.buf[off:off + self.world * nbytes].view(t.dtype).view((self.world,) + tuple(t.shape))
