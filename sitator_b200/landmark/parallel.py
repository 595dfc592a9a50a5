"""Frame-sharded runs: the few small collectives of the landmark path, over ``torch.distributed``.

One process per GPU; rank r holds the r-th contiguous block of frames.  Everything exchanged is tiny
next to NVLink bandwidth: the L x L Gram, per-landmark / per-cluster vectors,
a handful of scalars.  With the NCCL backend the tensors stay on the device; the ``gloo`` backend
(CPU tests of the host logic) stages them through host memory.
"""
import numpy as np


def default_comm():
    try:
        import torch.distributed as dist
    except Exception:
        return None
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    return Comm()


class Comm(object):
    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    # -- tensor collectives (in place) -----------------------------------------------------------
    def _reduce_(self, t, op):
        if self.backend == "nccl" or not t.is_cuda:
            self.dist.all_reduce(t, op=op, group=self.group)
        else:
            h = t.cpu()
            self.dist.all_reduce(h, op=op, group=self.group)
            t.copy_(h)
        return t

    def allreduce_sum_(self, t):
        return self._reduce_(t, self.dist.ReduceOp.SUM)

    def allreduce_max_u64_(self, t):
        """Keys are uint64 stored in int64 tensors: flip the sign bit so signed max orders them."""
        t ^= (-0x8000000000000000)
        self._reduce_(t, self.dist.ReduceOp.MAX)
        t ^= (-0x8000000000000000)
        return t

    def allreduce_min_u64_(self, t):
        t ^= (-0x8000000000000000)
        self._reduce_(t, self.dist.ReduceOp.MIN)
        t ^= (-0x8000000000000000)
        return t

    # -- host scalars / arrays ---------------------------------------------------------------------
    def _host_tensor(self, arr):
        import torch
        t = torch.as_tensor(arr)
        if self.backend == "nccl":
            t = t.cuda()
        return t

    def allreduce_sum_scalar(self, v):
        t = self._host_tensor(np.array([int(v)], dtype=np.int64))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def allreduce_sum_scalars(self, values):
        """Several integer sums in one collective."""
        t = self._host_tensor(np.array([int(v) for v in values], dtype=np.int64))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return [int(x) for x in t.cpu().numpy()]

    def allreduce_min_int(self, v):
        t = self._host_tensor(np.array([int(v)], dtype=np.int64))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return int(t.item())

    def allreduce_sum_numpy(self, arr):
        t = self._host_tensor(np.ascontiguousarray(arr))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def allgather_numpy(self, arr):
        """Stack one equally shaped array per rank: (world, ...)."""
        import torch
        t = self._host_tensor(np.ascontiguousarray(arr))
        out = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t, group=self.group)
        return np.stack([o.cpu().numpy() for o in out])

    def allgather_varlen(self, t, sizes=None):
        """Concatenate one 1-D tensor per rank, of different lengths, in rank order (result on ``t``'s device).
        Travels as bytes, so any dtype works on either backend (gloo has no int16)."""
        import torch
        t = t.contiguous()
        if sizes is None:
            sizes = self.allgather_numpy(np.array([t.numel()], dtype=np.int64))[:, 0]
        item = t.element_size()
        n_max = int(max(int(x) for x in sizes)) * item
        send = t if (self.backend == "nccl" or not t.is_cuda) else t.cpu()
        pad = torch.zeros((max(n_max, 1),), dtype=torch.uint8, device=send.device)
        pad[:t.numel() * item] = send.view(torch.uint8)
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad, group=self.group)
        parts = [o[:int(n) * item] for o, n in zip(out, sizes)]
        return torch.cat(parts).view(t.dtype).to(t.device)

    def exclusive_scan_int(self, v):
        """Sum of ``v`` over the lower ranks (frame offset of this rank's shard)."""
        allv = self.allgather_numpy(np.array([int(v)], dtype=np.int64))[:, 0]
        return int(allv[:self.rank].sum())

    # -- jump scan carry (SiteTrajectory._jumped_generator across shard boundaries) -------------------
    def jump_carry(self, dev_traj, unknown_as_jump):
        """Last known site of every atom before this rank's shard (or the previous frame's row when
        ``unknown_as_jump``), as an int64 CUDA tensor, and whether this shard starts the trajectory."""
        import torch
        traj = dev_traj
        F, M = traj.shape
        if unknown_as_jump:
            mine = traj[F - 1].cpu().numpy()
        else:
            known = traj >= 0
            # last known entry per column
            idx = torch.where(known, torch.arange(F, device=traj.device)[:, None], torch.full_like(traj, -1)).max(dim=0).values
            mine = torch.where(idx >= 0, traj.gather(0, idx.clamp(min=0)[None, :])[0], torch.full_like(idx, -1)).cpu().numpy()
        allv = self.allgather_numpy(mine.astype(np.int64))      # (world, M)
        carry = np.full(M, -1, dtype=np.int64)
        for r in range(self.rank):
            if unknown_as_jump:
                carry = allv[r]
            else:
                carry = np.where(allv[r] >= 0, allv[r], carry)
        return torch.as_tensor(carry, device=traj.device), int(self.rank == 0)
