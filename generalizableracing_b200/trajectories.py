"""rsl_rl.utils.split_and_pad_trajectories / unpad_trajectories (third party rsl-rl-lib 2.x, not vendored) on the
gr_traj_* kernels (csrc/traj.cu).  Used by RolloutStorage.reccurent_mini_batch_generator and by the recurrent policy's
Memory in batch mode (modules.ActorCriticRecurrent).  No CPU fallback."""
from __future__ import annotations

import torch

from . import _lib as B


class TrajectoryIndex:
    """(env, first step, length) of every trajectory of a [T, N] dones buffer, env-major; `offsets[n]` = trajectories of envs < n.
    One small device->host read (the number of trajectories: shapes depend on it, as in the reference's `.tolist()`)."""

    def __init__(self, dones: torch.Tensor, lib=None, boundaries=None):
        lib = lib or B.load()
        if dones.device.type != "cuda":
            raise RuntimeError("trajectory kernels run only on a CUDA device; there is no CPU fallback")
        d = dones.reshape(dones.shape[0], dones.shape[1])
        d = (d.view(torch.uint8) if d.dtype == torch.bool else d.to(torch.uint8)).contiguous()
        self.T, self.N = int(d.shape[0]), int(d.shape[1])
        dev = d.device
        self.offsets = torch.empty(self.N + 1, dtype=torch.int32, device=dev)
        self.env = torch.empty(self.T * self.N, dtype=torch.int32, device=dev)
        self.start = torch.empty_like(self.env)
        self.length = torch.empty_like(self.env)
        self._lib, self.device = lib, dev
        B.check(lib.gr_traj_index(d.data_ptr(), self.T, self.N, self.offsets.data_ptr(), self.env.data_ptr(), self.start.data_ptr(),
                                  self.length.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "gr_traj_index")
        pick = [self.N] if boundaries is None else list(boundaries) + [self.N]
        host = self.offsets[torch.tensor(pick, device=dev)].tolist()           # the one host read
        self.num_traj = host[-1]
        self.boundaries = host[:-1]

    def pad(self, tensor: torch.Tensor, first: int = 0, count: int = None, want_masks: bool = True):
        """tensor [T, N, D] -> padded [T, count, D] (+ bool masks [T, count]) for trajectories [first, first + count)."""
        count = self.num_traj - first if count is None else count
        src = tensor.reshape(self.T, self.N, -1)
        if src.dtype != torch.float32 or not src.is_contiguous():
            src = src.to(torch.float32).contiguous()
        D = int(src.shape[2])
        padded = torch.empty(self.T, count, D, device=self.device)
        masks = torch.empty(self.T, count, dtype=torch.uint8, device=self.device) if want_masks else None
        B.check(self._lib.gr_traj_pad(src.data_ptr(), self.T, self.N, D, self.env.data_ptr(), self.start.data_ptr(), self.length.data_ptr(), first, count,
                                      padded.data_ptr(), B.ptr(masks), torch.cuda.current_stream(self.device).cuda_stream), "gr_traj_pad")
        padded = padded.view(self.T, count, *tensor.shape[2:])
        return (padded, masks.view(torch.bool)) if want_masks else padded

    def hidden(self, saved: torch.Tensor, first: int, count: int) -> torch.Tensor:
        """saved [T, L, N, H] -> [L, count, H]: the hidden state each trajectory started from (rollout_storage.py:226-237)."""
        T, L, N, H = (int(x) for x in saved.shape)
        s = saved if saved.is_contiguous() else saved.contiguous()
        out = torch.empty(L, count, H, device=self.device)
        B.check(self._lib.gr_traj_hidden(s.data_ptr(), T, L, N, H, self.env.data_ptr(), self.start.data_ptr(), first, count, out.data_ptr(),
                                         torch.cuda.current_stream(self.device).cuda_stream), "gr_traj_hidden")
        return out


def split_and_pad_trajectories(tensor: torch.Tensor, dones: torch.Tensor):
    """Same contract as rsl_rl.utils.split_and_pad_trajectories: ([T, J, ...] zero padded, bool masks [T, J])."""
    return TrajectoryIndex(dones).pad(tensor)


class _Unpad(torch.autograd.Function):
    """unpad_trajectories with its adjoint (the pad of the cotangent): the recurrent policy back-propagates through it."""

    @staticmethod
    def forward(ctx, trajectories, masks):
        T, J = int(trajectories.shape[0]), int(trajectories.shape[1])
        D = int(trajectories.shape[-1])
        dev = trajectories.device
        m = masks.view(torch.uint8) if masks.dtype == torch.bool else masks.to(torch.uint8)
        m = m.contiguous()
        x = trajectories.reshape(T, J, D)
        x = x if (x.dtype == torch.float32 and x.is_contiguous()) else x.to(torch.float32).contiguous()
        n_valid = int(m.sum())                                  # (shape of the result; the reference's boolean indexing syncs too)
        Bn = n_valid // T
        out = torch.zeros(T, Bn, D, device=dev)
        scratch = torch.empty(2 * J + 1, dtype=torch.int32, device=dev)
        lib = B.load()
        B.check(lib.gr_traj_unpad(x.data_ptr(), m.data_ptr(), T, J, D, Bn, scratch.data_ptr(), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                "gr_traj_unpad")
        ctx.save_for_backward(scratch)
        ctx.dims = (T, J, D, Bn)
        return out

    @staticmethod
    def backward(ctx, g):
        (scratch,) = ctx.saved_tensors
        T, J, D, Bn = ctx.dims
        # adjoint of the scatter: gather g[t, env] back to (t', j) -- the pad kernel with (env, start, len) rebuilt from (len, cum)
        length, cum = scratch[:J], scratch[J:2 * J]
        env = torch.div(cum, T, rounding_mode="floor").to(torch.int32)
        start = (cum - env * T).to(torch.int32)
        gin = torch.empty(T, J, D, device=g.device)
        gc = g.contiguous()
        lib = B.load()
        B.check(lib.gr_traj_pad(gc.data_ptr(), T, Bn, D, env.contiguous().data_ptr(), start.contiguous().data_ptr(), length.contiguous().data_ptr(), 0, J,
                                gin.data_ptr(), None, torch.cuda.current_stream(g.device).cuda_stream), "gr_traj_pad")
        return gin, None


def unpad_trajectories(trajectories: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
    """Same contract as rsl_rl.utils.unpad_trajectories: [T, J, H] + masks [T, J] -> [T, B, H]; differentiable."""
    return _Unpad.apply(trajectories, masks)
