"""Drop-in ``RolloutStorage`` (standalone/rsl_rl/ext/storage/rollout_storage.py:12-254 of the reference).

Same constructor, tensors (``[T,N,.]``), ``Transition``, ``add_transitions``, ``compute_returns``,
``mini_batch_generator``, ``get_statistics``, ``clear`` and ``step``; the three hot methods are one or two
launches of libgracing.so instead of 9 copies / ~8T+6 elementwise ops / 9 gathers.  CUDA only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as B


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.privileged_observations = None
            self.actions = None
            self.privileged_actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None
            self.time_outs = None        # extension: lets add_transitions fuse the time-out bootstrap (ppo.py:89-92)
            self.gamma = 0.0

        def clear(self):
            self.__init__()

    def __init__(self, training_type, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, actions_shape, device="cuda:0", _lib=None):
        self.device = torch.device(device)
        if _lib is None:
            if self.device.type != "cuda":
                raise RuntimeError("RolloutStorage runs only on a CUDA device; there is no CPU fallback")
            _lib = B.load()
        self._lib = _lib
        if training_type != "rl":
            raise ValueError("only training_type='rl' (PPO) is built; distillation storage is out of scope")
        self.training_type = training_type
        self.obs_shape = obs_shape
        self.privileged_obs_shape = privileged_obs_shape
        self.actions_shape = actions_shape
        T, N, dev = num_transitions_per_env, num_envs, self.device
        if len(obs_shape) != 1 or len(actions_shape) != 1:
            raise ValueError("observation / action rows must be 1-D (widths that are multiples of 4 floats move as 128-bit words, others as scalars)")
        self.observations = torch.zeros(T, N, *obs_shape, device=dev)
        if privileged_obs_shape[0] is not None:
            self.privileged_observations = torch.zeros(T, N, *privileged_obs_shape, device=dev)
        else:
            self.privileged_observations = None
        self.rewards = torch.zeros(T, N, 1, device=dev)
        self.actions = torch.zeros(T, N, *actions_shape, device=dev)
        self.dones = torch.zeros(T, N, 1, device=dev).byte()
        self.actions_log_prob = torch.zeros(T, N, 1, device=dev)
        self.values = torch.zeros(T, N, 1, device=dev)
        self.returns = torch.zeros(T, N, 1, device=dev)
        self.advantages = torch.zeros(T, N, 1, device=dev)
        self.mu = torch.zeros(T, N, *actions_shape, device=dev)
        self.sigma = torch.zeros(T, N, *actions_shape, device=dev)
        self.num_transitions_per_env = T
        self.num_envs = N
        self.saved_hidden_states_a = None
        self.saved_hidden_states_c = None
        self.step = 0
        self._scratch = torch.zeros(int(self._lib.gr_gae_scratch_bytes(N)) // 8 + 1, dtype=torch.float64, device=dev)
        self.moments = torch.zeros(3, dtype=torch.float64, device=dev)     # (count, mean, M2) of the raw advantages
        self._keep = None
        self._desc_cache = None
        self._tr = B.GrTransition()

    # ------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    def pack_records(self, indices: torch.Tensor = None) -> torch.Tensor:
        """Transition records [T*N, 48]: the columns one PPO mini-batch row needs, side by side (include/gracing.h, GR_RECORD_FLOATS): packed
        once per iteration after compute_returns so that the update kernels read one 192-byte record per sampled row.  ``indices`` (the
        iteration's ``torch.randperm`` over [T*N], rollout_storage.py:165): record r holds transition ``indices[r]``, i.e. mini-batch i of every
        epoch is the contiguous slice ``records[i*mb:(i+1)*mb]``."""
        if getattr(self, "_records", None) is None:
            self._records = torch.empty(self.num_transitions_per_env * self.num_envs, B.GR_RECORD_FLOATS, device=self.device)
        if indices is None:
            B.check(self._lib.gr_storage_pack_records(C.byref(self._desc()), self._records.data_ptr(), self._stream()), "gr_storage_pack_records")
        else:
            if indices.dtype != torch.int64 or indices.numel() > self._records.shape[0] or not indices.is_contiguous() or indices.device != self._records.device:
                raise ValueError("pack_records: indices must be a contiguous int64 tensor of <= T*N transition ids on the storage's device")
            B.check(self._lib.gr_storage_pack_records_permuted(C.byref(self._desc()), indices.data_ptr(), indices.numel(), self._records.data_ptr(), self._stream()),
                    "gr_storage_pack_records_permuted")
        return self._records

    def _desc(self) -> B.GrStorage:
        """Argument struct of the buffers (built once: the class never re-binds its tensors; do not re-bind them either)."""
        if self._desc_cache is not None:
            return self._desc_cache
        s = B.GrStorage()
        s.obs = self.observations.data_ptr()
        s.critic_obs = None if self.privileged_observations is None else self.privileged_observations.data_ptr()
        s.actions, s.rewards, s.dones = self.actions.data_ptr(), self.rewards.data_ptr(), self.dones.data_ptr()
        s.values, s.log_prob, s.mu, s.sigma = self.values.data_ptr(), self.actions_log_prob.data_ptr(), self.mu.data_ptr(), self.sigma.data_ptr()
        s.returns, s.advantages = self.returns.data_ptr(), self.advantages.data_ptr()
        s.T, s.N = self.num_transitions_per_env, self.num_envs
        s.obs_dim = self.obs_shape[0]
        s.critic_dim = self.privileged_obs_shape[0] or 0
        s.act_dim = self.actions_shape[0]
        self._desc_cache = s
        self._p_desc = C.byref(s)
        return s

    def _f32(self, t, shape=None):
        if t.requires_grad:
            t = t.detach()
        if t.dtype != torch.float32 or t.device != self.device or not t.is_contiguous():
            t = t.to(self.device, torch.float32).contiguous()
        return t

    def add_transitions(self, transition: "RolloutStorage.Transition"):
        """rollout_storage.py:71-88 in one launch; when ``transition.time_outs`` is set the reward bootstrap
        ``r += gamma * V * time_out`` of PPO.process_env_step (ppo.py:89-92) is fused in."""
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        self._save_hidden_states(transition.hidden_states)
        tr = self._tr
        tr.critic_obs = None
        tr.time_outs = None
        keep = [self._f32(transition.observations), self._f32(transition.actions), self._f32(transition.rewards),
                self._f32(transition.values), self._f32(transition.actions_log_prob), self._f32(transition.action_mean),
                self._f32(transition.action_sigma)]
        tr.obs, tr.actions, tr.rewards, tr.values, tr.log_prob, tr.mu, tr.sigma = (k.data_ptr() for k in keep)
        if self.privileged_observations is not None:
            po = self._f32(transition.privileged_observations)
            keep.append(po)
            tr.critic_obs = po.data_ptr()
        d = transition.dones
        if d.dtype == torch.bool:
            d = d.view(torch.uint8)
        if d.dtype not in (torch.uint8, torch.int64) or d.device != self.device or not d.is_contiguous():
            d = d.to(self.device, torch.int64).contiguous()
        keep.append(d)
        tr.dones = d.data_ptr()
        tr.dones_is_int64 = int(d.dtype == torch.int64)
        if transition.time_outs is not None:
            to = transition.time_outs
            to = to.view(torch.uint8) if to.dtype == torch.bool else to.to(torch.uint8)
            to = to.to(self.device).contiguous()
            keep.append(to)
            tr.time_outs = to.data_ptr()
            tr.gamma = float(transition.gamma)
        self._keep = keep
        self._desc()
        rc = self._lib.gr_storage_add(self._p_desc, C.byref(tr), self.step, self._stream())
        if rc:
            B.check(rc, "gr_storage_add")
        self.step += 1

    def _save_hidden_states(self, hidden_states):
        # rollout_storage.py:90-108 (GRU states are wrapped into 1-tuples to match the LSTM format)
        if hidden_states is None or hidden_states == (None, None):
            return
        hid_a = hidden_states[0] if isinstance(hidden_states[0], tuple) else (hidden_states[0],)
        hid_c = hidden_states[1] if isinstance(hidden_states[1], tuple) else (hidden_states[1],)
        if self.saved_hidden_states_a is None:
            T = self.observations.shape[0]
            self.saved_hidden_states_a = [torch.zeros(T, *hid_a[i].shape, device=self.device) for i in range(len(hid_a))]
            self.saved_hidden_states_c = [torch.zeros(T, *hid_c[i].shape, device=self.device) for i in range(len(hid_c))]
        for i in range(len(hid_a)):
            self.saved_hidden_states_a[i][self.step].copy_(hid_a[i])
            self.saved_hidden_states_c[i][self.step].copy_(hid_c[i])

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam, normalize: bool = True):
        """rollout_storage.py:113-127: GAE scan + advantage normalisation (unbiased std + 1e-8).  With
        ``normalize=False`` the raw advantages and their (count, mean, M2) in ``self.moments`` are left for a
        cross-rank merge followed by :meth:`normalize_advantages`."""
        lv = self._f32(last_values)
        self._desc()
        B.check(self._lib.gr_compute_returns(self._p_desc, lv.data_ptr(), float(gamma), float(lam), self._scratch.data_ptr(),
                                             self.moments.data_ptr(), int(normalize), self._stream()), "gr_compute_returns")

    def normalize_advantages(self, moments: torch.Tensor = None):
        m = self.moments if moments is None else moments.to(self.device, torch.float64).contiguous()
        desc = self._desc()
        B.check(self._lib.gr_advantage_normalize(C.byref(desc), m.data_ptr(), self._stream()), "gr_advantage_normalize")

    def get_statistics(self):
        # rollout_storage.py:129-137 (logging helper; plain torch, not on the hot path)
        # every env's last stored step closes a trajectory (the reference marks it in the buffer itself, so do we); lengths are the gaps
        # between consecutive end markers of the env-major flattening
        self.dones[-1] = 1
        ends = self.dones[..., 0].t().reshape(-1).nonzero(as_tuple=False)[:, 0]
        starts = torch.cat((ends.new_full((1,), -1), ends[:-1]))
        return (ends - starts).float().mean(), self.rewards.mean()

    def mini_batch_generator(self, num_mini_batches, num_epochs=8, indices: torch.Tensor = None):
        """rollout_storage.py:152-191; the nine fancy-index gathers of one mini-batch are one launch."""
        batch_size = self.num_envs * self.num_transitions_per_env
        mb = batch_size // num_mini_batches
        if indices is None:
            indices = torch.randperm(num_mini_batches * mb, requires_grad=False, device=self.device)
        indices = indices.to(self.device, torch.int64).contiguous()
        dev = self.device
        od, ad = self.obs_shape[0], self.actions_shape[0]
        cd = self.privileged_obs_shape[0] or 0
        desc = self._desc()
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                idx = indices[i * mb:(i + 1) * mb]
                out = dict(obs=torch.empty(mb, od, device=dev), actions=torch.empty(mb, ad, device=dev), values=torch.empty(mb, 1, device=dev),
                           advantages=torch.empty(mb, 1, device=dev), returns=torch.empty(mb, 1, device=dev), log_prob=torch.empty(mb, 1, device=dev),
                           mu=torch.empty(mb, ad, device=dev), sigma=torch.empty(mb, ad, device=dev))
                if cd:
                    out["critic_obs"] = torch.empty(mb, cd, device=dev)
                g = B.GrMiniBatch()
                for k, v in out.items():
                    setattr(g, k, v.data_ptr())
                B.check(self._lib.gr_storage_gather(C.byref(desc), idx.data_ptr(), mb, C.byref(g), self._stream()), "gr_storage_gather")
                priv = out["critic_obs"] if cd else out["obs"]
                yield (out["obs"], priv, out["actions"], out["values"], out["advantages"], out["returns"], out["log_prob"],
                       out["mu"], out["sigma"], (None, None), None)

    def reccurent_mini_batch_generator(self, num_mini_batches, num_epochs=8):
        """rollout_storage.py:194-254 (the reference's spelling).  One trajectory index per call (gr_traj_index: a count per
        env, a scan, one fill), then per mini-batch one pad launch per observation group and one gather launch per saved
        hidden-state tensor; the remaining fields are views of the storage, as in the reference."""
        from .trajectories import TrajectoryIndex
        if self.training_type != "rl":
            raise ValueError("This function is only available for reinforcement learning training.")
        mb = self.num_envs // num_mini_batches
        idx = TrajectoryIndex(self.dones, lib=self._lib, boundaries=[i * mb for i in range(num_mini_batches + 1)])
        bounds = idx.boundaries
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                start, stop = i * mb, (i + 1) * mb
                first, count = bounds[i], bounds[i + 1] - bounds[i]
                obs_batch, masks_batch = idx.pad(self.observations, first, count)
                if self.privileged_observations is not None:
                    privileged_obs_batch = idx.pad(self.privileged_observations, first, count, want_masks=False)
                else:
                    privileged_obs_batch = obs_batch
                hid_a_batch = [idx.hidden(h, first, count) for h in (self.saved_hidden_states_a or [])]
                hid_c_batch = [idx.hidden(h, first, count) for h in (self.saved_hidden_states_c or [])]
                hid_a_batch = hid_a_batch[0] if len(hid_a_batch) == 1 else hid_a_batch
                hid_c_batch = hid_c_batch[0] if len(hid_c_batch) == 1 else hid_c_batch
                yield (obs_batch, privileged_obs_batch, self.actions[:, start:stop], self.values[:, start:stop], self.advantages[:, start:stop],
                       self.returns[:, start:stop], self.actions_log_prob[:, start:stop], self.mu[:, start:stop], self.sigma[:, start:stop],
                       (hid_a_batch, hid_c_batch), masks_batch)

    recurrent_mini_batch_generator = reccurent_mini_batch_generator
