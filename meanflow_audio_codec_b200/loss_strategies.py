"""Host-side mirror of trainers/{loss_strategies,noise_schedules,time_sampling,training_steps}.py for the
iMF path, dispatching to the fused CUDA step (``mfac_imf_loss_grad``).

ref: LossStrategy trainers/loss_strategies.py:27-47; ImprovedMeanFlowLoss :204-280;
LinearNoiseSchedule trainers/noise_schedules.py:52-88; MeanFlowTimeSampling trainers/time_sampling.py:79-135;
train_step trainers/training_steps.py:15-61.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod

import torch

from . import _lib
from .mlp_flow import ConditionalFlow, ParamTree, TrainState


class LinearNoiseSchedule:
    def __init__(self, noise_min: float = 0.001, noise_max: float = 0.999):
        self.noise_min, self.noise_max = noise_min, noise_max

    def interpolate(self, x0, x1, t):
        if t.ndim == 1:
            t = t[:, None]
        return (1.0 - t) * x0 + (self.noise_min + self.noise_max * t) * x1

    def compute_target(self, x0, x1):
        return self.noise_max * x1 - x0


class MeanFlowTimeSampling:
    def __init__(self, mean: float = -0.4, std: float = 1.0, data_proportion: float = 0.5):
        self.mean, self.std, self.data_proportion = mean, std, data_proportion

    def sample_time_pair(self, key, batch_size: int, dtype=torch.float32, device="cuda"):
        """(t, r) each [B,1], r <= t, first int(B*p) rows r = t (utils.py:36-45); ``key`` seeds torch's RNG."""
        gen = torch.Generator(device="cpu").manual_seed(int(key))
        n = torch.randn(2, batch_size, generator=gen, dtype=torch.float32)
        t = torch.sigmoid(n[0] * self.std + self.mean)
        r = torch.sigmoid(n[1] * self.std + self.mean)
        t, r = torch.maximum(t, r), torch.minimum(t, r)
        mask = torch.arange(batch_size) < int(batch_size * self.data_proportion)
        r = torch.where(mask, t, r)
        return t[:, None].to(device=device, dtype=dtype), r[:, None].to(device=device, dtype=dtype)


class LossStrategy(ABC):
    @abstractmethod
    def compute_loss(self, state: TrainState, key, x):
        """-> (loss, grads)"""


class ImprovedMeanFlowLoss(LossStrategy):
    """v_pred = u + (t - r) * stop_gradient(du/dt), weighted-L2 against noise_max*e - x.

    ``key`` is an integer seed (the reference passes a jax PRNGKey).  Like the reference
     the same key gives the same (e, t, r); pass ``step=`` to advance the counter
    stream, or ``noise=``, ``t=``, ``r=`` to pin the draws explicitly (parity tests).
    """

    def __init__(self, noise_schedule: LinearNoiseSchedule | None = None,
                 time_sampling: MeanFlowTimeSampling | None = None, use_weighted_loss: bool = True):
        self.noise_schedule = noise_schedule or LinearNoiseSchedule()
        self.time_sampling = time_sampling or MeanFlowTimeSampling()
        self.use_weighted_loss = use_weighted_loss
        self.last_aux = None

    def _config(self, seed: int, step: int, row_offset: int, step_dev=None) -> _lib.ImfConfig:
        ns, ts = self.noise_schedule, self.time_sampling
        return _lib.ImfConfig(ns.noise_min, ns.noise_max, ts.mean, ts.std, ts.data_proportion, 1e-3,
                              1 if self.use_weighted_loss else 0, int(seed) & (2 ** 64 - 1), int(step), int(row_offset),
                              None if step_dev is None else step_dev.data_ptr())

    def compute_loss(self, state: TrainState, key, x, *, noise=None, t=None, r=None, step: int | None = None,
                     row_offset: int = 0, return_aux: bool = False, step_tensor=None, grad_ready=None):
        """``step_tensor``: optional uint64 CUDA scalar read on the device as the RNG step (CUDA-graph replay).
        ``grad_ready(flat_slice)``: optional host callback, called as soon as the launches that finalise a slice of the
        flat gradient are enqueued (block by block, last block first, encoder last) -- data_parallel.py starts that
        bucket's all-reduce from it."""
        model: ConditionalFlow = state.model
        fp = model.flat_params(state.params)
        x = _lib.require_cuda(x, "x").to(torch.float32).contiguous()
        if x.ndim != 2 or x.shape[1] != model.noise_dimension:
            raise ValueError(f"x must be [B, {model.noise_dimension}], got {tuple(x.shape)}")
        B = x.shape[0]
        dev = x.device

        def prep(a, shape, name):
            if a is None:
                return None
            a = _lib.require_cuda(a, name).to(torch.float32).reshape(shape).contiguous()
            return a

        noise, t, r = prep(noise, (B, model.noise_dimension), "noise"), prep(t, (B,), "t"), prep(r, (B,), "r")
        if (t is None) != (r is None):
            raise ValueError("t and r must be given together")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        grads = torch.empty_like(fp.flat)
        aux_t, aux = {}, None
        cb_keep = None
        if grad_ready is not None and not return_aux:
            cb_keep = _lib.GRAD_READY_FN(lambda user, off, cnt: grad_ready(grads[off:off + cnt]))
            aux = _lib.ImfAux(*([None] * 7), cb_keep, None)
        if return_aux:
            D = model.noise_dimension
            aux_t = {k: torch.empty((B, D), dtype=torch.float32, device=dev) for k in ("v", "u", "dudt", "e")}
            aux_t.update({k: torch.empty((B,), dtype=torch.float32, device=dev) for k in ("per_example", "t", "r")})
            aux = _lib.ImfAux(*[aux_t[k].data_ptr() for k in ("v", "u", "dudt", "per_example", "e", "t", "r")])
        cfg = self._config(int(key) if not isinstance(key, torch.Tensor) else int(key.sum()), state.step if step is None else step,
                           row_offset, step_tensor)
        ws = model.workspace(_lib.WS_LOSS_GRAD, B, dev)
        ptr = lambda a: None if a is None else a.data_ptr()  # noqa: E731
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().mfac_imf_loss_grad(
                C.byref(model.dims), C.byref(cfg), fp.flat.data_ptr(), fp.shadow().data_ptr(), x.data_ptr(),
                ptr(noise), ptr(t), ptr(r), loss.data_ptr(), grads.data_ptr(),
                C.byref(aux) if aux is not None else None, B, ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                "imf_loss_grad")
        self.last_aux = aux_t if return_aux else None
        from .mlp_flow import FlatParams
        gtree = FlatParams(model, grads).tree()
        if return_aux:
            return loss, gtree, aux_t
        return loss, gtree


def train_step(state: TrainState, key, x, loss_strategy: LossStrategy | None = None, **kw):
    """(state, loss, key) -- trainers/training_steps.py:37-61.  The key is returned unchanged, as in the reference."""
    if loss_strategy is None:
        raise ValueError("only ImprovedMeanFlowLoss is implemented on this path; pass loss_strategy=")
    loss, grads = loss_strategy.compute_loss(state, key, x, **kw)
    state = state.apply_gradients(grads=grads)
    return state, loss, key
