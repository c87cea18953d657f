"""Host-side mirror of trainers/{loss_strategies,noise_schedules,time_sampling,training_steps}.py for the
iMF path, dispatching to the fused CUDA step (``mfac_imf_loss_grad``).

ref: LossStrategy trainers/loss_strategies.py:27-47; FlowMatchingLoss :50-112; MeanFlowLoss :115-201;
ImprovedMeanFlowLoss :204-280; LinearNoiseSchedule / UniformNoiseSchedule trainers/noise_schedules.py:52-115;
UniformTimeSampling / LogitNormalTimeSampling / MeanFlowTimeSampling trainers/time_sampling.py:32-135;
create_loss_strategy trainers/train.py:52-153; train_step trainers/training_steps.py:15-61.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod

import torch

from . import _lib
from .mlp_flow import ConditionalFlow, TrainState


class LinearNoiseSchedule:
    def __init__(self, noise_min: float = 0.001, noise_max: float = 0.999):
        self.noise_min, self.noise_max = noise_min, noise_max

    def interpolate(self, x0, x1, t):
        if t.ndim == 1:
            t = t[:, None]
        return (1.0 - t) * x0 + (self.noise_min + self.noise_max * t) * x1

    def compute_target(self, x0, x1):
        return self.noise_max * x1 - x0


class UniformNoiseSchedule(LinearNoiseSchedule):
    """(1-t) x0 + t x1, target x1 - x0 (noise_schedules.py:91-115) == LinearNoiseSchedule(0, 1)."""

    def __init__(self):
        super().__init__(0.0, 1.0)


class UniformTimeSampling:
    """t ~ U[0, 1) (time_sampling.py:32-42)."""

    def sample_time(self, key, batch_size: int, dtype=torch.float32, device="cuda"):
        gen = torch.Generator(device="cpu").manual_seed(int(key))
        return torch.rand(batch_size, 1, generator=gen, dtype=torch.float32).to(device=device, dtype=dtype)


class LogitNormalTimeSampling:
    """t = sigmoid(mean + std N(0,1)) (time_sampling.py:45-70, utils.py:32-33)."""

    def __init__(self, mean: float = -0.4, std: float = 1.0):
        self.mean, self.std = mean, std

    def sample_time(self, key, batch_size: int, dtype=torch.float32, device="cuda"):
        gen = torch.Generator(device="cpu").manual_seed(int(key))
        n = torch.randn(batch_size, 1, generator=gen, dtype=torch.float32)
        return torch.sigmoid(n * self.std + self.mean).to(device=device, dtype=dtype)


class MeanFlowTimeSampling(LogitNormalTimeSampling):
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


class _FusedLossStrategy(LossStrategy):
    """Shared host side of the three strategies: all of them run ``mfac_imf_loss_grad`` with a different ``method``.

    ``key`` is an integer seed (the reference passes a jax PRNGKey).  Like the reference
     the same key gives the same (e, t, r); pass ``step=`` to advance the counter
    stream, or ``noise=``, ``t=``, ``r=`` to pin the draws explicitly (parity tests).
    """

    method = _lib.LOSS_IMPROVED_MEAN_FLOW
    gamma = 0.5
    c = 1e-3
    use_weighted_loss = True
    last_aux = None

    def _schedule(self):
        return self.noise_schedule.noise_min, self.noise_schedule.noise_max

    def _config(self, seed: int, step: int, row_offset: int, step_dev=None, rows_r_equals_t: int = 0) -> _lib.ImfConfig:
        ts = self.time_sampling
        nmin, nmax = self._schedule()
        uniform = isinstance(ts, UniformTimeSampling)
        return _lib.ImfConfig(nmin, nmax, getattr(ts, "mean", -0.4), getattr(ts, "std", 1.0),
                              getattr(ts, "data_proportion", 0.5), self.c,
                              1 if self.use_weighted_loss else 0, int(seed) & (2 ** 64 - 1), int(step), int(row_offset),
                              None if step_dev is None else step_dev.data_ptr(), self.method, self.gamma,
                              1 if uniform else 0, int(rows_r_equals_t))

    def compute_loss(self, state: TrainState, key, x, *, noise=None, t=None, r=None, step: int | None = None,
                     row_offset: int = 0, return_aux: bool = False, step_tensor=None, grad_ready=None,
                     rows_r_equals_t: int = 0, _train: dict | None = None):
        """``rows_r_equals_t``: with explicit ``t``/``r``, a promise that the first so many rows have r == t (the rule
        ``sample_tr`` applies, utils.py:41-44) -- their u pass then doubles as their v pass; -1 disables that sharing for the
        internal draws too.  ``step_tensor``: optional uint64 CUDA scalar read on the device as the RNG step (CUDA-graph replay).
        ``grad_ready(flat_slice)``: optional host callback, called as soon as the launches that finalise a slice of the
        flat gradient are enqueued (block by block, last block first, encoder last) -- data_parallel.py starts that
        bucket's all-reduce from it."""
        model: ConditionalFlow = state.model
        fp = model.flat_params(state.params)
        from .input_pipeline import LazyTokens
        audio = None
        if isinstance(x, LazyTokens):
            if x.ndim == 2 and x.shape[1] == model.noise_dimension:
                audio = x            # the step tokenises in its own prologue (mfac_imf_*_audio)
            else:
                x = x.materialize()
        if audio is None:
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
            aux_t = {k: torch.zeros((B, D), dtype=torch.float32, device=dev) for k in ("v", "u", "dudt", "e")}
            aux_t.update({k: torch.empty((B,), dtype=torch.float32, device=dev) for k in ("per_example", "t", "r")})
            aux = _lib.ImfAux(*[aux_t[k].data_ptr() for k in ("v", "u", "dudt", "per_example", "e", "t", "r")])
        cfg = self._config(int(key) if not isinstance(key, torch.Tensor) else int(key.sum()), state.step if step is None else step,
                           row_offset, step_tensor, rows_r_equals_t)
        ws = model.workspace(_lib.WS_LOSS_GRAD, B, dev)
        ptr = lambda a: None if a is None else a.data_ptr()  # noqa: E731
        with torch.cuda.device(dev):
            xin = ([audio.audio.data_ptr(), int(audio.audio.shape[1]), audio.window_size, audio.hop_size] if audio is not None
                   else [x.data_ptr()])
            sfx = "_audio" if audio is not None else ""
            if _train is None:
                _lib.check(getattr(_lib.lib(), "mfac_imf_loss_grad" + sfx)(
                    C.byref(model.dims), C.byref(cfg), fp.flat.data_ptr(), fp.shadow().data_ptr(), *xin,
                    ptr(noise), ptr(t), ptr(r), loss.data_ptr(), grads.data_ptr(),
                    C.byref(aux) if aux is not None else None, B, ws.data_ptr(), ws.numel(), _lib.stream_ptr()),
                    "imf_loss_grad" + sfx)
            else:
                # fused step: the same schedule with AdamW (and, for world > 1, the all-reduce) applied slice by slice
                tx = state.tx
                opt = _lib.AdamWConfig(tx.learning_rate, tx.b1, tx.b2, tx.eps, tx.weight_decay)
                ct, sc = _train.get("count_tensor"), _train.get("scratch")
                _lib.check(getattr(_lib.lib(), "mfac_imf_train_step" + sfx)(
                    C.byref(model.dims), C.byref(cfg), C.byref(opt), fp.flat.data_ptr(), fp.shadow().data_ptr(),
                    state.opt_state["mu"].data_ptr(), state.opt_state["nu"].data_ptr(), int(state.opt_state["count"]),
                    ptr(ct), ptr(sc), *xin, ptr(noise), ptr(t), ptr(r), loss.data_ptr(), grads.data_ptr(),
                    C.byref(aux) if aux is not None else None, B, int(_train.get("world", 1)), ws.data_ptr(), ws.numel(),
                    _lib.stream_ptr()), "imf_train_step" + sfx)
                fp.mark_shadow_current()
        self.last_aux = aux_t if return_aux else None
        from .mlp_flow import FlatParams
        gtree = FlatParams(model, grads).tree()
        if return_aux:
            return loss, gtree, aux_t
        return loss, gtree


    def train_step_fused(self, state: TrainState, key, x, *, world: int = 1, count_tensor=None, scratch=None, **kw):
        """``mfac_imf_train_step``: loss, gradients and the AdamW update in ONE call (the update of each block's slice is
        enqueued as soon as the backward has finalised it).  ``world > 1`` sum-all-reduces every slice over the
        ``mfac_comm`` communicator first and scales by 1 / world.  Returns (new_state, loss, grads)."""
        out = self.compute_loss(state, key, x, _train=dict(world=world, count_tensor=count_tensor, scratch=scratch), **kw)
        if count_tensor is None:
            state.opt_state["count"] += 1
            state = TrainState(state.step + 1, state.apply_fn, state.params, state.tx, state.opt_state, state.model)
        return (state,) + tuple(out)


class ImprovedMeanFlowLoss(_FusedLossStrategy):
    """v_pred = u + (t - r) * stop_gradient(du/dt) with the network's own v = f(z, [t, 0]) as the JVP tangent,
    weighted-L2 against noise_max*e - x (loss_strategies.py:204-280)."""

    method = _lib.LOSS_IMPROVED_MEAN_FLOW

    def __init__(self, noise_schedule: LinearNoiseSchedule | None = None,
                 time_sampling: MeanFlowTimeSampling | None = None, use_weighted_loss: bool = True):
        self.noise_schedule = noise_schedule or LinearNoiseSchedule()
        self.time_sampling = time_sampling or MeanFlowTimeSampling()
        self.use_weighted_loss = use_weighted_loss
        self.last_aux = None


class MeanFlowLoss(_FusedLossStrategy):
    """u against e - x - clip(t - r) * stop_gradient(du/dt), JVP along (e - x, 1, 0), adaptive weight
    1 / (mean_D err^2 + c)^(1 - gamma) (loss_strategies.py:115-201).  Like the reference it interpolates with
    z = (1 - t) x + t e and target e - x whatever ``noise_schedule`` says (:161-166)."""

    method = _lib.LOSS_MEAN_FLOW

    def __init__(self, noise_schedule: LinearNoiseSchedule | None = None,
                 time_sampling: MeanFlowTimeSampling | None = None, gamma: float = 0.5, c: float = 1e-3):
        self.noise_schedule = noise_schedule or LinearNoiseSchedule()
        self.time_sampling = time_sampling or MeanFlowTimeSampling()
        self.gamma, self.c = gamma, c
        self.last_aux = None

    def _schedule(self):
        return 0.0, 1.0


class FlowMatchingLoss(_FusedLossStrategy):
    """Single time t, h = 0, no JVP: pred = f(z_t, [t, 0], encode(x)) against the schedule's target with
    weighted_l2_loss or plain MSE (loss_strategies.py:50-112).  ``r`` is ignored."""

    method = _lib.LOSS_FLOW_MATCHING

    def __init__(self, noise_schedule: LinearNoiseSchedule | None = None, time_sampling=None,
                 use_weighted_loss: bool = True):
        self.noise_schedule = noise_schedule or LinearNoiseSchedule()
        self.time_sampling = time_sampling or LogitNormalTimeSampling()
        self.use_weighted_loss = use_weighted_loss
        self.last_aux = None

    def compute_loss(self, state, key, x, *, t=None, r=None, **kw):
        return super().compute_loss(state, key, x, t=t, r=t if r is None else r, **kw)


def create_loss_strategy(config) -> LossStrategy:
    """Loss strategy from a ``TrainFlowConfig``-like object (trainers/train.py:52-153): same field names, defaults,
    fall-backs and error messages."""
    get = lambda k: getattr(config, k, None)  # noqa: E731
    name = get("loss_strategy")
    if name is None:
        name = "improved_mean_flow" if get("use_improved_mean_flow") else "flow_matching"
    ns_name = get("noise_schedule") or "linear"
    if ns_name == "linear":
        noise_schedule = LinearNoiseSchedule(get("noise_min") if get("noise_min") is not None else 0.001,
                                             get("noise_max") if get("noise_max") is not None else 0.999)
    elif ns_name == "uniform":
        noise_schedule = UniformNoiseSchedule()
    else:
        raise ValueError(f"Unknown noise_schedule: {ns_name}. Must be one of: 'linear', 'uniform'")
    mean = get("time_sampling_mean") if get("time_sampling_mean") is not None else -0.4
    std = get("time_sampling_std") if get("time_sampling_std") is not None else 1.0
    ts_name = get("time_sampling") or "logit_normal"
    if ts_name == "uniform":
        time_sampling = UniformTimeSampling()
    elif ts_name == "logit_normal":
        time_sampling = LogitNormalTimeSampling(mean=mean, std=std)
    elif ts_name == "mean_flow":
        dp = get("time_sampling_data_proportion")
        time_sampling = MeanFlowTimeSampling(mean=mean, std=std, data_proportion=dp if dp is not None else 0.5)
    else:
        raise ValueError(f"Unknown time_sampling: {ts_name}. Must be one of: 'uniform', 'logit_normal', 'mean_flow'")
    weighted = get("use_weighted_loss") if get("use_weighted_loss") is not None else True
    if name == "flow_matching":
        return FlowMatchingLoss(noise_schedule=noise_schedule, time_sampling=time_sampling, use_weighted_loss=weighted)
    if name in ("mean_flow", "improved_mean_flow"):
        if not isinstance(time_sampling, MeanFlowTimeSampling):
            time_sampling = MeanFlowTimeSampling(mean=get("time_sampling_mean") or -0.4, std=get("time_sampling_std") or 1.0,
                                                 data_proportion=get("time_sampling_data_proportion") or 0.5)
        if name == "mean_flow":
            return MeanFlowLoss(noise_schedule=noise_schedule, time_sampling=time_sampling,
                                gamma=get("gamma") if get("gamma") is not None else 0.5,
                                c=get("c") if get("c") is not None else 1e-3)
        return ImprovedMeanFlowLoss(noise_schedule=noise_schedule, time_sampling=time_sampling, use_weighted_loss=weighted)
    raise ValueError(f"Unknown loss_strategy: {name}. Must be one of: 'flow_matching', 'mean_flow', 'improved_mean_flow'")


def train_step(state: TrainState, key, x, loss_strategy: LossStrategy | None = None, **kw):
    """(state, loss, key) -- trainers/training_steps.py:37-61.  The key is returned unchanged, as in the reference."""
    if loss_strategy is None:
        loss_strategy = FlowMatchingLoss()   # the reference's default (training_steps.py:58-59)
    if isinstance(loss_strategy, _FusedLossStrategy) and isinstance(state, TrainState):
        state, loss, _ = loss_strategy.train_step_fused(state, key, x, **kw)
        return state, loss, key
    loss, grads = loss_strategy.compute_loss(state, key, x, **kw)
    state = state.apply_gradients(grads=grads)
    return state, loss, key
