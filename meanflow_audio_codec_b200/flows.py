"""Host-side mirrors of ``models/mlp_mixer.py`` (ConditionalMLPMixerFlow), ``models/conv_flow.py``
(ConditionalConvFlow) and ``models/factories.py`` (create_flow_model) on libmfac -- forward pass only.

    model = ConditionalMLPMixerFlow(noise_dimension, condition_dimension, num_blocks, latent_dimension)
    params = model.init(seed)["params"]                     # Flax-named tree of CUDA tensors
    out = model.apply({"params": params}, x, time, latents) # latents [B, num_latent_tokens, latent_dim] or None

Parameter trees carry the names Flax gives the reference modules (setup-defined submodules by attribute
name, ``@nn.compact`` ones auto-numbered in call order):
  mixer:   blocks_k/{input_proj, output_proj, mixer_block/Dense_0..5}, latent_proj
           Dense_0 = AdaLN 1, Dense_1/2 = token MLP, Dense_3 = AdaLN 2, Dense_4/5 = channel MLP
  convnet: blocks_k/{input_proj1, input_proj2, conditioning_layer, output_proj1, output_proj2,
           conv_block/{Conv_0 (3x3), Conv_1 (1x1 expand), GlobalResponseNormalization_0/{gamma,beta},
           Conv_2 (1x1 contract), layer_scale_gamma}}, latent_proj
ref: models/mlp_mixer.py:14-235, models/conv_flow.py:14-271, models/factories.py:106-148.
The reference never trains these two models (they have no ``encode``), so there is no loss path.

Encoder wiring ("smallest faithful choice"): ``init(key, with_encoder=True)`` adds the MLP flow's
``encoder`` subtree (mlp_flow.py:153-162: Dense -> GELU -> Dense, D -> (D + L) / 2 -> L) and
``apply(variables, x, method="encode")`` returns ``latents[B, 1, L]``, which the model's own ``latent_proj`` consumes when it was
built with ``num_latent_tokens=1``.  That is what lets ``MeanFlowCodec`` run MDCT -> encode -> sample -> IMDCT with any of the
three architectures.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref

import torch

from . import _lib


def _lecun(shape, fan_in, gen):
    w = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
    return w * (math.sqrt(1.0 / fan_in) / 0.87962566103423978)


def _dense(d_in, d_out, gen):
    return {"kernel": _lecun((d_in, d_out), d_in, gen), "bias": torch.zeros(d_out)}


def _to_device(tree, device):
    return {k: (_to_device(v, device) if isinstance(v, dict) else v.to(device)) for k, v in tree.items()}


class _WeightCache:
    """bf16 copies of the Dense kernels (what the tensor-core GEMMs read), refreshed when a leaf changes.

    Entries are keyed by the identity of the source tensor and hold a weak reference to it: a different tensor that
    happens to land on a freed address with the same shape (out-of-place parameter updates, a second checkpoint load)
    can never be mistaken for the cached one, and entries whose source died are dropped.  ``keep`` only pins the
    temporaries of ONE launch (``begin()`` clears it)."""

    def __init__(self):
        self._c = {}
        self.keep = []

    def begin(self):
        self.keep = []
        if len(self._c) > 4096:
            self._c = {k: v for k, v in self._c.items() if v[0]() is not None}

    def bf16(self, t: torch.Tensor) -> torch.Tensor:
        hit = self._c.get(id(t))
        if hit is None or hit[0]() is not t or hit[1] != t._version:
            copy = _lib.require_cuda(t, "kernel").to(torch.bfloat16).contiguous()
            hit = (weakref.ref(t), t._version, copy)
            self._c[id(t)] = hit
        return hit[2]

    def f32(self, t: torch.Tensor) -> torch.Tensor:
        t = _lib.require_cuda(t, "param")
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
            self.keep.append(t)
        return t

    def dense(self, p) -> _lib.Dense:
        return _lib.Dense(self.bf16(p["kernel"]).data_ptr(), self.f32(p["bias"]).data_ptr())


class _FlowBase:
    num_latent_tokens = 32

    def _check(self, x, time, latents):
        x = _lib.require_cuda(x, "x").to(torch.float32).contiguous()
        time = _lib.require_cuda(time, "time").to(torch.float32).contiguous()
        B = x.shape[0]
        if x.ndim != 2 or x.shape[1] != self.noise_dimension:
            raise ValueError(f"x must be [B, {self.noise_dimension}], got {tuple(x.shape)}")
        if tuple(time.shape) != (B, 2):
            raise ValueError(f"time must be [B, 2] = (t, h), got {tuple(time.shape)}")
        lat_flat = 0
        if latents is not None:
            latents = _lib.require_cuda(latents, "latents").to(torch.float32).reshape(B, -1).contiguous()
            lat_flat = latents.shape[1]
            if lat_flat != self.latent_flat:
                raise ValueError(f"latents must flatten to [B, {self.latent_flat}], got {tuple(latents.shape)}")
        return x, time, latents, B

    def __call__(self, variables, x, time=None, latents=None, method=None):
        return self.apply(variables, x, time, latents, method=method)

    # ---------------------------------------------------------------- encoder wiring
    def _encoder_model(self):
        from .mlp_flow import ConditionalFlow
        if getattr(self, "_enc_model", None) is None:
            # num_blocks = 0: the flat layout is the four encoder leaves and nothing else (csrc/imf_layout.cuh make_dims)
            self._enc_model, self._enc_fp = ConditionalFlow(self.noise_dimension, 2, 0, self.latent_dimension), None
        return self._enc_model

    def _init_encoder(self, gen, device) -> dict:
        return self._encoder_model().init(gen, device=device)["params"]["encoder"]

    def encode(self, params, x: torch.Tensor) -> torch.Tensor:
        """x[B, D] -> latents[B, 1, L] through the MLP encoder stored under ``params["encoder"]``."""
        if self.num_latent_tokens != 1:
            raise ValueError("encode() feeds latents[B, 1, L] to latent_proj: build the model with num_latent_tokens=1")
        if "encoder" not in params:
            raise KeyError("params has no 'encoder' subtree (use init(key, with_encoder=True))")
        enc = self._encoder_model()
        leaves = [params["encoder"]["encoder_mlp"][d][k] for d in ("dense1", "dense2") for k in ("bias", "kernel")]
        tag = tuple((id(t), t._version) for t in leaves)
        if self._enc_fp is None or self._enc_fp[0] != tag:
            self._enc_fp = (tag, enc.flat_params({"encoder": params["encoder"]}), leaves)   # leaves pinned: ids stay unique
        return enc.encode(self._enc_fp[1], x)[:, None, :]


class ConditionalMLPMixerFlow(_FlowBase):
    def __init__(self, noise_dimension: int, condition_dimension: int, num_blocks: int, latent_dimension: int,
                 token_mix_dim: int = 2048, channel_mix_dim: int = 2048, num_channels: int = 16,
                 num_latent_tokens: int = 32):
        self.noise_dimension, self.condition_dimension = int(noise_dimension), int(condition_dimension)
        self.num_blocks, self.latent_dimension = int(num_blocks), int(latent_dimension)
        self.token_mix_dim, self.channel_mix_dim, self.num_channels = int(token_mix_dim), int(channel_mix_dim), int(num_channels)
        self.num_latent_tokens = int(num_latent_tokens)
        self.spatial_size = int(math.sqrt(self.noise_dimension))          # mlp_mixer.py:121-122
        self.num_tokens = self.spatial_size * self.spatial_size
        self.latent_flat = self.num_latent_tokens * self.latent_dimension
        self._cache, self._ws = _WeightCache(), {}

    def dims(self) -> "_lib.MixerDims":
        return _lib.MixerDims(self.noise_dimension, self.condition_dimension, self.num_blocks, self.num_tokens,
                              self.num_channels, self.token_mix_dim, self.channel_mix_dim, self.latent_flat)

    def init(self, key=0, *args, device="cuda", with_encoder: bool = False, **kwargs) -> dict:
        gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
        D, Cd, T, ch = self.noise_dimension, self.condition_dimension, self.num_tokens, self.num_channels
        p = {}
        for k in range(self.num_blocks):
            p[f"blocks_{k}"] = {
                "input_proj": _dense(D, T * ch, gen),
                "mixer_block": {
                    "Dense_0": _dense(Cd, 2 * ch, gen), "Dense_1": _dense(T, self.token_mix_dim, gen),
                    "Dense_2": _dense(self.token_mix_dim, T, gen), "Dense_3": _dense(Cd, 2 * ch, gen),
                    "Dense_4": _dense(ch, self.channel_mix_dim, gen), "Dense_5": _dense(self.channel_mix_dim, ch, gen),
                },
                "output_proj": _dense(T * ch, D, gen),
            }
        p["latent_proj"] = _dense(self.latent_flat, Cd, gen)
        p = _to_device(p, device)
        if with_encoder:
            p["encoder"] = self._init_encoder(gen, device)
        return {"params": p}

    def apply(self, variables, x, time=None, latents=None, method=None):
        if method == "encode":
            return self.encode(variables["params"], x)
        if method is not None:
            raise ValueError(f"unknown method {method!r}")
        if time is None:
            raise TypeError("apply() missing required argument: 'time'")
        x, time, latents, B = self._check(x, time, latents)
        p = variables["params"]
        c = self._cache
        c.begin()
        blocks = (_lib.MixerBlockW * self.num_blocks)()
        for k in range(self.num_blocks):
            b, mb = p[f"blocks_{k}"], p[f"blocks_{k}"]["mixer_block"]
            blocks[k] = _lib.MixerBlockW(c.dense(b["input_proj"]), c.dense(mb["Dense_0"]), c.dense(mb["Dense_1"]),
                                         c.dense(mb["Dense_2"]), c.dense(mb["Dense_3"]), c.dense(mb["Dense_4"]),
                                         c.dense(mb["Dense_5"]), c.dense(b["output_proj"]))
        w = _lib.MixerWeights(blocks, c.dense(p["latent_proj"]))
        d = self.dims()
        l = _lib.lib()
        ws = self._ws.get((B, str(x.device)))
        if ws is None:
            n = l.mfac_mixer_workspace_bytes(C.byref(d), B)
            if n == 0:
                raise _lib.MfacError("mixer geometry not supported by libmfac (see mixer_dims_ok in csrc/flows.cu)")
            ws = torch.empty(n, dtype=torch.uint8, device=x.device)
            self._ws = {(B, str(x.device)): ws}
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(l.mfac_mixer_forward(C.byref(d), C.byref(w), x.data_ptr(), time.data_ptr(),
                                            None if latents is None else latents.data_ptr(), out.data_ptr(), B,
                                            ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "mixer_forward")
        return out


class ConditionalConvFlow(_FlowBase):
    def __init__(self, noise_dimension: int, condition_dimension: int, num_blocks: int, latent_dimension: int,
                 image_size: int = 28, use_grn: bool = True, num_latent_tokens: int = 32):
        self.noise_dimension, self.condition_dimension = int(noise_dimension), int(condition_dimension)
        self.num_blocks, self.latent_dimension = int(num_blocks), int(latent_dimension)
        self.image_size, self.use_grn, self.num_latent_tokens = int(image_size), bool(use_grn), int(num_latent_tokens)
        self.spatial_size = int(math.sqrt(self.noise_dimension))          # conv_flow.py:140
        self.channels = min(16, self.condition_dimension // 4)            # conv_flow.py:141
        self.bottleneck = 128                                             # conv_flow.py:144,158
        self.latent_flat = self.num_latent_tokens * self.latent_dimension
        if not self.use_grn:
            raise NotImplementedError("use_grn=False is not built (every reference config uses the default True)")
        self._cache, self._ws = _WeightCache(), {}

    def dims(self) -> "_lib.ConvDims":
        return _lib.ConvDims(self.noise_dimension, self.condition_dimension, self.num_blocks, self.spatial_size,
                             self.channels, self.bottleneck, self.latent_flat)

    def init(self, key=0, *args, device="cuda", with_encoder: bool = False, **kwargs) -> dict:
        gen = key if isinstance(key, torch.Generator) else torch.Generator().manual_seed(int(key))
        D, Cd, S, ch, bn = self.noise_dimension, self.condition_dimension, self.spatial_size, self.channels, self.bottleneck
        p = {}
        for k in range(self.num_blocks):
            p[f"blocks_{k}"] = {
                "input_proj1": _dense(D, bn, gen), "input_proj2": _dense(bn, S * S * ch, gen),
                "conditioning_layer": _dense(Cd, 2 * ch, gen),
                "conv_block": {
                    "Conv_0": {"kernel": _lecun((3, 3, ch, ch), 9 * ch, gen), "bias": torch.zeros(ch)},
                    "Conv_1": {"kernel": _lecun((1, 1, ch, 2 * ch), ch, gen), "bias": torch.zeros(2 * ch)},
                    "GlobalResponseNormalization_0": {"gamma": torch.zeros(2 * ch), "beta": torch.zeros(2 * ch)},
                    "Conv_2": {"kernel": _lecun((1, 1, 2 * ch, ch), 2 * ch, gen), "bias": torch.zeros(ch)},
                    "layer_scale_gamma": torch.full((ch,), 1e-6),
                },
                "output_proj1": _dense(S * S * ch, bn, gen), "output_proj2": _dense(bn, D, gen),
            }
        p["latent_proj"] = _dense(self.latent_flat, Cd, gen)
        p = _to_device(p, device)
        if with_encoder:
            p["encoder"] = self._init_encoder(gen, device)
        return {"params": p}

    def apply(self, variables, x, time=None, latents=None, method=None):
        if method == "encode":
            return self.encode(variables["params"], x)
        if method is not None:
            raise ValueError(f"unknown method {method!r}")
        if time is None:
            raise TypeError("apply() missing required argument: 'time'")
        x, time, latents, B = self._check(x, time, latents)
        p = variables["params"]
        c = self._cache
        c.begin()
        blocks = (_lib.ConvBlockW * self.num_blocks)()
        for k in range(self.num_blocks):
            b = p[f"blocks_{k}"]
            cb = b["conv_block"]
            grn = cb["GlobalResponseNormalization_0"]
            f = lambda t: c.f32(t).data_ptr()  # noqa: E731
            blocks[k] = _lib.ConvBlockW(c.dense(b["input_proj1"]), c.dense(b["input_proj2"]), c.dense(b["conditioning_layer"]),
                                        c.dense(b["output_proj1"]), c.dense(b["output_proj2"]),
                                        f(cb["Conv_0"]["kernel"]), f(cb["Conv_0"]["bias"]), f(cb["Conv_1"]["kernel"]),
                                        f(cb["Conv_1"]["bias"]), f(grn["gamma"]), f(grn["beta"]), f(cb["Conv_2"]["kernel"]),
                                        f(cb["Conv_2"]["bias"]), f(cb["layer_scale_gamma"]))
        w = _lib.ConvWeights(blocks, c.dense(p["latent_proj"]))
        d = self.dims()
        l = _lib.lib()
        ws = self._ws.get((B, str(x.device)))
        if ws is None:
            n = l.mfac_conv_workspace_bytes(C.byref(d), B)
            if n == 0:
                raise _lib.MfacError("convnet geometry not supported by libmfac (see conv_dims_ok in csrc/flows.cu)")
            ws = torch.empty(n, dtype=torch.uint8, device=x.device)
            self._ws = {(B, str(x.device)): ws}
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(l.mfac_conv_forward(C.byref(d), C.byref(w), x.data_ptr(), time.data_ptr(),
                                           None if latents is None else latents.data_ptr(), out.data_ptr(), B,
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "conv_forward")
        return out


def create_flow_model(config):
    """models/factories.py:106-148 -- dispatch on ``config.architecture`` (the reference validates the field but its
    trainer never reads it; this wires it)."""
    from .mlp_flow import ConditionalFlow
    arch = getattr(config, "architecture", None) or "mlp"
    kw = dict(noise_dimension=config.noise_dimension, condition_dimension=config.condition_dimension,
              num_blocks=config.num_blocks, latent_dimension=config.latent_dimension)
    if arch == "mlp":
        return ConditionalFlow(**kw)
    if arch == "convnet":
        return ConditionalConvFlow(image_size=int(config.noise_dimension ** 0.5), **kw)
    if arch == "mlp_mixer":
        return ConditionalMLPMixerFlow(**kw)
    raise ValueError(f"Unknown architecture: {arch}. Must be one of: 'mlp', 'convnet', 'mlp_mixer'")
