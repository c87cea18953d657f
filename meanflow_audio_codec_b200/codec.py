"""Codec inference pipeline: audio -> MDCT tokens -> encode -> few-NFE sample -> IMDCT overlap-add -> audio.

The reference has no end-to-end codec script; the nearest thing is sample -> detokenize inside the training loop
(trainers/train.py:364-387) around ``sample`` (evaluators/sampling.py:5-95).  This module packages the pipeline
BASELINE.json's configs[4] names ("codec inference sweep: MDCT -> 1/2-NFE iMF sample -> IMDCT on 10 s clips, batch
1-4096") behind one call, built only from the package's public pieces (``mdct``, ``ConditionalFlow.apply(method=
"encode")``, ``sample`` / ``sample_mean_flow``, ``imdct``):

    codec = MeanFlowCodec(model, params, window_size=512, hop_size=256)
    y = codec.reconstruct(audio_cuda, sampler="mf", nfe=1)             # device in, device out
    y = codec.reconstruct_host(audio_pinned, sampler="mf", nfe=1)      # host in, host out, streamed in sub-batches

Framing: a clip of T samples gives nf frames of N coefficients; the velocity network sees rows of
D = frames_per_row * N coefficients, so the clip is right-padded with zeros by whole hops until nf is a multiple of
frames_per_row (10 s at 44.1 kHz: 1721 -> 1722 frames -> 861 rows of 1024).  The output is cropped back to the
reference's detokenised length (nf - 1) * hop + 2N of the UNPADDED clip.

``reconstruct_host`` is the input/output staging of SURVEY.md section 8f-3 for inference: sub-batches travel through
two pinned staging slots per direction on dedicated copy streams, so the upload of sub-batch i+1 and the download of
sub-batch i-1 overlap the kernels of sub-batch i.
"""
from __future__ import annotations

import torch

from . import _lib
from .mdct import imdct, mdct, num_frames
from .flows import _FlowBase
from .mlp_flow import ConditionalFlow
from .sampling import sample, sample_mean_flow


class MeanFlowCodec:
    def __init__(self, model, params, window_size: int = 512, hop_size: int | None = None):
        """``model``: a ``ConditionalFlow``, or a ``ConditionalMLPMixerFlow`` / ``ConditionalConvFlow`` built with
        ``num_latent_tokens=1`` and initialised ``with_encoder=True`` (flows.py)."""
        if not isinstance(model, (ConditionalFlow, _FlowBase)):
            raise TypeError("model must be a ConditionalFlow, ConditionalMLPMixerFlow or ConditionalConvFlow")
        self.model, self.params = model, params
        self.N = int(window_size)
        self.hop = int(hop_size) if hop_size is not None else self.N // 2
        if model.noise_dimension % self.N:
            raise ValueError(f"noise_dimension ({model.noise_dimension}) must be a multiple of window_size ({self.N})")
        self.frames_per_row = model.noise_dimension // self.N
        self._fft_kw = dict(use_fft_threshold=self.N + 1)   # the direct cosine branch, stated explicitly

    # ------------------------------------------------------------------ geometry
    def geometry(self, T: int) -> dict:
        nf = num_frames(T, self.N, self.hop)
        t_pad, nf_pad = T, nf
        while nf_pad % self.frames_per_row:      # one more hop of (zero) samples adds exactly one frame once T >= N
            t_pad = self.N if t_pad < self.N else t_pad + self.hop
            nf_pad = num_frames(t_pad, self.N, self.hop)
            if t_pad == self.N and nf_pad % self.frames_per_row:
                t_pad += self.hop
                nf_pad = num_frames(t_pad, self.N, self.hop)
        return {"nf": nf, "nf_pad": nf_pad, "t_pad": t_pad, "rows_per_clip": nf_pad // self.frames_per_row,
                "out_len": (nf - 1) * self.hop + 2 * self.N}

    # ------------------------------------------------------------------ stages (device tensors)
    def tokens(self, audio: torch.Tensor, valid_length: int | None = None) -> torch.Tensor:
        """[B, T] -> model rows [B * rows_per_clip, D] (a view of the MDCT output, no copy).  ``valid_length``: the clips are
        ``valid_length`` samples long and ``audio`` is already zero-padded to ``geometry(valid_length)["t_pad"]`` columns (the
        host-streaming path stages straight into such a buffer, which saves the padding copy)."""
        _lib.require_cuda(audio, "audio")
        if audio.ndim != 2:
            raise ValueError(f"audio must be [B, T], got {tuple(audio.shape)}")
        g = self.geometry(int(audio.shape[1]) if valid_length is None else int(valid_length))
        if valid_length is not None and audio.shape[1] != g["t_pad"]:
            raise ValueError(f"pre-padded audio must have {g['t_pad']} columns, got {audio.shape[1]}")
        if g["t_pad"] != audio.shape[1]:
            audio = torch.nn.functional.pad(audio, (0, g["t_pad"] - audio.shape[1]))
        X = mdct(audio, self.N, self.hop, **self._fft_kw)            # [B, nf_pad, N]
        return X.view(-1, self.model.noise_dimension)

    def encode(self, audio: torch.Tensor, valid_length: int | None = None) -> torch.Tensor:
        """[B, T] -> latents [B * rows_per_clip, L]  ([.., 1, L] for the mixer / ConvNeXt flows)."""
        return self.model.apply({"params": self.params}, self.tokens(audio, valid_length), method="encode")

    def decode(self, latents: torch.Tensor, clips: int, T: int, sampler: str = "mf", nfe: int = 1, key: int = 0,
               guidance_scale: float = 1.0, noise=None) -> torch.Tensor:
        """latents [clips * rows_per_clip, L] (or [.., 1, L]) -> audio [clips, (nf - 1) * hop + 2N]."""
        g = self.geometry(T)
        D = self.model.noise_dimension
        if sampler == "mf":
            rows = sample_mean_flow(self.model.apply, D, self.params, key, latents, nfe=nfe, noise=noise)
        elif sampler == "heun":
            rows = sample(self.model.apply, D, self.params, key, latents=latents, n_steps=nfe,
                          guidance_scale=guidance_scale, noise=noise)
        else:
            raise ValueError(f"Unknown sampler: {sampler}. Must be one of: 'mf', 'heun'")
        y = imdct(rows.view(clips, g["nf_pad"], self.N), self.N, self.hop, **self._fft_kw)
        return y[:, :g["out_len"]]                                    # crop is a view

    def reconstruct(self, audio: torch.Tensor, sampler: str = "mf", nfe: int = 1, key: int = 0,
                    guidance_scale: float = 1.0, valid_length: int | None = None) -> torch.Tensor:
        lat = self.encode(audio, valid_length)
        T = int(audio.shape[1]) if valid_length is None else int(valid_length)
        return self.decode(lat, audio.shape[0], T, sampler, nfe, key, guidance_scale)

    # ------------------------------------------------------------------ host buffers, streamed
    def reconstruct_host(self, audio_host: torch.Tensor, out_host: torch.Tensor | None = None, sampler: str = "mf",
                         nfe: int = 1, key: int = 0, sub_batch: int | None = None, device=None) -> torch.Tensor:
        """Host in, host out.  ``audio_host`` [B, T] fp32 CPU tensor (pinned for full copy/compute overlap); returns a
        pinned CPU tensor [B, out_len].  Clips are processed ``sub_batch`` at a time through two device slots per
        direction; H2D, kernels and D2H of neighbouring sub-batches overlap on three streams.  ``sub_batch=None`` picks
        ``min(64, B, 2 sqrt(B))``: the first upload and the last download are not overlapped with anything, so few large chunks
        lose to more, smaller ones until the chunks get too small for the GEMMs (measured on B200 with 10 s clips: 16 clips 80 K
        audio-s/s at 8 per chunk against 70 K in one piece; 64 clips 108 K at 16 against 78 K at 64; 1024 clips best at 64)."""
        if audio_host.is_cuda:
            raise ValueError("audio_host must be a CPU tensor (use reconstruct() for device tensors)")
        if audio_host.ndim != 2 or audio_host.dtype != torch.float32:
            raise ValueError(f"audio_host must be fp32 [B, T], got {audio_host.dtype} {tuple(audio_host.shape)}")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        B, T = audio_host.shape
        g = self.geometry(T)
        if out_host is None:
            out_host = torch.empty((B, g["out_len"]), dtype=torch.float32).pin_memory()
        if sub_batch is None:
            sub_batch = self.auto_sub_batch(B)
        sb = max(1, min(int(sub_batch), B))
        compute = torch.cuda.current_stream(dev)
        up, down = self._streams(dev)
        # device slots are as wide as the padded clip and zero beyond T: uploads land in [:, :T], no padding copy afterwards
        x_dev = [torch.zeros((sb, g["t_pad"]), dtype=torch.float32, device=dev) for _ in range(2)]
        y_dev = [torch.empty((sb, g["out_len"]), dtype=torch.float32, device=dev) for _ in range(2)]
        uploaded = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]     # x slot free again
        produced = [torch.cuda.Event() for _ in range(2)]     # y slot holds a result
        drained = [torch.cuda.Event() for _ in range(2)]      # y slot copied out
        for e in consumed + drained:
            e.record(compute)
        chunks = [(i, min(B, i + sb)) for i in range(0, B, sb)]

        def upload(j):
            a, b = chunks[j]
            s = j & 1
            with torch.cuda.stream(up):
                up.wait_event(consumed[s])
                x_dev[s][:b - a, :T].copy_(audio_host[a:b], non_blocking=True)
                uploaded[s].record(up)

        upload(0)
        for j, (a, b) in enumerate(chunks):
            s = j & 1
            compute.wait_event(uploaded[s])
            if j + 1 < len(chunks):
                upload(j + 1)
            compute.wait_event(drained[s])
            y = self.reconstruct(x_dev[s][:b - a], sampler=sampler, nfe=nfe, key=key + j, valid_length=T)
            y_dev[s][:b - a].copy_(y)
            consumed[s].record(compute)
            produced[s].record(compute)
            with torch.cuda.stream(down):
                down.wait_event(produced[s])
                out_host[a:b].copy_(y_dev[s][:b - a], non_blocking=True)
                drained[s].record(down)
        compute.wait_stream(down)
        down.synchronize()
        return out_host

    @staticmethod
    def auto_sub_batch(B: int) -> int:
        """Clips per chunk of the host-streamed path: min(64, B, 2 sqrt(B)) (see ``reconstruct_host``)."""
        return max(1, min(64, int(B), int(round(2.0 * float(B) ** 0.5))))

    _stream_cache: dict = {}

    @classmethod
    def _streams(cls, dev):
        key = str(dev)
        if key not in cls._stream_cache:
            cls._stream_cache[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        return cls._stream_cache[key]
