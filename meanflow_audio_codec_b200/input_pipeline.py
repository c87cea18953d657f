"""Input staging for the training loop -- the build's counterpart of ``jnp.asarray(next(it))`` in the reference's hot loop
(trainers/train.py:333-341) -- and lazily tokenised batches.

``HostBatchStager`` moves host batches to the GPU through two pinned host slots and two device slots on a dedicated copy
stream, so the upload of batch i+1 overlaps the kernels of batch i:

    stager = HostBatchStager(device="cuda:0")
    for x in stager.stream(host_batches):          # numpy arrays or CPU tensors, all of one shape / dtype
        tokens = tok.tokenize(x).reshape(x.shape[0], -1)
        state, loss, key = train_step(state, key, tokens, strategy)

``LazyTokens`` is what ``MDCTTokenization(..., lazy=True).tokenize`` returns: the MDCT of a batch that has not been run yet.
``reshape(B, -1)`` keeps it lazy; handing it to a loss strategy makes the step tokenise inside its own prologue
(``mfac_imf_train_step_audio``: for short clips no token tensor is ever written to HBM); anything else calls
``materialize()`` and gets the ordinary token tensor.
"""
from __future__ import annotations

import torch

from . import _lib
from .mdct import MDCTConfig, mdct, num_frames


class LazyTokens:
    def __init__(self, audio: torch.Tensor, config: MDCTConfig, shape=None):
        if audio.ndim != 2:
            raise ValueError(f"lazy tokenisation takes mono batches [B, T], got {tuple(audio.shape)}")
        self.audio = _lib.require_cuda(audio, "audio").to(torch.float32).contiguous()
        self.config = config
        self.window_size = int(config.window_size)
        self.hop_size = int(config.hop_size) if config.hop_size is not None else self.window_size // 2
        B, T = self.audio.shape
        self.frames = num_frames(T, self.window_size, self.hop_size)
        self._full = (B, self.frames, self.window_size)
        self.shape = tuple(shape) if shape is not None else self._full
        self._tokens = None

    ndim = property(lambda self: len(self.shape))
    device = property(lambda self: self.audio.device)
    dtype = torch.float32

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        n = self._full[0] * self._full[1] * self._full[2]
        known = 1
        for v in shape:
            if v != -1:
                known *= int(v)
        shape = tuple(int(v) if v != -1 else n // max(known, 1) for v in shape)
        total = 1
        for v in shape:
            total *= v
        if total != n or shape[0] != self._full[0]:
            return self.materialize().reshape(shape)       # anything but a per-clip regrouping: just compute the tokens
        out = LazyTokens.__new__(LazyTokens)
        out.__dict__.update(self.__dict__)
        out.shape = shape
        return out

    def materialize(self) -> torch.Tensor:
        if self._tokens is None:
            self._tokens = mdct(self.audio, config=self.config)
        return self._tokens.reshape(self.shape)

    def __getattr__(self, name):          # any tensor method the lazy view does not know: act on the real tokens
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class HostBatchStager:
    def __init__(self, device=None, slots: int = 2):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.slots = max(2, int(slots))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._pinned, self._dev = [None] * self.slots, [None] * self.slots
        self._ready = [torch.cuda.Event() for _ in range(self.slots)]
        self._consumed = [torch.cuda.Event() for _ in range(self.slots)]
        self.bytes_uploaded = 0

    def _slot(self, i, like: torch.Tensor):
        s = i % self.slots
        if self._pinned[s] is None or self._pinned[s].shape != like.shape or self._pinned[s].dtype != like.dtype:
            self._pinned[s] = torch.empty(like.shape, dtype=like.dtype).pin_memory()
            self._dev[s] = torch.empty(like.shape, dtype=like.dtype, device=self.device)
        return s

    def _upload(self, i, batch):
        t = batch if isinstance(batch, torch.Tensor) else torch.from_numpy(batch)
        if t.is_cuda:
            raise ValueError("HostBatchStager stages HOST batches; device tensors need no staging")
        s = self._slot(i, t)
        if not (t.is_pinned() and t.is_contiguous()):
            self._consumed[s].synchronize()           # the pinned slot is about to be overwritten by the host
            self._pinned[s].copy_(t)
            t = self._pinned[s]
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._consumed[s])     # the step that last read this device slot has finished
            self._dev[s].copy_(t, non_blocking=True)
            self._ready[s].record(self.copy_stream)
        self.bytes_uploaded += t.numel() * t.element_size()
        return s

    def stream(self, host_batches):
        """Yields device tensors, one per host batch; a yielded tensor is valid until the next one is requested."""
        it = iter(host_batches)
        cur = torch.cuda.current_stream(self.device)
        for e in self._consumed:
            e.record(cur)
        try:
            nxt = next(it)
        except StopIteration:
            return
        i = 0
        s = self._upload(i, nxt)
        while True:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready[s])
            try:
                nxt = next(it)
                s_next = self._upload(i + 1, nxt)
            except StopIteration:
                nxt, s_next = None, None
            yield self._dev[s]
            self._consumed[s].record(torch.cuda.current_stream(self.device))
            if nxt is None:
                return
            i, s = i + 1, s_next
