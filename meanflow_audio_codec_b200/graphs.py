"""CUDA-graph capture of the training step -- this build's counterpart of wrapping the reference's ``train_step``
in ``jax.jit`` (the reference leaves it un-jitted).

    step = GraphedTrainStep(state, loss_strategy, tokenization, example_batch)
    loss = step(batch)            # batch: CUDA tensor with the example's shape; loss: 0-d CUDA tensor (static buffer)

One replay = tokenise + iMF loss/grad + AdamW (+ bf16 shadow refresh): ~180-200 kernel launches (for small batches a DAG over the
library's side streams) submitted as ONE graph launch, which is what matters at the config-faithful batch of 128 where the step is launch-bound.  The step counter
(RNG stream offset, AdamW bias correction) lives in device memory and advances inside the graph, so consecutive
replays draw fresh (e, t, r) exactly like consecutive eager steps with ``step=state.step``.
ref: trainers/training_steps.py:15-61, trainers/train.py:333-347.
"""
from __future__ import annotations

import torch

from .loss_strategies import LossStrategy
from .mlp_flow import TrainState


class GraphedTrainStep:
    def __init__(self, state: TrainState, loss_strategy: LossStrategy, tokenization, example: torch.Tensor, key: int = 0,
                 warmup: int = 2):
        if not example.is_cuda:
            raise ValueError("example batch must be a CUDA tensor")
        self.state, self.strategy, self.tok, self.key = state, loss_strategy, tokenization, key
        dev = example.device
        self.x = example.clone()
        self.count = torch.full((), int(state.opt_state["count"]), dtype=torch.int64, device=dev)   # read as uint64 by libmfac
        self.scratch = torch.zeros(2, dtype=torch.float32, device=dev)
        self.steps_run = 0
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # warm-up outside capture: table uploads, workspace allocation, attributes
            snap = [t.clone() for t in (self._flat(), state.opt_state["mu"], state.opt_state["nu"], self.count)]
            for _ in range(warmup):
                self._body()
            for t, s in zip((self._flat(), state.opt_state["mu"], state.opt_state["nu"], self.count), snap):
                t.copy_(s)                        # warm-up must not advance training
            self.state.model.flat_params(self.state.params).shadow()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # The captured launches hold RAW pointers into the model's loss/grad workspace.  ConditionalFlow.workspace keeps one
        # buffer per kind and drops it when another batch size asks (validation, a last partial batch): pin the one the
        # graph uses so that eviction there can never free memory the replays still write to.
        from . import _lib
        rows = self.x.shape[0]
        self._ws_ref = state.model.workspace(_lib.WS_LOSS_GRAD, rows, dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        if state.model.workspace(_lib.WS_LOSS_GRAD, rows, dev).data_ptr() != self._ws_ref.data_ptr():
            raise RuntimeError("loss/grad workspace changed during capture")

    def _flat(self):
        return self.state.model.flat_params(self.state.params).flat

    def _body(self):
        x = self.x
        if self.tok is not None:
            x = self.tok.tokenize(x)
            x = x.reshape(x.shape[0], -1)
        if hasattr(self.strategy, "train_step_fused"):
            _, loss, _ = self.strategy.train_step_fused(self.state, self.key, x, step_tensor=self.count, count_tensor=self.count,
                                                        scratch=self.scratch)
            return loss
        loss, grads = self.strategy.compute_loss(self.state, self.key, x, step_tensor=self.count)
        self.state.apply_gradients(grads=grads, count_tensor=self.count, scratch=self.scratch)
        return loss

    def __call__(self, batch: torch.Tensor) -> torch.Tensor:
        if batch.data_ptr() != self.x.data_ptr():
            self.x.copy_(batch, non_blocking=True)
        self.graph.replay()
        self.steps_run += 1
        self.state.step += 1
        self.state.opt_state["count"] += 1
        return self.loss


class GraphedCodec:
    """One CUDA-graph launch per ``MeanFlowCodec.reconstruct`` call of a fixed shape -- for serving a few clips at a time, where the
    eager call (≈45 launches) is bound by the host's launch rate, not by the GPU (1 clip of 10 s: 0.32 ms eager).

        run = GraphedCodec(codec, clips=1, T=441000, sampler="mf", nfe=1)
        y = run(audio)          # audio [clips, T] CUDA tensor -> [clips, out_len] (static buffer, overwritten by the next call)

    ``fresh_noise=True`` draws new initial noise before every replay (one extra launch); otherwise the noise is the one ``key``
    selects, as for the eager call with that key.  MLP ``ConditionalFlow`` models only (the mixer / ConvNeXt wrappers keep a single
    workspace per batch size, which a captured graph cannot pin).
    """

    def __init__(self, codec, clips: int, T: int, sampler: str = "mf", nfe: int = 1, key: int = 0, fresh_noise: bool = False,
                 device=None):
        from . import _lib
        from .mlp_flow import ConditionalFlow
        if not isinstance(codec.model, ConditionalFlow):
            raise TypeError("GraphedCodec captures the MLP ConditionalFlow pipeline only")
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.codec, self.T, self.clips = codec, int(T), int(clips)
        self.sampler, self.nfe, self.key = sampler, int(nfe), int(key)
        g = codec.geometry(self.T)
        rows = self.clips * g["rows_per_clip"]
        self.x = torch.zeros((self.clips, g["t_pad"]), dtype=torch.float32, device=dev)     # zero beyond T: the padding frames
        self.noise = torch.empty((rows, codec.model.noise_dimension), dtype=torch.float32, device=dev) if fresh_noise else None
        if self.noise is not None:
            self.noise.normal_()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):            # warm-up outside capture: tables, workspaces, function attributes, bf16 shadow
            for _ in range(2):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the captured launches hold raw pointers into these two workspaces: keep them alive and notice a swap
        self._ws_ref = (codec.model.workspace(_lib.WS_FORWARD, rows, dev), codec.model.workspace(_lib.WS_SAMPLE, rows, dev))
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.y = self._body()
        now = (codec.model.workspace(_lib.WS_FORWARD, rows, dev), codec.model.workspace(_lib.WS_SAMPLE, rows, dev))
        if any(a.data_ptr() != b.data_ptr() for a, b in zip(now, self._ws_ref)):
            raise RuntimeError("codec workspaces changed during capture")

    def _body(self):
        lat = self.codec.encode(self.x, valid_length=self.T)
        return self.codec.decode(lat, self.clips, self.T, sampler=self.sampler, nfe=self.nfe, key=self.key, noise=self.noise)

    def __call__(self, audio: torch.Tensor) -> torch.Tensor:
        if tuple(audio.shape) != (self.clips, self.T):
            raise ValueError(f"audio must be [{self.clips}, {self.T}], got {tuple(audio.shape)}")
        self.x[:, :self.T].copy_(audio, non_blocking=True)
        if self.noise is not None:
            self.noise.normal_()
        self.graph.replay()
        return self.y
