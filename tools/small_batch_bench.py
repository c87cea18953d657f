#!/usr/bin/env python
"""Step time at the config-faithful batch sizes, eager and as one CUDA graph, with the concurrent schedule on and off.
usage (GPU box): python tools/small_batch_bench.py [batches...]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import meanflow_audio_codec_b200 as m
from meanflow_audio_codec_b200 import _lib

batches = [int(a) for a in sys.argv[1:]] or [128, 256, 1024, 4096]
T = 784


def make(B):
    model = m.ConditionalFlow(noise_dimension=1024, condition_dimension=128, num_blocks=8, latent_dimension=256)
    state = m.TrainState.create(apply_fn=model.apply, params=model.init(42)["params"], tx=m.adamw(1e-4, 1e-4))
    strat = m.ImprovedMeanFlowLoss()
    tok = m.MDCTTokenization(512, 256)
    x = 0.1 * torch.randn(B, T, device="cuda")
    return state, strat, tok, x


def timeit(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for conc in (0, 4096):
    _lib.set_concurrency_max_rows(conc)
    for B in batches:
        state, strat, tok, x = make(B)
        box = {"s": state}

        def eager():
            xx = tok.tokenize(x).reshape(B, -1)
            box["s"], loss, _ = m.train_step(box["s"], 0, xx, strat)
        ms_e = timeit(eager, 30)
        l0 = _lib.launches(); eager(); nl = _lib.launches() - l0
        g = m.GraphedTrainStep(box["s"], strat, tok, x, key=0)
        ms_g = timeit(lambda: g(x), 100)
        print(f"conc_rows={conc:5d} B={B:5d}: eager {ms_e:7.3f} ms  graph {ms_g:7.3f} ms  ({B / ms_g * 1e3:9.0f} samples/s, {nl} launches)", flush=True)
        del state, strat, tok, x, g, box
        torch.cuda.empty_cache()
