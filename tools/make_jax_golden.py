#!/usr/bin/env python
"""Pins the iMF oracle to the reference's own JAX code -- run this wherever jax / flax / optax import.

    python tools/make_jax_golden.py [--ref /root/reference | baseline/_ref] [--out tests/golden/imf_jax.npz]

JAX is not installable in the build container (no wheel, no network), so every iMF row of the
oracle is "parity unpinned" until this script has been run once on a machine that has the reference's dependencies
(jax 0.4.38, flax 0.10.4, optax 0.2.5 per the reference's uv.lock).  It imports the UNMODIFIED reference package, feeds
it explicit draws (the reference's own ``jax.random.normal`` / ``sample_time_pair`` are replaced by functions that
return the arrays stored in the fixture, because its threefry stream cannot be reproduced elsewhere -- SURVEY.md R6)
and stores what the reference computes:

  * forward ``model.apply`` and ``method="encode"``                         models/mlp_flow.py:153-230
  * ``ImprovedMeanFlowLoss.compute_loss`` -> loss, gradients                  trainers/loss_strategies.py:227-280
  * three ``train_step`` / optax.adamw updates                                trainers/training_steps.py:15-34
  * ``sample(n_steps = 1, 2)``                                                evaluators/sampling.py:5-95

for (a) the reference's own test model (test/test_improved_mean_flow.py:34-39: noise 8, cond 32, latent 64, 2 blocks)
and (b) BASELINE configs[0] (D = 1024, L = 256, C = 128, 8 blocks, batch 8).  Parameters are NOT stored: both sides
rebuild them from ``oracle.imf_np.init_params(seed)`` (the arrays are handed to Flax as its param tree), so the fixture
stays small (config (b) stores gradient norms per leaf and a strided sample of the flat gradient instead of 113 MB).

``tests/test_imf_jax_golden.py`` consumes the file when it exists: the fp64 oracle must reproduce it on the CPU
(``-m "not gpu"``) and the CUDA path on the GPU (``-m gpu``).  ``bench.py --impl reference`` uses the same imports
to time the reference's jitted step when they work.
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path
from unittest import mock

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CASES = {
    # name: (D, L, C, nb, B, seed)
    "test_model": (8, 64, 32, 2, 4, 11),
    "config1": (1024, 256, 128, 8, 8, 42),
}


def draws(D, B, seed):
    """The explicit (x, e, t, r, sampler noise) of one case -- also what the consuming test regenerates."""
    from oracle import imf_np
    rng = np.random.default_rng(seed + 1000)
    x = (0.5 * rng.standard_normal((B, D))).astype(np.float32)
    e = rng.standard_normal((B, D)).astype(np.float32)
    t, r = imf_np.sample_tr_from_normals(rng.standard_normal(B).astype(np.float32), rng.standard_normal(B).astype(np.float32))
    z0 = rng.standard_normal((B, D)).astype(np.float32)
    return x, e, t.astype(np.float32).reshape(B, 1), r.astype(np.float32).reshape(B, 1), z0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=None, help="directory that contains the meanflow_audio_codec package")
    ap.add_argument("--out", default=str(ROOT / "tests" / "golden" / "imf_jax.npz"))
    args = ap.parse_args()
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    for cand in ([args.ref] if args.ref else [str(ROOT / "baseline" / "_ref"), "/root/reference"]):
        if cand and (Path(cand) / "meanflow_audio_codec").is_dir():
            sys.path.insert(0, cand)
            break
    import jax
    import jax.numpy as jnp
    import optax
    from meanflow_audio_codec.evaluators.sampling import sample
    from meanflow_audio_codec.models import ConditionalFlow, TrainState
    from meanflow_audio_codec.trainers.loss_strategies import ImprovedMeanFlowLoss
    from meanflow_audio_codec.trainers.noise_schedules import LinearNoiseSchedule
    from meanflow_audio_codec.trainers.time_sampling import MeanFlowTimeSampling
    from meanflow_audio_codec.trainers.training_steps import train_step

    from oracle import imf_np

    out = {"jax_version": np.array(jax.__version__)}
    for name, (D, L, C, nb, B, seed) in CASES.items():
        p_np = imf_np.init_params(D, L, C, nb, seed=seed, bias_scale=0.05)
        tree = jax.tree_util.tree_map(jnp.asarray, imf_np.to_tree(p_np))
        model = ConditionalFlow(noise_dimension=D, condition_dimension=C, latent_dimension=L, num_blocks=nb)
        x, e, t, r, z0 = draws(D, B, seed)
        xj = jnp.asarray(x)
        lat = model.apply({"params": tree}, xj, method="encode")
        th = jnp.concatenate([jnp.asarray(t), jnp.asarray(t - r)], axis=-1)
        out[f"{name}/latents"] = np.asarray(lat)
        out[f"{name}/forward"] = np.asarray(model.apply({"params": tree}, jnp.asarray(e), th, lat))
        out[f"{name}/forward_no_latents"] = np.asarray(model.apply({"params": tree}, jnp.asarray(e), th))

        class FixedTimes(MeanFlowTimeSampling):
            def sample_time_pair(self, key, batch_size, dtype=jnp.float32):
                return jnp.asarray(t, dtype), jnp.asarray(r, dtype)

        strat = ImprovedMeanFlowLoss(LinearNoiseSchedule(0.001, 0.999), FixedTimes(-0.4, 1.0, 0.5), True)
        state = TrainState.create(apply_fn=model.apply, params=tree, tx=optax.adamw(1e-4, weight_decay=1e-4))
        def fixed(arr):
            return lambda key, shape=(), dtype=jnp.float32: jnp.asarray(arr, dtype).reshape(shape)

        with mock.patch.object(jax.random, "normal", fixed(e)):
            loss, grads = strat.compute_loss(state, jax.random.PRNGKey(0), xj)
            flat = np.concatenate([np.asarray(g).reshape(-1) for g in jax.tree_util.tree_leaves(grads)])
            out[f"{name}/loss"] = np.asarray(loss)
            out[f"{name}/grad_norm"] = np.asarray(np.linalg.norm(flat.astype(np.float64)))
            out[f"{name}/grad_leaf_norms"] = np.array([np.linalg.norm(np.asarray(g, np.float64)) for g in jax.tree_util.tree_leaves(grads)])
            stride = max(1, flat.size // 65536)
            out[f"{name}/grad_stride"] = np.array(stride)
            out[f"{name}/grad_sample"] = flat[::stride].copy()
            losses = []
            st = state
            for _ in range(3):       # the reference re-draws the same (e, t, r) every step (R6): fixed draws ARE its behaviour
                st, l, _ = train_step(st, jax.random.PRNGKey(0), xj, strat)
                losses.append(float(l))
            out[f"{name}/train3_losses"] = np.array(losses)
            pflat = np.concatenate([np.asarray(q).reshape(-1) for q in jax.tree_util.tree_leaves(st.params)])
            p0 = np.concatenate([np.asarray(q).reshape(-1) for q in jax.tree_util.tree_leaves(tree)])
            out[f"{name}/train3_param_delta_sample"] = (pflat - p0)[::stride].copy()
        with mock.patch.object(jax.random, "normal", fixed(z0)):
            for n in (1, 2):
                out[f"{name}/sample_heun_{n}"] = np.asarray(sample(model.apply, D, tree, jax.random.PRNGKey(1), latents=lat, n_steps=n))
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(args.out, **out)
    print(f"wrote {args.out}: {len(out)} arrays, jax {jax.__version__}")


if __name__ == "__main__":
    main()
