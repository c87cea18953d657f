#!/usr/bin/env python
"""bench.py -- iMF training throughput (samples/s) of the B200-native hot path, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl mfac|reference] [--batch B] [--noise-dimension T]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one full training step of the reference's hot loop (trainers/train.py:333-347) on one
batch of synthetic audio: MDCT tokenisation -> ImprovedMeanFlowLoss.compute_loss (v pass, u pass +
JVP, loss, backward) -> gradient all-reduce (N > 1) -> AdamW.  Workload = BASELINE.json configs[1]
(method=improved_mean_flow, architecture=mlp, dataset=audio, tokenization=mdct) with the substitution
SURVEY.md R4 forces: the shipped noise_dimension=196608 needs 309 G parameters per block and cannot be
instantiated anywhere, so noise_dimension is a flag (default 784 samples -> 2 MDCT frames -> D=1024, the
geometry of the one runnable config); every other hyper-parameter is the config's.  The per-GPU batch is
a flag too: the config's batch_size=128 is launch/optimizer-bound on a B200, so the default is 37888 = 2 x 148 x 128
rows: the full-batch passes run two whole waves of 128-row GEMM tiles per SM and the half-batch passes (the rows with
r != t, data_proportion = 0.5) exactly one.  The config-faithful 128 and 18944 are reported next to it under "sweep".

Keys beyond the base contract: "roofline" (tcgen05 GEMM family, per-launch CUDA events), "cpu_baseline"
(oracle port on the host cores), "e2e" (host buffers through the public Python API), "clocks", "codec"
(MDCT -> encode -> 1-NFE mean-flow sample -> IMDCT audio-seconds/s at 1 GPU), "sweep".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

CFG = dict(condition_dimension=128, latent_dimension=256, num_blocks=8, base_lr=1e-4, weight_decay=1e-4,
           window_size=512, hop_size=256, seed=42)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sustained=float(d["bf16_tflops_sustained"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def csrc_hash() -> str:
    """sha256 over the CUDA sources the shipped libmfac.so is built from (sorted by name): an ncu traffic figure is only
    quoted when it was captured on exactly these sources."""
    import hashlib
    h = hashlib.sha256()
    d = ROOT / "meanflow_audio_codec_b200" / "csrc"
    for f in sorted(list(d.glob("*.cu")) + list(d.glob("*.cuh")) + [ROOT / "include" / "mfac.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def ncu_traffic():
    """(bytes per launch | None, note): profiles/ncu_traffic.json is refused when its recorded source hash is not HEAD's."""
    tf = ROOT / "profiles" / "ncu_traffic.json"
    if not tf.exists():
        return None, "no ncu capture committed"
    try:
        d = json.loads(tf.read_text())
    except Exception as ex:  # noqa: BLE001
        return None, f"unreadable ncu_traffic.json: {ex}"
    have, want = d.get("csrc_hash"), csrc_hash()
    if have != want:
        return None, f"stale capture refused: ncu_traffic.json was taken on csrc {have}, this build is {want}"
    return d.get("gemm_tcgen05_dram_bytes_per_launch"), d.get("source")


def token_dim(T: int) -> tuple[int, int]:
    N, hop = CFG["window_size"], CFG["hop_size"]
    nf = 1 if T < N else (T - N) // hop + 1
    return nf, nf * N


def flops_per_sample(D: int) -> float:
    """B * (3 * 2 * P_enc + 5 * 2 * P_dec) GEMM FLOPs."""
    L, C, nb = CFG["latent_dimension"], CFG["condition_dimension"], CFG["num_blocks"]
    I, He = L + D, (D + L) // 2
    p_enc = D * He + He * L
    p_dec = nb * (C * C + C * (2 * I + D) + I * I + I * D)
    return 3 * 2 * p_enc + 5 * 2 * p_dec


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self) -> dict:
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- reference arm
def cpu_reference_step_factory(T: int, B: int):
    """The reference's CPU path for one step, restated (oracle/): MDCT tokenisation -> iMF loss+grads -> AdamW.
    JAX is not installable in this image (no wheel, no network), so this is the torch-CPU port ("kind": "port")."""
    import numpy as np
    import torch
    from oracle import imf_np, imf_torch, mdct_np
    nf, D = token_dim(T)
    p = {k: torch.from_numpy(v) for k, v in imf_np.init_params(D, CFG["latent_dimension"], CFG["condition_dimension"],
                                                               CFG["num_blocks"], seed=CFG["seed"]).items()}
    mu = {k: torch.zeros_like(v) for k, v in p.items()}
    nu = {k: torch.zeros_like(v) for k, v in p.items()}
    g = torch.Generator().manual_seed(42)
    x_raw = (0.1 * torch.randn(B, T, generator=g)).numpy()
    w = torch.from_numpy(mdct_np.window_2n(CFG["window_size"], np.float32))
    Cb = torch.from_numpy(mdct_np.cosine_basis(CFG["window_size"], np.float32))
    N, hop = CFG["window_size"], CFG["hop_size"]
    state = {"count": 0}

    def step():
        xr = torch.from_numpy(x_raw)
        need = (nf - 1) * hop + 2 * N
        xr = torch.nn.functional.pad(xr, (0, max(0, need - T)))
        tok = torch.stack([(xr[:, i * hop:i * hop + 2 * N] * w) @ Cb for i in range(nf)], 1).reshape(B, D)
        e = torch.randn(B, D, generator=g)
        n2 = torch.randn(2, B, generator=g)
        t = torch.sigmoid(n2[0] - 0.4); r = torch.sigmoid(n2[1] - 0.4)
        t, r = torch.maximum(t, r), torch.minimum(t, r)
        r = torch.where(torch.arange(B) < B // 2, t, r)
        loss, grads, _ = imf_torch.imf_loss_and_grads(p, tok, e, t[:, None], r[:, None])
        imf_torch.adamw_step(p, grads, mu, nu, state["count"], lr=CFG["base_lr"], wd=CFG["weight_decay"])
        state["count"] += 1
        return float(loss)

    return step


def jax_reference_step_factory(T: int, B: int):
    """The reference's OWN code (kind "reference"): its mdct (direct branch forced), ImprovedMeanFlowLoss.
    compute_loss and TrainState.apply_gradients wrapped in ONE jax.jit (the shipped loop is un-jitted, R7; jitting it is the
    fair baseline BASELINE.md section 2 prescribes), on the host cores.  Raises ImportError where jax / flax / optax or the
    reference package (baseline/_ref, /root/reference) are absent -- the caller then falls back to the torch port."""
    os.environ.setdefault("JAX_PLATFORMS", "cpu")
    for cand in (ROOT / "baseline" / "_ref", Path("/root/reference")):
        if (cand / "meanflow_audio_codec").is_dir() and str(cand) not in sys.path:
            sys.path.insert(0, str(cand))
    import jax
    import jax.numpy as jnp
    import numpy as np
    import optax
    from meanflow_audio_codec.models import ConditionalFlow, TrainState
    from meanflow_audio_codec.preprocessing.mdct import mdct
    from meanflow_audio_codec.trainers.loss_strategies import ImprovedMeanFlowLoss
    from meanflow_audio_codec.trainers.noise_schedules import LinearNoiseSchedule
    from meanflow_audio_codec.trainers.time_sampling import MeanFlowTimeSampling
    nf, D = token_dim(T)
    model = ConditionalFlow(noise_dimension=D, condition_dimension=CFG["condition_dimension"],
                            latent_dimension=CFG["latent_dimension"], num_blocks=CFG["num_blocks"])
    key = jax.random.PRNGKey(CFG["seed"])
    k1, k2, k3 = jax.random.split(key, 3)
    x0, t0 = jnp.zeros((B, D), jnp.float32), jnp.zeros((B, 2), jnp.float32)
    params_enc = model.init(k1, x0, method="encode")["params"]                     # trainers/train.py:246-259
    params_dec = model.init(k2, x0, t0, jnp.zeros((B, CFG["latent_dimension"]), jnp.float32))["params"]
    params = {**params_dec, "encoder": params_enc["encoder"]}
    state = TrainState.create(apply_fn=model.apply, params=params,
                              tx=optax.adamw(learning_rate=CFG["base_lr"], weight_decay=CFG["weight_decay"]))
    strat = ImprovedMeanFlowLoss(LinearNoiseSchedule(0.001, 0.999), MeanFlowTimeSampling(-0.4, 1.0, 0.5), True)
    x_raw = jnp.asarray(0.1 * np.random.default_rng(42).standard_normal((B, T)).astype(np.float32))

    @jax.jit
    def jstep(state, key, x_raw):
        tok = mdct(x_raw, window_size=CFG["window_size"], hop_size=CFG["hop_size"], use_fft_threshold=10 ** 9).reshape(B, -1)
        loss, grads = strat.compute_loss(state, key, tok)
        return state.apply_gradients(grads=grads), loss

    box = {"state": state, "key": k3}

    def step():
        box["key"], sub = jax.random.split(box["key"])
        box["state"], loss = jstep(box["state"], sub, x_raw)
        return float(loss)     # device sync every step, like trainers/train.py:347

    return step


def time_cpu_reference(T: int, B: int, steps: int, warmup: int, budget_s: float = 25.0, try_jax: bool = True):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    kind, why = "port", None
    step = None
    if try_jax:
        try:
            step = jax_reference_step_factory(T, B)
            step()                     # compile
            kind = "reference"
        except Exception as ex:  # noqa: BLE001 -- ImportError here; anything else must not kill the bench either
            step, why = None, f"{type(ex).__name__}: {ex}"[:160]
    if step is None:
        step = cpu_reference_step_factory(T, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    what = ("the reference's own jitted JAX step (mdct direct branch + ImprovedMeanFlowLoss + optax.adamw), JAX_PLATFORMS=cpu"
            if kind == "reference" else "torch-CPU restatement of the reference step (oracle/), fp32")
    out = dict(value=B * done / dt, unit="samples/s", cores=os.cpu_count() if kind == "reference" else torch.get_num_threads(),
               kind=kind, sample=f"{done} steps of batch {B} ({what}, T={T})", ms_per_step=dt / done * 1e3)
    if why:
        out["reference_unavailable"] = why
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nf, D = token_dim(args.noise_dimension)
    B = min(args.batch, 256)
    cb = time_cpu_reference(args.noise_dimension, B, max(1, args.steps), min(args.warmup, 1), budget_s=90.0)
    line = {
        "metric": "imf_train_samples_per_sec", "value": cb["value"], "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, B, D, nf),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "reference_unavailable") if k in cb},
        "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["config"]["sample"] = (f"each CPU step is a bounded sample of {B} rows of the workload (the GPU arm's per-GPU batch is "
                                f"{args.batch}); our arm quotes the ratio at the matching batch under 'matching_batch'")
    print(json.dumps(line))
    return 0


def workload_config(args, B, D, nf):
    return {
        "workload": (f"imf_train_step method=improved_mean_flow architecture=mlp dataset=audio(synthetic 0.1*randn) "
                     f"tokenization=mdct(N=512,hop=256) noise_dimension={args.noise_dimension}->nf={nf},D={D} "
                     f"L=256 C=128 blocks=8 per_gpu_batch={B}"),
        "substitution": "config noise_dimension=196608 is not instantiable (SURVEY.md R4); batch_size=128 reported under sweep",
        "global_batch": B * args.gpus, "per_gpu_batch": B, "parallelism": f"dp{args.gpus}",
        "l2": "working set per step (fp32 params+grads+AdamW moments+activations) exceeds the 126 MB L2",
    }


# --------------------------------------------------------------------------------------------- our arm
def run_mfac(args):
    import torch
    import meanflow_audio_codec_b200 as m
    from meanflow_audio_codec_b200 import _lib
    from meanflow_audio_codec_b200.data_parallel import DataParallel, train_step_dp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl mfac needs a CUDA device (no CPU fallback)")
    dp = DataParallel()
    world, rank = dp.world, dp.rank
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(dp.local_rank)
    dev = torch.device("cuda", dp.local_rank)
    if dp.enabled and os.environ.get("MFAC_DP_LIBCOMM", "0") == "1":
        # Optional: small batches then exchange gradient buckets through libmfac's own communicator (mfac_comm_*) from inside
        # the fused step.  Not the default: measured on 8 x B200 it is no faster than one torch.distributed all-reduce after the
        # backward (128 rows / GPU: 1.64 vs 1.54 ms / step; 4096 rows: 3.27 vs 3.23 ms), and a second live communicator alone cost
        # the run-ahead loop 2 ms / step at 37 888 rows (16.98 vs 15.06 ms; its proxy threads compete with 8 launch threads).
        dp.init_library_comm()
    T = args.noise_dimension
    nf, D = token_dim(T)
    pk = peaks()

    def make(B, Tm=None):
        Tm = T if Tm is None else Tm
        Dm = token_dim(Tm)[1]
        model = m.ConditionalFlow(noise_dimension=Dm, condition_dimension=CFG["condition_dimension"],
                                  num_blocks=CFG["num_blocks"], latent_dimension=CFG["latent_dimension"])
        # same weights on every rank; the wide geometries of the noise-dimension sweep (up to 1.05 G parameters) are drawn on the device
        params = model.init(CFG["seed"], device=dev, on_device=Dm > 2048)["params"]
        state = m.TrainState.create(apply_fn=model.apply, params=params, tx=m.adamw(CFG["base_lr"], CFG["weight_decay"]))
        strat = m.ImprovedMeanFlowLoss(m.LinearNoiseSchedule(0.001, 0.999), m.MeanFlowTimeSampling(-0.4, 1.0, 0.5), True)
        # lazy: the step tokenises inside its own prologue (mfac_imf_train_step_audio) -- no [B, nf, N] token tensor in HBM
        tok = m.MDCTTokenization(window_size=CFG["window_size"], hop_size=CFG["hop_size"], lazy=os.environ.get("MFAC_BENCH_EAGER_TOKENS") != "1")
        g = torch.Generator(device=dev).manual_seed(42 + rank)
        x_raw = 0.1 * torch.randn(B, Tm, device=dev, generator=g)
        return model, state, strat, tok, x_raw

    def step_fn_eager(state, strat, tok, x_raw, key=0):
        x = tok.tokenize(x_raw).reshape(x_raw.shape[0], -1)
        state, loss, _ = train_step_dp(dp, state, key, x, strat)
        return state, loss

    step_fn = step_fn_eager

    def timed(B, steps, warmup, sample_clocks=False, graphed=False, Tm=None, profile=False):
        model, state, strat, tok, x_raw = make(B, Tm)
        if graphed:   # the whole step (tokenise + loss/grad + AdamW) as ONE CUDA-graph launch; single GPU only
            gstep = m.GraphedTrainStep(state, strat, tok, x_raw, key=0)

            def step_fn(state, strat, tok, x_raw, key=0):  # noqa: F811
                return state, gstep(x_raw)
        else:
            step_fn = step_fn_eager
        for _ in range(warmup):
            state, loss = step_fn(state, strat, tok, x_raw)
        torch.cuda.synchronize()
        dp.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(dp.local_rank) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            state, loss = step_fn(state, strat, tok, x_raw)
        e1.record()
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        dp.barrier()
        if sampler:
            sampler.__exit__()
        ms = dp.max_over_ranks(e0.elapsed_time(e1), dev)
        launches = (_lib.launches() - l0) / steps
        gemm_flops = None
        if profile:   # executed tensor FLOPs of one more (untimed) step, counted per GEMM launch
            _lib.profile_enable(True)
            state, loss = step_fn(state, strat, tok, x_raw)
            gemm_flops = _lib.profile_collect()["gemm_tcgen05"]["work"]
            _lib.profile_enable(False)
        return dict(ms_total=ms, ms_per_step=ms / steps, wall_ms=(w1 - w0) * 1e3, launches=launches, loss=float(loss),
                    clocks=sampler.summary() if sampler else None, objs=(model, state, strat, tok, x_raw), gemm_flops=gemm_flops)

    B = args.batch
    main = timed(B, args.steps, args.warmup, sample_clocks=True)
    value = B * world * args.steps / (main["ms_total"] * 1e-3)
    model, state, strat, tok, x_raw = main["objs"]

    # ---- end to end through the public API with HOST buffers (pinned): H2D of the raw audio + D2H of the loss per step
    # Input pipeline: two pinned host batches, uploaded on a copy stream into two device buffers so that the upload of
    # step i+1 overlaps the compute of step i; the loss is read back to the host every step.
    x_host = [x_raw.cpu().pin_memory(), x_raw.flip(0).cpu().pin_memory()]
    stager = m.HostBatchStager(device=dev)     # the package's input pipeline (input_pipeline.py)

    def e2e_loop(n):
        nonlocal state
        lv = None
        for x_d in stager.stream(x_host[i & 1] for i in range(n)):
            state, loss = step_fn(state, strat, tok, x_d)
            lv = float(loss)  # device -> host read every step, like trainers/train.py:347
        return lv

    e2e_loop(2)
    torch.cuda.synchronize(); dp.barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = dp.max_over_ranks(time.perf_counter() - t0, dev)
    dp.barrier()
    e2e = {"value": B * world * args.steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": int(x_host[0].numel() * 4),
           "d2h_bytes_per_step": 4, "api": "HostBatchStager.stream (pinned host audio, double-buffered upload on a copy stream) -> "
           "MDCTTokenization(lazy).tokenize -> train_step (ImprovedMeanFlowLoss, fused mfac_imf_train_step_audio); loss read back every step"}

    # ---- roofline of the dominant kernel family (tcgen05 GEMM): per-launch CUDA events on the launch stream
    roofline, families = None, None
    if True:
        _lib.profile_enable(True)
        nprof = max(1, min(3, args.steps))
        for _ in range(nprof):
            state, loss = step_fn(state, strat, tok, x_raw)
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        gm = prof["gemm_tcgen05"]
        if gm["launches"]:
            achieved = gm["work"] / (gm["ms"] * 1e-3) / 1e12
            traffic, traffic_note = ncu_traffic()
            roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all fused-epilogue instantiations)",
                        "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                        "frac": achieved / pk["bf16_sustained"], "traffic": traffic, "traffic_source": traffic_note,
                        "peak_source": pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                        "launches_per_step": gm["launches"] / nprof, "ms_per_step_in_kernel": gm["ms"] / nprof,
                        "share_of_step": gm["ms"] / nprof / main["ms_per_step"],
                        "flops_per_step_measured": gm["work"] / nprof, "flops_per_step_survey": flops_per_sample(D) * B}
        families = {k: {"launches_per_step": v["launches"] / nprof, "ms_per_step": v["ms"] / nprof,
                        "achieved": (v["work"] / (v["ms"] * 1e-3) / (1e12 if k == "gemm_tcgen05" else 1e9)) if v["ms"] > 0 else None,
                        "unit": "TFLOP/s" if k == "gemm_tcgen05" else "GB/s"} for k, v in prof.items() if v["launches"]}

    sweep, codec, cpu_baseline, matching, scaling_small = None, None, None, None, None
    main_ms, main_wall, main_loss, main_launches, main_clocks = (main["ms_per_step"], main["wall_ms"], main["loss"], main["launches"],
                                                                main["clocks"])
    if world > 1 and not args.quick:
        # the exchange step where it hurts: the config-faithful 128 rows per GPU and a tensor-bound 4096
        main = None
        model = state = strat = tok = x_raw = None
        torch.cuda.empty_cache()
        scaling_small = {}
        for b in (128, 4096):
            r = timed(b, max(10, args.steps), max(3, args.warmup))
            scaling_small[f"per_gpu_batch_{b}"] = {"samples_per_s": b * world * max(10, args.steps) / (r["ms_total"] * 1e-3),
                                                   "ms_per_step": r["ms_per_step"], "global_batch": b * world}
            del r
            torch.cuda.empty_cache()
    if world == 1 and not args.quick:
        # config-faithful batch and a few others, same protocol (shorter)
        sweep = {}
        for b in args.sweep:
            if b == B:
                continue
            ks = max(3, args.steps // 2)
            r = timed(b, ks, max(3, args.warmup))
            sweep[f"per_gpu_batch_{b}"] = {"samples_per_s": b * ks / (r["ms_total"] * 1e-3), "ms_per_step": r["ms_per_step"],
                                           "gpu_launches": r["launches"],
                                           "tensor_frac_whole_step_reference_flops": flops_per_sample(D) * b / (r["ms_per_step"] * 1e-3) / 1e12 / pk["bf16_sustained"]}
            del r
            torch.cuda.empty_cache()
            if b <= 1024:   # launch-bound sizes: the same step replayed as one CUDA graph
                r = timed(b, ks * 4, max(3, args.warmup), graphed=True)
                sweep[f"per_gpu_batch_{b}_cuda_graph"] = {"samples_per_s": b * ks * 4 / (r["ms_total"] * 1e-3),
                                                          "ms_per_step": r["ms_per_step"], "graph_launches_per_step": 1}
                del r
                torch.cuda.empty_cache()
        # the audio config's other hyper-parameters with tractable noise dimensions.  At D = 1024
        # the step is HBM-bound; the wider geometries are where the tensor roof binds.
        main = None   # release the headline run's 17 GB workspace first
        model = state = strat = tok = x_raw = None
        torch.cuda.empty_cache()
        noise_sweep = {}
        # batches = whole waves of 256-row tile pairs on 74 SM pairs (18 944 = 74 pairs, 9472 = 37), like the headline's 37 888
        for Tn, bn in ((2048, 18944), (4096, 9472)):
            try:
                r = timed(bn, 3, 3, Tm=Tn, profile=True)
                nfn, Dn = token_dim(Tn)
                noise_sweep[f"noise_dimension_{Tn}"] = {
                    "D": Dn, "frames": nfn, "per_gpu_batch": bn, "ms_per_step": r["ms_per_step"],
                    "samples_per_s": bn * 3 / (r["ms_total"] * 1e-3),
                    "tensor_frac_whole_step": r["gemm_flops"] / (r["ms_per_step"] * 1e-3) / 1e12 / pk["bf16_sustained"],
                    "tensor_frac_whole_step_reference_flops": flops_per_sample(Dn) * bn / (r["ms_per_step"] * 1e-3) / 1e12 / pk["bf16_sustained"]}
                del r
            except Exception as ex:  # noqa: BLE001 -- a sweep point must not cost the headline line
                noise_sweep[f"noise_dimension_{Tn}"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
            torch.cuda.empty_cache()
        sweep["noise_dimension"] = noise_sweep
        codec = codec_bench(m, _lib, dev, pk)
        codec["other_architectures_forward"] = flows_bench(m, dev)
        codec["cpu_baseline"] = codec_cpu_baseline()
        cpu_baseline = time_cpu_reference(T, 128, 50, 1, budget_s=20.0)
        cpu_baseline.pop("ms_per_step", None)
        g128 = sweep.get("per_gpu_batch_128_cuda_graph") or sweep.get("per_gpu_batch_128")
        if g128:
            matching = {"batch": 128, "gpu_samples_per_s": g128["samples_per_s"], "cpu_samples_per_s": cpu_baseline["value"],
                        "cpu_kind": cpu_baseline["kind"], "ratio": g128["samples_per_s"] / cpu_baseline["value"],
                        "note": "the config's own batch_size on both arms (device-resident, one CUDA-graph launch per step on the GPU)"}

    if rank == 0:
        line = {
            "metric": "imf_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, B, D, nf),
            "e2e": e2e, "gpu_launches": main_launches, "clocks": main_clocks,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            # executed tensor FLOPs of a step (counted per GEMM launch) over the whole step time; the "reference_flops"
            # figure divides the FLOPs the reference's schedule would spend (three full network evaluations +
            # tangent + backward) -- the step shares one evaluation between u and v on the rows with r == t
            "tensor_frac_whole_step": ((roofline["flops_per_step_measured"] if roofline else flops_per_sample(D) * B)
                                       / (main_ms * 1e-3) / 1e12 / pk["bf16_sustained"]),
            "tensor_frac_whole_step_reference_flops": flops_per_sample(D) * B / (main_ms * 1e-3) / 1e12 / pk["bf16_sustained"],
            # SURVEY.md section 8d config 2 (the audio config's hyper-parameters at tractable noise dimensions): where the tensor
            # roof binds, whole step on executed FLOPs
            "tensor_frac_whole_step_by_noise_dimension": ({k: v.get("tensor_frac_whole_step") for k, v in sweep["noise_dimension"].items()}
                                                          if sweep and "noise_dimension" in sweep else None),
            "kernel_families": families, "sweep": sweep, "codec": codec, "matching_batch": matching,
            "scaling_small": scaling_small, "csrc_hash": csrc_hash(),
            "wall_ms_per_step": main_wall / args.steps, "loss": main_loss,
        }
        print(json.dumps(line))
    dp.destroy()
    return 0


CODEC_T, CODEC_CLIP_SECONDS = 441000, 10.0


def codec_flops_per_clip(nfe_evals: int) -> float:
    """one encode + ``nfe_evals`` velocity evaluations per model row, 861 rows per 10 s clip."""
    D, L, C, nb = 1024, CFG["latent_dimension"], CFG["condition_dimension"], CFG["num_blocks"]
    I, He = L + D, (D + L) // 2
    p_enc = D * He + He * L
    p_dec = nb * (C * C + C * (2 * I + D) + I * I + I * D)
    return 861 * 2.0 * (p_enc + nfe_evals * p_dec)


def codec_cpu_baseline(clips: int = 2, reps: int = 3):
    """The codec pipeline on the host cores through the oracle port (torch fp32): direct-cosine MDCT as a matmul (what the
    reference's einsum does), MLP encode, 1-NFE mean-flow sample, IMDCT + overlap-add.  Bounded sample: a few clips."""
    import numpy as np
    import torch
    from oracle import imf_np, imf_torch, mdct_np
    torch.set_num_threads(os.cpu_count() or 1)
    N, hop, D = CFG["window_size"], CFG["hop_size"], 1024
    T = CODEC_T + hop                                    # 1722 frames = 861 rows
    nf = (T - N) // hop + 1
    p = {k: torch.from_numpy(v) for k, v in imf_np.init_params(D, CFG["latent_dimension"], CFG["condition_dimension"],
                                                               CFG["num_blocks"], seed=CFG["seed"]).items()}
    w = torch.from_numpy(mdct_np.window_2n(N, np.float32))
    Cb = torch.from_numpy(mdct_np.cosine_basis(N, np.float32))
    g = torch.Generator().manual_seed(42)
    x = 0.1 * torch.randn(clips, CODEC_T, generator=g)

    def run():
        xp = torch.nn.functional.pad(x, (0, (nf - 1) * hop + 2 * N - CODEC_T))
        frames = xp.unfold(1, 2 * N, hop)[:, :nf]                             # [clips, nf, 2N]
        X = (frames * w) @ Cb                                                 # [clips, nf, N]
        rows = X.reshape(-1, D)
        lat = imf_torch.encode(p, rows)
        rec = imf_torch.mf_sample(p, lat, torch.randn(rows.shape, generator=g), nfe=1)
        Y = (rec.reshape(clips, nf, N) @ Cb.T) * (2.0 / N) * w                # [clips, nf, 2N]
        out = torch.zeros(clips, (nf - 1) * hop + 2 * N)
        for i in range(nf):
            out[:, i * hop:i * hop + 2 * N] += Y[:, i]
        return out

    with torch.no_grad():
        run()
        t0 = time.perf_counter()
        for _ in range(reps):
            run()
        dt = (time.perf_counter() - t0) / reps
    return {"value": clips * CODEC_CLIP_SECONDS / dt, "unit": "audio-s/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} passes over {clips} clips of 10 s (torch-CPU restatement: MDCT matmul, encode, 1-NFE sample, IMDCT+OLA)"}


def codec_bench(m, _lib, dev, pk, batches=(1, 4, 16, 64, 256, 1024, 4096), quick=False):
    """BASELINE configs[4]: MDCT -> encode -> 1-/2-NFE mean-flow (and Heun n=1) sample -> IMDCT on 10 s synthetic 44.1 kHz
    clips, batch 1..4096, through ``MeanFlowCodec`` (the package's public call).  Per batch and sampler: audio-seconds/s with
    the clips resident in HBM ("value") and from pinned HOST buffers with the reconstructed audio copied back to the host
    ("e2e"; sub-batches of <= 256 clips stream through two slots per direction).  Plus the MDCT / IMDCT kernel rooflines
    (HBM) and the pipeline's tensor roofline at the quoted size."""
    import torch
    from meanflow_audio_codec_b200.codec import MeanFlowCodec
    T, N, hop, D = CODEC_T, 512, 256, 1024
    model = m.ConditionalFlow(D, CFG["condition_dimension"], CFG["num_blocks"], CFG["latent_dimension"])
    params = model.init(CFG["seed"], device=dev)["params"]
    codec = MeanFlowCodec(model, params, N, hop)
    out = {"workload": "10 s clips @ 44.1 kHz (T=441000), N=512 hop=256 -> 1721 (+1 pad) frames -> 861 rows of D=1024 per clip; "
                       "MLP flow L=256 C=128 blocks=8; synthetic 0.1*randn audio, random-init weights"}
    time.sleep(1.0)   # the kernel rooflines come first and after a pause: the GEMM-heavy legs before and after leave the part power-capped
    # each kernel timed alone, back to back (per-launch CUDA events on the launch stream); 256 clips = 451 MB in, 902 MB of
    # coefficients (far beyond the 126 MB L2) is the quoted size, 1024 clips shows the large-batch end of the sweep
    for nclips in (256, 1024):
        x = 0.1 * torch.randn(nclips, T, device=dev)
        for _ in range(3):
            X = m.mdct(x, N, hop, use_fft_threshold=N + 1)
            y = m.imdct(X, N, hop, use_fft_threshold=N + 1)
        torch.cuda.synchronize()
        _lib.profile_enable(True)
        for _ in range(10):
            X = m.mdct(x, N, hop, use_fft_threshold=N + 1)
        for _ in range(10):
            y = m.imdct(X, N, hop, use_fft_threshold=N + 1)
        prof = _lib.profile_collect()
        _lib.profile_enable(False)
        for fam in ("mdct512", "imdct512"):
            v = prof[fam]
            gbs = v["work"] / (v["ms"] * 1e-3) / 1e9
            out[fam if nclips == 256 else f"{fam}_{nclips}clips"] = {
                "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
                "clips": nclips, "ms_per_launch": v["ms"] / v["launches"]}
        del x, X, y
        torch.cuda.empty_cache()

    samplers = {"mf1": ("mf", 1, 1), "mf2": ("mf", 2, 2), "heun1": ("heun", 1, 2)}   # name -> (sampler, nfe, network evaluations)
    SUB = 256
    res_cap = 1024                                   # clips kept resident / pinned at once; larger batches cycle through them
    gen = torch.Generator(device=dev).manual_seed(42)
    # resident clips sit in the codec's staging layout: rows as wide as the padded clip (one extra hop of zeros -> an even frame
    # count), so neither leg pays a padding copy; the pinned host copy holds the T real samples per clip
    t_pad = codec.geometry(T)["t_pad"]
    x_res = torch.zeros(res_cap if max(batches) >= res_cap else max(batches), t_pad, device=dev)
    x_res[:, :T] = 0.1 * torch.randn(x_res.shape[0], T, device=dev, generator=gen)
    x_pin = x_res[:, :T].cpu().pin_memory()
    y_pin = torch.empty((x_pin.shape[0], codec.geometry(T)["out_len"]), dtype=torch.float32).pin_memory()
    sweep = {}
    for Bc in batches:
        entry = {}
        for name, (smp, nfe, evals) in samplers.items():
            if quick and name != "mf1":
                continue

            def run_device():
                done = 0
                while done < Bc:
                    n = min(SUB, Bc - done)
                    a = done % x_res.shape[0]
                    y = codec.reconstruct(x_res[a:a + n], sampler=smp, nfe=nfe, key=done, valid_length=T)
                    done += n
                return y

            def run_host():
                done = 0
                while done < Bc:
                    n = min(x_pin.shape[0], Bc - done)
                    # reconstruct_host picks its own sub-batch (min(64, B, 2 sqrt(B)) clips): H2D / kernels / D2H of neighbouring
                    # sub-batches overlap, and the un-overlapped first upload / last download stay short
                    codec.reconstruct_host(x_pin[:n], out_host=y_pin[:n], sampler=smp, nfe=nfe, key=done, sub_batch=None, device=dev)
                    done += n

            iters = 3 if Bc <= 256 else 1
            run_device()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                run_device()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            run_host()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                run_host()
            torch.cuda.synchronize()
            ms_h = (time.perf_counter() - t0) * 1e3 / iters
            fl = codec_flops_per_clip(evals) * Bc
            entry[name] = {"audio_seconds_per_s": Bc * CODEC_CLIP_SECONDS / (ms * 1e-3), "ms": ms,
                           "e2e_audio_seconds_per_s": Bc * CODEC_CLIP_SECONDS / (ms_h * 1e-3), "e2e_ms": ms_h,
                           "tensor_frac": fl / (ms * 1e-3) / 1e12 / pk["bf16_sustained"]}
        entry["h2d_bytes"] = Bc * T * 4
        entry["d2h_bytes"] = Bc * codec.geometry(T)["out_len"] * 4
        sweep[f"clips_{Bc}"] = entry
    out["sweep"] = sweep
    out["sweep_note"] = (f"batches > {SUB} clips run as sub-batches of {SUB}; batches > {res_cap} cycle through {res_cap} resident / pinned clips "
                         "(14 GB of pinned host memory for 4096 clips is not needed to time the stream); e2e = pinned host audio in, "
                         "reconstructed audio back in pinned host memory, H2D / kernels / D2H overlapped on three streams")
    q = sweep.get("clips_256") or sweep[sorted(sweep)[-1]]
    out["pipeline_roofline"] = {"bound": "tensor", "clips": 256, "sampler": "mf1", "flops_per_clip": codec_flops_per_clip(1),
                                "hbm_bytes_per_clip_mdct_imdct": 4 * (CODEC_T + 2 * 1721 * 512 + 441344),
                                "achieved": q["mf1"]["tensor_frac"] * pk["bf16_sustained"], "peak": pk["bf16_sustained"],
                                "unit": "TFLOP/s", "frac": q["mf1"]["tensor_frac"]}
    # serving a few clips at a time: the eager call is host-launch-bound there, GraphedCodec submits the pipeline as one graph
    try:
        graphed = {}
        for clips in (1, 4):
            run = m.GraphedCodec(codec, clips=clips, T=T, sampler="mf", nfe=1)
            a = x_res[:clips, :T].contiguous()
            for _ in range(3):
                run(a)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                run(a)
            e1.record()
            torch.cuda.synchronize()
            gms = e0.elapsed_time(e1) / 20
            graphed[f"clips_{clips}"] = {"ms": gms, "audio_seconds_per_s": clips * 10.0 / (gms * 1e-3)}
            del run
        out["mf1_cuda_graph"] = graphed
    except Exception as ex:  # noqa: BLE001 -- an extra line must not cost the headline
        out["mf1_cuda_graph"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    for k in ("clips_16", "clips_64", "clips_256"):        # round-1 key layout, kept for continuity
        if k in sweep:
            out[k] = {"audio_seconds_per_s": sweep[k]["mf1"]["audio_seconds_per_s"], "ms": sweep[k]["mf1"]["ms"]}
    del x_res, x_pin, y_pin
    torch.cuda.empty_cache()
    return out


def flows_bench(m, dev):
    """BASELINE configs[2]/[3]: forward-only velocity throughput of the mixer and convnet flows at D=1024."""
    import torch
    out = {}
    for name, cls, Bf in (("mlp_mixer", m.ConditionalMLPMixerFlow, 256), ("convnet", m.ConditionalConvFlow, 2048)):
        model = cls(1024, CFG["condition_dimension"], CFG["num_blocks"], CFG["latent_dimension"])
        params = model.init(CFG["seed"], device=dev)["params"]
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randn(Bf, 1024, device=dev, generator=g)
        t = torch.rand(Bf, 2, device=dev, generator=g)
        lat = torch.randn(Bf, 32, CFG["latent_dimension"], device=dev, generator=g)
        for _ in range(3):
            y = model.apply({"params": params}, x, t, lat)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = model.apply({"params": params}, x, t, lat)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[name] = {"rows_per_s": Bf / (ms * 1e-3), "ms": ms, "batch": Bf, "blocks": CFG["num_blocks"], "D": 1024}
        del model, params, x, lat, y
        torch.cuda.empty_cache()
        # the codec pipeline with this architecture (SURVEY.md section 8f-1 encoder wiring: MLP encoder -> latents[B, 1, L] -> the
        # model's own latent_proj): MDCT -> encode -> 1-NFE mean-flow jump -> IMDCT on 10 s clips, device-resident
        try:
            clips = 16
            cm = cls(1024, CFG["condition_dimension"], CFG["num_blocks"], CFG["latent_dimension"], num_latent_tokens=1)
            cp = cm.init(CFG["seed"], device=dev, with_encoder=True)["params"]
            codec = m.MeanFlowCodec(cm, cp, window_size=512, hop_size=256)
            gg = codec.geometry(441000)
            a = 0.1 * torch.randn(clips, gg["t_pad"], device=dev, generator=g)
            a[:, 441000:] = 0
            for _ in range(2):
                yy = codec.reconstruct(a, sampler="mf", nfe=1, valid_length=441000)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                yy = codec.reconstruct(a, sampler="mf", nfe=1, valid_length=441000)
            e1.record()
            torch.cuda.synchronize()
            cms = e0.elapsed_time(e1) / 3
            out[name]["codec_mf1"] = {"clips": clips, "rows": clips * gg["rows_per_clip"], "ms": cms,
                                      "audio_seconds_per_s": clips * 10.0 / (cms * 1e-3), "finite": bool(torch.isfinite(yy).all())}
            del cm, cp, codec, a, yy
        except Exception as ex:  # noqa: BLE001 -- an extra line must not cost the headline
            out[name]["codec_mf1"] = {"error": f"{type(ex).__name__}: {ex}"[:200]}
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="mfac", choices=["mfac", "reference"])
    ap.add_argument("--batch", type=int, default=37888, help="per-GPU batch (2 x 148 SMs x 128-row tiles)")
    ap.add_argument("--noise-dimension", type=int, default=784, help="raw samples per example (T)")
    ap.add_argument("--sweep", type=int, nargs="*", default=[128, 1024, 4096, 16384, 18944],
                    help="other per-GPU batches (SURVEY.md section 8d: 128 = the config's own, 1024, 4096, 16384; 18944 = one wave of tile pairs)")
    ap.add_argument("--quick", action="store_true", help="skip sweep / codec / cpu baseline")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "mfac":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_mfac(args)


if __name__ == "__main__":
    sys.exit(main())
