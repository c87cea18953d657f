"""NumPy restatement of the reference's MLP velocity network, iMF loss, AdamW and samplers.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: no
JAX in the build container and the reference stores no numbers for this part;
this file is cross-checked against ``oracle/imf_torch.py`` (autograd /
``torch.func.jvp``) and against the reference's two property tests.

Everything here is written as the explicit recurrences the CUDA path
implements, so intermediate tensors can be
compared one by one.  ``dtype`` selects fp64 (exact oracle) or fp32.

Reference lines followed:
  sinusoidal_embedding   <- utils.py:5-13
  gelu (tanh form)       <- models/mlp_flow.py:29  (jax.nn.gelu(approximate=True))
  mlp / encode           <- models/mlp_flow.py:12-55,153-162
  block / forward        <- models/mlp_flow.py:63-117,164-230
  layer norm             <- flax nn.LayerNorm(use_scale=False,use_bias=False): eps 1e-6,
                            var = max(0, E[x^2]-E[x]^2)  (use_fast_variance default)
  sample_tr              <- utils.py:32-45
  schedule               <- trainers/noise_schedules.py:69-88
  weighted_l2_loss       <- utils.py:16-25
  imf loss               <- trainers/loss_strategies.py:227-280
  adamw                  <- optax 0.2.5 adamw defaults, trainers/train.py:236
  heun sample            <- evaluators/sampling.py:42-95
  mean-flow 1/2-NFE      <- documentation/research/improved_meanflow/improved_meanflow_key_eqn.md:311-318
"""
from __future__ import annotations

import math

import numpy as np

LN_EPS = 1e-6
K0 = math.sqrt(2.0 / math.pi)
K1 = 0.044715


# ------------------------------------------------------------------ params
def param_shapes(D: int, L: int, C: int, nb: int) -> list[tuple[str, tuple[int, ...]]]:
    """Flat parameter order = jax tree_flatten order of the Flax tree for nb <= 10
    (dict keys sorted; 'bias' < 'kernel'); blocks are in numeric order."""
    I = L + D
    He = (D + L) // 2
    out = []
    for k in range(nb):
        p = f"blocks_{k}"
        out += [
            (f"{p}/conditioning_layer/dense1/bias", (C,)),
            (f"{p}/conditioning_layer/dense1/kernel", (C, C)),
            (f"{p}/conditioning_layer/dense2/bias", (2 * I + D,)),
            (f"{p}/conditioning_layer/dense2/kernel", (C, 2 * I + D)),
            (f"{p}/mlp/dense1/bias", (I,)),
            (f"{p}/mlp/dense1/kernel", (I, I)),
            (f"{p}/mlp/dense2/bias", (D,)),
            (f"{p}/mlp/dense2/kernel", (I, D)),
        ]
    out += [
        ("encoder/encoder_mlp/dense1/bias", (He,)),
        ("encoder/encoder_mlp/dense1/kernel", (D, He)),
        ("encoder/encoder_mlp/dense2/bias", (L,)),
        ("encoder/encoder_mlp/dense2/kernel", (He, L)),
    ]
    return out


def init_params(D, L, C, nb, seed=42, dtype=np.float32, bias_scale=0.0):
    """lecun_normal kernels (truncated normal, variance 1/fan_in), zero (or small random) biases."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, shape in param_shapes(D, L, C, nb):
        if name.endswith("kernel"):
            std = math.sqrt(1.0 / shape[0]) / 0.87962566103423978
            w = rng.standard_normal(shape)
            bad = np.abs(w) > 2.0
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            params[name] = (w * std).astype(dtype)
        else:
            params[name] = (bias_scale * rng.standard_normal(shape)).astype(dtype)
    return params


def flatten(params, D, L, C, nb) -> np.ndarray:
    return np.concatenate([np.asarray(params[n]).reshape(-1) for n, _ in param_shapes(D, L, C, nb)])


def unflatten(flat, D, L, C, nb):
    out, o = {}, 0
    for n, s in param_shapes(D, L, C, nb):
        sz = int(np.prod(s))
        out[n] = np.asarray(flat[o:o + sz]).reshape(s)
        o += sz
    return out


def to_tree(params):
    """Flat-name dict -> nested Flax-style dict."""
    tree = {}
    for name, v in params.items():
        d = tree
        parts = name.split("/")
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = v
    return tree


def from_tree(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(from_tree(v, prefix + k + "/"))
        else:
            out[prefix + k] = v
    return out


def dims_of(params):
    nb = len({n.split("/")[0] for n in params if n.startswith("blocks_")})
    C = params["blocks_0/conditioning_layer/dense1/kernel"].shape[0]
    I, D = params["blocks_0/mlp/dense2/kernel"].shape
    return D, I - D, C, nb


# ------------------------------------------------------------------ primitives
def gelu(a):
    return 0.5 * a * (1.0 + np.tanh(K0 * (a + K1 * a ** 3)))


def dgelu(a):
    th = np.tanh(K0 * (a + K1 * a ** 3))
    return 0.5 * (1.0 + th) + 0.5 * a * (1.0 - th * th) * K0 * (1.0 + 3.0 * K1 * a * a)


def sinusoidal_embedding(x, dim, max_period=10000.0):
    half = dim // 2
    dt = x.dtype
    freqs = np.exp(-dt.type(math.log(max_period)) * np.arange(half, dtype=dt) / dt.type(half)).astype(dt)
    args = x[:, None] * freqs[None]
    return np.concatenate([np.cos(args), np.sin(args)], axis=-1), freqs


def d_sinusoidal_embedding(x, dim, xdot):
    """d/ds of embedding(x + s*xdot)."""
    _, freqs = sinusoidal_embedding(x, dim)
    args = x[:, None] * freqs[None]
    fx = freqs[None] * xdot[:, None]
    return np.concatenate([-np.sin(args) * fx, np.cos(args) * fx], axis=-1)


def layer_norm(c):
    mu = c.mean(-1, keepdims=True)
    var = np.maximum(0.0, (c * c).mean(-1, keepdims=True) - mu * mu)
    rstd = 1.0 / np.sqrt(var + c.dtype.type(LN_EPS))
    return (c - mu) * rstd, mu, rstd


def dense(p, name, x):
    return x @ p[name + "/kernel"] + p[name + "/bias"]


# ------------------------------------------------------------------ model
def encode(p, x):
    a = dense(p, "encoder/encoder_mlp/dense1", x)
    return dense(p, "encoder/encoder_mlp/dense2", gelu(a))


def forward(p, x, time, latents=None, xdot=None, tdot=None, hdot=None, cache=None):
    """u = f(x, [t,h], latents); optionally the forward-mode tangent along (xdot, tdot, hdot).

    Returns u, or (u, udot) when xdot is given.  ``cache`` (a list) receives the
    per-block saved tensors the backward needs.
    """
    D, L, C, nb = dims_of(p)
    dt = x.dtype
    B = x.shape[0]
    lat = np.zeros((B, L), dtype=dt) if latents is None else latents.astype(dt)
    t, h = time[:, 0], time[:, 1]
    e_t, _ = sinusoidal_embedding(t, C)
    e_h, _ = sinusoidal_embedding(h, C)
    cond = e_t + e_h
    tangent = xdot is not None
    if tangent:
        cdot = d_sinusoidal_embedding(t, C, tdot) + d_sinusoidal_embedding(h, C, hdot)
        xd = xdot.astype(dt)
    I = L + D
    for k in range(nb):
        pre = f"blocks_{k}"
        c = np.concatenate([lat, x], axis=-1)
        n, mu, rstd = layer_norm(c)
        a_c = dense(p, pre + "/conditioning_layer/dense1", cond)
        g_c = gelu(a_c)
        m = dense(p, pre + "/conditioning_layer/dense2", g_c)
        s1, sh, s2 = m[:, :I], m[:, I:2 * I], m[:, 2 * I:]
        hin = (1.0 + s1) * n + sh
        a = dense(p, pre + "/mlp/dense1", hin)
        g = gelu(a)
        o = dense(p, pre + "/mlp/dense2", g)
        x_new = o * (1.0 + s2) / dt.type(nb) + x
        if tangent:
            cd = np.concatenate([np.zeros((B, L), dtype=dt), xd], axis=-1)
            nd = (cd - cd.mean(-1, keepdims=True) - n * (n * cd).mean(-1, keepdims=True)) * rstd
            md = (dgelu(a_c) * (cdot @ p[pre + "/conditioning_layer/dense1/kernel"])) @ p[pre + "/conditioning_layer/dense2/kernel"]
            s1d, shd, s2d = md[:, :I], md[:, I:2 * I], md[:, 2 * I:]
            hind = s1d * n + (1.0 + s1) * nd + shd
            ad = hind @ p[pre + "/mlp/dense1/kernel"]
            od = (dgelu(a) * ad) @ p[pre + "/mlp/dense2/kernel"]
            xd = (od * (1.0 + s2) + o * s2d) / dt.type(nb) + xd
        if cache is not None:
            cache.append(dict(x=x, n=n, rstd=rstd, cond=cond, a_c=a_c, g_c=g_c, s1=s1, s2=s2,
                              hin=hin, a=a, g=g, o=o))
        x = x_new
    return (x, xd) if tangent else x


# ------------------------------------------------------------------ loss pieces
def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def sample_tr_from_normals(nt, nr, mean=-0.4, std=1.0, data_proportion=0.5):
    """utils.py:36-45 with the two N(0,1) draws passed in explicitly."""
    dt = nt.dtype
    t = sigmoid(nt * dt.type(std) + dt.type(mean))
    r = sigmoid(nr * dt.type(std) + dt.type(mean))
    t, r = np.maximum(t, r), np.minimum(t, r)
    B = t.shape[0]
    mask = np.arange(B) < int(B * data_proportion)
    r = np.where(mask, t, r)
    return t.reshape(B, 1), r.reshape(B, 1)


def interpolate(x, e, t, noise_min=0.001, noise_max=0.999):
    dt = x.dtype
    return (1.0 - t) * x + (dt.type(noise_min) + dt.type(noise_max) * t) * e


def target_of(x, e, noise_max=0.999):
    return x.dtype.type(noise_max) * e - x


def weighted_l2(delta, p=1.0, c=1e-3):
    s = (delta * delta).sum(-1)
    w = 1.0 / (s + delta.dtype.type(c)) ** p
    return (w * s).mean(), s, w


def imf_forward(p, x, e, t, r, method="improved_mean_flow", gamma=0.5, c=1e-3, use_weighted_loss=True):
    """Forward part of the three loss strategies for given (e, t, r).  Returns dict of tensors.

    improved_mean_flow  loss_strategies.py:227-275  z, target from LinearNoiseSchedule(0.001, 0.999); tangent seed = the
                        network's own v = f(z, [t, 0]); weighted_l2_loss (p = 1, sum over D)
    mean_flow           loss_strategies.py:144-199  z = (1-t) x + t e, target = e - x; tangent seed = target (no v pass);
                        (t - r) clipped to [0, 1]; adaptive weight 1 / (mean_D delta^2 + c)^(1 - gamma)
    flow_matching       loss_strategies.py:74-111   single time t (h = 0), no JVP; weighted_l2_loss or plain MSE
    """
    dt = x.dtype
    one = np.ones_like(t[:, 0])
    zero = np.zeros_like(t)
    lat = encode(p, x)
    cache = []
    v = None
    if method == "improved_mean_flow":
        z, tgt = interpolate(x, e, t), target_of(x, e)
        v = forward(p, z, np.concatenate([t, zero], -1), lat)
        # d/ds of th = [t, t - r] along (tdot=1, rdot=0) is [1, 1]
        u, dudt = forward(p, z, np.concatenate([t, t - r], -1), lat, xdot=v, tdot=one, hdot=one, cache=cache)
        tmr = t - r
    elif method == "mean_flow":
        z, tgt = (1.0 - t) * x + t * e, e - x
        u, dudt = forward(p, z, np.concatenate([t, t - r], -1), lat, xdot=tgt, tdot=one, hdot=one, cache=cache)
        tmr = np.clip(t - r, 0.0, 1.0)
    elif method == "flow_matching":
        z, tgt = interpolate(x, e, t), target_of(x, e)
        u, dudt = forward(p, z, np.concatenate([t, zero], -1), lat, xdot=np.zeros_like(z), tdot=0 * one, hdot=0 * one, cache=cache)
        tmr = zero
    else:
        raise ValueError(method)
    v_pred = u + tmr * dudt
    delta = v_pred - tgt
    s = (delta * delta).sum(-1)
    D = x.shape[1]
    if method == "mean_flow":
        dsq = s / dt.type(D)
        w = 1.0 / (dsq + dt.type(c)) ** dt.type(1.0 - gamma)
        loss = (w * dsq).mean()
        gscale = w / dt.type(D)                    # d loss / d delta = 2 gscale delta / B
    elif use_weighted_loss:
        w = 1.0 / (s + dt.type(c))
        loss = (w * s).mean()
        gscale = w
    else:
        w = np.ones_like(s)
        loss = (delta * delta).mean()
        gscale = w / dt.type(D)
    return dict(z=z, target=tgt, latents=lat, v=v, u=u, dudt=dudt, v_pred=v_pred, delta=delta,
                loss=loss, per_example=s, weights=w, gscale=gscale, cache=cache)


def imf_loss_and_grads(p, x, e, t, r, method="improved_mean_flow", gamma=0.5, c=1e-3, use_weighted_loss=True):
    """(loss, grads, aux): reverse-mode recurrences of SURVEY.md row L5 (all three loss strategies: the gradient always
    flows through the primal network output and the encoder only).

    Gradient flows through the primal ``u`` and through ``encode`` only: ``v`` enters
    as a JVP tangent (jax.jvp tangents carry no cotangent back to params unless the
    tangent itself depends on params -- it does (v = f(params)), but the product
    (t-r)*dudt is stop-gradiented, loss_strategies.py:270) and ``dudt`` is stop-gradiented.
    """
    D, L, C, nb = dims_of(p)
    aux = imf_forward(p, x, e, t, r, method=method, gamma=gamma, c=c, use_weighted_loss=use_weighted_loss)
    dt = x.dtype
    B = x.shape[0]
    I = L + D
    grads = {}
    g_x = (2.0 * aux["gscale"][:, None] * aux["delta"] / dt.type(B)).astype(dt)
    g_lat = np.zeros((B, L), dtype=dt)
    for k in reversed(range(nb)):
        pre = f"blocks_{k}"
        cch = aux["cache"][k]
        n, rstd, s1, s2, a, g, o, hin = (cch[q] for q in ("n", "rstd", "s1", "s2", "a", "g", "o", "hin"))
        g_o = g_x * (1.0 + s2) / dt.type(nb)
        g_s2 = g_x * o / dt.type(nb)
        grads[pre + "/mlp/dense2/kernel"] = g.T @ g_o
        grads[pre + "/mlp/dense2/bias"] = g_o.sum(0)
        g_a = (g_o @ p[pre + "/mlp/dense2/kernel"].T) * dgelu(a)
        grads[pre + "/mlp/dense1/kernel"] = hin.T @ g_a
        grads[pre + "/mlp/dense1/bias"] = g_a.sum(0)
        g_hin = g_a @ p[pre + "/mlp/dense1/kernel"].T
        g_s1 = g_hin * n
        g_sh = g_hin
        g_n = g_hin * (1.0 + s1)
        g_c = (g_n - g_n.mean(-1, keepdims=True) - n * (g_n * n).mean(-1, keepdims=True)) * rstd
        g_lat = g_lat + g_c[:, :L]
        g_x = g_x + g_c[:, L:]
        g_m = np.concatenate([g_s1, g_sh, g_s2], axis=-1)
        grads[pre + "/conditioning_layer/dense2/kernel"] = cch["g_c"].T @ g_m
        grads[pre + "/conditioning_layer/dense2/bias"] = g_m.sum(0)
        g_ac = (g_m @ p[pre + "/conditioning_layer/dense2/kernel"].T) * dgelu(cch["a_c"])
        grads[pre + "/conditioning_layer/dense1/kernel"] = cch["cond"].T @ g_ac
        grads[pre + "/conditioning_layer/dense1/bias"] = g_ac.sum(0)
    a_e = dense(p, "encoder/encoder_mlp/dense1", x)
    grads["encoder/encoder_mlp/dense2/kernel"] = gelu(a_e).T @ g_lat
    grads["encoder/encoder_mlp/dense2/bias"] = g_lat.sum(0)
    g_ae = (g_lat @ p["encoder/encoder_mlp/dense2/kernel"].T) * dgelu(a_e)
    grads["encoder/encoder_mlp/dense1/kernel"] = x.T @ g_ae
    grads["encoder/encoder_mlp/dense1/bias"] = g_ae.sum(0)
    aux["g_latents"] = g_lat
    return aux["loss"], grads, aux


# ------------------------------------------------------------------ optimiser
def adamw_step(params, grads, mu, nu, count, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    """optax.adamw: scale_by_adam -> add_decayed_weights (all params, no mask) -> -lr.
    ``count`` is the number of steps already taken (optax increments before bias correction)."""
    c = count + 1
    bc1 = 1.0 - b1 ** c
    bc2 = 1.0 - b2 ** c
    new_p, new_mu, new_nu = {}, {}, {}
    for k in params:
        g = grads[k]
        dt = params[k].dtype
        m = dt.type(b1) * mu[k] + dt.type(1 - b1) * g
        v = dt.type(b2) * nu[k] + dt.type(1 - b2) * g * g
        upd = (m / dt.type(bc1)) / (np.sqrt(v / dt.type(bc2)) + dt.type(eps)) + dt.type(wd) * params[k]
        new_p[k] = params[k] - dt.type(lr) * upd
        new_mu[k], new_nu[k] = m, v
    return new_p, new_mu, new_nu


# ------------------------------------------------------------------ samplers
def heun_sample(p, latents, noise, n_steps, guidance_scale=1.0):
    """evaluators/sampling.py:42-95 with the initial noise passed in.  Reproduces the two quirks:
    grid linspace(1,0,n) (spacing 1/(n-1)) with step dt=1/n, and k2 evaluated at t-dt (t<0 at the end)."""
    dt = noise.dtype
    x = noise.copy()
    B = x.shape[0]
    step = dt.type(1.0 / float(n_steps))
    ts = np.linspace(1.0, 0.0, n_steps, dtype=dt)

    def f(xx, tt):
        tp = np.stack([np.full(B, tt, dtype=dt), np.zeros(B, dtype=dt)], -1)
        if guidance_scale == 1.0:
            return forward(p, xx, tp, latents)
        gs = dt.type(guidance_scale)
        return gs * forward(p, xx, tp, latents) + (1.0 - gs) * forward(p, xx, tp, None)

    for tt in ts:
        k1 = f(x, tt)
        k2 = f(x - step * k1, tt - step)
        x = x - (step / 2.0) * (k1 + k2)
    return x


def mf_sample(p, latents, noise, nfe=1):
    """Mean-flow few-step rule x_r = x_t - (t - r) u(x_t, [t, t-r]) on the uniform grid 1 -> 0."""
    dt = noise.dtype
    x = noise.copy()
    B = x.shape[0]
    grid = np.linspace(1.0, 0.0, nfe + 1, dtype=dt)
    for i in range(nfe):
        t, r = grid[i], grid[i + 1]
        tp = np.stack([np.full(B, t, dtype=dt), np.full(B, t - r, dtype=dt)], -1)
        x = x - (t - r) * forward(p, x, tp, latents)
    return x
