"""CPU tests: the C-ABI library loads, exports every symbol include/mfac.h declares, and its host-only
entry points and the Python mirrors' argument checking work without a GPU (no compute calls here)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import meanflow_audio_codec_b200 as m
from meanflow_audio_codec_b200 import _lib
from oracle import imf_np, mdct_np

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "mfac.h").read_text()
    return sorted(set(re.findall(r"MFAC_API[^;(]*?\b(mfac_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mfac.h but not exported by libmfac.so"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes and header drifted apart"


def test_version_and_status_strings(lib):
    assert lib.mfac_version() >= 100
    assert lib.mfac_status_string(0) == b"success"
    assert b"workspace" in lib.mfac_status_string(-3)
    assert lib.mfac_status_string(-1002) != b"unknown status"  # CUDA error passthrough


@pytest.mark.parametrize("T,N,hop", [(784, 512, 256), (100, 512, 256), (441000, 512, 256), (1024, 256, 128), (3000, 576, 288)])
def test_frame_arithmetic_matches_reference_rule(lib, T, N, hop):
    nf = lib.mfac_mdct_num_frames(T, N, hop)
    assert nf == mdct_np.num_frames(T, N, hop) == m.mdct.__globals__["num_frames"](T, N, hop)
    assert lib.mfac_imdct_length(nf, N, hop) == mdct_np.padded_length(nf, N, hop)
    assert lib.mfac_mdct_num_frames(0, N, hop) < 0 and lib.mfac_imdct_length(1, 0, hop) < 0


@pytest.mark.parametrize("D,L,Cd,nb", [(1024, 256, 128, 8), (8, 64, 32, 2), (6, 64, 32, 2), (3584, 256, 128, 8)])
def test_param_layout_matches_flax_tree_order(lib, D, L, Cd, nb):
    dims = _lib.MlpDims(D, L, Cd, nb)
    shapes = imf_np.param_shapes(D, L, Cd, nb)
    total = sum(int(np.prod(s)) for _, s in shapes)
    assert lib.mfac_mlp_param_count(C.byref(dims)) == total
    model = m.ConditionalFlow(D, Cd, nb, L)
    assert model.param_count() == total
    off = 0
    sl = model.leaf_slices()
    for i, (name, shp) in enumerate(shapes):
        path = tuple(name.split("/"))
        assert sl[path] == (off, shp)
        block = -1 if path[0] == "encoder" else int(path[0].split("_")[1])
        which = i % 8 if block >= 0 else i - 8 * nb
        o, r, c = C.c_int64(), C.c_int64(), C.c_int64()
        assert lib.mfac_mlp_param_offset(C.byref(dims), block, which, C.byref(o), C.byref(r), C.byref(c)) == 0
        assert o.value == off and r.value * c.value == int(np.prod(shp))
        off += int(np.prod(shp))
    assert lib.mfac_workspace_bytes(_lib.WS_LOSS_GRAD, C.byref(dims), 128) > lib.mfac_workspace_bytes(_lib.WS_FORWARD, C.byref(dims), 128) > 0
    assert lib.mfac_mlp_shadow_bytes(C.byref(dims)) >= 2 * (total - (sum(int(np.prod(s)) for n, s in shapes if n.endswith("bias"))))


def test_encoder_only_layout(lib):
    """num_blocks = 0: the MLP encoder on its own (what the mixer / ConvNeXt flows use for ``method="encode"``)."""
    D, L = 1024, 256
    dims = _lib.MlpDims(D, L, 2, 0)
    He = (D + L) // 2
    assert lib.mfac_mlp_param_count(C.byref(dims)) == He + D * He + L + He * L
    assert lib.mfac_workspace_bytes(_lib.WS_FORWARD, C.byref(dims), 16) > 0
    assert lib.mfac_workspace_bytes(_lib.WS_LOSS_GRAD, C.byref(dims), 16) == 0      # no velocity network to train
    assert lib.mfac_workspace_bytes(_lib.WS_SAMPLE, C.byref(dims), 16) == 0
    enc = m.ConditionalFlow(D, 2, 0, L)
    assert list(enc.leaf_slices())[0][0] == "encoder" and enc.param_count() == He + D * He + L + He * L


def test_bad_dims_are_rejected(lib):
    bad = _lib.MlpDims(8, 64, 31, 2)  # odd condition dimension
    assert lib.mfac_mlp_param_count(C.byref(bad)) < 0
    assert lib.mfac_workspace_bytes(0, C.byref(bad), 4) == 0
    with pytest.raises(ValueError):
        m.ConditionalFlow(8, 31, 2, 64)


def test_python_mirror_raises_like_the_reference_before_any_compute():
    with pytest.raises(TypeError):          # mdct.py:189-190
        m.mdct(np.zeros(100), 512)
    with pytest.raises(ValueError):         # mdct.py:191-192
        m.mdct(torch.zeros(()), 512)
    with pytest.raises(ValueError):         # mdct.py:460-461
        m.mdct(torch.zeros(100), 0)
    with pytest.raises(ValueError):         # mdct.py:462-463
        m.mdct(torch.zeros(100), 512, -3)
    with pytest.raises(ValueError):         # mdct.py:249-250
        m.imdct(torch.zeros(512), 512)
    with pytest.raises(TypeError):
        m.imdct([[0.0] * 512], 512)
    with pytest.raises(ValueError):
        m.MDCTConfig(window_size=-1)
    cfg = m.MDCTConfig(window_size=512)
    assert cfg.hop_size == 256              # mdct.py:77-78
    tok = m.MDCTTokenization(config=cfg)
    with pytest.raises(ValueError):         # tokenization.py:94
        tok.tokenize(torch.zeros(2, 3, 4, 5))
    with pytest.raises(ValueError):         # tokenization.py:105-106
        tok.detokenize(torch.zeros(2, 512))
    assert m.compute_token_shape(tok, 784, "mnist") == (2, 512)
    assert m.compute_tokenized_dimension(tok, 784, "mnist") == 1024
    assert m.compute_token_shape(tok, 441000, "audio") == (1721, 512)
    with pytest.raises(ValueError):
        m.compute_token_shape(tok, 784, "imagenet")


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors (no oracle / torch fallback)."""
    with pytest.raises(m.MfacError):
        m.mdct(torch.zeros(2, 1000), 512)
    with pytest.raises(m.MfacError):
        m.imdct(torch.zeros(2, 3, 512), 512)
    src = "".join(p.read_text() for p in (ROOT / "meanflow_audio_codec_b200").glob("*.py"))
    assert "import oracle" not in src and "from oracle" not in src


def test_tokenization_factory_from_reference_config():
    class Cfg:
        tokenization_strategy = "mdct"
        tokenization_config = {"window_size": 512, "hop_size": 256}
    tok = m.create_tokenization_strategy(Cfg)
    assert (tok.config.window_size, tok.config.hop_size) == (512, 256)
    Cfg.tokenization_strategy = None
    assert m.create_tokenization_strategy(Cfg) is None
    Cfg.tokenization_strategy = "wavelet"
    with pytest.raises(ValueError):
        m.create_tokenization_strategy(Cfg)


def test_imf_config_struct_matches_header():
    """The ctypes mirror of MfacImfConfig has the header's fields in the header's order (and C's layout)."""
    text = (ROOT / "include" / "mfac.h").read_text()
    body = re.search(r"typedef struct MfacImfConfig \{(.*?)\} MfacImfConfig;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [re.search(r"(\w+)\s*$", v.strip()).group(1) for d in body.split(";") if d.strip() for v in d.split(",")]
    assert names == [f[0] for f in _lib.ImfConfig._fields_]
    assert C.sizeof(_lib.ImfConfig) == 6 * 4 + 4 + 4 + 3 * 8 + 8 + 4 + 4 + 4 + 4 + 8  # floats, flag, pad, u64s, ptr, 3 x 4, pad, i64
    assert (_lib.LOSS_IMPROVED_MEAN_FLOW, _lib.LOSS_MEAN_FLOW, _lib.LOSS_FLOW_MATCHING) == (0, 1, 2)
    for name, val in (("MFAC_LOSS_IMPROVED_MEAN_FLOW", 0), ("MFAC_LOSS_MEAN_FLOW", 1), ("MFAC_LOSS_FLOW_MATCHING", 2)):
        assert re.search(rf"{name} = {val}\b", text)


def test_create_loss_strategy_follows_the_reference_factory():
    """trainers/train.py:52-153: names, defaults, fall-backs and errors."""
    cfg = lambda **kw: type("Cfg", (), kw)()  # noqa: E731
    s = m.create_loss_strategy(cfg(use_improved_mean_flow=True))
    assert isinstance(s, m.ImprovedMeanFlowLoss) and isinstance(s.time_sampling, m.MeanFlowTimeSampling)
    assert (s.noise_schedule.noise_min, s.noise_schedule.noise_max) == (0.001, 0.999)
    s = m.create_loss_strategy(cfg(use_improved_mean_flow=False))
    assert isinstance(s, m.FlowMatchingLoss) and type(s.time_sampling) is m.LogitNormalTimeSampling
    s = m.create_loss_strategy(cfg(loss_strategy="flow_matching", noise_schedule="uniform", time_sampling="uniform",
                                   use_weighted_loss=False))
    assert isinstance(s.noise_schedule, m.UniformNoiseSchedule) and isinstance(s.time_sampling, m.UniformTimeSampling)
    c = s._config(1, 2, 3)
    assert (c.noise_min, c.noise_max, c.uniform_time, c.use_weighted_loss, c.method) == (0.0, 1.0, 1, 0, _lib.LOSS_FLOW_MATCHING)
    s = m.create_loss_strategy(cfg(loss_strategy="mean_flow", gamma=0.25, c=1e-2, time_sampling_data_proportion=0.75,
                                   time_sampling="mean_flow"))
    assert isinstance(s, m.MeanFlowLoss) and (s.gamma, s.c, s.time_sampling.data_proportion) == (0.25, 1e-2, 0.75)
    c = s._config(1, 2, 3)
    assert (c.noise_min, c.noise_max, c.method) == (0.0, 1.0, _lib.LOSS_MEAN_FLOW)   # loss_strategies.py:161-166
    assert abs(c.gamma - 0.25) < 1e-7 and abs(c.loss_c - 1e-2) < 1e-9
    for bad in (dict(loss_strategy="rectified"), dict(noise_schedule="cosine"), dict(time_sampling="beta")):
        with pytest.raises(ValueError):
            m.create_loss_strategy(cfg(**bad))
    t = m.LogitNormalTimeSampling().sample_time(0, 1000, device="cpu")
    assert t.shape == (1000, 1) and 0.38 < float(t.mean()) < 0.46
    u = m.UniformTimeSampling().sample_time(0, 1000, device="cpu")
    assert u.shape == (1000, 1) and 0.45 < float(u.mean()) < 0.55


def test_reference_config_file_drives_the_factories():
    """BASELINE.json configs[0] in the reference's own file format (tests/golden/config_imf_mlp_mnist_mdct.json, copied by
    tests/golden/make_golden.py) through the three factories the reference's train_flow uses (trainers/train.py:178-264)."""
    from meanflow_audio_codec_b200.config import load_config
    cfg = load_config(ROOT / "tests" / "golden" / "config_imf_mlp_mnist_mdct.json")
    assert cfg.batch_size == 128 and cfg.loss_strategy is None and cfg.time_sampling is None
    tok = m.create_tokenization_strategy(cfg)
    assert (tok.config.window_size, tok.config.hop_size) == (512, 256)
    D = m.compute_tokenized_dimension(tok, cfg.noise_dimension, cfg.dataset)
    assert D == 1024 and m.compute_token_shape(tok, cfg.noise_dimension, cfg.dataset) == (2, 512)
    model = m.create_flow_model(load_config({**cfg.to_dict(), "noise_dimension": D}))
    assert isinstance(model, m.ConditionalFlow) and model.param_count() == 28_262_272   # 28.26 M
    strat = m.create_loss_strategy(cfg)
    assert isinstance(strat, m.ImprovedMeanFlowLoss) and strat.use_weighted_loss and strat.c == 1e-3
    assert (strat.time_sampling.mean, strat.time_sampling.std, strat.time_sampling.data_proportion) == (-0.4, 1.0, 0.5)
    opt = m.adamw(cfg.base_lr, cfg.weight_decay)
    assert (opt.learning_rate, opt.weight_decay, opt.b1, opt.b2, opt.eps) == (1e-4, 1e-4, 0.9, 0.999, 1e-8)


def test_jax_ffi_handlers_type_check_against_the_c_abi():
    """The XLA FFI unit cannot be built here (no jaxlib headers), but it can be compiled with -fsyntax-only against a minimal
    stand-in for xla/ffi/api/ffi.h: every handler body -- above all its call into include/mfac.h -- is type-checked, and each
    handler's parameter count must match its Ctx/Arg/Ret/Attr chain.  Also: every registered target has a handler symbol."""
    import re
    import shutil
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    cc = root / "meanflow_audio_codec_b200" / "jax_ffi" / "mfac_jax_ffi.cc"
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the build image"
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", f"-I{root / 'tests' / 'stubs'}",
                        f"-I{root / 'tests' / 'stubs' / 'cuda'}", f"-I{root / 'include'}", str(cc)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    symbols = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", cc.read_text()))
    init = (root / "meanflow_audio_codec_b200" / "jax_ffi" / "__init__.py").read_text()
    registered = set(re.findall(r'\("mfac_\w+", "(\w+)"\)', init))
    assert registered == symbols and len(symbols) >= 8, (registered, symbols)
