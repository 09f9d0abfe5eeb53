"""Parity of the CUDA reverse-diffusion path (through the C ABI) against the committed golden vectors of the unmodified
reference (tests/golden/*.npz) and against the CPU oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): 1e-3 relative fp32 per mel bin -- metric in tests/parity.py."""
import glob
import os

import numpy as np
import pytest
import torch

import dex_oracle as O
from dexb200.manifest import DecoderCfg
from dexb200.synth import synth_decoder_weights, synth_inputs
from parity import REL_TOL, per_bin_violation, tensor_rel_err

pytestmark = pytest.mark.gpu

# decoder fixtures only (stft_ / tiv_ / tv_ / lf0_ fixtures belong to the front-end and pre-loop tests)
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if os.path.basename(p).startswith(("dex_", "gedex_")))
IDS = [os.path.basename(p)[:-4] for p in GOLD]
_engines = {}


def get_engine(variant, live, gemm_engine, n_spks=None):
    from dexb200.engine import ReverseDiffusion
    if n_spks is None:
        n_spks = DecoderCfg.make(variant).n_spks
    key = (variant, live, gemm_engine, n_spks)
    if key not in _engines:
        cfg = DecoderCfg.make(variant, n_spks=n_spks)
        eng = ReverseDiffusion(cfg, gemm_engine=gemm_engine)
        eng.load_state_dict(synth_decoder_weights(cfg, seed=100, live=live))
        _engines[key] = eng
    return _engines[key]


def load_case(path):
    g = np.load(path)
    meta = [int(v) for v in g["meta"]]
    B, T, Ts, steps, ragged, live, seed = meta[:7]
    variant = str(g["variant"])
    cfg = DecoderCfg.make(variant, n_spks=meta[7] if len(meta) > 7 else None)     # 8th entry: multi-speaker GeDEX-TTS
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed, ragged=bool(ragged))
    cond = None
    if variant == "dex":
        cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"])
    elif cfg.n_spks > 1:
        cond = dict(spk=inp["spk"])
    return g, cfg, inp, cond, steps, bool(live)


def to_cuda(cond):
    if cond is None:
        return None
    if "spk" in cond:
        return dict(spk=cond["spk"].cuda())
    return dict(sty=cond["sty"].cuda(), sty_lengths=cond["sty_lengths"].cuda(), ref_skips=[r.cuda() for r in cond["ref_skips"]])


@pytest.mark.parametrize("gemm_engine", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_first_network_call_matches_oracle(path, gemm_engine):
    """D(x_0; sigma_0) of the first sampler step against the CPU oracle (one full denoiser evaluation)."""
    g, cfg, inp, cond, steps, live = load_case(path)
    eng = get_engine(cfg.variant, live, gemm_engine, cfg.n_spks)
    w = synth_decoder_weights(cfg, seed=100, live=live)
    ts = O.sigma_schedule(steps)
    x0 = (inp["z"] / float(g["temperature"]) + inp["mu"]) * ts[0]
    with torch.no_grad():
        ref = O.edm_precond(w, O.make_cfg(cfg.variant, n_spks=cfg.n_spks), x0, ts[0], inp["mask"], inp["mu"], cond=cond)
    out = eng.denoise_once(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps, 0, cond=to_cuda(cond)).cpu()
    assert torch.isfinite(out).all()
    assert tensor_rel_err(out, ref) < 2e-4
    # the network output F_x itself (D = c_skip x + c_out F): sigma_0 = 80 makes c_skip tiny, so this is the strict check
    assert per_bin_violation(out, ref) < REL_TOL


@pytest.mark.parametrize("gemm_engine", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("path", GOLD, ids=IDS)
def test_trajectory_matches_reference_golden(path, gemm_engine):
    """Full reverse diffusion against the output of the unmodified reference (golden fixture)."""
    g, cfg, inp, cond, steps, live = load_case(path)
    eng = get_engine(cfg.variant, live, gemm_engine, cfg.n_spks)
    x0 = inp["z"] / float(g["temperature"]) + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps, cond=to_cuda(cond)).cpu()
    y_ref = torch.from_numpy(g["y"])
    assert y.shape == y_ref.shape and torch.isfinite(y).all()
    assert per_bin_violation(y, y_ref) < REL_TOL
    assert eng.launches > 0


@pytest.mark.parametrize("path", [p for p in GOLD if "b2r" in p], ids=[i for i in IDS if "b2r" in i])
def test_host_entry_point_equals_device_entry_point(path):
    g, cfg, inp, cond, steps, live = load_case(path)
    eng = get_engine(cfg.variant, live, 0, cfg.n_spks)
    x0 = inp["z"] / float(g["temperature"]) + inp["mu"]
    y_dev = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps, cond=to_cuda(cond)).cpu()
    y_host = eng.sample_host(x0, inp["mask"], inp["mu"], steps, cond=cond)
    # the pipeline is deterministic (fixed-order reductions; the only atomics left add fp32 partials in double precision):
    # both entry points run the same graph and must agree bit for bit
    assert torch.equal(y_host, y_dev)
    y_dev2 = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps, cond=to_cuda(cond)).cpu()
    assert torch.equal(y_dev2, y_dev)


def test_tcgen05_takes_every_gemm():
    """The product engine must not silently run GEMMs on CUDA cores for the shipped configurations."""
    for variant in ("dex", "gedex"):
        eng = get_engine(variant, True, 0)
        cfg = DecoderCfg.make(variant)
        inp = synth_inputs(cfg, 1, 64, Ts=20, seed=3)
        cond = None
        if variant == "dex":
            cond = to_cuda(dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"]))
        eng.sample(inp["z"].cuda(), inp["mask"].cuda(), inp["mu"].cuda(), 2, cond=cond)
        assert eng.simt_fallbacks == 0


@pytest.mark.parametrize("variant,T", [("gedex", 200), ("dex", 120)])
def test_fifty_step_trajectory_matches_oracle(variant, T):
    """The headline setting (50 Euler steps, temperature 1.5): errors compound over the trajectory, so this is the case
    that decides whether split-bf16 x3 is precise enough.  CPU oracle on the same seeded inputs."""
    cfg = DecoderCfg.make(variant)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    inp = synth_inputs(cfg, 1, T, Ts=40, seed=77)
    cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"]) if variant == "dex" else None
    steps = 50
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    with torch.no_grad():
        y_ref = O.reverse_diffusion(w, O.make_cfg(variant), inp["z"], inp["mask"], inp["mu"], steps, temperature=1.5, cond=cond)
    eng = get_engine(variant, True, 0)
    x0 = inp["z"] / 1.5 + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps, cond=to_cuda(cond)).cpu()
    v = per_bin_violation(y, y_ref)
    print(f"50-step {variant} T={T}: per-bin violation {v:.3e} (tol {REL_TOL:g})")
    assert v < REL_TOL


def test_style_reference_longer_than_the_fused_key_tile():
    """Reference utterances beyond 511 frames (~5.9 s at hop 256): the TV adaptor leaves the fused cross-attention kernel (one 512-key
    tile) for the GEMM -> softmax -> GEMM route; synthesize.py:96-99 passes the whole reference mel as `sty` with no cap."""
    cfg = DecoderCfg.make("dex")
    w = synth_decoder_weights(cfg, seed=100, live=True)
    inp = synth_inputs(cfg, 2, 64, Ts=700, seed=19, ragged=True)
    cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"])
    with torch.no_grad():
        y_ref = O.reverse_diffusion(w, O.make_cfg("dex"), inp["z"], inp["mask"], inp["mu"], 3, temperature=1.5, cond=cond)
    eng = get_engine("dex", True, 0)
    x0 = inp["z"] / 1.5 + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), 3, cond=to_cuda(cond)).cpu()
    assert per_bin_violation(y, y_ref) < REL_TOL
    assert eng.simt_fallbacks == 0


# ---- LibriTTS decoder sizes (DEX-TTS/config/LibriTTS/base.yaml:65,77: decoder.dim 128, dit.hidden_size 384 -> head dim 192) -------
# Different code runs there: conv_in / final kernels at 128 channels, 256-channel GroupNorm (4 chunks per group, two n-tiles: per-tile
# statistics), the CUDA-core LinearAttention context at C = 256, the GEMM -> softmax -> GEMM routes of the TV adaptor (256 channels)
# and of the DiT attention (head dim 192), and the positional convolution with 48 channels per group (four taps per K chunk).
LIBRI = dict(dim=128, hidden=384)


def _libri_engine(gemm_engine=0):
    from dexb200.engine import ReverseDiffusion
    cfg = DecoderCfg.make("dex", **LIBRI)
    eng = ReverseDiffusion(cfg, gemm_engine=gemm_engine)
    eng.load_state_dict(synth_decoder_weights(cfg, seed=100, live=True))
    return cfg, eng


def test_libritts_sizes_match_reference_fixture():
    path = os.path.join(os.path.dirname(__file__), "golden", "libri_dex_b1.npz")
    g = np.load(path)
    B, T, Ts, steps, ragged, live, seed = [int(v) for v in g["meta"][:7]]
    assert [int(v) for v in g["dims"]] == [128, 384]
    cfg, eng = _libri_engine()
    inp = synth_inputs(cfg, B, T, Ts=Ts, seed=seed, ragged=bool(ragged))
    cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"])
    x0 = inp["z"] / float(g["temperature"]) + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps, cond=to_cuda(cond)).cpu()
    v = per_bin_violation(y, torch.from_numpy(g["y"]))
    print(f"libri_dex_b1 (dim 128 / hidden 384): per-bin violation {v:.2e} vs the unmodified reference, launches {eng.launches}")
    assert v < REL_TOL
    eng.close()


@pytest.mark.parametrize("B,T,Ts", [(2, 128, 259), (1, 512, 100)])
def test_libritts_sizes_match_oracle(B, T, Ts):
    cfg, eng = _libri_engine()
    inp = synth_inputs(cfg, B, T, Ts=Ts, seed=900 + T, ragged=B > 1)
    cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"])
    w = synth_decoder_weights(cfg, seed=100, live=True)
    steps = 3
    with torch.no_grad():
        ref = O.reverse_diffusion(w, O.make_cfg("dex", **LIBRI), inp["z"], inp["mask"], inp["mu"], steps, temperature=1.5, cond=cond)
    x0 = inp["z"] / 1.5 + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), steps, cond=to_cuda(cond)).cpu()
    v = per_bin_violation(y, ref)
    print(f"LibriTTS sizes B={B} T={T} Ts={Ts}: per-bin violation {v:.2e}")
    assert v < REL_TOL
    eng.close()
