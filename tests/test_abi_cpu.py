"""CPU-side checks: the C-ABI library loads, exports every symbol include/dexb200.h declares (and nothing is bound that
the header does not declare), and the host-side module mirrors the reference decoder's parameter tree."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dexb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(dexb_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from dexb200 import lib
    assert os.path.exists(lib.LIB_PATH), "build the extension first: python __graft_entry__.py"
    names = header_symbols()
    assert len(names) >= 14
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), n
    assert names == set(lib.SYMBOLS), names ^ set(lib.SYMBOLS)
    lib.load()


def test_no_compute_without_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dexb200.engine import ReverseDiffusion
    from dexb200.manifest import DecoderCfg
    with pytest.raises(RuntimeError):
        ReverseDiffusion(DecoderCfg.make("dex"))


def test_sigma_schedule_matches_oracle():
    import dex_oracle as O
    from dexb200.engine import edm_sigmas
    for n in (2, 3, 10, 50, 100):
        assert torch.equal(edm_sigmas(n), O.sigma_schedule(n))


@pytest.mark.parametrize("variant", ["dex", "gedex"])
def test_module_state_dict_matches_manifest(variant):
    from dexb200.manifest import DecoderCfg, decoder_manifest
    from dexb200.model import Diffusion, GeDiffusion
    dit = dict(patch_size=3 if variant == "dex" else 7, stride_size=2 if variant == "dex" else 4, hidden_size=256, depth=4,
               num_heads=2, mlp_ratio=2, conv_pos=16, conv_pos_groups=8)
    cls = Diffusion if variant == "dex" else GeDiffusion
    m = cls(n_feats=80, dim=64, dit_cfg=dit, dim_mults=[1, 2], model_type="dit", n_spks=0 if variant == "dex" else 1)
    sd = m.state_dict()
    man = decoder_manifest(DecoderCfg.make(variant))
    assert len(sd) == 2 * len(man)
    for e in man:
        assert tuple(sd["denoise_fn." + e.name].shape) == tuple(e.shape)
        assert sd["precond_model.model." + e.name].data_ptr() == sd["denoise_fn." + e.name].data_ptr()
    # zero-initialised tensors of the reference (adaLN-Zero, Rezero gates) are zero here too
    assert float(sd["denoise_fn.vit.blocks.0.adaLN_modulation.1.weight"].abs().max()) == 0.0
    assert float(sd["denoise_fn.downs.0.2.fn.g"]) == 0.0
    # the training branch delegates to the reference's own modules (reference_twin.py); without a checkout it says what to do
    import dexb200.model.reference_twin as RT
    RT.set_reference_dir(None)
    os.environ.pop("DEXB_REFERENCE_DIR", None)
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.abspath(__file__)))
    try:
        with pytest.raises(RuntimeError, match="DEXB_REFERENCE_DIR"):
            if variant == "dex":
                m(None, None, None, None, None, None, None, infer=False)
            else:
                m(None, None, None, infer=False)
    finally:
        os.chdir(cwd)


def test_align_entry_points_validate_arguments_before_touching_the_gpu():
    """dexb_align_*: null pointers and bad shapes come back as -1 with a message in dexb_last_error (no CUDA call is made first,
    so this runs without a device)."""
    import ctypes
    from dexb200 import lib
    L = lib.load()
    buf = (ctypes.c_float * 16)()
    ibuf = (ctypes.c_int64 * 4)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    ip = ctypes.cast(ibuf, ctypes.c_void_p)
    assert L.dexb_align_lengths(None, p, 1, 4, 1.0, p, ip, ip, None) == -1
    assert b"null" in L.dexb_last_error()
    assert L.dexb_align_lengths(p, p, 0, 4, 1.0, p, ip, ip, None) == -1
    assert L.dexb_align_lengths(p, p, 1, 4, 0.0, p, ip, ip, None) == -1          # length_scale must be positive
    assert b"length_scale" in L.dexb_last_error()
    assert L.dexb_align_expand(p, p, ip, None, 1, 4, 80, 8, None, p, p, None) == -1
    assert L.dexb_align_expand(p, p, ip, p, 1, 4, 80, 0, None, p, p, None) == -1
    assert b"Ty = 0" in L.dexb_last_error()


def test_text_encoder_handle_fails_loudly_without_gpu():
    """dexb_text_create touches the device first: on a box without one it returns an error code and a message, never a handle."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes
    from dexb200 import lib
    L = lib.load()
    h = ctypes.c_void_p()
    rc = L.dexb_text_create(149, 80, 192, 1024, 256, 2, 8, 3, 1, 0, ctypes.byref(h))
    assert rc != 0 and not h.value and len(L.dexb_last_error()) > 0
    from dexb200.model import TextEncoder
    with pytest.raises(RuntimeError):
        TextEncoder(n_vocab=149, n_feats=80, n_channels=192, filter_channels=1024, filter_channels_dp=256, n_heads=2, n_layers=8,
                    kernel_size=3, p_dropout=0.1, use_softmax=True, use_decay=False).cuda_engine()


def test_mas_entry_point_validates_arguments_before_touching_the_gpu():
    import ctypes
    from dexb200 import lib
    L = lib.load()
    buf = (ctypes.c_float * 16)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert L.dexb_mas_maximum_path(None, p, 1, 2, 2, p, p, None) == -1 and b"null" in L.dexb_last_error()
    assert L.dexb_mas_maximum_path(p, p, 1, 5000, 2, p, p, None) == -1 and b"Tx = 5000" in L.dexb_last_error()
