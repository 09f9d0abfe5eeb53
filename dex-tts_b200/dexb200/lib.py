"""ctypes binding of libdexb200.so (C ABI: include/dexb200.h).

The library is built in-tree by ``dex-tts_b200/build.sh`` (``__graft_entry__.build()``).  There is no fallback:
if the shared object is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdexb200.so")

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class DexbConfig(ctypes.Structure):
    _fields_ = [("variant", ctypes.c_int), ("dim", ctypes.c_int), ("hidden", ctypes.c_int), ("depth", ctypes.c_int),
                ("heads", ctypes.c_int), ("mlp_hidden", ctypes.c_int), ("patch", ctypes.c_int), ("stride", ctypes.c_int),
                ("conv_pos", ctypes.c_int), ("conv_pos_groups", ctypes.c_int), ("n_feats", ctypes.c_int),
                ("pe_scale", ctypes.c_float), ("gemm_engine", ctypes.c_int), ("nsplit", ctypes.c_int),
                ("n_spks", ctypes.c_int), ("spk_emb_dim", ctypes.c_int)]


class DexbCond(ctypes.Structure):
    _fields_ = [("sty_dev", ctypes.c_void_p), ("sty_len_dev", ctypes.c_void_p), ("ref_skips_dev", ctypes.c_void_p * 6),
                ("Tr", ctypes.c_int), ("spk_dev", ctypes.c_void_p)]


# every symbol include/dexb200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "dexb_last_error": (ctypes.c_char_p, []),
    "dexb_version": (ctypes.c_int, []),
    "dexb_create": (ctypes.c_int, [ctypes.POINTER(DexbConfig), ctypes.POINTER(ctypes.c_void_p)]),
    "dexb_destroy": (None, [ctypes.c_void_p]),
    "dexb_load_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, c_int64_p, ctypes.c_int]),
    "dexb_finalize_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_plan": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_float_p,
                                 ctypes.POINTER(ctypes.c_size_t)]),
    "dexb_reverse_diffusion": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.POINTER(DexbCond), ctypes.c_void_p]),
    "dexb_reverse_diffusion_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                                   ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_denoise_once": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.POINTER(DexbCond), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_gemm_test": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_gemm_bench": (ctypes.c_int, [ctypes.c_int] * 14 + [c_float_p]),
    "dexb_attn_test": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p]),
    "dexb_stft_mel": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_debug_tap": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                      ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    "dexb_profile_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_void_p]),
    "dexb_last_launch_count": (ctypes.c_long, [ctypes.c_void_p]),
    "dexb_simt_fallbacks": (ctypes.c_int, [ctypes.c_void_p]),
    "dexb_tiv_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "dexb_tiv_destroy": (None, [ctypes.c_void_p]),
    "dexb_tiv_load_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, c_int64_p, ctypes.c_int]),
    "dexb_tiv_finalize_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_tiv_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p]),
    "dexb_tiv_last_launch_count": (ctypes.c_long, [ctypes.c_void_p]),
    "dexb_tv_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_float, ctypes.POINTER(ctypes.c_void_p)]),
    "dexb_tv_destroy": (None, [ctypes.c_void_p]),
    "dexb_tv_load_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, c_int64_p, ctypes.c_int]),
    "dexb_tv_finalize_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_tv_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_tv_last_launch_count": (ctypes.c_long, [ctypes.c_void_p]),
    "dexb_lf0_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "dexb_lf0_destroy": (None, [ctypes.c_void_p]),
    "dexb_lf0_load_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, c_int64_p, ctypes.c_int]),
    "dexb_lf0_finalize_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_lf0_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_lf0_last_launch_count": (ctypes.c_long, [ctypes.c_void_p]),
    "dexb_style_fuse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]),
    "dexb_text_create": (ctypes.c_int, [ctypes.c_int] * 10 + [ctypes.POINTER(ctypes.c_void_p)]),
    "dexb_text_destroy": (None, [ctypes.c_void_p]),
    "dexb_text_load_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, c_int64_p, ctypes.c_int]),
    "dexb_text_finalize_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_text_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_text_last_launch_count": (ctypes.c_long, [ctypes.c_void_p]),
    "dexb_text_set_layer_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "dexb_text_copy_stream": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_mas_maximum_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_align_lengths": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_void_p,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_align_expand": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_voc_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_int32_p, ctypes.c_int, c_int32_p, ctypes.c_int, c_int32_p, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_void_p)]),
    "dexb_voc_destroy": (None, [ctypes.c_void_p]),
    "dexb_voc_load_weight": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, c_int64_p, ctypes.c_int]),
    "dexb_voc_finalize_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_voc_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "dexb_voc_last_launch_count": (ctypes.c_long, [ctypes.c_void_p]),
}

_lib = None


def load():
    """Load libdexb200.so and bind every declared symbol.  Raises RuntimeError when the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                           "g.build()' or bash dex-tts_b200/build.sh).  dexb200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)            # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().dexb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")
