"""Training fall-through of the drop-in models: ``compute_loss`` and ``decoder(..., infer=False)``.

The CUDA path is inference only.  So that ``main.py train`` keeps working under the drop-in ``model`` package, the training entry
points DELEGATE to the reference's own ``model`` package -- supplied by the caller, never vendored here -- instantiated from the same
``cfg`` and running on the drop-in's OWN ``nn.Parameter`` / buffer objects (the two module trees have identical names, see
tests/test_tts_module_cpu.py), so gradients, optimiser state, EMA copies and ``state_dict`` all live on the drop-in:

    DeXTTS.compute_loss   -> <ref>/model/tts.py:76-153      GeDEXTTS.compute_loss -> GeDEX-TTS/model/tts.py:58-121
    Diffusion.forward(infer=False) -> <ref>/model/diffusion.py:252-254 (EDMLoss, edm.py:22-68)

Where the reference lives: ``set_reference_dir(path)``, else ``$DEXB_REFERENCE_DIR``, else the current directory if it is a checkout
(``./model/tts.py`` exists -- the situation when the reference's own ``main.py`` runs with the drop-in first on ``sys.path``).
Its one native dependency, the Cython Monotonic Alignment Search, is replaced by the CUDA kernel of this package
(``dexb200.model.monotonic_align.maximum_path``, csrc/mas.cu).  Everything else of the training path is the reference's plain PyTorch.
"""
import importlib
import os
import sys
import types

_ref_dir = None
_modules = {}


def set_reference_dir(path):
    """Directory of the reference checkout to train with (the one that contains ``model/tts.py``), e.g. ``.../DEX-TTS``."""
    global _ref_dir
    _ref_dir = os.path.abspath(path) if path else None
    _modules.clear()


def reference_dir():
    for cand in (_ref_dir, os.environ.get("DEXB_REFERENCE_DIR"), os.getcwd()):
        if cand and os.path.isfile(os.path.join(cand, "model", "tts.py")) and os.path.isfile(os.path.join(cand, "model", "edm.py")):
            return os.path.abspath(cand)
    raise RuntimeError("the training path (compute_loss / infer=False) delegates to the reference's own `model` package: point "
                       "dexb200.model.reference_twin.set_reference_dir() or $DEXB_REFERENCE_DIR at the DEX-TTS / GeDEX-TTS checkout "
                       "(the directory that contains model/tts.py).  The CUDA path itself is inference only.")


def _import_reference(name):
    """Import ``model.<name>`` of the reference checkout as the reference wrote it (its files use absolute ``model.*`` imports) while
    another package called ``model`` -- normally this drop-in -- stays the one the rest of the process sees."""
    root = reference_dir()
    key = (root, name)
    if key in _modules:
        return _modules[key]
    from . import monotonic_align as mas
    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    for k in saved:
        del sys.modules[k]
    pkg = types.ModuleType("model")
    pkg.__path__ = [os.path.join(root, "model")]
    pkg.monotonic_align = mas
    sys.modules["model"] = pkg
    sys.modules["model.monotonic_align"] = mas          # tts.py:7 `from model import monotonic_align` -> the CUDA kernel
    try:
        mod = importlib.import_module("model." + name)
    finally:
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _modules[key] = mod
    return mod


class _Cfg(dict):
    """Attribute + item access, as the reference's constructors use both (``cfg.n_spks = 0``, ``**cfg.tv_encoder``)."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__

    @staticmethod
    def wrap(c):
        if isinstance(c, dict):
            return _Cfg({k: _Cfg.wrap(v) for k, v in c.items()})
        return c


def _tie(own, twin):
    """Make every parameter / buffer of ``twin`` the SAME tensor object as the equally named one of ``own``."""
    for kind in ("_parameters", "_buffers"):
        named = own.named_parameters(remove_duplicate=False) if kind == "_parameters" else own.named_buffers(remove_duplicate=False)
        src = dict(named)
        tw = twin.named_parameters(remove_duplicate=False) if kind == "_parameters" else twin.named_buffers(remove_duplicate=False)
        names = [n for n, _ in tw]
        missing = [n for n in names if n not in src]
        if missing:
            raise RuntimeError(f"reference module has {kind[1:]} the drop-in does not: {missing[:5]} ...")
        for n in names:
            mod = twin
            *path, leaf = n.split(".")
            for p in path:
                mod = getattr(mod, p)
            getattr(mod, kind)[leaf] = src[n]


def tts_twin(own, cfg, variant):
    """The reference ``DeXTTS`` / ``GeDEXTTS`` built from ``cfg`` and tied to the drop-in model ``own``."""
    mod = _import_reference("tts")
    cls = mod.DeXTTS if variant == "dex" else mod.GeDEXTTS
    twin = cls(_Cfg.wrap(cfg))
    conf = getattr(getattr(getattr(twin, "encoder", None), "encoder", None), "config", None)
    if conf is not None and not hasattr(conf, "use_cache"):
        conf.use_cache = True                 # transformers >= 5 dropped the attribute retnet.py:79 reads
    _tie(own, twin)
    return twin


def diffusion_twin(own, ctor_kwargs):
    """The reference ``Diffusion`` (decoder alone) tied to the drop-in decoder ``own``."""
    mod = _import_reference("diffusion")
    kw = dict(ctor_kwargs)
    kw["dit_cfg"] = _Cfg.wrap(kw["dit_cfg"])
    twin = mod.Diffusion(**kw)
    _tie(own, twin)
    return twin
