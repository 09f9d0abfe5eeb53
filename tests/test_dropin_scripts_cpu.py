"""The reference's entry scripts import unchanged under the drop-in packages (BASELINE.json north_star: "drops in under synthesize.py
and main.py unchanged").

A fresh interpreter gets ``dex-tts_b200/dropin`` (packages literally named ``model``, ``audio`` and ``hifigan``) and ``dex-tts_b200`` in front of a
COPY of the reference checkout on ``sys.path``, plus empty stand-ins for the third-party modules this image does not have (matplotlib,
neptune, soundfile, ... -- none of them is on the accelerated path), and executes ``import main`` / ``import synthesize``: every
``from model ...`` / ``import audio`` line of the scripts and of ``src/{dataset,train,evaluation,utils}.py`` must resolve to dexb200.
Runs in the build container only (the GPU box has no /root/reference)."""
import os
import shutil
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.environ.get("DEX_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "DEX-TTS", "model")),
                                reason="the reference checkout is only present in the build container")

STUBS = {
    "matplotlib/__init__.py": "", "matplotlib/pyplot.py": "",
    "neptune.py": "", "soundfile.py": "", "resampy.py": "", "pyworld.py": "", "jiwer.py": "", "tgt.py": "",
    "librosa/__init__.py": "from . import effects, util, filters\n", "librosa/effects.py": "", "librosa/util.py": "def pad_center(*a, **k): raise NotImplementedError\ndef tiny(*a, **k): raise NotImplementedError\n",
    "librosa/filters.py": "def mel(*a, **k): raise NotImplementedError\n",
    "unidecode.py": "def unidecode(s): return s\n",
    "inflect.py": "class engine:\n    def number_to_words(self, *a, **k): return ''\n",
    "resemblyzer.py": "class VoiceEncoder: pass\ndef normalize_volume(*a, **k): pass\ndef trim_long_silences(*a, **k): pass\n",
    "g2p_en.py": "class G2p: pass\n",
    "soxr.py": "",                       # transformers.audio_utils imports it as soon as a module called librosa is importable
}


def _copy_reference(variant, dst):
    src = os.path.join(REF_ROOT, variant)
    keep = (".py", ".yaml", ".txt")
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ("checkpoints", "syn_samples", "filelists", "preprocess", "__pycache__")]
        for f in files:
            if f.endswith(keep) or base.endswith("resources"):
                rel = os.path.relpath(os.path.join(base, f), src)
                os.makedirs(os.path.dirname(os.path.join(dst, rel)) or dst, exist_ok=True)
                shutil.copyfile(os.path.join(base, f), os.path.join(dst, rel))


@pytest.mark.parametrize("variant,cls", [("DEX-TTS", "DeXTTS"), ("GeDEX-TTS", "GeDEXTTS")])
def test_main_and_synthesize_import_under_the_dropin(tmp_path, variant, cls):
    ref = tmp_path / "ref"
    stubs = tmp_path / "stubs"
    _copy_reference(variant, str(ref))
    for rel, body in STUBS.items():
        p = stubs / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(body)
    # the copy keeps the reference's own model/ and audio/ directories: the drop-ins must win because they come first on the path
    assert (ref / "model" / "tts.py").exists() and (ref / "audio" / "stft.py").exists()
    code = textwrap.dedent(f"""
        import sys
        import transformers                      # the reference pins transformers 4.35.2; this image has 5.x, which dropped two names its
        for name in ("Wav2Vec2Tokenizer", "top_k_top_p_filtering"):      # metric / text-encoder files import (SURVEY.md 8c)
            if not hasattr(transformers, name):
                setattr(transformers, name, None)
        import main, synthesize
        import model, audio, dexb200.model as M, dexb200.audio as A
        import hifigan, dexb200.hifigan as H, src.utils              # get_vocoder builds hifigan.Generator (src/utils.py:251-281)
        assert hifigan.Generator is H.Generator and src.utils.hifigan.Generator is H.Generator and hifigan.AttrDict is H.AttrDict
        assert model.{cls} is M.{cls}, model.__file__
        assert main.fix_len_compatibility is M.fix_len_compatibility
        assert synthesize.{cls} is M.{cls}
        if hasattr(synthesize, "Audio"):         # DEX-TTS only: GeDEX-TTS has no reference audio
            assert synthesize.Audio.stft.TacotronSTFT is A.stft.TacotronSTFT and synthesize.Audio.tools.get_mel_from_wav is A.tools.get_mel_from_wav
        else:
            assert "{variant}" == "GeDEX-TTS"
        import src.dataset, src.train, src.evaluation
        from dexb200.model.augmentation import Augment
        if hasattr(src.dataset, "Augment"):
            assert src.dataset.Augment is Augment
        else:
            assert "{variant}" == "GeDEX-TTS"
        assert src.train.{cls} is M.{cls} and src.evaluation.{cls} is M.{cls}
        assert 'dexb200' in sys.modules['model.utils'].fix_len_compatibility.__module__
        print('ok')
    """)
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dex-tts_b200", "dropin"), os.path.join(ROOT, "dex-tts_b200"),
                                                       str(stubs), str(ref)]))
    # cwd is NOT the checkout: `python -c` (like `python main.py`) puts the working / script directory first on sys.path, where the
    # checkout's own model/ would shadow the drop-in -- which is why dexb200.run exists (next test)
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), (out.stdout[-1500:], out.stderr[-3000:])
    # the launcher: `python -m dexb200.run <checkout>/synthesize.py --help` runs the unmodified script as __main__ from inside the
    # checkout with the drop-ins in front (argparse prints the script's own options and exits 0)
    env2 = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "dex-tts_b200"), str(stubs)]))
    out = subprocess.run([sys.executable, "-m", "dexb200.run", str(ref / "synthesize.py"), "--help"], env=env2, cwd=str(tmp_path),
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "--n_timesteps" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
    probe = ref / "probe_model.py"
    probe.write_text("import model, audio\nprint(model.__file__)\nprint(audio.__file__)\n")
    out = subprocess.run([sys.executable, "-m", "dexb200.run", str(probe)], env=env2, cwd=str(tmp_path), capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0 and out.stdout.count(os.path.join("dex-tts_b200", "dropin")) == 2, (out.stdout, out.stderr[-2000:])
