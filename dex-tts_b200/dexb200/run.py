"""Run one of the reference's entry scripts, unmodified, on the drop-in packages:

    python -m dexb200.run /path/to/DEX-TTS/synthesize.py --weight_path ... --input_text "..."
    python -m dexb200.run /path/to/DEX-TTS/main.py test --config ...

``python synthesize.py`` itself puts the script's directory first on ``sys.path``, so the checkout's own ``model/`` and ``audio/``
directories would shadow any drop-in.  This launcher reproduces what the interpreter does for a script (``sys.argv``, working
directory = the checkout, the checkout on the path) but with ``dex-tts_b200/dropin`` -- the packages literally named ``model`` and
``audio`` (DEX-TTS/synthesize.py:11,15; main.py:14; src/dataset.py:10-11; src/train.py:19; src/evaluation.py:15) -- in front of it,
and tells the training fall-through where the reference's own ``model`` package lives (``reference_twin.set_reference_dir``)."""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        raise SystemExit(f"dexb200.run: no such script: {argv[0]}")
    checkout = os.path.dirname(script)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))            # .../dex-tts_b200
    for p in (checkout, here, os.path.join(here, "dropin")):                      # inserted last = searched first
        while p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in [m for m in sys.modules if m.split(".")[0] in ("model", "audio")]:
        del sys.modules[name]
    from dexb200.model import reference_twin
    reference_twin.set_reference_dir(checkout)
    os.chdir(checkout)                                                            # the scripts use ./config, ./resources, ./checkpoints
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
