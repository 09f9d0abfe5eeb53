"""Run network calls of a workload un-graphed (plain launches) so ncu / the launch list see individual kernels.
usage: python tools/prof_net_call.py [C2] [n_calls]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-tts_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from dexb200.engine import ReverseDiffusion  # noqa: E402
from dexb200.manifest import DecoderCfg  # noqa: E402
from dexb200.synth import synth_decoder_weights, synth_inputs  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
n_calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
variant, B, T, Ts, n_steps = WORKLOADS[wl]
cfg = DecoderCfg.make(variant)
eng = ReverseDiffusion(cfg)
eng.load_state_dict(synth_decoder_weights(cfg, seed=100, live=True))
inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=1234)
cond = None
if variant == "dex":
    cond = dict(sty=inp["sty"].cuda(), sty_lengths=inp["sty_lengths"].cuda(), ref_skips=[r.cuda() for r in inp["ref_skips"]])
from dexb200.engine import edm_sigmas  # noqa: E402
step = n_steps // 2
sig = float(edm_sigmas(n_steps)[step])
x = (inp["mu"] + inp["z"] * sig).cuda()                  # a plausible sampler state at that step: signal + sigma * noise
for i in range(n_calls):
    out = eng.denoise_once(x, inp["mask"].cuda(), inp["mu"].cuda(), n_steps, step, cond=cond)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
