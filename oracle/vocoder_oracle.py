"""CPU oracle for the step after the path: the HiFi-GAN v1 generator the reference vocodes with (mel -> waveform).

TEST INFRASTRUCTURE ONLY -- never imported by the product path (``dex-tts_b200/``); see the header of ``dex_oracle.py``.
SURVEY.md §8f rank 3.  There is NO CUDA side for this stage yet: this file and tests/golden/voc_*.npz are the pinned oracle the
next round can build against.

Functional restatement over a flat ``{name: tensor}`` dict keyed by the generator's ``state_dict`` names AFTER
``remove_weight_norm()`` -- the state ``get_vocoder`` leaves it in (DEX-TTS/src/utils.py:251-281) -- of

    Generator.forward     DEX-TTS/hifigan/models.py:157-173   (upsample rates 8,8,2,2 / kernels 16,16,4,4, 512 initial channels,
                                                               DEX-TTS/hifigan/config.json:12-16)
    ResBlock.forward      DEX-TTS/hifigan/models.py:96-103    (kernels 3,7,11 x dilations 1,3,5)

Parity pin: outputs of the unmodified reference ``Generator`` with seeded weights, generated in the build container by
oracle/make_golden_vocoder.py and committed as tests/golden/voc_*.npz; tests/test_vocoder_oracle.py replays them.
"""
import torch.nn.functional as F
import torch

UPSAMPLE_RATES = (8, 8, 2, 2)
UPSAMPLE_KERNELS = (16, 16, 4, 4)
RESBLOCK_KERNELS = (3, 7, 11)
RESBLOCK_DILATIONS = (1, 3, 5)
INITIAL_CHANNELS = 512
LRELU_SLOPE = 0.1


def vocoder_manifest(n_mels=80):
    """[(name, shape)] of the weight-norm-free generator (``remove_weight_norm`` re-registers each weight after its bias, so the
    upstream key order differs; ``load_state_dict`` goes by name)."""
    out = [("conv_pre.weight", (INITIAL_CHANNELS, n_mels, 7)), ("conv_pre.bias", (INITIAL_CHANNELS,))]
    for i, k in enumerate(UPSAMPLE_KERNELS):
        ci, co = INITIAL_CHANNELS >> i, INITIAL_CHANNELS >> (i + 1)
        out.extend([(f"ups.{i}.weight", (ci, co, k)), (f"ups.{i}.bias", (co,))])       # ConvTranspose1d: (in, out, k)
    for i in range(len(UPSAMPLE_RATES)):
        ch = INITIAL_CHANNELS >> (i + 1)
        for j, k in enumerate(RESBLOCK_KERNELS):
            r = i * len(RESBLOCK_KERNELS) + j
            for grp in ("convs1", "convs2"):
                for d in range(len(RESBLOCK_DILATIONS)):
                    out.extend([(f"resblocks.{r}.{grp}.{d}.weight", (ch, ch, k)), (f"resblocks.{r}.{grp}.{d}.bias", (ch,))])
    ch = INITIAL_CHANNELS >> len(UPSAMPLE_RATES)
    out.extend([("conv_post.weight", (1, ch, 7)), ("conv_post.bias", (1,))])
    return out


def synth_vocoder_weights(seed=100):
    """Seeded weights at a scale that keeps activations O(1) through the 4 stages (the reference's N(0, 0.01) init would give ~0)."""
    w = {}
    for n, (name, shape) in enumerate(vocoder_manifest()):
        g = torch.Generator()
        g.manual_seed(seed * 7919 + n)
        if name.endswith("bias"):
            w[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan = (shape[1] if not name.startswith("ups.") else shape[0]) * shape[2]
            if name.startswith("ups."):
                fan = shape[0] * shape[2] / UPSAMPLE_RATES[int(name.split(".")[1])]        # taps that reach one output sample
            gain = 0.25 if name.startswith("conv_pre") else 0.45 if ".convs2." in name else 0.3 if name.startswith("conv_post") else 1.4
            w[name] = torch.randn(shape, generator=g) * (gain / fan ** 0.5)
    return w


def _resblock(w, p, x, k):
    """ResBlock.forward, models.py:96-103."""
    for d, dil in enumerate(RESBLOCK_DILATIONS):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, w[f"{p}.convs1.{d}.weight"], w[f"{p}.convs1.{d}.bias"], dilation=dil, padding=(k * dil - dil) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, w[f"{p}.convs2.{d}.weight"], w[f"{p}.convs2.{d}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def hifigan_generator(w, mel):
    """Generator.forward, models.py:157-173: mel (B, 80, T) -> waveform (B, 1, 256 T)."""
    x = F.conv1d(mel, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    nk = len(RESBLOCK_KERNELS)
    for i, (u, k) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, w[f"ups.{i}.weight"], w[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(RESBLOCK_KERNELS):
            y = _resblock(w, f"resblocks.{i * nk + j}", x, rk)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                                   # models.py:169: default slope 0.01 here, not LRELU_SLOPE
    x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
    return torch.tanh(x)
