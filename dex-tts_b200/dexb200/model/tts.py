"""``DeXTTS`` / ``GeDEXTTS`` -- the top-level models of the reference with every inference stage on hand-written sm_100a CUDA.

Replace (same constructor argument ``cfg`` = the ``model:`` block of config/*/base.yaml plus ``n_vocab``, same ``forward`` signature and
return value, same ``state_dict`` keys -- ``load_state_dict(ckpt['state_dict'])`` as in synthesize.py:68-72 works unchanged):
    DEX-TTS/model/tts.py:12-74      class DeXTTS    (``from model import DeXTTS``, DEX-TTS/synthesize.py:11,67,105)
    GeDEX-TTS/model/tts.py:14-56    class GeDEXTTS  (GeDEX-TTS/synthesize.py)

``forward`` is the reference's own sequence of calls with each stage replaced by its C-ABI counterpart: LF0 / TV encoders and the style
fusion (``dexb_lf0_*``, ``dexb_tv_*``, ``dexb_style_fuse``), TIV encoder (``dexb_tiv_*``), text encoder (``dexb_text_*``), duration /
alignment glue (``dexb_align_*``) and the reverse-diffusion loop (``dexb_reverse_diffusion``).  PyTorch only allocates the tensors and
builds the three sequence masks.  ``compute_loss`` keeps the reference signature and delegates to the reference's own PyTorch
training code running on this model's parameters, with the Monotonic Alignment Search on the CUDA kernel (``reference_twin.py``).
"""
import torch
import torch.nn as nn

from .diffusion import Diffusion, GeDiffusion
from .ref_encoder import LF0Encoder, TIVEncoder, TVEncoder, style_fusion
from .text_encoder import GeTextEncoder, TextEncoder
from .utils import align_durations, sequence_mask


def _get(cfg, key):
    return cfg[key] if isinstance(cfg, dict) else getattr(cfg, key)


class DeXTTS(nn.Module):
    """DEX-TTS/model/tts.py:12-74."""

    def __init__(self, cfg):
        super().__init__()
        self._cfg = cfg
        self.n_spks = 0                                                       # tts.py:16 forces cfg.n_spks = 0
        self.n_feats = _get(cfg, "n_feats")
        tv, dec = _get(cfg, "tv_encoder"), _get(cfg, "decoder")
        self.tv_encoder = TVEncoder(**tv)
        self.lf0_encoder = LF0Encoder(**_get(cfg, "lf0_encoder"))
        self.tiv_encoder = TIVEncoder(**_get(cfg, "tiv_encoder"))
        self.encoder = TextEncoder(**_get(cfg, "encoder"), n_vocab=_get(cfg, "n_vocab"), n_feats=self.n_feats, n_spks=0,
                                   spk_emb_dim=_get(cfg, "spk_emb_dim"))
        self.decoder = Diffusion(**dec, dit_cfg=_get(cfg, "dit"), n_feats=self.n_feats, n_spks=0, spk_emb_dim=_get(cfg, "spk_emb_dim"))
        self.conv_sty = nn.Conv1d(_get(tv, "c_out_g"), _get(dec, "dim") * 2, 1, 1)                     # tts.py:31

    @torch.no_grad()
    def forward(self, x, x_lengths, ref, ref_lengths, sty, sty_lengths, lf0, lf0_lengths, n_timesteps, temperature=1.0, spk=None,
                length_scale=1.0):
        ref_mask = torch.unsqueeze(sequence_mask(ref_lengths, ref.size(2)), 1).to(ref.dtype)         # tts.py:38-40
        lf0_mask = torch.unsqueeze(sequence_mask(lf0_lengths, lf0.size(1)), 1).to(lf0.dtype)
        sty_mask = torch.unsqueeze(sequence_mask(sty_lengths, sty.size(2)), 1).to(sty.dtype)

        lf0_enc, lf0_dec = self.lf0_encoder(lf0, lf0_mask)                                           # :42
        sty_enc, sty_dec, _ = self.tv_encoder(sty, sty_mask)                                         # :43
        sty_enc, sty_dec = style_fusion(self.conv_sty, sty_enc, sty_dec, sty_mask, lf0_enc, lf0_dec, lf0_mask)   # :45-49

        ref, ref_skips = self.tiv_encoder(ref, ref_mask)                                             # :50
        mu_x, logw, x_mask = self.encoder(x, x_lengths, sty_enc, spk=None)                           # :51

        mu_y, y_mask, attn, _, y_max_length = align_durations(logw, x_mask, mu_x, length_scale)      # :55-68
        enc_out = mu_y[:, :, :y_max_length]                                                          # :69

        dec_out = self.decoder(mu_y, y_mask, mu_y, ref_skips, ref_lengths, sty_dec, sty_lengths, temperature=temperature,
                               n_timesteps=n_timesteps, spk=spk, infer=True)                         # :71
        dec_out = dec_out[:, :, :y_max_length]
        return enc_out, dec_out, attn[:, :, :y_max_length]                    # :74 (upstream slices the 4-D attn along Tx here)

    def _training_twin(self):
        if getattr(self, "_twin", None) is None:
            from .reference_twin import tts_twin
            object.__setattr__(self, "_twin", tts_twin(self, self._cfg, "dex"))     # not a registered sub-module: no duplicate keys
        self._twin.train(self.training)
        return self._twin

    def compute_loss(self, x, x_lengths, y, y_lengths, ref, ref_lengths, sty, sty_lengths, lf0, lf0_lengths, spk=None, out_size=None,
                     mask_ratio=0):
        """-> (dur_loss, prior_loss, diff_loss, vq_loss), DEX-TTS/model/tts.py:76-153.  Training is not a CUDA path of this package:
        the reference's own modules compute it on this model's parameters (see reference_twin.py)."""
        return self._training_twin().compute_loss(x, x_lengths, y, y_lengths, ref, ref_lengths, sty, sty_lengths, lf0, lf0_lengths,
                                                  spk=spk, out_size=out_size, mask_ratio=mask_ratio)


class GeDEXTTS(nn.Module):
    """GeDEX-TTS/model/tts.py:14-56."""

    def __init__(self, cfg):
        super().__init__()
        self._cfg = cfg
        self.n_spks = _get(cfg, "n_spks")
        self.n_feats = _get(cfg, "n_feats")
        if self.n_spks > 1:
            self.spk_emb = nn.Embedding(self.n_spks, _get(cfg, "spk_emb_dim"))                       # tts.py:22-23
        self.encoder = GeTextEncoder(**_get(cfg, "encoder"), n_vocab=_get(cfg, "n_vocab"), n_feats=self.n_feats, n_spks=self.n_spks,
                                     spk_emb_dim=_get(cfg, "spk_emb_dim"))
        self.decoder = GeDiffusion(**_get(cfg, "decoder"), dit_cfg=_get(cfg, "dit"), n_feats=self.n_feats, n_spks=self.n_spks,
                                   spk_emb_dim=_get(cfg, "spk_emb_dim"))

    @torch.no_grad()
    def forward(self, x, x_lengths, n_timesteps, temperature=1.0, spk=None, length_scale=1.0):
        if self.n_spks > 1:
            spk = self.spk_emb(spk)                                                                  # tts.py:30-31
        mu_x, logw, x_mask = self.encoder(x, x_lengths, spk=spk)                                     # :34
        mu_y, y_mask, attn, _, y_max_length = align_durations(logw, x_mask, mu_x, length_scale)      # :36-50
        enc_out = mu_y[:, :, :y_max_length]
        dec_out = self.decoder(mu_y, y_mask, mu_y, temperature=temperature, n_timesteps=n_timesteps, spk=spk, infer=True)   # :53
        dec_out = dec_out[:, :, :y_max_length]
        return enc_out, dec_out, attn[:, :, :y_max_length]

    def _training_twin(self):
        if getattr(self, "_twin", None) is None:
            from .reference_twin import tts_twin
            object.__setattr__(self, "_twin", tts_twin(self, self._cfg, "gedex"))
        self._twin.train(self.training)
        return self._twin

    def compute_loss(self, x, x_lengths, y, y_lengths, spk=None, out_size=None, mask_ratio=0):
        """-> (dur_loss, prior_loss, diff_loss), GeDEX-TTS/model/tts.py:58-121, computed by the reference's own modules on this
        model's parameters (see reference_twin.py)."""
        return self._training_twin().compute_loss(x, x_lengths, y, y_lengths, spk=spk, out_size=out_size, mask_ratio=mask_ratio)
