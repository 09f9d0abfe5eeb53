"""Host-side driver of the CUDA reverse-diffusion engine (libdexb200.so).

``ReverseDiffusion`` mirrors what ``Diffusion.forward(infer=True)`` does in the reference
(DEX-TTS/model/diffusion.py:250-259 -> ablation_sampler, DEX-TTS/model/edm.py:104-211): it owns one C handle, feeds it
the ``denoise_fn.*`` tensors of a state dict, plans a (B, T, Ts, n_steps) problem and runs whole trajectories on the
current CUDA stream.  PyTorch is used for device memory and streams only.
"""
import ctypes

import torch

from . import lib as _lib
from .manifest import decoder_manifest


def edm_sigmas(num_steps, sigma_min=0.002, sigma_max=80.0, rho=7, device="cpu"):
    """t_0..t_N of ablation_sampler(discretization='edm') in fp32, same op order as DEX-TTS/model/edm.py:136,152,179-180."""
    step_indices = torch.arange(num_steps, device=device)
    sigma_steps = (sigma_max ** (1 / rho) + step_indices / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    t_steps = torch.as_tensor(sigma_steps)
    return torch.cat([t_steps, torch.zeros_like(t_steps[:1])]).float().cpu()


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class ReverseDiffusion:
    """One handle = one (device, decoder weights) pair.  Not thread-safe."""

    def __init__(self, cfg, gemm_engine=0, nsplit=3):
        if not torch.cuda.is_available():
            raise RuntimeError("dexb200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.cfg = cfg
        self.L = _lib.load()
        self.device = torch.cuda.current_device()          # the handle's allocations, streams and launches belong to this device
        c = _lib.DexbConfig(variant=1 if cfg.variant == "dex" else 0, dim=cfg.dim, hidden=cfg.hidden, depth=cfg.depth,
                            heads=cfg.heads, mlp_hidden=int(cfg.hidden * cfg.mlp_ratio), patch=cfg.patch, stride=cfg.stride,
                            conv_pos=cfg.conv_pos, conv_pos_groups=cfg.conv_pos_groups, n_feats=cfg.n_feats,
                            pe_scale=float(cfg.pe_scale), gemm_engine=gemm_engine, nsplit=nsplit,
                            n_spks=int(cfg.n_spks), spk_emb_dim=int(cfg.spk_emb_dim))
        self.multi_spk = cfg.variant == "gedex" and cfg.n_spks > 1
        h = ctypes.c_void_p()
        _lib.check(self.L.dexb_create(ctypes.byref(c), ctypes.byref(h)), "dexb_create")
        self.h = h
        self.plan_key = None
        self.workspace_bytes = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.dexb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -----------------------------------------------------------------------------------
    def load_state_dict(self, sd, prefix="denoise_fn."):
        """Copy every manifest tensor (``prefix + name``) to the handle and pack it.  Tensors may live on any device."""
        dev = torch.device("cuda", self.device)
        with torch.cuda.device(self.device):
            staged = []
            for e in decoder_manifest(self.cfg):
                key = prefix + e.name
                if key not in sd:
                    raise RuntimeError(f"state dict is missing '{key}'")
                t = sd[key].detach().to(device=dev, dtype=torch.float32).contiguous()
                if tuple(t.shape) != tuple(e.shape):
                    raise RuntimeError(f"'{key}' has shape {tuple(t.shape)}, expected {tuple(e.shape)}")
                staged.append((e, t))
            # the casts / copies above were enqueued on torch's current stream; dexb_load_weight copies synchronously on the legacy
            # stream, which does not order against a non-blocking torch stream: finish them first
            torch.cuda.current_stream().synchronize()
            for e, t in staged:
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                _lib.check(self.L.dexb_load_weight(self.h, e.name.encode(), _ptr(t), shape, t.dim()), f"dexb_load_weight({e.name})")
            torch.cuda.synchronize()
            _lib.check(self.L.dexb_finalize_weights(self.h, _stream()), "dexb_finalize_weights")
        self.plan_key = None

    def _check_device(self, t):
        if not t.is_cuda or t.device.index != self.device:
            raise RuntimeError(f"this dexb200 handle lives on cuda:{self.device} but got a tensor on {t.device}; build one engine (model "
                               "copy) per device")

    # ---- plan --------------------------------------------------------------------------------------
    def plan(self, B, T, Ts, n_steps, Tr=None):
        Tr = Ts if Tr is None else Tr
        key = (int(B), int(T), int(Ts), int(Tr), int(n_steps))
        if key == self.plan_key:
            return
        sig = edm_sigmas(n_steps).contiguous()
        self.sigmas = sig
        ws = ctypes.c_size_t(0)
        with torch.cuda.device(self.device):
            _lib.check(self.L.dexb_plan(self.h, key[0], key[1], key[2], key[3], key[4],
                                        sig.numpy().ctypes.data_as(_lib.c_float_p), ctypes.byref(ws)), "dexb_plan")
        self.workspace_bytes = ws.value
        self.plan_key = key

    def _cond(self, cond, B):
        if self.multi_spk:                       # GeDEX-TTS speaker embedding (GeDEX-TTS/model/tts.py:30-31,53)
            if cond is None or cond.get("spk") is None:
                raise RuntimeError("multi-speaker GeDEX-TTS needs cond['spk'] (B, spk_emb_dim)")
            spk = cond["spk"].detach().float().reshape(B, self.cfg.spk_emb_dim).contiguous()
            c = _lib.DexbCond()
            c.spk_dev = spk.data_ptr()
            return c, [spk]
        if self.cfg.variant != "dex":
            return None, []
        sty = cond["sty"].float().contiguous()
        sl = cond["sty_lengths"].to(device=sty.device, dtype=torch.int32).contiguous()
        refs = [r.float().contiguous() for r in cond["ref_skips"]]
        if len(refs) != 6:
            raise RuntimeError("DEX-TTS expects 6 reference skip tensors")
        c = _lib.DexbCond()
        c.sty_dev = sty.data_ptr()
        c.sty_len_dev = sl.data_ptr()
        for i, r in enumerate(refs):
            c.ref_skips_dev[i] = r.data_ptr()
        c.Tr = refs[0].shape[-1]
        return c, [sty, sl] + refs

    # ---- run ---------------------------------------------------------------------------------------
    def sample(self, x0, mask, mu, n_steps, cond=None):
        """x0 = z / temperature + mu (B,80,T) -> generated mel (B,80,T).  mask (B,1,T) or (B,T).  All CUDA fp32."""
        B, F, T = x0.shape
        self._check_device(x0)
        Ts = cond["sty"].shape[-1] if self.cfg.variant == "dex" else 0
        Tr = cond["ref_skips"][0].shape[-1] if self.cfg.variant == "dex" else 0
        self.plan(B, T, Ts, n_steps, Tr)
        with torch.cuda.device(self.device):
            x = x0.detach().float().contiguous().clone()
            mu = mu.detach().float().contiguous()
            m = mask.detach().float().reshape(B, T).contiguous()
            c, keep = self._cond(cond, B)
            _lib.check(self.L.dexb_reverse_diffusion(self.h, _ptr(x), _ptr(mu), _ptr(m), ctypes.byref(c) if c is not None else None,
                                                     _stream()), "dexb_reverse_diffusion")
        self._keep = keep + [mu, m]
        return x

    def sample_host(self, x0, mask, mu, n_steps, cond=None):
        """Same trajectory through the HOST-buffer entry point: pinned/pageable CPU tensors in, CPU tensor out
        (host<->device copies happen inside the C call, which synchronises the stream)."""
        B, F, T = x0.shape
        Ts = cond["sty"].shape[-1] if self.cfg.variant == "dex" else 0
        Tr = cond["ref_skips"][0].shape[-1] if self.cfg.variant == "dex" else 0
        self.plan(B, T, Ts, n_steps, Tr)
        x = x0.float().contiguous().clone()
        if x0.is_pinned():
            x = x.pin_memory()
        mu = mu.float().contiguous()
        m = mask.float().reshape(B, T).contiguous()
        sty_p = sl_p = None
        refs_arr = None
        Tr = 0
        keep = []
        if self.cfg.variant == "dex":
            sty = cond["sty"].float().contiguous()
            sl = cond["sty_lengths"].to(torch.int32).contiguous()
            refs = [r.float().contiguous() for r in cond["ref_skips"]]
            refs_arr = (ctypes.c_void_p * 6)(*[r.data_ptr() for r in refs])
            sty_p, sl_p, Tr = _ptr(sty), _ptr(sl), refs[0].shape[-1]
            keep = [sty, sl] + refs
        spk_p = None
        if self.multi_spk:
            spk = cond["spk"].float().reshape(B, self.cfg.spk_emb_dim).contiguous()
            spk_p = _ptr(spk)
            keep.append(spk)
        _lib.check(self.L.dexb_reverse_diffusion_host(self.h, _ptr(x), _ptr(mu), _ptr(m), sty_p, sl_p, refs_arr, Tr, spk_p,
                                                      _stream()), "dexb_reverse_diffusion_host")
        del keep
        return x

    def denoise_once(self, x, mask, mu, n_steps, step, cond=None):
        """D(x; sigma_step) of EDMPrecond (edm.py:88-98) for unit parity; x is the sampler state at that step."""
        B, F, T = x.shape
        Ts = cond["sty"].shape[-1] if self.cfg.variant == "dex" else 0
        Tr = cond["ref_skips"][0].shape[-1] if self.cfg.variant == "dex" else 0
        self.plan(B, T, Ts, n_steps, Tr)
        x = x.detach().float().contiguous()
        mu = mu.detach().float().contiguous()
        m = mask.detach().float().reshape(B, T).contiguous()
        out = torch.empty_like(x)
        c, keep = self._cond(cond, B)
        _lib.check(self.L.dexb_denoise_once(self.h, _ptr(x), _ptr(mu), _ptr(m), ctypes.byref(c) if c is not None else None,
                                            int(step), _ptr(out), _stream()), "dexb_denoise_once")
        torch.cuda.synchronize()
        del keep
        return out

    def debug_tap(self, name):
        """fp32 (B, C, H, W) copy of an internal activation of the last ``denoise_once`` call (names: include/dexb200.h)."""
        C, H, W = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(self.L.dexb_debug_tap(self.h, name.encode(), None, ctypes.byref(C), ctypes.byref(H), ctypes.byref(W), _stream()),
                   "dexb_debug_tap")
        out = torch.empty(self.plan_key[0], C.value, H.value, W.value, device="cuda", dtype=torch.float32)
        _lib.check(self.L.dexb_debug_tap(self.h, name.encode(), _ptr(out), ctypes.byref(C), ctypes.byref(H), ctypes.byref(W), _stream()),
                   "dexb_debug_tap")
        torch.cuda.synchronize()
        return out

    def profile_step(self, step=0):
        """[(tag, ms, gflop)] for every launch of one network call (un-graphed, events around each launch)."""
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(self.L.dexb_profile_step(self.h, int(step), buf, len(buf), _stream()), "dexb_profile_step")
        out = []
        for line in buf.value.decode().splitlines():
            tag, ms, gf = line.split("\t")
            out.append((tag, float(ms), float(gf)))
        return out

    @property
    def launches(self):
        return int(self.L.dexb_last_launch_count(self.h))

    @property
    def simt_fallbacks(self):
        return int(self.L.dexb_simt_fallbacks(self.h))


def gemm_test(a, w, bias=None, taps=(1, 1), off=(0, 0), in_stride=1, engine=0, nsplit=3):
    """Unit entry of the implicit-GEMM engine.  a (nimg,H,W,K), w (KH*KW,N,K), bias (N) -> (nimg,ceil(H/s),ceil(W/s),N)."""
    L = _lib.load()
    a = a.float().contiguous()
    w = w.float().contiguous()
    nimg, H, W, K = a.shape
    N = w.shape[1]
    oh, ow = (H + in_stride - 1) // in_stride, (W + in_stride - 1) // in_stride
    out = torch.zeros(nimg, oh, ow, N, device=a.device, dtype=torch.float32)
    b = bias.float().contiguous() if bias is not None else None
    torch.cuda.synchronize()
    _lib.check(L.dexb_gemm_test(engine, nsplit, _ptr(a), nimg, H, W, K, _ptr(w), N, taps[0], taps[1], off[0], off[1], in_stride,
                                _ptr(b) if b is not None else None, _ptr(out), _stream()), "dexb_gemm_test")
    return out


def attn_test(qkv, heads):
    """Unit entry of the fused attention kernel: qkv (B, N, 3*hid) -> (B, N, hid)."""
    L = _lib.load()
    qkv = qkv.float().contiguous()
    B, N, C3 = qkv.shape
    hid = C3 // 3
    out = torch.zeros(B, N, hid, device=qkv.device, dtype=torch.float32)
    torch.cuda.synchronize()
    _lib.check(L.dexb_attn_test(_ptr(qkv), B, N, heads, hid, _ptr(out), _stream()), "dexb_attn_test")
    return out


def stft_mel(wav, window, mel_basis, n_fft=1024, hop=256, return_energy=False):
    """wav (B,S) in [-1,1] on CUDA -> log-mel (B, n_mels, 1 + S // hop) [, energy (B, 1 + S // hop)].  Mirrors
    TacotronSTFT.mel_spectrogram (DEX-TTS/audio/stft.py:159-178)."""
    L = _lib.load()
    wav = wav.float().contiguous()
    B, S = wav.shape
    n_mels = mel_basis.shape[0]
    out = torch.empty(B, n_mels, S // hop + 1, device=wav.device, dtype=torch.float32)
    energy = torch.empty(B, S // hop + 1, device=wav.device, dtype=torch.float32) if return_energy else None
    window = window.float().contiguous()
    mel_basis = mel_basis.float().contiguous()
    _lib.check(L.dexb_stft_mel(_ptr(wav), B, S, _ptr(window), _ptr(mel_basis), n_fft, hop, n_mels, _ptr(out),
                               _ptr(energy) if energy is not None else None, _stream()), "dexb_stft_mel")
    return (out, energy) if return_energy else out
