#!/usr/bin/env python
"""Benchmark of the reverse-diffusion hot path (BASELINE.json metric: mel-frames/sec through 50-step reverse diffusion).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C1|C3|C4|C5]

A bench "step" is ONE complete reverse diffusion (all sampler steps) of one batch of synthetic utterances.
Default workload = BASELINE.json configs[1] ("C2"): DEX-TTS, batch 8 per GPU, 50 sampler steps, 80x512 mel, style
length 259, synthetic inputs and seeded random ("live") weights of the reference architecture.
  value : mel-frames/s with inputs resident in HBM (CUDA events around K trajectories, max over ranks)
  e2e   : the same metric through the public Python API, ``dexb200.model.Diffusion.forward(..., infer=True)`` (the module the
          reference attaches as ``model.decoder``), with HOST inputs: pinned host tensors -> device copies -> loop -> mel back on
          the host, all inside the timed region (+ the NCCL gather of the mels at N > 1)
  parity: before anything is timed, sample 0 of the workload runs 2 sampler steps on the CUDA path and on the CPU oracle; the
          per-bin violation (tolerance 1e-3) is printed in the line
  roofline     : the dominant kernel class (tcgen05 implicit-GEMM), timed live with CUDA events around its launches
  cpu_baseline : the REFERENCE's own modules (``baseline/_ref``: unmodified copy of DEX-TTS|GeDEX-TTS/model, staged by
                 ``__graft_entry__.build()``) on this box's host cores, ``Diffusion.forward(infer=True)`` for a bounded number of
                 sampler steps of the same batch, extrapolated to the full step count (kind "reference"; falls back to the
                 oracle port, kind "port", when the staged copy is absent)
`--impl reference` times that CPU path as the main line (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "dex-tts_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from dexb200.manifest import DecoderCfg  # noqa: E402
from dexb200.synth import synth_decoder_weights, synth_inputs  # noqa: E402

WORKLOADS = {
    # name: (variant, B per GPU, T, Ts, sampler steps)
    "C1": ("gedex", 1, 200, 0, 10),
    "C2": ("dex", 8, 512, 259, 50),
    "C3": ("dex", 32, 512, 259, 100),
    "C4": ("gedex", 32, 512, 0, 50),
    "C5": ("dex", 8, 2000, 259, 50),
}
METRIC = "mel-frames/sec through 50-step reverse diffusion"


def algorithmic_flops(variant, T, Ts):
    """SURVEY.md 8(d) closed form (2*MAC) of one denoiser call for one utterance."""
    import dex_oracle as O
    return O.flops_per_sample_step(O.make_cfg(variant), T, Ts)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def port_run(variant, B, T, Ts, n_steps, sample_steps, repeats):
    """Time the CPU oracle (port) on `sample_steps` sampler steps of the workload's batch; returns (sec per sampler step, cores)."""
    import dex_oracle as O
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    cfg = DecoderCfg.make(variant)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=1234)
    cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"]) if variant == "dex" else None
    ocfg = O.make_cfg(variant)
    ts = O.sigma_schedule(n_steps)
    x = (inp["z"] / 1.5 + inp["mu"]) * ts[0]
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            xx = x
            for i in range(sample_steps):
                den = O.edm_precond(w, ocfg, xx, ts[i], inp["mask"], inp["mu"], cond=cond)
                xx = xx + (ts[i + 1] - ts[i]) * ((1 / ts[i]) * xx - 1 / ts[i] * den)
            dt = (time.perf_counter() - t0) / sample_steps
            best = dt if best is None else min(best, dt)
    return best, cores


_REF = {}


def reference_root():
    """Where the unmodified reference sources are: the copy staged by __graft_entry__.build() (travels to the GPU box), else the
    checkout itself (build container)."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "DEX-TTS", "model", "diffusion.py")):
            return cand
    return None


def reference_decoder(variant):
    """The reference's own ``Diffusion`` (DEX-TTS/model/diffusion.py:238 / GeDEX-TTS/model/diffusion.py:209) with the seeded synthetic
    weights, built exactly like oracle/make_golden.py builds it for the fixtures (shims: a timm stub, nothing edited)."""
    if variant in _REF:
        return _REF[variant]
    root = reference_root()
    if root is None:
        return None
    os.environ["DEX_REFERENCE_ROOT"] = root
    import ref_loader
    ref_loader.REF_ROOT = root
    import contextlib
    import io
    cfg = DecoderCfg.make(variant)
    dec_cfg = dict(dim=cfg.dim, pe_scale=cfg.pe_scale, dim_mults=[1, 2], model_type="dit", precond="edm", loss_type="base")
    dit_cfg = dict(in_channels=3, patch_size=cfg.patch, stride_size=cfg.stride, overlap=True, hidden_size=cfg.hidden, depth=cfg.depth,
                   num_heads=cfg.heads, mlp_ratio=cfg.mlp_ratio, out_channels=1, conv_pos=cfg.conv_pos,
                   conv_pos_groups=cfg.conv_pos_groups, use_decoder=False, mask_type="time_random")
    with contextlib.redirect_stdout(io.StringIO()):                  # the constructor prints a banner; stdout carries the JSON line
        dec, mod = ref_loader.build_reference_decoder(variant, dec_cfg, dit_cfg)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    sd = dict(w)
    sd.update({k.replace("denoise_fn.", "precond_model.model."): v for k, v in w.items()})
    dec.load_state_dict(sd, strict=True)
    orig = mod.EDMPrecond.forward

    def fwd(self, x, sigma, *a, **k):        # reference bug at B > 1: sigma arrives 0-dim; per-sample math unchanged
        return orig(self, x, sigma.reshape(-1).expand(x.shape[0]), *a, **k)
    mod.EDMPrecond.forward = fwd
    _REF[variant] = dec
    return dec


def cpu_reference_run(variant, B, T, Ts, n_steps, sample_steps, repeats):
    """Time the reference's own ``Diffusion.forward(infer=True)`` for `sample_steps` sampler steps of the workload's batch on all host
    cores; returns (sec per sampler step, cores, kind).  Falls back to the oracle port when the reference copy is not staged."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    dec = reference_decoder(variant)
    if dec is None:
        dt, cores = port_run(variant, B, T, Ts, n_steps, sample_steps, repeats)
        return dt, cores, "port"
    cfg = DecoderCfg.make(variant)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=1234)
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            if variant == "dex":
                dec(inp["mu"], inp["mask"], inp["mu"], inp["ref_skips"], inp["ref_lengths"], inp["sty"], inp["sty_lengths"],
                    n_timesteps=sample_steps, infer=True, temperature=1.5)
            else:
                dec(inp["mu"], inp["mask"], inp["mu"], n_timesteps=sample_steps, infer=True, temperature=1.5)
            dt = (time.perf_counter() - t0) / sample_steps
            best = dt if best is None else min(best, dt)
    return best, cores, "reference"


def parity_check(eng, variant, B, T, Ts, seed):
    """Sample 0 of the workload, 2 sampler steps: CUDA path vs the CPU oracle (test infrastructure used as the checker only)."""
    import dex_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from parity import REL_TOL, per_bin_violation
    cfg = DecoderCfg.make(variant)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=seed)
    cond = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"]) if variant == "dex" else None
    cond_d = dict(sty=cond["sty"].cuda(), sty_lengths=cond["sty_lengths"].cuda(), ref_skips=[r.cuda() for r in cond["ref_skips"]]) if cond else None
    x0 = inp["z"] / 1.5 + inp["mu"]
    y = eng.sample(x0.cuda(), inp["mask"].cuda(), inp["mu"].cuda(), 2, cond=cond_d)[:1].cpu()
    c1 = dict(sty=cond["sty"][:1], sty_lengths=cond["sty_lengths"][:1], ref_skips=[r[:1] for r in cond["ref_skips"]]) if cond else None
    torch.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    with torch.no_grad():
        ref = O.reverse_diffusion(w, O.make_cfg(variant), inp["z"][:1], inp["mask"][:1], inp["mu"][:1], 2, temperature=1.5, cond=c1)
    v = per_bin_violation(y, ref)
    rms_err = float((y.double() - ref.double()).pow(2).mean().sqrt() / ref.double().pow(2).mean().sqrt())
    return {"per_bin_violation": v, "rms_rel_err": rms_err, "tol": REL_TOL, "ok": bool(v < REL_TOL),
            "checked": f"sample 0 of the timed batch (B={B}, T={T}), 2 sampler steps, CUDA path vs oracle/dex_oracle.py on the host"}


def build_decoder(variant, gemm_engine=0, nsplit=3):
    """The drop-in decoder module (``model.decoder`` of DeXTTS / GeDEXTTS) with the seeded synthetic weights, on the current GPU."""
    from dexb200.model import Diffusion, GeDiffusion
    cfg = DecoderCfg.make(variant)
    dit = dict(in_channels=3, patch_size=cfg.patch, stride_size=cfg.stride, overlap=True, hidden_size=cfg.hidden, depth=cfg.depth,
               num_heads=cfg.heads, mlp_ratio=cfg.mlp_ratio, out_channels=1, conv_pos=cfg.conv_pos, conv_pos_groups=cfg.conv_pos_groups,
               use_decoder=False, mask_type="time_random")
    cls = Diffusion if variant == "dex" else GeDiffusion
    dec = cls(n_feats=cfg.n_feats, dim=cfg.dim, dit_cfg=dit, model_type="dit", dim_mults=[1, 2], n_spks=cfg.n_spks,
              spk_emb_dim=cfg.spk_emb_dim, pe_scale=cfg.pe_scale, gemm_engine=gemm_engine, nsplit=nsplit)
    w = synth_decoder_weights(cfg, seed=100, live=True)
    sd = dict(w)
    sd.update({k.replace("denoise_fn.", "precond_model.model."): v for k, v in w.items()})
    dec.load_state_dict(sd, strict=True)
    return dec.cuda().eval(), cfg


def run_workload(args, name, rank, world, local_rank, dist, full):
    """One bench line for workload `name`.  full = roofline / cpu_baseline / parity legs as well (the headline workload)."""
    variant, B, T, Ts, n_steps = WORKLOADS[name]
    wl = f"{name}: {'DEX-TTS' if variant == 'dex' else 'GeDEX-TTS'} B={B}/GPU T={T} (80x{T} mel) Ts={Ts} {n_steps} sampler steps"
    config = {"workload": wl, "batch_per_gpu": B, "mel_frames": T, "sampler_steps": n_steps, "style_len": Ts,
              "parallelism": f"dp{world}", "l2": "per-step working set (hundreds of MB of activations) exceeds the 126 MB L2",
              "weights": "seeded random init of the reference architecture, zero-initialised tensors re-drawn (live)"}
    steps, warmup = (args.steps, max(3, args.warmup)) if full else (2, 3)
    dec, cfg = build_decoder(variant, args.gemm_engine, args.nsplit)
    eng = dec.cuda_engine()
    parity = parity_check(eng, variant, B, T, Ts, 1234) if (rank == 0 and not args.no_parity) else None
    inp = synth_inputs(cfg, B, T, Ts=max(Ts, 1), seed=1234 + rank)
    x0 = inp["z"] / 1.5 + inp["mu"]
    cond_h = dict(sty=inp["sty"], sty_lengths=inp["sty_lengths"], ref_skips=inp["ref_skips"]) if variant == "dex" else None
    dev = lambda t: t.cuda(non_blocking=True)
    cond_d = dict(sty=dev(cond_h["sty"]), sty_lengths=dev(cond_h["sty_lengths"]), ref_skips=[dev(r) for r in cond_h["ref_skips"]]) \
        if cond_h else None
    x0_d, mask_d, mu_d = dev(x0), dev(inp["mask"]), dev(inp["mu"])
    gathered = torch.empty(world * B, 80, T, device="cuda") if world > 1 else None
    audio_d = win_d = fb_d = None
    if name == "C3":
        # config 3 also runs the reference-audio front-end (STFT -> mel -> log) on 3 s of synthetic audio per utterance
        from dexb200.audio.stft import slaney_mel_basis
        from dexb200.engine import stft_mel
        g = torch.Generator().manual_seed(99 + rank)
        audio_d = dev(torch.rand(B, 66150, generator=g) - 0.5)
        win_d = dev(torch.hann_window(1024, periodic=True))
        fb_d = dev(torch.from_numpy(slaney_mel_basis(22050, 1024, 80, 0.0, 8000.0)))
        # ... and the whole style stage of DeXTTS.forward (tts.py:42-50) on that mel: TIV encoder -> the loop's `ref_skips`;
        # TV encoder + LF0 encoder (synthetic log-F0 contour: pitch extraction is CPU pre-processing upstream) -> style fusion +
        # conv_sty -> the loop's `sty`
        from dexb200.model import LF0Encoder, TIVEncoder, TVEncoder, style_fusion
        from dexb200.synth import synth_conv_sty_weights, synth_lf0, synth_lf0_weights, synth_tiv_weights, synth_tv_weights
        tiv = TIVEncoder(c_in=80, c_out=64, num_layer=6, c_h=128)
        tiv.load_state_dict(synth_tiv_weights(prefix=""), strict=True)
        tiv = tiv.cuda().eval()
        tv = TVEncoder(c_in=80, c_out=192, c_out_g=192, num_layer=6, c_h=128, n_emb=512, commit_w=0.25)
        tv.load_state_dict(synth_tv_weights(prefix=""), strict=True)
        tv = tv.cuda().eval()
        lf0e = LF0Encoder(c_h=192, c_out=192, c_out_g=192, num_layer=2, c_in=1)
        lf0e.load_state_dict(synth_lf0_weights(prefix=""), strict=True)
        lf0e = lf0e.cuda().eval()
        conv_sty = torch.nn.Conv1d(192, 128, 1, 1)
        cw = synth_conv_sty_weights()
        conv_sty.load_state_dict({"weight": cw["conv_sty.weight"], "bias": cw["conv_sty.bias"]})
        conv_sty = conv_sty.cuda().eval()
        n_ref = 66150 // 256 + 1
        ref_mask_d = torch.ones(B, 1, n_ref, device="cuda")
        lf0_d = dev(synth_lf0(B, n_ref, seed=77 + rank)["lf0"])
        config["stft"] = ("every bench step runs dexb_stft_mel on (B, 66150) synthetic audio -> (B, 80, 259) log-mel -> dexb_tiv_forward "
                          "(ref_skips), dexb_tv_forward + dexb_lf0_forward (synthetic log-F0) + dexb_style_fuse (sty) -> the loop")

    def one_pass():
        cond = cond_d
        if audio_d is not None:
            mel = stft_mel(audio_d, win_d, fb_d)
            _, skips = tiv(mel, ref_mask_d)
            z_before, z_dec, _ = tv(mel, ref_mask_d)
            lf0_enc, lf0_dec = lf0e(lf0_d, ref_mask_d)
            _, sty = style_fusion(conv_sty, z_before, z_dec, ref_mask_d, lf0_enc, lf0_dec, ref_mask_d, want_sty_enc=False)
            cond = dict(cond_d, ref_skips=skips, sty=sty)
        y = eng.sample(x0_d, mask_d, mu_d, n_steps, cond=cond)
        if world > 1:
            dist.all_gather_into_tensor(gathered, y)        # the path's only collective: finished mels (SURVEY.md 8e,
        return y                                            # dexb200.parallel.gather_mels does the same for ragged shards)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        y = one_pass()
    barrier()
    assert torch.isfinite(y).all(), "non-finite mel"
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        one_pass()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per = ms / steps
    value = world * B * T / (ms_per * 1e-3)
    launches = eng.launches * steps
    if audio_d is not None:                               # C3: + the STFT kernel, the three encoders and the fusion of every step
        launches += (1 + tiv.cuda_engine().launches + tv.cuda_engine().launches + lf0e.cuda_engine().launches + 2) * steps

    # ---- e2e: the public Python API (Diffusion.forward, the module the reference calls as model.decoder) on HOST inputs:
    #      pinned host tensors -> device, loop, (gather,) mel -> pinned host tensor, everything inside the timed region
    pin = lambda t: t.contiguous().pin_memory()
    mask_p, mu_p = pin(inp["mask"]), pin(inp["mu"])
    out_p = torch.empty(B, 80, T).pin_memory()
    cond_p = dict(sty=pin(cond_h["sty"]), sty_lengths=pin(cond_h["sty_lengths"]), ref_lengths=pin(inp["ref_lengths"]),
                  ref_skips=[pin(r) for r in cond_h["ref_skips"]]) if cond_h else None
    h2d = sum(t.numel() * t.element_size() for t in (mask_p, mu_p))
    if cond_p:
        h2d += sum(t.numel() * t.element_size() for t in [cond_p["sty"], cond_p["sty_lengths"], cond_p["ref_lengths"]] + cond_p["ref_skips"])
    d2h = out_p.numel() * 4

    def e2e_pass():
        mu_ = mu_p.cuda(non_blocking=True)
        mask_ = mask_p.cuda(non_blocking=True)
        if variant == "dex":
            y_ = dec(mu_, mask_, mu_, [r.cuda(non_blocking=True) for r in cond_p["ref_skips"]], cond_p["ref_lengths"].cuda(non_blocking=True),
                     cond_p["sty"].cuda(non_blocking=True), cond_p["sty_lengths"].cuda(non_blocking=True), n_timesteps=n_steps, infer=True,
                     temperature=1.5)
        else:
            y_ = dec(mu_, mask_, mu_, n_timesteps=n_steps, infer=True, temperature=1.5)
        if world > 1:
            dist.all_gather_into_tensor(gathered, y_)
        out_p.copy_(y_, non_blocking=True)
        torch.cuda.synchronize()
        return out_p

    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_pass()
    e2e_s = (time.perf_counter() - t0) / steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * B * T / e2e_s
    clk = clocks.stop() if rank == 0 else None

    pk = peaks()
    flops_traj = algorithmic_flops(variant, T, Ts) * B * n_steps
    whole = {"algorithmic_tflop_per_step": flops_traj * 1e-12, "achieved_tflops": flops_traj * 1e-12 / (ms_per * 1e-3),
             "frac_of_sustained_bf16_peak": flops_traj * 1e-12 / (ms_per * 1e-3) / pk["tf_sust"]}
    line = {"metric": METRIC, "value": value, "unit": "mel-frames/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (split-bf16 x3 MMA)",
            "data": "synthetic", "config": config, "rtf": (ms_per * 1e-3) / (world * B * T * 256 / 22050.0),
            "e2e": {"value": e2e_val, "unit": "mel-frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "api": "dexb200.model.Diffusion.forward(infer=True) on pinned host tensors"
                                                       + (" + all_gather_into_tensor of the mels" if world > 1 else "")},
            "parity": parity, "gpu_launches": launches, "simt_fallback_gemms_per_net_call": eng.simt_fallbacks, "clocks": clk,
            "whole_step": whole, "workspace_gb": eng.workspace_bytes / 2 ** 30}
    if not full:
        del dec
        torch.cuda.empty_cache()
        return line

    # ---- dominant kernel class, timed live: CUDA events around every launch of un-graphed network calls
    roof = None
    if rank == 0:
        agg = {}
        reps = 3
        for r in range(reps):
            for tag, ms_k, gf in eng.profile_step(step=n_steps // 2):
                a = agg.setdefault(tag, [0, 0.0, 0.0])
                a[0] += 1; a[1] += ms_k; a[2] += gf
        tot = sum(a[1] for a in agg.values())
        gem = [(t, a) for t, a in agg.items() if t.startswith("gemm:")]
        g_ms, g_gf, g_n = sum(a[1] for _, a in gem), sum(a[2] for _, a in gem), sum(a[0] for _, a in gem)
        if args.profile:
            for t, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                print(f"  {t:46s} n={a[0] // reps:3d} {a[1] / reps:8.3f} ms/call-of-net {a[2] / reps:9.2f} GFLOP "
                      f"{(a[2] / a[1]) if a[1] > 0 else 0:8.1f} TFLOP/s {100 * a[1] / tot:5.1f}%", file=sys.stderr)
        ach = g_gf / g_ms if g_ms > 0 else 0.0                 # GFLOP / ms = TFLOP/s
        at = [(t, a) for t, a in agg.items() if t.startswith("attn_fwd_kernel")]
        a_ms, a_gf, a_n = sum(a[1] for _, a in at), sum(a[2] for _, a in at), sum(a[0] for _, a in at)
        ncu = {}
        for cand in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
            try:
                ncu = json.load(open(os.path.join(ROOT, "profiles", cand)))
                break
            except Exception:
                pass
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel<BLOCK_N> + conv_pair_kernel<BLOCK_N> (tcgen05 implicit GEMM: all linear contractions; 3x3 convolutions on cta_group::2 CTA pairs)",
                "achieved": ach, "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": ach / pk["tf_sust"],
                "traffic": ncu.get("gemm_tc_kernel", {}).get("avg_dram_bytes_per_launch") if name == "C2" else None,
                "traffic_source": ncu.get("source") if name == "C2" else None,
                "ncu_tensor_pipe_active_pct": ncu.get("gemm_tc_kernel", {}).get("time_weighted_tensor_pipe_active_pct"),
                "launches_per_net_call": g_n // reps, "avg_launch_ms": g_ms / max(g_n, 1), "share_of_step": g_ms / tot if tot else None,
                "peak_source": f"{pk['src']} sustained dense bf16 (MEASURED_PEAKS.json)",
                "note": f"algorithmic fp32 FLOPs; the kernel issues {args.nsplit} bf16 MMAs per product (split-bf16 for the 1e-3 parity "
                        f"bound), so the attainable fraction is <= 1/{args.nsplit}"}
        if a_ms > 0:
            a_ach = a_gf / a_ms
            roof["attention"] = {"kernel": "attn_fwd_kernel (fused one-pass attention with lazy rescaling, Q/P in tensor memory; DiT blocks + TV adaptor "
                                           "+ linear-attention contexts)",
                                 "bound": "tensor", "achieved": a_ach, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                                 "frac": a_ach / pk["tf_sust"], "launches_per_net_call": a_n // reps,
                                 "avg_launch_ms": a_ms / max(a_n, 1), "share_of_step": a_ms / tot if tot else None,
                                 "ncu_tensor_pipe_active_pct": ncu.get("attn_fwd_kernel(dit)", {}).get("tensor_pipe_active_pct"),
                                 "note": "algorithmic 4*N*Nk*d FLOPs per head; issued MMA work is 3x that (split-bf16)"}
        if audio_d is not None:                                # C3: the STFT kernel alone against the HBM roofline
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ts_ = []
            for _ in range(10):
                flush.zero_()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(); stft_mel(audio_d, win_d, fb_d); b_.record()
                torch.cuda.synchronize()
                ts_.append(a_.elapsed_time(b_))
            ts_.sort()
            nb = B * (4 * 66150 + 4 * 80 * (66150 // 256 + 1))
            roof["stft"] = {"kernel": "k_stft_mel", "bound": "hbm", "achieved": nb / (ts_[len(ts_) // 2] * 1e-3) / 1e9, "peak": pk["hbm"],
                            "unit": "GB/s", "frac": nb / (ts_[len(ts_) // 2] * 1e-3) / 1e9 / pk["hbm"], "ms": ts_[len(ts_) // 2],
                            "algorithmic_bytes": nb, "note": "latency / FFT-compute bound at this size: 11 MB of algorithmic traffic is "
                            "~2 us of HBM time; cold L2 (256 MB flush before every timed launch)"}
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        sec_step, cores, kind = cpu_reference_run(variant, B, T, Ts, n_steps, args.cpu_sample_steps, 1)
        what = ("the reference's own Diffusion.forward(infer=True) (baseline/_ref, unmodified DEX-TTS/GeDEX-TTS model package; sigma broadcast "
                "to (B,) for B > 1)") if kind == "reference" else "oracle/dex_oracle.py (port)"
        cpu = {"value": B * T / (sec_step * n_steps), "unit": "mel-frames/s", "cores": cores, "kind": kind,
               "sample": f"{args.cpu_sample_steps} of {n_steps} sampler steps of one batch (B={B}, T={T}) on the host cores, extrapolated "
                         f"x{n_steps}/{args.cpu_sample_steps}; {what}; torch CPU fp32, {cores} threads"}
    line["roofline"] = roof
    line["cpu_baseline"] = cpu
    if rank == 0 and world == 1 and name == "C2" and not args.no_pipeline:
        line["pipeline"] = pipeline_stages(B, T)
    return line


def pipeline_stages(B, T):
    """Not part of the metric: device time of the stage BEHIND the loop for the same batch -- the HiFi-GAN v1 vocoder on the tcgen05 GEMM
    engine (dexb_voc_forward, one CUDA graph) turning the B x T mel frames into B x 256 T samples -- so the line shows what the loop's
    share of mel -> waveform is.  Never fails the bench: errors are reported in the key."""
    try:
        from dexb200.hifigan.models import VocoderEngine
        from dexb200.synth import HIFIGAN_V1_CFG, synth_vocoder_weights
        eng = VocoderEngine(HIFIGAN_V1_CFG)
        eng.load_state_dict(synth_vocoder_weights())
        mel = (torch.randn(B, 80, T, generator=torch.Generator().manual_seed(3)) * 1.5 - 4.0).cuda()
        for _ in range(3):
            wav = eng.forward(mel)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            wav = eng.forward(mel)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        out = {"vocoder_ms": ms, "vocoder_launches": eng.launches, "audio_seconds": B * T * 256 / 22050.0,
               "vocoder_rtf": ms * 1e-3 / (B * T * 256 / 22050.0), "finite": bool(torch.isfinite(wav).all()),
               "what": f"hifigan.Generator (HiFi-GAN v1) on the mel of the same batch (B={B}, T={T}), seeded random weights"}
        eng.close()
        return out
    except Exception as e:                                   # pragma: no cover
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dexb200")
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--cpu-sample-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the `pipeline` key (vocoder device time for the same batch)")
    ap.add_argument("--other-workloads", default=None, help="comma list of extra workloads attached under `other_workloads` "
                                                            "(default: C4,C5 -- the north star's 8-GPU configs -- when N == 8)")
    ap.add_argument("--profile", action="store_true", help="print the per-launch breakdown of one network call to stderr")
    ap.add_argument("--gemm-engine", type=int, default=0)
    ap.add_argument("--nsplit", type=int, default=3)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    variant, B, T, Ts, n_steps = WORKLOADS[args.workload]

    # -------------------------------------------------------------------------------- reference arm: the reference's own CPU path
    if args.impl == "reference":
        if rank != 0:
            return
        wl = f"{args.workload}: {'DEX-TTS' if variant == 'dex' else 'GeDEX-TTS'} B={B}/GPU T={T} (80x{T} mel) Ts={Ts} {n_steps} sampler steps"
        config = {"workload": wl, "batch_per_gpu": B, "mel_frames": T, "sampler_steps": n_steps, "style_len": Ts,
                  "parallelism": f"dp{world}", "l2": "n/a (host)", "weights": "seeded random init of the reference architecture, zero-initialised tensors re-drawn (live)"}
        K = max(1, args.steps)
        times, kind, cores = [], "port", 1
        t_begin = time.perf_counter()
        for i in range(args.warmup + K):
            dt, cores, kind = cpu_reference_run(variant, B, T, Ts, n_steps, args.cpu_sample_steps, 1)
            if i >= args.warmup:
                times.append(dt)
            if time.perf_counter() - t_begin > 150 and times:      # keep the whole arm within a few minutes
                break
        sec_step = sum(times) / len(times)
        sec_traj = sec_step * n_steps
        val = B * T / sec_traj
        what = ("the reference's own Diffusion.forward(infer=True) from baseline/_ref (unmodified model package; timm stub; sigma broadcast to "
                "(B,) for B > 1)") if kind == "reference" else "oracle/dex_oracle.py (port of the reference path)"
        sample = (f"{args.cpu_sample_steps} of {n_steps} sampler steps of the full batch (B={B}, T={T}) per bench step, "
                  f"extrapolated x{n_steps}/{args.cpu_sample_steps}; {what}; torch CPU fp32, {cores} threads")
        port_dt, _ = port_run(variant, B, T, Ts, n_steps, args.cpu_sample_steps, 1) if kind == "reference" else (sec_step, cores)
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "mel-frames/s", "n_gpus": args.gpus, "steps": len(times),
                "warmup": args.warmup, "ms_per_step": sec_traj * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "mel-frames/s", "cores": cores, "kind": kind, "sample": sample},
                "port_cross_check": {"value": B * T / (port_dt * n_steps), "unit": "mel-frames/s", "kind": "port",
                                     "what": "oracle/dex_oracle.py on the same sample"},
                "e2e": {"value": val, "unit": "mel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # -------------------------------------------------------------------------------- CUDA arm
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = run_workload(args, args.workload, rank, world, local_rank, dist, full=True)
    others = args.other_workloads if args.other_workloads is not None else ("C4,C5" if world == 8 and args.workload == "C2" else "")
    extra = {}
    for name in [n for n in others.split(",") if n]:
        extra[name] = run_workload(args, name, rank, world, local_rank, dist, full=False)
    if rank == 0:
        if extra:
            line["other_workloads"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
