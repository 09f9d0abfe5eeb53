"""Parity metric of SURVEY.md §8(d) / BASELINE.json north_star: 1e-3 relative fp32 per mel bin."""
import torch

REL_TOL = 1e-3


def per_bin_violation(y, y_ref, mask=None):
    """max over (b, f, t) of |y - y_ref| / max(|y_ref|, RMS_t(y_ref[b, f]))  (<= REL_TOL passes).
    Element-wise relative error alone is ill-posed at zero crossings, so each bin's RMS over time floors the scale."""
    y, y_ref = y.double(), y_ref.double()
    if mask is not None:
        y, y_ref = y * mask, y_ref * mask
    rms = y_ref.pow(2).mean(dim=-1, keepdim=True).sqrt()
    scale = torch.maximum(y_ref.abs(), rms).clamp_min(1e-12)
    return float(((y - y_ref).abs() / scale).max())


def tensor_rel_err(a, b):
    """max|a-b| / RMS(b): for intermediate activations."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.pow(2).mean().sqrt().clamp_min(1e-12))
