"""Unit parity of the implicit-GEMM engine (both the tcgen05 and the CUDA-core kernel) through the C ABI
(dexb_gemm_test) against a float64 CPU convolution of the same operands."""
import zlib

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# name: (nimg, H, W, K, N, taps, off, stride)
CASES = {
    "lin_64x64": (2, 2, 256, 64, 64, (1, 1), (0, 0), 1),
    "conv3_64_ragged": (2, 10, 200, 64, 64, (3, 3), (-1, -1), 1),
    "conv3_128": (1, 12, 128, 128, 128, (3, 3), (-1, -1), 1),
    "conv3_256to64": (1, 8, 136, 256, 64, (3, 3), (-1, -1), 1),
    "conv3_stride2": (2, 20, 256, 64, 64, (3, 3), (-1, -1), 2),
    "lin_qkv": (1, 1, 300, 256, 768, (1, 1), (0, 0), 1),
    "lin_fc2": (1, 1, 520, 512, 256, (1, 1), (0, 0), 1),
    "taps2x2_n32": (2, 9, 70, 64, 32, (2, 2), (-1, 0), 1),
    "lin_n260": (1, 3, 130, 128, 260, (1, 1), (0, 0), 1),
    # several tiles per persistent CTA (more tiles than SMs): ring / accumulator hand-over between tiles, both MMA issuers
    "conv3_64_many": (4, 80, 256, 64, 64, (3, 3), (-1, -1), 1),
    "conv3_128_many": (3, 40, 256, 128, 128, (3, 3), (-1, -1), 1),
    "conv3_256to64_many": (2, 40, 384, 256, 64, (3, 3), (-1, -1), 1),
    "lin_256_many": (1, 1, 40000, 256, 256, (1, 1), (0, 0), 1),
    "lin_k64_many": (3, 64, 256, 64, 128, (1, 1), (0, 0), 1),
    "conv3_stride2_many": (4, 80, 512, 64, 64, (3, 3), (-1, -1), 2),
    # an odd number of 128-pixel tiles: the 3x3 convolution stays on single CTAs (halo mode of gemm_tc_kernel); every even case above
    # runs on CTA pairs (conv_pair_kernel)
    "conv3_64_odd": (1, 5, 128, 64, 64, (3, 3), (-1, -1), 1),
    "conv3_128_odd": (3, 7, 100, 128, 128, (3, 3), (-1, -1), 1),
    # pairs whose two tiles lie in different images (one tile per image row block) and n-tiles > 1
    "conv3_256_pairs_across_images": (6, 1, 128, 128, 256, (3, 3), (-1, -1), 1),
}


def reference(a, w, bias, taps, off, stride):
    """out[img, oy, ox, n] = sum_{ty,tx,k} a[img, oy*s + ty + offH, ox*s + tx + offW, k] * w[ty*KW+tx, n, k] + bias[n]."""
    nimg, H, W, K = a.shape
    KH, KW = taps
    N = w.shape[1]
    x = a.double().permute(0, 3, 1, 2)                                    # NCHW
    wt = w.double().reshape(KH, KW, N, K).permute(2, 3, 0, 1)             # (N, K, KH, KW)
    oh, ow = (H + stride - 1) // stride, (W + stride - 1) // stride
    # pad so that tap (0,0) of output (0,0) reads input (offH, offW)
    pt, pl = -off[0], -off[1]
    pb = max(0, (oh - 1) * stride + KH - 1 + off[0] - (H - 1))
    pr = max(0, (ow - 1) * stride + KW - 1 + off[1] - (W - 1))
    xp = F.pad(x, (pl, pr, pt, pb))
    y = F.conv2d(xp, wt, bias.double(), stride=stride)[:, :, :oh, :ow]
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("engine", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("name", list(CASES))
def test_gemm_engine(name, engine):
    from dexb200.engine import gemm_test
    nimg, H, W, K, N, taps, off, stride = CASES[name]
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    a = torch.randn(nimg, H, W, K, generator=g)
    w = torch.randn(taps[0] * taps[1], N, K, generator=g) / (K * taps[0] * taps[1]) ** 0.5
    b = torch.randn(N, generator=g)
    ref = reference(a, w, b, taps, off, stride)
    out = gemm_test(a.cuda(), w.cuda(), b.cuda(), taps=taps, off=off, in_stride=stride, engine=engine, nsplit=3).cpu().double()
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item() / ref.pow(2).mean().sqrt().item()
    print(f"gemm {name} engine {engine}: max err / rms = {err:.3e}")
    # bf16x3 operands carry ~16 mantissa bits (1.3e-5 max/rms for these shapes in exact arithmetic); the tensor core's fp32
    # accumulation adds a K-dependent term on top, so the bound is 6e-5 (plain bf16 lands at ~4e-3, tf32 at ~5e-4)
    assert err < 6e-5, f"{name} engine {engine}: max err / rms = {err:.3e}"


def test_gemm_single_split_is_bf16():
    """nsplit = 1 (plain bf16 operands) must land at bf16 accuracy -- checks that the hi/lo paths are really distinct."""
    from dexb200.engine import gemm_test
    nimg, H, W, K, N, taps, off, stride = CASES["lin_64x64"]
    g = torch.Generator().manual_seed(5)
    a = torch.randn(nimg, H, W, K, generator=g)
    w = torch.randn(1, N, K, generator=g) / K ** 0.5
    b = torch.zeros(N)
    ref = reference(a, w, b, taps, off, stride)
    for engine in (1, 0):
        out = gemm_test(a.cuda(), w.cuda(), b.cuda(), engine=engine, nsplit=1).cpu().double()
        err = (out - ref).abs().max().item() / ref.pow(2).mean().sqrt().item()
        if engine == 0:
            assert 1e-4 < err < 5e-2, err
