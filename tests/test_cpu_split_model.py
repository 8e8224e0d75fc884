"""CPU: the arithmetic model behind the two fp32-parity matmul modes, independent of any kernel.

x = hi + lo, products hi*hi + hi*lo + lo*hi (lo*lo dropped).  bf16 split: hi = bf16_rn(x), lo = bf16_rn(x - hi);
TF32 split: hi = x truncated to 19 bits, lo = x - hi truncated by the tensor core to 19 bits.  The bounds asserted here
are the ones quoted in csrc/clb_gemm_tc4.cu and DESIGN.md; the kernels themselves are tested on the GPU."""
import torch


def _bf16_split(x):
    hi = x.to(torch.bfloat16).to(torch.float32)
    lo = (x - hi).to(torch.bfloat16).to(torch.float32)
    return hi, lo


def _tf32_split(x):
    mask = torch.tensor(-8192, dtype=torch.int32)                      # 0xFFFFE000
    trunc = lambda t: (t.view(torch.int32) & mask).view(torch.float32)
    hi = trunc(x)
    return hi, trunc(x - hi)


def test_operand_residuals():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1 << 16, generator=g) * torch.exp(torch.randn(1 << 16, generator=g))
    hi, lo = _bf16_split(x)
    assert torch.equal((x - hi) - lo, x - hi - lo)                      # x - hi is exact in fp32
    r = ((x - hi - lo).abs() / x.abs()).max().item()
    assert r <= 2.0 ** -16                                              # two round-to-nearest steps of 8 bits each
    hi, lo = _tf32_split(x)
    r = ((x - hi - lo).abs() / x.abs()).max().item()
    assert r <= 2.0 ** -20                                              # two truncations of 11 bits each


def test_split_dot_product_error():
    """K = 4608 (VGG-11's largest reduction): the three kept products reproduce the fp32 dot product to ~1e-6 of the
    natural scale sqrt(sum (a_i b_i)^2) for both splits -- far inside north_star's 1e-4."""
    g = torch.Generator().manual_seed(1)
    worst = {"bf16": 0.0, "tf32": 0.0}
    for _ in range(64):
        a = torch.relu(torch.randn(4608, generator=g))                  # post-ReLU activations
        b = torch.randn(4608, generator=g) * (2.0 / 4608) ** 0.5        # kaiming weights
        exact = (a.double() * b.double()).sum()
        scale = (a.double() * b.double()).pow(2).sum().sqrt()
        for name, split in (("bf16", _bf16_split), ("tf32", _tf32_split)):
            ah, al = split(a)
            bh, bl = split(b)
            got = (ah.double() * bh.double()).sum() + (ah.double() * bl.double()).sum() + (al.double() * bh.double()).sum()
            worst[name] = max(worst[name], abs((got - exact) / scale).item())
    assert worst["bf16"] <= 2e-5 and worst["tf32"] <= 2e-6, worst
    assert worst["bf16"] > worst["tf32"]                                # the bf16 split is the coarser representation
