"""Direct checks of the two tcgen05 kernels through the C ABI against a plain torch fp32 gather + matmul of the SAME
bf16 operands (so the only difference is fp32 summation order): forward / dgrad gather-GEMM (`scn_conv_forward`,
csrc/conv_tc.cu) and weight gradient (`scn_conv_wgrad`, csrc/wgrad_tc.cu).  Covers the shapes the layer tests do not
reach: row counts that are not multiples of 128, a single offset, tables without any pair, Cin != Cout, the widest
reference layers (160 / 192 channels), 256 channels, several CTAs per offset.  Tolerance 2e-3 relative (north-star bar;
observed ~3e-3 of bf16 output rounding for the forward is why the forward compares against the bf16-rounded reference)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1, 32, 32), (200, 1, 64, 64), (129, 3, 64, 32), (1000, 27, 32, 32), (3000, 27, 64, 96), (2500, 27, 96, 96),
          (2000, 27, 128, 128), (1500, 27, 160, 160), (1500, 27, 192, 192), (1200, 8, 160, 192), (900, 27, 256, 256),
          (70000, 27, 64, 64), (40000, 8, 32, 64)]


def make(n, K, cin, cout, density, seed):
    from sparseeventid_b200.scn import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    n_pad = ops.pad128(n)
    nbr = torch.full((K, n_pad), -1, dtype=torch.int32, device="cuda")
    mask = torch.rand(K, n, device="cuda", generator=g) < density
    idx = torch.randint(0, n, (K, n), device="cuda", dtype=torch.int32, generator=g)
    nbr[:, :n] = torch.where(mask, idx, torch.full_like(idx, -1))
    x = torch.randn(n, cin, device="cuda", generator=g).bfloat16()
    d = torch.randn(n, cout, device="cuda", generator=g).bfloat16()
    w = (torch.randn(K, cin, cout, device="cuda", generator=g) / cin ** 0.5).bfloat16().float().contiguous()
    return nbr, x, d, w


@pytest.mark.parametrize("density", [0.3, 0.0, 1.0])
@pytest.mark.parametrize("n,K,cin,cout", SHAPES)
def test_wgrad_kernel_matches_torch(n, K, cin, cout, density):
    from sparseeventid_b200 import _lib as L
    from sparseeventid_b200.scn import ops
    if density == 1.0 and n > 5000:
        pytest.skip("dense large case adds nothing")
    nbr, x, d, _ = make(n, K, cin, cout, density, seed=n + K)
    dw = ops.conv_wgrad(x, d, nbr, n, cin, cout, L.PREC_BF16)
    ref = torch.zeros(K, cin, cout, device="cuda")
    xf, df = x.float(), d.float()
    for k in range(K):
        j = nbr[k, :n].long()
        m = j >= 0
        if bool(m.any()):
            ref[k] = xf[j[m]].t() @ df[m]
    scale = float(ref.abs().max())
    if scale == 0.0:
        assert float(dw.abs().max()) == 0.0
    else:
        assert float((dw - ref).abs().max()) <= 2e-3 * scale


@pytest.mark.parametrize("density", [0.3, 0.0])
@pytest.mark.parametrize("n,K,cin,cout", SHAPES)
def test_forward_kernel_matches_torch(n, K, cin, cout, density):
    from sparseeventid_b200 import _lib as L
    from sparseeventid_b200.scn import ops
    nbr, x, _, w = make(n, K, cin, cout, density, seed=7 * n + K)
    bias = torch.randn(cout, device="cuda")
    bp = ops.prep_weights(w, False, False, L.PREC_BF16, torch.bfloat16)
    out = ops.conv_forward(x, nbr, n, cin, cout, bp, bias, L.PREC_BF16, torch.bfloat16)
    ref = bias.expand(n, cout).clone()
    xf = x.float()
    for k in range(K):
        j = nbr[k, :n].long()
        m = j >= 0
        if bool(m.any()):
            ref[m] += xf[j[m]] @ w[k]
    ref = ref.bfloat16().float()                     # the kernel stores bf16
    scale = float(ref.abs().max())
    # one bf16 ulp (2^-8 relative to the element) where fp32 summation order flips a rounding
    assert float((out.float() - ref).abs().max()) <= 2 ** -7 * scale
    assert float((out.float() - ref).norm() / ref.norm().clamp_min(1e-20)) <= 2e-3
