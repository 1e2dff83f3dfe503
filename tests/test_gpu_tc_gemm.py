"""GPU unit tests of the tcgen05/TMA/TMEM GEMM core through the C-ABI debug entry
(fsmg_debug_gemm): fp16 operands, fp32 accumulation, against torch fp32 matmul of the same
fp16-rounded inputs, for K-major (NT) and MN-major (TN-over-tokens) operands, ragged edges and
split-K."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [
    # (M, N, K)
    (128, 256, 64), (128, 128, 64), (256, 512, 512), (200, 300, 136), (45, 800, 248), (1440, 2048, 512),
    (130, 10008, 512), (2304, 512, 10008), (512, 2048, 9000), (77, 40, 72),
    # long-K plain-store GEMMs take the 256 x 512 pair tiles (two N = 256 MMAs per K step): exact, ragged N, ragged M
    (1000, 512, 4096), (300, 1000, 4104), (3000, 1024, 8192),
]


@pytest.fixture(scope="module")
def lib(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fsmg import _lib
    return _lib.load()


def run_gemm(lib, torch, m, n, k, mn, simt=0, seed=0):
    from fsmg import _lib
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.rand((k, m) if mn else (m, k), device="cuda", generator=g) - 0.5).half()
    b = (torch.rand((k, n) if mn else (n, k), device="cuda", generator=g) - 0.5).half()
    c = torch.full((m, n), float("nan"), device="cuda", dtype=torch.float32)
    _lib.check(lib.fsmg_debug_gemm(m, n, k, a.data_ptr(), b.data_ptr(), c.data_ptr(), int(mn), int(mn), simt,
                                   torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = (a.float().t() @ b.float()) if mn else (a.float() @ b.float().t())
    return c, ref


@pytest.mark.parametrize("mn", [False, True], ids=["k_major", "mn_major"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tcgen05_gemm_matches_fp32_matmul(lib, shape, mn):
    import torch
    m, n, k = shape
    if mn and (m % 8 or n % 8):
        pytest.skip("MN-major operands need leading dimensions that are multiples of 8")
    if not mn and k % 8:
        pytest.skip("K-major operands need K % 8 == 0")
    c, ref = run_gemm(lib, torch, m, n, k, mn)
    err = (c - ref).abs().max().item()
    assert torch.isfinite(c).all()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


def test_simt_gemm_matches_too(lib):
    import torch
    for mn in (False, True):
        c, ref = run_gemm(lib, torch, 200, 304, 136, mn, simt=1)
        assert (c - ref).abs().max().item() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [(0, 32, 6), (0, 128, 6), (1, 8, 6), (1, 32, 6), (2, 1, 6), (2, 2, 6), (2, 4, 3)],
                         ids=lambda v: "mode%d_p%d_w%d" % v)
@pytest.mark.parametrize("shape", [(5, 37), (300, 1000), (777, 10001)], ids=lambda s: "%dx%d" % s)
def test_softmax_grad_pass_matches_torch(lib, variant, shape):
    """dlogits = softmax(logits) - onehot(y) in place over fp16 logits + bias-gradient column sums (reference
    lstm_baseline.py:70-75 differentiated): every kernel variant against fp32 torch, ragged sizes included."""
    import torch
    from fsmg import _lib
    rows, v1 = shape
    ld = (v1 + 15) // 16 * 16
    g = torch.Generator(device="cuda").manual_seed(rows * 131 + v1)
    logits = (torch.randn(rows, ld, device="cuda", generator=g) * 3).half()
    y = torch.randint(0, v1, (rows,), device="cuda", dtype=torch.int32, generator=g)
    lse = torch.logsumexp(logits[:, :v1].float(), dim=1).contiguous()
    want = torch.softmax(logits[:, :v1].float(), dim=1)
    want[torch.arange(rows, device="cuda"), y.long()] -= 1.0
    db = torch.zeros(v1, device="cuda")
    work = logits.clone()
    mode, param, waves = variant
    _lib.check(lib.fsmg_debug_softmax_grad(rows, v1, ld, work.data_ptr(), lse.data_ptr(), y.data_ptr(), 0.5, db.data_ptr(),
                                           mode, param, waves, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = work[:, :v1].float()
    assert torch.allclose(got, want, atol=2e-3, rtol=2e-3), float((got - want).abs().max())
    assert float(work[:, v1:].float().abs().max() if ld > v1 else 0.0) == 0.0       # padding columns are zeroed
    assert torch.allclose(db, 0.5 * want.sum(0), atol=2e-3 * max(1.0, rows ** 0.5), rtol=1e-2)
