"""CPU emulation of the arithmetic of the fused softmax gradient (csrc/tc_gemm.cuh "XF"): what the GPU path stores
and multiplies, restated in NumPy with the same roundings, against the fp64 softmax - onehot of the reference's loss
(lstm_baseline.py:70-75: sparse softmax cross-entropy through xw_plus_b).

    logits GEMM epilogue:   e[r, v]   = fp16( exp(z[r, v] - cmax[r, v // 16]) )        (the chunk's largest is exactly 1)
    consumer GEMM operand:  dl[r, v]  = fp16( e[r, v] * fp16(exp(cmax[r, v // 16] - lse[r])) ) - (v == y[r])     (fp16 subtract)

The test pins the error budget DESIGN.md states: every dlogits element within 2^-9 relative of the exact probability (three fp16
roundings) or 2^-24 absolute (fp16 subnormal floor), and the scheme is at least as accurate as the fp16-logits route it replaced,
whose error grows with |logit|.  No GPU, no library call: this is the specification the -m gpu tests in test_gpu_xf.py hold the
kernels to."""
import numpy as np
import pytest


def _fused_dlogits(z, y):
    rows, V = z.shape
    nc = (V + 15) // 16
    zp = np.full((rows, nc * 16), -np.inf, np.float32)
    zp[:, :V] = z
    zc = zp.reshape(rows, nc, 16)
    cmax = zc.max(axis=2)
    e = np.exp(zc - cmax[:, :, None]).astype(np.float16)
    m = z.max(axis=1, keepdims=True)
    lse = (m[:, 0] + np.log(np.exp(z.astype(np.float64) - m).sum(axis=1))).astype(np.float32)
    s = np.exp(cmax - lse[:, None]).astype(np.float16)
    dl = (e.astype(np.float32) * s.astype(np.float32)[:, :, None]).astype(np.float16).reshape(rows, nc * 16)[:, :V].copy()
    dl[np.arange(rows), y] = (dl[np.arange(rows), y].astype(np.float32) - 1.0).astype(np.float16)
    return dl, e, cmax


def _logit16_dlogits(z, y):
    """The route this replaces: fp16 logits stored, softmax - onehot formed from them in fp32, rounded to fp16."""
    z16 = z.astype(np.float16).astype(np.float32)
    m = z.max(axis=1, keepdims=True)
    lse = (m[:, 0] + np.log(np.exp(z.astype(np.float64) - m).sum(axis=1))).astype(np.float32)
    p = np.exp(z16 - lse[:, None])
    p[np.arange(z.shape[0]), y] -= 1.0
    return p.astype(np.float16)


def _exact(z, y):
    z64 = z.astype(np.float64)
    p = np.exp(z64 - z64.max(axis=1, keepdims=True))
    p /= p.sum(axis=1, keepdims=True)
    d = p.copy()
    d[np.arange(z.shape[0]), y] -= 1.0
    return p, d


@pytest.mark.parametrize("scale,shift", [(0.3, 0.0), (2.0, 0.0), (2.0, 9.0), (6.0, -4.0)],
                         ids=["init_like", "trained", "large_positive_logits", "peaked"])
def test_fused_softmax_gradient_arithmetic_is_within_its_stated_error(scale, shift):
    rng = np.random.RandomState(5)
    rows, V = 96, 1001                                  # ragged last chunk (1001 = 62 * 16 + 9)
    z = (rng.randn(rows, V) * scale + shift).astype(np.float32)
    y = rng.randint(0, V, rows)
    z[::5, y[::5]] += 10.0                              # confident rows
    dl, e, cmax = _fused_dlogits(z, y)
    p, d = _exact(z, y)
    assert np.isfinite(dl.astype(np.float32)).all()
    assert float(e.max()) == 1.0 and float(e.astype(np.float32).min()) >= 0.0
    err = np.abs(dl.astype(np.float64) - d)
    # three fp16 roundings (e, scale, product): <= 3 * 2^-11 relative of the probability, or the fp16 subnormal quantum; the one-hot
    # element adds the rounding of (p - 1) to fp16: half an ulp at magnitude <= 1
    tol = np.maximum(3.0 * 2.0 ** -11 * p, 2.0 ** -24)
    tol[np.arange(rows), y] += 2.0 ** -11
    assert (err <= tol).all(), float((err / tol).max())
    # and it is no worse than the route it replaced, whose fp16 LOGIT rounding is an absolute 2^-11 |z| in the exponent
    old = np.abs(_logit16_dlogits(z, y).astype(np.float64) - d)
    assert err.max() <= old.max() * 1.05 + 2.0 ** -24
    if abs(shift) + scale >= 6.0:
        assert err.max() < 0.6 * old.max()              # large |logit|: the stored-exponential scheme is clearly better


def test_fused_scheme_bias_gradient_column_sums_match_fp64():
    rng = np.random.RandomState(6)
    rows, V = 512, 333
    z = (rng.randn(rows, V) * 1.5).astype(np.float32)
    y = rng.randint(0, V, rows)
    dl, _, _ = _fused_dlogits(z, y)
    _, d = _exact(z, y)
    got = dl.astype(np.float32).sum(axis=0)             # the dWs transform sums the fp16 operand in fp32
    want = d.sum(axis=0)
    assert np.abs(got - want).max() < 1e-3 * max(1.0, np.abs(want).max())
