"""GPU tests of the fused softmax gradient (csrc/tc_gemm.cuh, "XF"): the logits GEMM stores e = exp(logit - chunk max) and the two
projection-backward GEMMs rebuild dlogits = softmax - onehot (reference lstm_baseline.py:70-75 + tf.gradients through
sequence_loss / xw_plus_b) on their A operand in shared memory.

* the transform GEMMs on their own (fsmg_debug_gemm_xf) against a torch fp32 contraction of the same fp16-rounded dlogits,
  K-major (dH) and MN-major (dWs^T + bias gradient), ragged edges, split-K / stream-K plans, two N tiles;
* the engine with FSMG_FUSED_SG=1 against FSMG_FUSED_SG=0 (the in-place HBM pass) on the same batch at BASELINE configs[1]
  dimensions: same loss, same gradients within the fp16 rounding of dlogits.
"""
import math
import os

import numpy as np
import pytest

from oracle import lstm_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fsmg import _lib
    return _lib.load()


def _case(torch, rows, V, seed):
    """Random logits with a few peaked rows -> (E fp16 [rows, Vp] NaN-padded, cmaxT fp32 [n_c16, rows_pad], lse, y, dl fp16)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(rows, V, generator=g, device="cuda") * 2.0
    y = torch.randint(0, V, (rows,), generator=g, device="cuda", dtype=torch.int32)
    peaked = torch.arange(0, rows, 7, device="cuda")
    z[peaked, y[peaked].long()] += 12.0                      # targets with p ~ 1 and everything else tiny
    z[torch.arange(3, rows, 11, device="cuda"), 5] -= 40.0   # a logit far below its chunk maximum (e underflows to 0)
    nc = (V + 15) // 16
    zp = torch.full((rows, nc * 16), float("-inf"), device="cuda")
    zp[:, :V] = z
    cmax = zp.view(rows, nc, 16).max(dim=2).values                                   # [rows, nc]
    e = torch.exp(zp.view(rows, nc, 16) - cmax[:, :, None]).view(rows, nc * 16)[:, :V].half()
    lse = torch.logsumexp(z.double(), dim=1).float()
    s = torch.exp(cmax - lse[:, None]).half()                                        # per (row, chunk) scale, fp16 like the kernel's
    dl = e * s.repeat_interleave(16, dim=1)[:, :V]                                   # fp16 product, rounded once
    dl[torch.arange(rows, device="cuda"), y.long()] -= 1.0
    Vp = (V + 7) // 8 * 8
    E = torch.full((rows, Vp), float("nan"), device="cuda", dtype=torch.float16)
    E[:, :V] = e
    rows_pad = (rows + 63) // 64 * 64
    cmaxT = torch.full((nc, rows_pad), float("nan"), device="cuda")
    cmaxT[:, :rows] = cmax.t()
    return E, cmaxT, lse, y, dl


@pytest.mark.parametrize("shape", [(300, 512, 4708), (5760, 512, 10001), (1000, 1024, 5000), (20000, 512, 4200), (4500, 448, 4104)],
                         ids=lambda s: "x".join(map(str, s)))
def test_dh_gemm_rebuilds_softmax_gradient_on_k_major_operand(lib, shape):
    import torch
    from fsmg import _lib
    rows, n, V = shape
    E, cmaxT, lse, y, dl = _case(torch, rows, V, seed=rows + V)
    g = torch.Generator(device="cuda").manual_seed(1)
    Kp = E.shape[1]
    B = torch.zeros((n, Kp), device="cuda", dtype=torch.float16)
    B[:, :V] = (torch.rand((n, V), device="cuda", generator=g) - 0.5).half()
    C = torch.full((rows, n), float("nan"), device="cuda")
    _lib.check(lib.fsmg_debug_gemm_xf(rows, n, V, E.data_ptr(), Kp, B.data_ptr(), Kp, C.data_ptr(), 0, cmaxT.data_ptr(),
                                      cmaxT.shape[1], lse.data_ptr(), y.data_ptr(), 1.0, 0, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = dl.float() @ B[:, :V].float().t()
    assert torch.isfinite(C).all()
    err = (C - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("shape", [(4708, 512, 4200), (10001, 512, 5760), (4708, 1024, 11520), (10001, 512, 18432), (4104, 448, 4500)],
                         ids=lambda s: "x".join(map(str, s)))
def test_dws_gemm_rebuilds_softmax_gradient_on_mn_major_operand_and_sums_the_bias_gradient(lib, shape):
    import torch
    from fsmg import _lib
    V, n, rows = shape
    E, cmaxT, lse, y, dl = _case(torch, rows, V, seed=rows + V + 1)
    g = torch.Generator(device="cuda").manual_seed(2)
    B = (torch.rand((rows, n), device="cuda", generator=g) - 0.5).half()          # hs [tokens, H]
    C = torch.zeros((V, n), device="cuda")
    db = torch.zeros(V + 8, device="cuda")
    alpha = 0.125
    _lib.check(lib.fsmg_debug_gemm_xf(V, n, rows, E.data_ptr(), E.shape[1], B.data_ptr(), n, C.data_ptr(), 1, cmaxT.data_ptr(),
                                      cmaxT.shape[1], lse.data_ptr(), y.data_ptr(), alpha, db.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = alpha * (dl.float().t() @ B.float())
    assert torch.isfinite(C).all() and torch.isfinite(db).all()
    err = (C - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), err
    db_ref = alpha * dl.float().sum(dim=0)
    assert (db[V:] == 0).all()
    err_b = (db[:V] - db_ref).abs().max().item()
    assert err_b < 1e-3 * max(1.0, db_ref.abs().max().item()), err_b


def _grads(torch, cfg, tok, fused, seed=1234):
    from fsmg.engine import Engine
    old = os.environ.get("FSMG_FUSED_SG")
    os.environ["FSMG_FUSED_SG"] = "1" if fused else "0"
    try:
        eng = Engine(cfg, max_seqs=tok.shape[0], device="cuda:0")
    finally:
        if old is None:
            del os.environ["FSMG_FUSED_SG"]
        else:
            os.environ["FSMG_FUSED_SG"] = old
    eng.init_params(seed)
    eng.forward_backward(eng._stage(tok), tok.size)
    launches = eng.last_launch_count()
    n_p = eng.n_params
    out = {k: v.double().clone() for k, v in eng.param_views("grads").items()}
    extras = eng.grads[n_p:n_p + 2].double().cpu().numpy()
    loss = eng.train_host(tok)              # reported loss = mean NLL before the update
    eng.close()
    return loss, out, extras, launches


@pytest.mark.parametrize("n_ep,dims", [(1, (10000, 512, 512, 128)), (32, (10000, 512, 512, 128)), (1, (4708, 1024, 1024, 256))],
                         ids=["cfg1_episode", "cfg1_full_batch", "cfg2_midi_episode"])
def test_engine_fused_softmax_gradient_equals_the_hbm_pass(lib, n_ep, dims):
    import torch
    v, e, h, t = dims
    cfg = dict(name="lstm_baseline", input_size=v, embedding_size=e, hidden_size=h, n_layers=1, max_len=t, lr=5e-3,
               n_decay=10000, max_grad_norm=5)
    tok = O.synthetic_tokens(np.random.RandomState(31), (n_ep * 45, t), v, "zipf")
    loss_f, g_f, x_f, l_f = _grads(torch, cfg, tok, True)
    loss_u, g_u, x_u, l_u = _grads(torch, cfg, tok, False)
    assert l_f < l_u, (l_f, l_u)            # the fused route really ran: one launch less per chunk
    assert abs(loss_f - loss_u) < 2e-4 * abs(loss_u), (loss_f, loss_u)
    np.testing.assert_allclose(x_f, x_u, rtol=2e-3)
    for k in g_u:
        scale = float(g_u[k].abs().max()) + 1e-20
        err = float((g_f[k] - g_u[k]).abs().max())
        assert math.isfinite(err) and err < 2e-3 * scale, (k, err, scale)
