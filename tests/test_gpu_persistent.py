"""GPU tests of the persistent-RNN kernels (W_hh resident in shared memory, one launch for all T
steps) against (a) the per-step route of the same engine and (b) the fp64 oracle."""
import numpy as np
import pytest

from oracle import lstm_oracle as O

pytestmark = pytest.mark.gpu

FLAG_PER_STEP = 2  # FSMG_FLAG_SIMT_RECURRENT: per-step launches, tcgen05 GEMMs


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def engines(cfg, n):
    from fsmg.engine import Engine
    a = Engine(cfg, max_seqs=n, device="cuda:0", flags=0)
    b = Engine(cfg, max_seqs=n, device="cuda:0", flags=FLAG_PER_STEP)
    params = O.glorot_init(cfg, 21)
    a.load_params(params)
    b.load_params(params)
    return a, b, params


CASES = [
    # (N, T, E, H, layers)          what it exercises
    (45, 6, 64, 64, 1),             # U=32, several small groups
    (45, 5, 128, 128, 2),           # two layers
    (300, 4, 96, 512, 1),           # H=512: 16 CTAs/group, 128 KB resident slice, MT=1
    (1440, 3, 64, 512, 1),          # BASELINE configs[1] batch: 9 groups x 160 rows, MT=2
    (1000, 3, 64, 512, 1),          # ragged: 8 full groups of 112 rows + one of 104, sub-groups of 32 / 24 rows per CTA
    (2400, 3, 64, 512, 1),          # larger than one launch can hold: batch slicing
    (45, 4, 64, 1024, 1),           # H=1024: U=16, 64 CTAs per group
    (300, 4, 64, 128, 2),           # N*T > V': layer 0 per word (pre-activation table + segment-sum gradients) under a second layer
    (300, 4, 100, 128, 1),          # E % 8 != 0: table forward, per-token gradient path backward
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "N%d_T%d_E%d_H%d_L%d" % c)
def test_persistent_matches_per_step_route_and_oracle(torch_cuda, case):
    n, t, e, h, layers = case
    cfg = dict(name="lstm_baseline", input_size=500, embedding_size=e, hidden_size=h, n_layers=layers, max_len=t,
               lr=5e-3, n_decay=10000, max_grad_norm=5)
    a, b, params = engines(cfg, n)
    tok = O.synthetic_tokens(np.random.RandomState(3), (n, t), 500, "zipf")
    _, nll_a = a.eval_host(tok, return_nll=True)
    _, nll_b = b.eval_host(tok, return_nll=True)
    assert np.max(np.abs(nll_a - nll_b) / nll_b) < 2e-4
    if n <= 300:
        ref = O.per_token_nll(params, tok, 500, np.float64)
        assert np.max(np.abs(nll_a - ref) / ref) < 1e-3
    # backward: identical gradient buffers up to rounding
    for eng in (a, b):
        eng.forward_backward(eng._stage(tok), tok.size)
    ga, gb = a.export("grads"), b.export("grads")
    for k in ga:
        scale = np.abs(gb[k]).max() + 1e-12
        assert np.abs(ga[k] - gb[k]).max() < 5e-3 * scale, k
    ea, eb = a.grads[a.n_params:a.n_params + 2].cpu().numpy(), b.grads[b.n_params:b.n_params + 2].cpu().numpy()
    np.testing.assert_allclose(ea, eb, rtol=2e-3)
    # a few optimizer steps stay together
    la = [a.train_host(tok) for _ in range(3)]
    lb = [b.train_host(tok) for _ in range(3)]
    np.testing.assert_allclose(la, lb, rtol=5e-4)
    a.close()
    b.close()


def test_persistent_route_is_one_launch_per_direction(torch_cuda):
    cfg = dict(name="lstm_baseline", input_size=500, embedding_size=64, hidden_size=128, n_layers=1, max_len=16)
    a, b, _ = engines(cfg, 45)
    tok = O.synthetic_tokens(np.random.RandomState(3), (45, 16), 500, "zipf")
    a.forward_backward(a._stage(tok), tok.size)
    b.forward_backward(b._stage(tok), tok.size)
    assert a.last_launch_count() + 2 * (16 - 1) <= b.last_launch_count()
