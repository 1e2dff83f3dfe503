"""GPU tests at BASELINE.json's FULL sizes.  The per-token NLL of the complete 1440 x 128 batch of configs[1] is compared with
the fp32 oracle (torch-CPU restatement, oracle/torch_ref.py, forward only, a few seconds per 90 sequences); the training
step at full size rests on (a) the oracle on ONE episode at the full model dimensions and (b) size-independent properties
of the path that tie the full batch to that episode:

* sequences are independent (reference lstm_baseline.py:50-55: zero initial state per row): the NLL of an episode's
  rows does not depend on what else is in the batch;
* the loss is the plain mean of the per-token NLL (lstm_baseline.py:70-75): loss * N * T == sum(nll);
* gradients are linear in the batch: the gradient of the full batch equals the sum of the per-episode gradients
  taken at the same loss scale ("checksum of checksums") — this is also what the data-parallel all-reduce relies on;
* greedy decoding is deterministic and identical for every song (lstm_baseline.py:135-156: support set unused).
"""
import numpy as np
import pytest

from oracle import lstm_oracle as O

pytestmark = pytest.mark.gpu

EPISODE = 45        # batch_size * (support + query) sequences (5shot.yaml, lstm_baseline.yaml:15)


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _cfg(v, e, h, t):
    return dict(name="lstm_baseline", input_size=v, embedding_size=e, hidden_size=h, n_layers=1, max_len=t, lr=5e-3,
                n_decay=10000, max_grad_norm=5)


def test_cfg1_full_batch_nll_rows_independent_and_match_oracle_episode(torch_cuda):
    """configs[1]: 32 episodes x 45 sequences x 128 tokens, V=10k, E=H=512."""
    from fsmg.engine import Engine
    from oracle.torch_ref import TorchRef
    torch = torch_cuda
    cfg = _cfg(10000, 512, 512, 128)
    n = 32 * EPISODE
    params = O.glorot_init(cfg, 1234)
    tok = O.synthetic_tokens(np.random.RandomState(11), (n, 128), 10000, "zipf")
    eng = Engine(cfg, max_seqs=n, device="cuda:0")
    eng.load_params(params)
    mean_full, nll_full = eng.eval_host(tok, return_nll=True)
    assert nll_full.shape == (n, 128) and np.isfinite(nll_full).all()
    # loss == mean of the per-token NLL
    assert abs(mean_full - float(nll_full.astype(np.float64).mean())) < 1e-6 * mean_full
    # rows of episode 7 inside the full batch == the same rows evaluated alone (different group partition, same per-row arithmetic)
    sl = slice(7 * EPISODE, 8 * EPISODE)
    _, nll_ep = eng.eval_host(tok[sl], return_nll=True)
    assert np.max(np.abs(nll_ep - nll_full[sl]) / nll_full[sl]) < 1e-5
    # the oracle on that episode at the full model dimensions (fp32 torch-CPU restatement)
    ref = TorchRef(params, cfg, torch.float32).per_token_nll(tok[sl])
    assert np.max(np.abs(nll_full[sl] - ref) / ref) < 1e-3
    # permuting the batch permutes the result
    perm = np.random.RandomState(5).permutation(n)
    _, nll_perm = eng.eval_host(tok[perm], return_nll=True)
    assert np.max(np.abs(nll_perm - nll_full[perm]) / nll_full[perm]) < 1e-5
    eng.close()


def test_cfg1_full_batch_per_token_nll_matches_fp32_oracle_on_every_row(torch_cuda):
    """All 184 320 tokens of the configs[1] batch (32 episodes x 45 sequences x 128, V=10k, E=H=512): per-token NLL within
    1e-3 relative of the fp32 reference path (BASELINE.json north_star), no sampling of rows."""
    from fsmg.engine import Engine
    from oracle.torch_ref import TorchRef
    torch = torch_cuda
    cfg = _cfg(10000, 512, 512, 128)
    n = 32 * EPISODE
    params = O.glorot_init(cfg, 1234)
    tok = O.synthetic_tokens(np.random.RandomState(21), (n, 128), 10000, "zipf")
    tok[5, 40:] = 0                                  # a zero-padded sequence: pad id 0 is scored (base_loader.py:59-61)
    eng = Engine(cfg, max_seqs=n, device="cuda:0")
    eng.load_params(params)
    _, nll = eng.eval_host(tok, return_nll=True)
    ref = TorchRef(params, cfg, torch.float32)
    worst = 0.0
    for r0 in range(0, n, 2 * EPISODE):              # the oracle in slices (fp32 logits of a slice: 460 MB)
        want = ref.per_token_nll(tok[r0:r0 + 2 * EPISODE])
        worst = max(worst, float(np.max(np.abs(nll[r0:r0 + 2 * EPISODE] - want) / want)))
    assert worst < 1e-3, worst
    eng.close()


def test_cfg1_full_batch_gradient_is_sum_of_episode_gradients(torch_cuda):
    from fsmg.engine import Engine
    cfg = _cfg(10000, 512, 512, 128)
    n_ep = 32
    n = n_ep * EPISODE
    tok = O.synthetic_tokens(np.random.RandomState(12), (n, 128), 10000, "zipf")
    eng = Engine(cfg, max_seqs=n, device="cuda:0")
    eng.init_params(1234)
    eng.forward_backward(eng._stage(tok), tok.size)
    g_full = eng.grads.clone()
    acc = torch_cuda.zeros_like(g_full, dtype=torch_cuda.float64)
    for e in range(n_ep):
        eng.forward_backward(eng._stage(tok[e * EPISODE:(e + 1) * EPISODE]), tok.size)    # same loss scale: 1 / global tokens
        acc += eng.grads.double()
    # trailing scalars: sum(nll) and the per-occurrence embedding-gradient square norm are additive too
    n_p = eng.n_params
    extras_full, extras_sum = g_full[n_p:n_p + 2].double().cpu().numpy(), acc[n_p:n_p + 2].cpu().numpy()
    np.testing.assert_allclose(extras_full, extras_sum, rtol=2e-4)
    eng.grads.copy_(g_full)
    views_full = {k: v.double() for k, v in eng.param_views("grads").items()}     # .double() copies
    eng.grads.copy_(acc.float())
    views_sum = {k: v.double() for k, v in eng.param_views("grads").items()}
    for k in views_full:
        scale = float(views_sum[k].abs().max()) + 1e-20
        err = float((views_full[k] - views_sum[k]).abs().max())
        assert err < 2e-3 * scale, (k, err, scale)
    eng.close()


def test_cfg1_full_batch_training_step_loss_and_episode_oracle_step(torch_cuda):
    """The reported loss of a full-size optimizer step equals the pre-step mean NLL; and one full-dimension episode step
    (what the reference's train.py does per call) follows the oracle's loss trajectory."""
    from fsmg.engine import Engine
    from oracle.torch_ref import TorchRef
    torch = torch_cuda
    cfg = _cfg(10000, 512, 512, 128)
    n = 32 * EPISODE
    params = O.glorot_init(cfg, 1234)
    tok = O.synthetic_tokens(np.random.RandomState(13), (n, 128), 10000, "zipf")
    eng = Engine(cfg, max_seqs=n, device="cuda:0")
    eng.load_params(params)
    before = eng.eval_host(tok)
    loss = eng.train_host(tok)
    assert abs(loss - before) < 1e-5 * before
    after = eng.eval_host(tok)
    assert after < before                      # one clipped Adam step on the same batch lowers its loss
    # episode-sized steps at full model dimensions against the oracle (3 updates)
    eng.load_params(params)
    eng.global_step = 0
    eng.adam_m.zero_()
    eng.adam_v.zero_()
    ref = TorchRef(params, cfg, torch.float32)
    ep = tok[:EPISODE]
    got = [eng.train_host(ep) for _ in range(3)]
    want = [ref.train_step(ep) for _ in range(3)]
    np.testing.assert_allclose(got, want, rtol=1e-3)
    eng.close()


def test_cfg2_midi_full_dims_episode_matches_oracle(torch_cuda):
    """configs[2]: MIDI-event vocabulary (4708 ids), seq_len=256, E=H=1024, one 5-shot episode."""
    from fsmg.engine import Engine
    from oracle.torch_ref import TorchRef
    torch = torch_cuda
    cfg = _cfg(4708, 1024, 1024, 256)
    params = O.glorot_init(cfg, 1234)
    tok = O.synthetic_tokens(np.random.RandomState(14), (EPISODE, 256), 4708, "uniform")
    eng = Engine(cfg, max_seqs=EPISODE, device="cuda:0")
    eng.load_params(params)
    ref = TorchRef(params, cfg, torch.float32)
    _, nll = eng.eval_host(tok, return_nll=True)
    want = ref.per_token_nll(tok)
    assert np.max(np.abs(nll - want) / want) < 1e-3
    got = [eng.train_host(tok) for _ in range(2)]
    exp = [ref.train_step(tok) for _ in range(2)]
    np.testing.assert_allclose(got, exp, rtol=1e-3)
    eng.close()


def test_cfg4_full_size_greedy_generation_is_identical_across_songs_and_matches_oracle(torch_cuda):
    """configs[4]: 256 songs x 512 tokens, E=H=1024, V=4708, on-device greedy sampling."""
    from fsmg.engine import Engine
    cfg = _cfg(4708, 1024, 1024, 256)
    params = O.glorot_init(cfg, 1234)
    eng = Engine(cfg, max_seqs=256, device="cuda:0")
    eng.load_params(params)
    out = eng.sample_host(256, 512)
    assert out.shape == (256, 512) and out.min() >= 0 and out.max() <= 4708
    assert (out == out[0:1]).all()             # zero state + start word + argmax: every song is the same sequence
    from test_gpu_parity import assert_greedy
    assert_greedy(params, out[0], O.sample_greedy(params, 512, np.float64))   # all 512 tokens, teacher-forced past any tie
    eng.close()
