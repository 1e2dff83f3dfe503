"""Unigram baseline (reference src/models/unigram_model.py): CPU tests of the oracle restatement, GPU parity of the CUDA
path (fsmg_unigram_step / fsmg_unigram_argmax through the reference-facing plugin class) against it."""
import numpy as np
import pytest

from oracle import lstm_oracle as O
from oracle.unigram_oracle import UnigramOracle


class _Ep:
    def __init__(self, s, q):
        self.support, self.query = s, q


def _episodes(n, vocab, max_len, seed=3, kind="zipf"):
    rng = np.random.RandomState(seed)
    return [_Ep(*O.synthetic_episode(rng, 5, 5, 4, max_len, vocab, kind)) for _ in range(n)]


def test_oracle_uniform_prior_and_counting():
    """alpha = 1 everywhere: the first loss is log(V) whatever the words are; counts grow by one per fed word
    (tokens[:, :-1] of support AND query); eval scores tokens[:, 1:] of the query only."""
    v, t = 50, 12
    eps = _episodes(3, v, t)
    m = UnigramOracle(v)
    assert abs(m.eval(eps[0].query) - np.log(v)) < 1e-6
    loss0 = m.train(eps[0].support, eps[0].query)
    assert abs(loss0 - np.log(v)) < 1e-6          # loss on the counts BEFORE the update
    fed = np.concatenate([eps[0].support.reshape(-1, t)[:, :-1], eps[0].query.reshape(-1, t)[:, :-1]]).reshape(-1)
    want = 1.0 + np.bincount(fed, minlength=v)
    assert np.array_equal(m.word_count, want.astype(np.float32))
    y = eps[1].query.reshape(-1, t)[:, 1:].reshape(-1)
    assert abs(m.eval(eps[1].query) - float(-np.mean(np.log(want[y] / want.sum())))) < 1e-6
    assert m.sample(4) == [int(np.argmax(want))] * 4


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(50, 12, 4), (10000, 50, 6), (4708, 256, 3)], ids=["tiny", "lyrics_v10k", "midi_t256"])
def test_gpu_unigram_matches_oracle(built_lib, shape):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from train.train import load_model_from_config
    v, t, n = shape
    cfg = dict(name="unigram_model", model_module_name="models.unigram_model", model_class_name="UnigramModel",
               input_size=v, max_len=t, batch_size=5, support_size=5, query_size=4, seed=1234)
    model = load_model_from_config(cfg)          # the reference's registry
    model.recover_or_init("")
    ref = UnigramOracle(v)
    eps = _episodes(n, v, t, kind="zipf" if v != 4708 else "uniform")
    for ep in eps:
        got, want = model.eval(ep), ref.eval(ep.query)
        assert abs(got - want) < 2e-6 * max(1.0, abs(want))
        got, want = model.train(ep), ref.train(ep.support, ep.query)
        assert abs(got - want) < 2e-6 * max(1.0, abs(want))
        assert np.array_equal(model.word_count, ref.word_count)          # integer counts: bit-exact
    assert model.sample(None, 5) == ref.sample(5)


@pytest.mark.gpu
def test_gpu_unigram_large_block_path_and_checkpoint(built_lib, tmp_path):
    """> 65 536 words per call takes the three-kernel path; counts and loss still match; save / recover round trip."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from models.unigram_model import UnigramModel
    v, t = 10000, 128
    cfg = dict(name="unigram_model", input_size=v, max_len=t)
    model = UnigramModel(cfg)
    model.recover_or_init("")
    ref = UnigramOracle(v)
    rng = np.random.RandomState(9)
    sup = O.synthetic_tokens(rng, (32, 25, t), v, "zipf")        # 32 episodes' worth in one call: 1440 x 127 words
    qry = O.synthetic_tokens(rng, (32, 20, t), v, "zipf")
    for _ in range(2):
        got, want = model.train(_Ep(sup, qry)), ref.train(sup, qry)
        assert abs(got - want) < 5e-6 * abs(want)
        assert np.array_equal(model.word_count, ref.word_count)
    assert abs(model.eval(_Ep(sup, qry)) - ref.eval(qry)) < 5e-6 * ref.eval(qry)
    path = model.save(str(tmp_path))
    assert path.endswith("unigram_model-0.npz")
    other = UnigramModel(cfg)
    other.recover_or_init(str(tmp_path))
    assert np.array_equal(other.word_count, ref.word_count)
    other.recover_or_init(str(tmp_path), only_load_trainable_vars=True)     # word_count is not trainable: stays at alpha
    assert np.array_equal(other.word_count, np.ones(v, np.float32))
    with pytest.raises(Exception):
        model.train(_Ep(np.full((1, 1, t), v, np.int32), qry[:1, :1]))      # id == input_size is out of range for this model
