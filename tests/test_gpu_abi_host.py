"""GPU tests of the C-ABI exactly as INTEGRATION.md §2 presents it to a maintainer: plain ctypes on libfsmg.so, host numpy
buffers in, python scalars out (`fsmg_create -> fsmg_bind -> fsmg_refresh_weights -> fsmg_train_host / fsmg_eval_host /
fsmg_sample_host`), checked against the CPU oracle; plus the single-kernel parity hooks of the boundary
(`fsmg_debug_prep_tokens` against the REFERENCE's own input/target shift, `fsmg_gather_token_rows` against numpy indexing)
and the checkpoint / summary semantics of the plugin class (reference tf_model.py:28-129, lstm_baseline.py:106-111,126-131).
"""
import ctypes as C
import glob
import os
from pathlib import Path

import numpy as np
import pytest

from oracle import lstm_oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def torch_cuda(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


class fsmg_config(C.Structure):      # copied from INTEGRATION.md, not imported from the package
    _fields_ = [("vocab", C.c_int32), ("embed", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32),
                ("max_len", C.c_int32), ("max_seqs", C.c_int32), ("n_decay", C.c_int32), ("flags", C.c_int32),
                ("lr", C.c_float), ("max_grad_norm", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("reserved", C.c_float)]


class fsmg_param_info(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_int64), ("rows", C.c_int32), ("cols", C.c_int32)]


class RawHandle(object):
    """INTEGRATION.md §2, literally: ctypes + torch tensors used ONLY as device allocations."""

    def __init__(self, torch, lib_path, config, max_seqs):
        self.torch = torch
        lib = self.lib = C.CDLL(str(lib_path))
        lib.fsmg_param_count.restype = lib.fsmg_grad_count.restype = lib.fsmg_workspace_bytes.restype = C.c_int64
        lib.fsmg_last_error.restype = C.c_char_p
        for fn in (lib.fsmg_param_count, lib.fsmg_grad_count, lib.fsmg_workspace_bytes, lib.fsmg_num_params, lib.fsmg_destroy):
            fn.argtypes = [C.c_void_p]
        lib.fsmg_param_info_at.argtypes = [C.c_void_p, C.c_int, C.POINTER(fsmg_param_info)]
        lib.fsmg_bind.argtypes = [C.c_void_p] * 6 + [C.c_int64]
        lib.fsmg_refresh_weights.argtypes = [C.c_void_p, C.c_void_p]
        lib.fsmg_train_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_float), C.c_void_p]
        lib.fsmg_eval_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        lib.fsmg_sample_host.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        cfg = fsmg_config(vocab=config['input_size'], embed=config['embedding_size'], hidden=config['hidden_size'],
                          layers=config['n_layers'], max_len=config['max_len'], max_seqs=max_seqs, n_decay=config['n_decay'],
                          flags=0, lr=config['lr'], max_grad_norm=config['max_grad_norm'], beta1=0.9, beta2=0.999, eps=1e-8)
        self.h = C.c_void_p()
        assert lib.fsmg_create(C.byref(cfg), b"lstm_baseline", C.byref(self.h)) == 0
        P = lambda n, dt=torch.float32: torch.zeros(n, dtype=dt, device="cuda")
        self.params, self.grads = P(lib.fsmg_param_count(self.h)), P(lib.fsmg_grad_count(self.h))
        self.adam_m, self.adam_v = P(lib.fsmg_param_count(self.h)), P(lib.fsmg_param_count(self.h))
        self.ws = P(lib.fsmg_workspace_bytes(self.h) + 256, torch.uint8)
        vp = lambda t: C.c_void_p(t.data_ptr())
        assert lib.fsmg_bind(self.h, vp(self.params), vp(self.grads), vp(self.adam_m), vp(self.adam_v),
                             C.c_void_p((self.ws.data_ptr() + 255) // 256 * 256), C.c_int64(lib.fsmg_workspace_bytes(self.h))) == 0

    def write_params(self, named):
        for i in range(self.lib.fsmg_num_params(self.h)):
            pi = fsmg_param_info()
            assert self.lib.fsmg_param_info_at(self.h, i, C.byref(pi)) == 0
            arr = np.ascontiguousarray(named[pi.name.decode()], dtype=np.float32).reshape(-1)
            assert arr.size == pi.rows * pi.cols
            self.params[pi.offset: pi.offset + arr.size].copy_(self.torch.from_numpy(arr))
        assert self.lib.fsmg_refresh_weights(self.h, None) == 0

    def read_params(self):
        out = {}
        flat = self.params.cpu().numpy()
        for i in range(self.lib.fsmg_num_params(self.h)):
            pi = fsmg_param_info()
            self.lib.fsmg_param_info_at(self.h, i, C.byref(pi))
            out[pi.name.decode()] = flat[pi.offset: pi.offset + pi.rows * pi.cols].copy()
        return out

    def close(self):
        self.torch.cuda.synchronize()
        self.lib.fsmg_destroy(self.h)


def test_host_abi_train_eval_sample_follow_the_oracle(torch_cuda, built_lib):
    cfg = dict(name="lstm_baseline", input_size=200, embedding_size=32, hidden_size=32, n_layers=1, max_len=12, lr=5e-3,
               n_decay=10000, max_grad_norm=5)
    params = O.glorot_init(cfg, 99)
    rng = np.random.RandomState(0)
    sup, qry = O.synthetic_episode(rng, 5, 5, 4, 12, 200)
    raw = RawHandle(torch_cuda, built_lib, cfg, 45)
    raw.write_params(params)
    lib, h, T = raw.lib, raw.h, 12
    out = C.c_float()
    # LSTMBaseline.eval(episode): query set only, mean NLL + the per-token parity quantity
    query = np.ascontiguousarray(qry.reshape(-1, T), dtype=np.int32)
    nll = np.empty(query.shape, np.float32)
    assert lib.fsmg_eval_host(h, query.ctypes.data_as(C.c_void_p), query.shape[0], C.byref(out), nll.ctypes.data_as(C.c_void_p), None) == 0
    want_nll = O.per_token_nll(params, query, 200, np.float64)
    assert np.max(np.abs(nll - want_nll) / want_nll) < 1e-3
    assert abs(out.value - want_nll.mean()) < 1e-3 * want_nll.mean()
    # LSTMBaseline.train(episode): support rows then query rows, one optimizer step per call, pre-update loss back
    tokens = np.concatenate([sup.reshape(-1, T), qry.reshape(-1, T)]).astype(np.int32)
    state = O.TrainState(params, cfg, np.float64)
    for step in range(4):
        want = O.train_step(state, tokens)
        rc = lib.fsmg_train_host(h, tokens.ctypes.data_as(C.c_void_p), tokens.shape[0], C.c_int64(step), C.byref(out), None)
        assert rc == 0, lib.fsmg_last_error()
        assert abs(out.value - want) < 1e-3 * want
    got = raw.read_params()
    for k, v in state.params.items():
        # compare the 4-step UPDATE, norm-wise (Adam turns a near-zero gradient whose sign differs by rounding into a full-size step)
        upd = np.linalg.norm(v - params[k]) + 1e-12
        assert np.linalg.norm(got[k].reshape(v.shape) - v) < 0.05 * upd, k
    # LSTMBaseline.sample(support_set, num): greedy, support ignored, ids on the host
    num = 20
    ids = np.empty((2, num), np.int32)
    assert lib.fsmg_sample_host(h, 2, num, ids.ctypes.data_as(C.c_void_p), None) == 0
    assert (ids[0] == ids[1]).all()
    new_params = {k: v.reshape(np.asarray(params[k]).shape) for k, v in got.items()}
    from test_gpu_parity import assert_greedy
    assert_greedy(new_params, ids[0], O.sample_greedy(new_params, num, np.float64))
    # error behaviour: status codes + fsmg_last_error, nothing thrown across the ABI
    assert lib.fsmg_train_host(h, tokens.ctypes.data_as(C.c_void_p), 46, C.c_int64(9), C.byref(out), None) != 0
    assert b"max_seqs" in lib.fsmg_last_error()
    bad = tokens.copy()
    bad[3, 5] = 201                                              # outside [0, V]: TensorFlow would raise InvalidArgument
    assert lib.fsmg_eval_host(h, bad.ctypes.data_as(C.c_void_p), 45, C.byref(out), None, None) != 0
    assert b"outside" in lib.fsmg_last_error()
    bad[3, 5] = 200                                              # V itself (the start word) is a legal row of the tables
    assert lib.fsmg_eval_host(h, bad.ctypes.data_as(C.c_void_p), 45, C.byref(out), None, None) == 0
    raw.close()


def test_device_input_target_shift_equals_the_reference(torch_cuda):
    """prep_tokens_kernel (csrc/simt_kernels.cuh) == convert_tokens_to_input_and_target(tokens, start_word=V) of the
    UNMODIFIED reference (models/base_model.py:63-86), bit for bit, on the committed outputs of the reference
    (tests/golden/reference_shift.npz, written by running the reference in the build container)."""
    torch = torch_cuda
    from fsmg import _lib
    from fsmg.engine import Engine
    g = np.load(GOLD / "reference_shift.npz")
    i = 0
    while f"tok{i}" in g.files:
        tok, want_x, want_y, start = g[f"flat{i}"], g[f"x{i}"], g[f"y{i}"], int(g[f"start{i}"])
        n, T = tok.shape
        eng = Engine(dict(name="lstm_baseline", input_size=start, embedding_size=8, hidden_size=8, n_layers=1, max_len=T), max_seqs=n,
                     device="cuda:0")
        dev = torch.from_numpy(np.ascontiguousarray(tok, dtype=np.int32)).cuda()
        x, y = torch.empty(n * T, dtype=torch.int32, device="cuda"), torch.empty(n * T, dtype=torch.int32, device="cuda")
        _lib.check(eng.lib.fsmg_debug_prep_tokens(eng.h, dev.data_ptr(), n, x.data_ptr(), y.data_ptr(), eng._stream()))
        # the device layout is time-major [T, n]
        assert np.array_equal(x.cpu().numpy().reshape(T, n).T, want_x)
        assert np.array_equal(y.cpu().numpy().reshape(T, n).T, want_y)
        cnt = C.c_int64(-1)
        _lib.check(eng.lib.fsmg_token_range_errors(eng.h, C.byref(cnt), eng._stream()))
        assert cnt.value == 0
        eng.close()
        i += 1
    assert i == 5


def test_out_of_range_token_ids_are_reported_not_silently_clamped(torch_cuda):
    torch = torch_cuda
    from fsmg import FsmgError, _lib
    from fsmg.engine import Engine
    eng = Engine(dict(name="lstm_baseline", input_size=50, embedding_size=8, hidden_size=8, n_layers=1, max_len=6), max_seqs=4, device="cuda:0")
    eng.init_params(0)
    tok = np.full((4, 6), 7, np.int32)
    tok[2, 3] = 51
    with pytest.raises(FsmgError):
        eng.eval_host(tok)
    with pytest.raises(FsmgError):
        eng.train_host(-tok)
    # device-pointer entry point: clamped for memory safety, counted in the device flag
    dev = torch.from_numpy(tok).cuda()
    eng.forward_nll(dev)
    cnt = C.c_int64(-1)
    _lib.check(eng.lib.fsmg_token_range_errors(eng.h, C.byref(cnt), eng._stream()))
    assert cnt.value == 1
    _lib.check(eng.lib.fsmg_token_range_errors(eng.h, C.byref(cnt), eng._stream()))
    assert cnt.value == 0
    eng.close()


def test_device_gather_of_episode_rows_is_numpy_indexing(torch_cuda):
    """fsmg_gather_token_rows(corpus, ids) == corpus_host[ids] bit for bit (the rows the reference's
    get_episode copies song by song, data/episode.py:62-74), including repeated and boundary rows, odd row lengths."""
    torch = torch_cuda
    from fsmg import _lib
    lib = _lib.load()
    rng = np.random.RandomState(4)
    for n_rows, T, n_ids in ((1000, 128, 1440), (7, 5, 45), (33, 257, 9), (2, 1, 64)):
        corpus = rng.randint(0, 10001, size=(n_rows, T)).astype(np.int32)
        ids = rng.randint(0, n_rows, size=n_ids).astype(np.int32)
        ids[0], ids[-1] = n_rows - 1, 0
        d_corpus, d_ids = torch.from_numpy(corpus).cuda(), torch.from_numpy(ids).cuda()
        out = torch.full((n_ids, T), -1, dtype=torch.int32, device="cuda")
        _lib.check(lib.fsmg_gather_token_rows(d_corpus.data_ptr(), n_rows, T, d_ids.data_ptr(), n_ids, out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream))
        assert np.array_equal(out.cpu().numpy(), corpus[ids])
    # and through the sampler + engine: the staged batch of an IndexedEpisode is the host episode's rows
    from data.episode import load_sampler_from_config
    from fsmg.engine import Engine
    data = dict(dataset="synthetic_lyrics", dataset_path=".", split="train", batch_size=5, support_size=5, query_size=4, max_len=12,
                synthetic_vocab=200, synthetic_artists=9, synthetic_songs_per_artist=10, seed=5, device_episodes=True)
    samp = load_sampler_from_config(data)
    eng = Engine(dict(name="lstm_baseline", input_size=200, embedding_size=8, hidden_size=8, n_layers=1, max_len=12), max_seqs=45, device="cuda:0")
    ep = samp.get_episode()
    ids = np.concatenate([ep.support_ids.reshape(-1), ep.query_ids.reshape(-1)])
    staged = eng._stage_indexed(ep.corpus_device, ids).cpu().numpy()
    assert np.array_equal(staged, np.concatenate([ep.support.reshape(-1, 12), ep.query.reshape(-1, 12)]))
    assert np.array_equal(staged, samp.corpus_host[ids])
    eng.close()


class _Ep:
    def __init__(self, s, q):
        self.support, self.query = s, q


def _plugin_config(tmpdir, **over):
    cfg = dict(name="lstm_baseline", model_module_name="models.lstm_baseline", model_class_name="LSTMBaseline",
               input_size=200, embedding_size=32, hidden_size=32, n_layers=1, max_len=12, lr=5e-3, n_decay=10000,
               max_grad_norm=5, batch_size=5, support_size=5, query_size=4, seed=1234, checkpt_dir=str(tmpdir))
    cfg.update(over)
    return cfg


def test_summaries_use_the_reference_tags(torch_cuda, tmp_path):
    """'Train/loss' per train call and 'Eval/Avg_NLL' per eval call (reference lstm_baseline.py:106-111, 126-131), written
    under config['checkpt_dir'] (tf_model.py:84-88)."""
    pytest.importorskip("tensorboard")
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    from train.train import load_model_from_config
    model = load_model_from_config(_plugin_config(tmp_path))          # tensorboard defaults to on, like the reference
    rng = np.random.RandomState(0)
    ep = _Ep(*O.synthetic_episode(rng, 5, 5, 4, 12, 200))
    losses = [model.train(ep) for _ in range(3)]
    nlls = [model.eval(ep) for _ in range(2)]
    model._summary_writer.flush()
    acc = EventAccumulator(str(tmp_path))
    acc.Reload()
    assert set(acc.Tags()["scalars"]) == {"Train/loss", "Eval/Avg_NLL"}
    tr, ev = acc.Scalars("Train/loss"), acc.Scalars("Eval/Avg_NLL")
    assert [e.step for e in tr] == [0, 1, 2] and [e.step for e in ev] == [0, 1]
    np.testing.assert_allclose([e.value for e in tr], losses, rtol=1e-6)
    np.testing.assert_allclose([e.value for e in ev], nlls, rtol=1e-6)


def test_saver_keeps_ten_checkpoints_and_restores_optimistically(torch_cuda, tmp_path):
    """tf.train.Saver(max_to_keep=10) (reference tf_model.py:96-97) and optimistic_restore (tf_model.py:28-75): only
    variables whose name AND shape match are restored; the others keep their fresh initialisation."""
    from train.train import load_model_from_config
    cfg = _plugin_config(tmp_path, tensorboard=False)
    rng = np.random.RandomState(0)
    ep = _Ep(*O.synthetic_episode(rng, 5, 5, 4, 12, 200))
    a = load_model_from_config(cfg)
    a.recover_or_init("")
    for _ in range(13):
        a.train(ep)
        a.save(str(tmp_path))
    kept = sorted(int(p.rsplit("-", 1)[1][:-4]) for p in glob.glob(str(tmp_path / "lstm_baseline" / "lstm_baseline-*.npz")))
    assert kept == list(range(4, 14))                                  # the ten most recent
    assert 'lstm_baseline-13.npz' in open(tmp_path / "lstm_baseline" / "checkpoint").read()
    trained = a.get_params()
    # same names, different hidden size: embedding [201, 32] and softmax_b [201] match, kernel / bias / softmax_w do not
    b = load_model_from_config(_plugin_config(tmp_path, tensorboard=False, hidden_size=48))
    fresh = load_model_from_config(_plugin_config(tmp_path / "none", tensorboard=False, hidden_size=48))
    fresh.recover_or_init("")
    b.recover_or_init(str(tmp_path))
    pb, pf = b.get_params(), fresh.get_params()
    same = ("lstm_baseline/embedding", "lstm_baseline/softmax_b")
    for k in same:
        np.testing.assert_array_equal(pb[k], trained[k])
        assert not np.array_equal(pb[k], pf[k])
    for k in pb:
        if k not in same:
            assert pb[k].shape != trained[k].shape
            np.testing.assert_array_equal(pb[k], pf[k])                # untouched: the seed's Glorot draw
    assert b.global_step == 13                                         # global_step is a scalar: name and shape match
    assert np.isfinite(b.train(ep))
    # only_load_trainable_vars: weights yes, Adam slots / global_step no (tf_model.py:112-125)
    c = load_model_from_config(cfg)
    c.recover_or_init(str(tmp_path), only_load_trainable_vars=True)
    assert c.global_step == 0 and float(c.engine.adam_m.abs().max()) == 0.0
    for k, v in c.get_params().items():
        np.testing.assert_array_equal(v, trained[k])
    # recover_or_init on a model that already holds weights restores, it does not re-draw
    before = a.get_params()
    a.recover_or_init(str(tmp_path / "nowhere"))
    for k, v in a.get_params().items():
        np.testing.assert_array_equal(v, before[k])
    assert a.global_step == 13


def test_in_graph_events_do_not_change_the_step_and_fire_in_order(torch_cuda):
    """fsmg_set_stage_events / fsmg_set_loss_event: caller-owned events recorded inside forward_backward (plain launches on the
    first call, external event-record nodes of the captured graph afterwards).  The step's results must be the same with and
    without them (up to the order of the fp32 gradient REDs), the loss read back behind the loss event must be the final one,
    and the parameter ranges must tile the flat buffer in TF get_vars() order."""
    torch = torch_cuda
    from fsmg import _lib
    from fsmg.engine import Engine
    cfg = dict(name="lstm_baseline", input_size=500, embedding_size=64, hidden_size=64, n_layers=1, max_len=16, lr=5e-3, n_decay=10000,
               max_grad_norm=5)
    params = O.glorot_init(cfg, 3)
    tok = O.synthetic_tokens(np.random.RandomState(2), (45, 16), 500, "zipf")
    os.environ["FSMG_EARLY_LOSS"] = "0"
    try:
        plain = Engine(cfg, max_seqs=45, device="cuda:0")
    finally:
        os.environ.pop("FSMG_EARLY_LOSS", None)
    evented = Engine(cfg, max_seqs=45, device="cuda:0")          # default: loss event set
    assert evented._early_loss and not plain._early_loss
    ev_soft, ev_emb = torch.cuda.Event(), torch.cuda.Event()
    for ev in (ev_soft, ev_emb):
        ev.record()
    _lib.check(evented.lib.fsmg_set_stage_events(evented.h, ev_soft.cuda_event, ev_emb.cuda_event, 4))
    plain.load_params(params)
    evented.load_params(params)
    state = O.TrainState(params, cfg, np.float64)
    for step in range(5):                                          # call 1: plain launches, call 2: capture, then graph replays
        want = O.train_step(state, tok)
        a, b = plain.train_host(tok), evented.train_host(tok)
        assert abs(a - b) <= 1e-5 * abs(a) and abs(b - want) < 1e-3 * want
        # the softmax / embedding slices are final when their events have fired, and they fire before the step ends
        ev_soft.synchronize()
        ev_emb.synchronize()
    torch.cuda.synchronize()
    init = Engine(cfg, max_seqs=1, device="cuda:0")
    init.load_params(params)
    update = float((plain.params - init.params).norm())
    assert float((plain.params - evented.params).norm()) < 1e-3 * update      # same five Adam steps up to the order of the gradient REDs
    init.close()
    ranges = []
    for which in range(4):
        b, e = C.c_int64(), C.c_int64()
        _lib.check(evented.lib.fsmg_param_range(evented.h, which, C.byref(b), C.byref(e)))
        ranges.append((b.value, e.value))
    assert ranges[0][0] == 0 and ranges[0][1] == ranges[1][0] and ranges[1][1] == ranges[2][0] and ranges[2][1] == ranges[3][0]
    assert ranges[3] == (evented.n_params, evented.n_params + 8)
    names = [i["name"].rsplit("/", 1)[-1] for i in evented.infos]
    assert names == ["embedding", "kernel", "bias", "softmax_w", "softmax_b"]
    offs = [i["offset"] for i in evented.infos]
    assert offs[0] == ranges[0][0] and offs[1] == ranges[1][0] and offs[3] == ranges[2][0]
    plain.close()
    evented.close()
