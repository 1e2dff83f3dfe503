"""CPU tests: the C-ABI library loads and exports every symbol include/fsmg.h declares (no compute
calls without a GPU), the host-only parts of the ABI work, and the host logic (sampler, config
merge, plugin registry plumbing, data-parallel protocol over gloo) behaves like the reference."""
import ctypes as C
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import yaml

ROOT = Path(__file__).resolve().parents[1]
PKG = ROOT / "few-shot-music-generation_b200"


def test_library_exports_every_declared_symbol(built_lib):
    header = (ROOT / "include" / "fsmg.h").read_text()
    declared = set(re.findall(r"\b(fsmg_[a-z_]+)\s*\(", header))
    declared -= {"fsmg_config", "fsmg_handle", "fsmg_param_info"}
    assert len(declared) >= 20
    lib = C.CDLL(str(built_lib))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in fsmg.h but not exported"
    from fsmg import _lib
    assert set(_lib.SYMBOLS) == declared
    assert _lib.load().fsmg_abi_version() == 1


def test_host_only_abi_param_layout(built_lib):
    from fsmg import _lib
    lib = _lib.load()
    cfg = _lib.fsmg_config(vocab=333, embed=50, hidden=36, layers=2, max_len=9, max_seqs=45, n_decay=10000, flags=0,
                           lr=5e-3, max_grad_norm=5.0, beta1=0.9, beta2=0.999, eps=1e-8, reserved=0)
    h = C.c_void_p()
    assert lib.fsmg_create(C.byref(cfg), b"lstm_baseline", C.byref(h)) == 0
    names, end = [], 0
    for i in range(lib.fsmg_num_params(h)):
        pi = _lib.fsmg_param_info()
        assert lib.fsmg_param_info_at(h, i, C.byref(pi)) == 0
        names.append((pi.name.decode(), pi.rows, pi.cols))
        assert pi.offset >= end and pi.offset % 64 == 0
        end = pi.offset + pi.rows * pi.cols
    # TF get_vars() order and shapes
    assert names == [
        ("lstm_baseline/embedding", 334, 50),
        ("lstm_baseline/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/kernel", 86, 144),
        ("lstm_baseline/rnn/multi_rnn_cell/cell_0/basic_lstm_cell/bias", 144, 1),
        ("lstm_baseline/rnn/multi_rnn_cell/cell_1/basic_lstm_cell/kernel", 72, 144),
        ("lstm_baseline/rnn/multi_rnn_cell/cell_1/basic_lstm_cell/bias", 144, 1),
        ("lstm_baseline/softmax_w", 36, 334),
        ("lstm_baseline/softmax_b", 334, 1),
    ]
    assert lib.fsmg_param_count(h) >= end and lib.fsmg_grad_count(h) == lib.fsmg_param_count(h) + 8
    assert lib.fsmg_workspace_bytes(h) > 0
    # error behaviour: calls before bind fail with a message instead of crashing
    assert lib.fsmg_refresh_weights(h, None) != 0 and b"bound" in lib.fsmg_last_error()
    bad = _lib.fsmg_config(vocab=0, embed=1, hidden=1, layers=1, max_len=1, max_seqs=1)
    h2 = C.c_void_p()
    assert lib.fsmg_create(C.byref(bad), b"x", C.byref(h2)) != 0
    lib.fsmg_destroy(h)


def _plan(lib, m, n, k, split=1, narrow=0):
    from fsmg import _lib
    out = (C.c_int32 * 8)()
    _lib.check(lib.fsmg_debug_plan(m, n, k, split, narrow, out))
    return dict(zip(("bn", "cl", "grid", "m_tiles", "n_tiles", "k_splits", "streamk", "units"), list(out)))


def test_gemm_plans_the_design_rests_on(built_lib):
    """Host-side launch planning (csrc/tc_gemm.cuh tc_plan) for the shapes of BASELINE.json's configs, no device needed: which
    GEMMs get the 256 x 512 cta_group::2 pair tiles — the only tiles that carry the operand transform of the fused softmax gradient
    (DESIGN.md §5 G.c) — and how the decode step's 256-row GEMMs are cut (§5 S)."""
    from fsmg import _lib
    lib = _lib.load()
    # configs[1]: 18 432-row chunk, H = 512, V' = 10 001
    dh = _plan(lib, 18432, 512, 10001)
    assert (dh["bn"], dh["cl"], dh["grid"], dh["m_tiles"], dh["k_splits"], dh["streamk"]) == (512, 2, 144, 72, 1, 0)   # 72 whole pair tiles, one wave
    dws = _plan(lib, 10001, 512, 18432)
    assert (dws["bn"], dws["cl"], dws["grid"], dws["streamk"]) == (512, 2, 148, 1) and dws["units"] * 74 >= 40 * 288   # stream-K over 40 x 288 k blocks
    lse = _plan(lib, 18432, 10001, 512, split=0)
    assert (lse["bn"], lse["cl"], lse["grid"], lse["m_tiles"], lse["n_tiles"], lse["k_splits"]) == (256, 2, 148, 72, 40, 1)
    # one episode at the same model dimensions (what the reference trains per call) and configs[2] (H = 1024, V' = 4709): still pair tiles
    assert _plan(lib, 5760, 512, 10001)["bn"] == 512 and _plan(lib, 10001, 512, 5760)["bn"] == 512
    assert _plan(lib, 11520, 1024, 4709)["bn"] == 512 and _plan(lib, 4709, 1024, 11520)["bn"] == 512
    # small models fall back to the in-place pass: no pair-tile plan
    assert _plan(lib, 720, 64, 301)["bn"] != 512 and _plan(lib, 301, 64, 720)["bn"] != 512
    # decode step (configs[4]: 256 songs, K = 3 x 1024 split-fp16): 128-wide tiles -> plain 2-way split of K that fills the 74 pairs
    for n, tiles in ((4096, 32), (4709, 37)):
        wide, narrow = _plan(lib, 256, n, 3072), _plan(lib, 256, n, 3072, narrow=1)
        assert (narrow["bn"], narrow["n_tiles"], narrow["k_splits"], narrow["streamk"]) == (128, tiles, 2, 0)
        assert narrow["grid"] == 2 * 2 * tiles and narrow["grid"] <= 148
        assert wide["bn"] == 256 and (wide["k_splits"] >= 4 or wide["streamk"] == 1)
        # bytes of fp32 partial tiles combined with REDs: splits x 256 x N x 4
        cuts_wide = wide["k_splits"] if not wide["streamk"] else (wide["grid"] // 2) / wide["n_tiles"]
        assert 2 * 256 * n * 4 < 0.6 * cuts_wide * 256 * n * 4


def test_engine_refuses_to_run_without_cuda(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fsmg import FsmgError
    from fsmg.engine import Engine
    with pytest.raises(FsmgError):
        Engine(dict(input_size=10, embedding_size=8, hidden_size=8, n_layers=1, max_len=4), max_seqs=4)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from fsmg import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.FsmgError):
        _lib.load()


def test_product_path_never_imports_oracle():
    """oracle/ is test infrastructure: nothing shipped under the package may import or exec it."""
    files = list((PKG / "fsmg").glob("*.py")) + list((PKG / "src").rglob("*.py")) + list((PKG / "csrc").glob("*"))
    assert len(files) > 8
    for path in files:
        text = path.read_text(errors="ignore")
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), path
        assert "oracle/" not in text and "lstm_oracle" not in text and "torch_ref" not in text, path


def _merged_config(data, task, model):
    cfgdir = PKG / "src" / "config"
    cfg = yaml.safe_load(open(cfgdir / data))
    cfg.update(yaml.safe_load(open(cfgdir / task)))
    cfg.update(yaml.safe_load(open(cfgdir / model)))
    return cfg


def test_config_merge_keeps_reference_keys():
    cfg = _merged_config("lyrics.yaml", "5shot.yaml", "lstm_baseline.yaml")
    for key in ("name", "model_module_name", "model_class_name", "n_train", "n_decay", "print_every_n", "val_every_n",
                "n_val", "n_test", "n_samples", "lr", "max_grad_norm", "batch_size", "embedding_size", "n_layers",
                "hidden_size", "query_size", "support_size", "seed", "dataset", "dataset_path", "splits", "max_len"):
        assert key in cfg, key
    assert (cfg["embedding_size"], cfg["hidden_size"], cfg["batch_size"], cfg["max_len"]) == (250, 200, 5, 50)
    assert (cfg["support_size"], cfg["query_size"], cfg["seed"]) == (5, 4, 1234)
    assert cfg["model_module_name"] == "models.lstm_baseline" and cfg["model_class_name"] == "LSTMBaseline"
    b200 = _merged_config("synthetic_lyrics.yaml", "5shot.yaml", "lstm_baseline_b200_lyrics.yaml")
    assert (b200["hidden_size"], b200["embedding_size"], b200["max_len"], b200["episodes_per_step"]) == (512, 512, 128, 32)


def test_episode_sampler_shapes_determinism_and_disjointness():
    from data.episode import load_sampler_from_config
    cfg = _merged_config("synthetic_lyrics_small.yaml", "5shot.yaml", "lstm_baseline_cpu_ref.yaml")
    cfg["split"] = "train"
    s1, s2 = load_sampler_from_config(cfg), load_sampler_from_config(cfg)
    e1, e2 = s1.get_episode(), s2.get_episode()
    assert e1.support.shape == (5, 5, 32) and e1.query.shape == (5, 4, 32) and e1.support.dtype == np.int32
    np.testing.assert_array_equal(e1.support, e2.support)
    np.testing.assert_array_equal(e1.query, e2.query)
    assert s1.get_num_unique_words() == 10000 and e1.support.max() < 10000
    # support and query songs of one artist are drawn without replacement
    for b in range(5):
        rows = np.concatenate([e1.support[b], e1.query[b]])
        assert len({r.tobytes() for r in rows}) == 9
    with pytest.raises(RuntimeError):
        load_sampler_from_config({k: v for k, v in cfg.items() if k != "max_len"})


def test_npy_corpus_reads_reference_token_caches(tmp_path):
    from data.episode import load_sampler_from_config
    rng = np.random.RandomState(0)
    for a in range(12):
        d = tmp_path / f"artist{a}"
        d.mkdir()
        for s in range(10):
            np.save(d / f"song{s}.txt.16.npy", rng.randint(0, 99, size=16).astype(np.int32))
    cfg = dict(dataset="lyrics", dataset_path=str(tmp_path), max_len=16, split="train", batch_size=5, support_size=5,
               query_size=4, seed=1)
    ep = load_sampler_from_config(cfg).get_episode()
    assert ep.support.shape == (5, 5, 16) and ep.query.shape == (5, 4, 16)


def test_base_model_token_helpers_match_reference_semantics():
    from models.base_model import BaseModel, convert_tokens_to_input_and_target
    tok = np.arange(24).reshape(2, 3, 4)
    x, y = convert_tokens_to_input_and_target(tok, start_word=99)
    assert x.shape == (6, 4) and (x[:, 0] == 99).all() and (x[:, 1:] == tok.reshape(6, 4)[:, :-1]).all()
    assert (y == tok.reshape(6, 4)).all()
    x, y = convert_tokens_to_input_and_target(tok)
    assert x.shape == (6, 3) and (y == tok.reshape(6, 4)[:, 1:]).all()
    with pytest.raises(NotImplementedError):
        BaseModel({"name": "m"}).train(None)


def test_train_cli_parses_reference_flags():
    sys.path.insert(0, str(PKG / "src"))
    from train.train import build_parser
    a = build_parser().parse_args(["--data", "d.yaml", "--model", "m.yaml", "--task", "t.yaml", "--checkpt_dir", "c", "--init_dir", "i"])
    assert (a.data, a.model, a.task, a.checkpt_dir, a.init_dir) == ("d.yaml", "m.yaml", "t.yaml", "c", "i")


DP_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["FSMG_ROOT"])
from oracle import lstm_oracle as O
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
cfg = dict(name="lstm_baseline", input_size=40, embedding_size=7, hidden_size=6, n_layers=1, max_len=5, lr=5e-3, n_decay=10000, max_grad_norm=0.05)
params = O.glorot_init(cfg, 9, np.float64)
tok = O.synthetic_tokens(np.random.RandomState(3), (8, 5), 40, "uniform")
rank = dist.get_rank()
shard = tok[rank * 4:(rank + 1) * 4]
state = O.TrainState(params, cfg, np.float64)
names = sorted(params)
for step in range(3):
    x, y = O.shift_inputs(shard, 40)
    nll, _, cache = O.forward(state.params, x, y, np.float64, keep_cache=True)
    grads, occ = O.backward(state.params, cache, np.float64, loss_denominator=tok.size)
    # the engine's protocol: ONE all-reduce of [flat grads | sum nll | occ sqnorm]
    flat = torch.from_numpy(np.concatenate([grads[k].reshape(-1) for k in names] + [np.array([nll.sum(), occ])]))
    dist.all_reduce(flat)
    flat = flat.numpy(); off = 0; red = {}
    for k in names:
        n = grads[k].size; red[k] = flat[off:off + n].reshape(grads[k].shape); off += n
    O.apply_clip_adam(state, red, float(flat[-1]))
if rank == 0:
    single = O.TrainState(params, cfg, np.float64)
    for step in range(3):
        loss = O.train_step(single, tok)
    for k in names:
        np.testing.assert_allclose(state.params[k], single.params[k], rtol=1e-9, atol=1e-13)
    print("DP_OK")
dist.barrier()
'''


def test_data_parallel_protocol_world2_gloo(tmp_path):
    """world_size-2 gloo run of the engine's DP protocol (flat grads + 2 scalars, one all-reduce,
    redundant clip+Adam) with the oracle standing in for the kernels: equals the 1-process step."""
    script = tmp_path / "dp_worker.py"
    script.write_text(DP_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=port, FSMG_ROOT=str(ROOT), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "DP_OK" in outs[0]


EVAL_WORKER = r'''
import os, sys, numpy as np, torch.distributed as dist
for p in os.environ["FSMG_PATHS"].split(os.pathsep):
    sys.path.insert(0, p)
from data.episode import load_sampler_from_config
from train.train import evaluate
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["PORT"], rank=rank, world_size=2)
cfg = dict(dataset="synthetic_lyrics", dataset_path=".", split="val", batch_size=5, support_size=5, query_size=4, max_len=8,
           synthetic_vocab=300, synthetic_artists=12, synthetic_songs_per_artist=11, seed=4)
class Model:
    def __init__(self): self.seen = []
    def eval(self, ep):
        self.seen.append(int(ep.query.sum()))
        return float(ep.query.sum() % 1009) / 7.0
m = Model()
got = evaluate(m, load_sampler_from_config(dict(cfg)), 9, rank, 2, True)
assert len(m.seen) == (5 if rank == 0 else 4)          # 9 episodes sharded 5 + 4, each scored exactly once
single = Model()
want = evaluate(single, load_sampler_from_config(dict(cfg)), 9)
assert m.seen == single.seen[rank::2]                   # the SAME episodes the single-process loop scores
assert abs(got - want) < 1e-12, (got, want)
own = evaluate(Model(), load_sampler_from_config(dict(cfg, seed=4 + rank)), 9, rank, 2, False)   # per-rank streams: still 9 episodes in total
assert np.isfinite(own)
if rank == 0:
    print("EVAL_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_sharded_evaluate_equals_the_single_process_loop_world2_gloo(tmp_path):
    """train.evaluate under torchrun (reference src/train/train.py:27-33): the episodes of one evaluation are sharded over the
    ranks of a world_size-2 gloo group and the all-reduced mean equals the reference loop over the same sampler stream."""
    script = tmp_path / "eval_worker.py"
    script.write_text(EVAL_WORKER)
    port = str(31500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=port, OMP_NUM_THREADS="1",
                   FSMG_PATHS=os.pathsep.join([str(ROOT), str(PKG), str(PKG / "src")]))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "EVAL_OK" in outs[0]


def test_package_synthetic_generator_matches_oracle_generator():
    """bench.py's product arm draws its episodes from data.synthetic (no oracle import); the tests draw theirs
    from the oracle.  Same seed -> same episodes, so parity runs and bench runs see identical inputs."""
    from data import synthetic as S
    from oracle import lstm_oracle as O
    for kind, vocab in (("zipf", 10000), ("uniform", 4708)):
        a = S.synthetic_episode(np.random.RandomState(5), 5, 5, 4, 32, vocab, kind)
        b = O.synthetic_episode(np.random.RandomState(5), 5, 5, 4, 32, vocab, kind)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_device_episode_sampler_draws_the_same_episodes_as_the_host_sampler():
    """Same seed -> same artists and songs: IndexedEpisode.support/.query (materialised from the index sets) equal the
    arrays EpisodeSampler.get_episode builds; row order of the train batch is support rows then query rows."""
    from data.device_episode import DeviceEpisodeSampler
    from data.episode import load_sampler_from_config
    cfg = dict(dataset="synthetic_lyrics", dataset_path=".", split="train", batch_size=5, support_size=5, query_size=4,
               max_len=16, synthetic_vocab=300, synthetic_artists=12, synthetic_songs_per_artist=11, seed=77)
    host = load_sampler_from_config(dict(cfg))
    dev = load_sampler_from_config(dict(cfg, device_episodes=True))
    assert isinstance(dev, DeviceEpisodeSampler)
    for _ in range(5):
        a, b = host.get_episode(), dev.get_episode()
        assert b.support_ids.shape == (5, 5) and b.query_ids.shape == (5, 4)
        assert np.array_equal(a.support, b.support) and np.array_equal(a.query, b.query)
        assert np.array_equal(dev.corpus_host[b.support_ids.reshape(-1)], a.support.reshape(-1, 16))
    # wrapping an existing sampler continues its RNG stream
    wrapped = DeviceEpisodeSampler.from_sampler(host)
    nxt_host = load_sampler_from_config(dict(cfg))
    for _ in range(5):
        nxt_host.get_episode()
    a, b = nxt_host.get_episode(), wrapped.get_episode()
    assert np.array_equal(a.support, b.support) and np.array_equal(a.query, b.query)
