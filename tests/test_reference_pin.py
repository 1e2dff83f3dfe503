"""Parity pins that touch the REFERENCE ITSELF (CPU, no GPU): every part of the reference's path that runs without TensorFlow
is executed unmodified (tests/refpin.py) and compared bit-exactly with this repo's host code and with the oracle.

Two layers:
* ``*_live``  — import /root/reference here and now (build container only; skipped where it does not exist);
* ``*_golden`` — the committed outputs of the same reference calls (tests/golden/reference_*, written by
  tests/golden/make_reference_golden.py), which travel to the GPU box.

What stays unpinned: the TensorFlow graph arithmetic (src/models/lstm_baseline.py:38-87) and the two loss goldens of
src/train/test_seed.py:45-65, which need TF 1.x, the Google-Drive datasets and TF's initialiser RNG stream.
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest

import refpin
from data import dataset_meta as DM
from data import midi_events as M
from data.lyrics_vocab import LyricsVocab
from models import base_model as repo_base_model
from oracle import lstm_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
live = pytest.mark.skipif(not refpin.available(), reason="reference sources not present (GPU box)")


# ---------------------------------------------------------------------------------------------------------------------
# a1: flatten_first_two_dims / convert_tokens_to_input_and_target  (reference models/base_model.py:57-86)
# ---------------------------------------------------------------------------------------------------------------------
def _shift_cases():
    g = np.load(GOLD / "reference_shift.npz")
    i = 0
    while f"tok{i}" in g.files:
        yield {k: g[f"{k}{i}"] for k in ("tok", "x", "y", "xn", "yn", "flat", "start")}
        i += 1


def test_input_target_shift_golden():
    n = 0
    for c in _shift_cases():
        start = int(c["start"])
        assert np.array_equal(repo_base_model.flatten_first_two_dims(c["tok"]), c["flat"])
        x, y = repo_base_model.convert_tokens_to_input_and_target(c["tok"], start)
        assert np.array_equal(x, c["x"]) and np.array_equal(y, c["y"])
        x, y = repo_base_model.convert_tokens_to_input_and_target(c["tok"])
        assert np.array_equal(x, c["xn"]) and np.array_equal(y, c["yn"])
        # the oracle's restatement (what every NLL parity test feeds) and the flattening the plugin class uses
        ox, oy = O.shift_inputs(c["flat"], start)
        assert np.array_equal(ox, c["x"]) and np.array_equal(oy, c["y"])
        n += 1
    assert n == 5


@live
def test_input_target_shift_live():
    ref = refpin.load()
    rng = np.random.RandomState(0)
    for _ in range(50):
        b, s, t = rng.randint(1, 6), rng.randint(1, 10), rng.randint(1, 40)
        v = int(rng.randint(2, 20000))
        tok = rng.randint(0, v, size=(b, s, t)).astype(np.int32)
        rx, ry = ref.base_model.convert_tokens_to_input_and_target(tok, start_word=v)
        x, y = repo_base_model.convert_tokens_to_input_and_target(tok, v)
        assert np.array_equal(x, rx) and np.array_equal(y, ry)
        ox, oy = O.shift_inputs(ref.base_model.flatten_first_two_dims(tok), v)
        assert np.array_equal(ox, rx) and np.array_equal(oy, ry)
        rx, ry = ref.base_model.convert_tokens_to_input_and_target(tok)
        x, y = repo_base_model.convert_tokens_to_input_and_target(tok)
        assert np.array_equal(x, rx) and np.array_equal(y, ry)


@live
def test_plugin_token_assembly_equals_reference_feed_live():
    """LSTMBaseline.train feeds concat(flatten(support), flatten(query)) shifted with start word V
    (reference lstm_baseline.py:91-103); the repo assembles the same rows (src/models/lstm_baseline.py _train_tokens) and
    shifts on the device.  Host-side equality against the reference's own helpers."""
    ref = refpin.load()
    rng = np.random.RandomState(1)
    sup, qry = O.synthetic_episode(rng, 5, 5, 4, 32, 10000)
    X, Y = ref.base_model.convert_tokens_to_input_and_target(sup, 10000)
    Xq, Yq = ref.base_model.convert_tokens_to_input_and_target(qry, 10000)
    want_x, want_y = np.concatenate([X, Xq]), np.concatenate([Y, Yq])
    tokens = O.episode_train_tokens(sup, qry)
    ox, oy = O.shift_inputs(tokens, 10000)
    assert np.array_equal(ox, want_x) and np.array_equal(oy, want_y)


# ---------------------------------------------------------------------------------------------------------------------
# f-4: MIDI event pipeline  (reference data/midi_loader.py:62-399)
# ---------------------------------------------------------------------------------------------------------------------
def _decoded_rows(tokens):
    return [[n.family * 8, n.start, n.end, n.pitch, n.velocity] for n in M.notes_from_tokens(tokens)]


def test_midi_tokenizer_golden():
    """160 seeded random multi-track songs (pedals, drums, same-family clashes, family 16): ids bit-identical to the
    reference's MIDILoader.tokenize, decoded notes identical to its detokenize."""
    blob = json.loads((GOLD / "reference_midi.json").read_text())
    rng = np.random.RandomState(blob["seed"])
    import sys
    sys.path.insert(0, str(GOLD))
    from make_reference_golden import song_digest
    n_tokens = 0
    for case in blob["cases"]:
        song = refpin.random_song(rng)
        assert song_digest(song) == case["digest"], "random_song changed: regenerate tests/golden/reference_midi.json"
        notes, ccs = refpin.to_repo_notes(song)
        got = M.tokenize_notes(notes, ccs)
        assert got == case["tokens"]
        assert _decoded_rows(case["tokens"][:400]) == case["decoded"]
        n_tokens += len(got)
    assert len(blob["cases"]) == 160 and n_tokens > 10000
    assert max(max(c["tokens"]) for c in blob["cases"] if c["tokens"]) < M.NUM_TOKENS


@live
def test_midi_tokenizer_live_400_random_songs():
    ref = refpin.load()
    assert ref.midi_loader.MIDILoader(50).get_num_tokens() == M.NUM_TOKENS == 4708
    rng = np.random.RandomState(77)
    for _ in range(400):
        song = refpin.random_song(rng, max_tracks=6, max_notes=60)
        want = refpin.reference_tokenize(ref, song)
        notes, ccs = refpin.to_repo_notes(song)
        assert M.tokenize_notes(notes, ccs) == want
        midi = ref.midi_loader.MIDILoader(50).detokenize(np.asarray(want[:300], dtype=np.int64))
        dec = [[inst.program, n.start, n.end, n.pitch, n.velocity] for inst in midi.instruments for n in inst.notes]
        assert _decoded_rows(want[:300]) == dec


@live
def test_token_cache_contract_live(tmp_path):
    """base_loader.py:52-64: '<song>.<max_len>.npy' holds max_len ids, zero-padded / truncated; the repo's corpus reader
    returns exactly those rows."""
    ref = refpin.load()
    from data.episode import load_npy_corpus
    rng = np.random.RandomState(5)

    class Loader(ref.midi_loader.MIDILoader):
        def read(self, path):           # file parsing is the one thing the stub cannot do
            return refpin.to_reference_midi(ref, songs[os.path.basename(path)])

    songs, want = {}, {}
    loader = Loader(max_len=24)
    for a in range(3):
        os.makedirs(tmp_path / ("artist%d" % a))
        for s in range(4):
            name = "a%d_s%d.mid" % (a, s)
            songs[name] = refpin.random_song(rng, max_notes=5 if s == 0 else 40)
            want[(a, name)] = loader.load(str(tmp_path / ("artist%d" % a) / name))      # tokenises and persists the cache
    corpus = load_npy_corpus(str(tmp_path), 24, 4, "train", props=(1, 0, 0), seed=0, dataset="midi")
    assert len(corpus) == 3
    for name, rows in zip(corpus.artist_names, corpus.artists):
        a = int(name[-1])
        for s, row in enumerate(rows):
            ref_row = want[(a, "a%d_s%d.mid" % (a, s))]
            assert row.dtype == np.int32 and np.array_equal(row, ref_row) and len(row) == 24


# ---------------------------------------------------------------------------------------------------------------------
# f-4: lyrics word ids  (reference data/lyrics_loader.py:36-95)
# ---------------------------------------------------------------------------------------------------------------------
def test_lyrics_vocab_golden(tmp_path):
    g = json.loads((GOLD / "reference_lyrics.json").read_text())
    path = tmp_path / "word_ids.csv"
    v = LyricsVocab(str(path), tokenizer=lambda text: text.split())
    assert [v.tokenize(t) for t in g["texts"]] == g["ids"]
    assert path.read_text() == g["word_ids_csv"]
    assert [v.detokenize(i) for i in g["ids"]] == g["detokenized"]
    assert LyricsVocab(str(path)).get_num_tokens() == g["num_tokens"]


@live
def test_lyrics_vocab_live(tmp_path):
    ref = refpin.load()
    rng = np.random.RandomState(9)
    words = ["love", "n't", "'s", ",", ".", "!", "do", "baby", "'cause", "(", ")", "yeah", "``", "''", "I", "-", "a,b", "x"]
    meta = ref.dataset.Metadata(str(tmp_path), "few_shot_metadata_lyrics_50")
    loader = ref.lyrics_loader.LyricsLoader(50, metadata=meta, tokenizer=lambda t: t.split())
    mine = LyricsVocab(str(tmp_path / "mine.csv"), tokenizer=lambda t: t.split())
    for _ in range(40):
        text = " ".join(rng.choice(words, size=rng.randint(0, 30)))
        ids = loader.tokenize(text)
        assert mine.tokenize(text) == ids
        assert mine.detokenize(ids) == loader.detokenize(ids)
    meta.close()
    assert (tmp_path / "mine.csv").read_text() == open(os.path.join(meta.dir, "word_ids.csv")).read()
    assert mine.get_num_tokens() == loader.get_num_tokens()
    assert DM.highest_word_id(DM.Metadata(str(tmp_path), "few_shot_metadata_lyrics_50", create=False)) + 1 == loader.get_num_tokens()


# ---------------------------------------------------------------------------------------------------------------------
# f-4: split persistence and valid_songs.csv  (reference data/dataset.py:22-232)
# ---------------------------------------------------------------------------------------------------------------------
def test_dataset_files_golden(tmp_path):
    g = json.loads((GOLD / "reference_dataset.json").read_text())
    for ci, case in enumerate(g["cases"]):
        # the split rule, given the artists in the order the reference collected them
        eligible = [a for a in case["listdir"] if len(g["artists"][a]) >= g["min_songs"]]
        splits = DM.split_artists(eligible, case["proportions"], case["seed"])
        for split in DM.SPLITS:
            assert splits[split] == list(case["splits"][split].keys())
        # writing: same bytes as the reference's files
        meta = DM.Metadata(str(tmp_path), "w%d" % ci)
        DM.write_splits(meta, splits)
        for split in DM.SPLITS:
            assert open(meta.path(split + ".csv")).read() == case["files"][split + ".csv"]
        for line in case["files"]["valid_songs.csv"].splitlines():
            a, s = line.split(",", 1)
            DM.append_valid_song(meta, DM.unquote(a), DM.unquote(s))
        meta.close()
        assert open(meta.path("valid_songs.csv")).read() == case["files"]["valid_songs.csv"]
        # reading: the reference's files give back its artists and songs
        rmeta = DM.Metadata(str(tmp_path), "r%d" % ci)
        for name, content in case["files"].items():
            open(rmeta.path(name), "w").write(content)
        valid = DM.read_valid_songs(rmeta)
        for split in DM.SPLITS:
            assert DM.read_split(rmeta, split) == list(case["splits"][split].keys())
            for artist, songs in case["splits"][split].items():
                assert sorted(valid[artist]) == songs
        assert "bad one.mid" not in valid["tool"] and "few" in valid       # below min_songs: validated, but in no split
        assert DM.read_split(DM.Metadata(str(tmp_path), "none%d" % ci), "train") is None


@live
def test_corpus_prepared_by_the_reference_is_consumed_with_its_split_live(tmp_path):
    """A tree whose metadata directory was written by the reference's Dataset class: load_npy_corpus takes the reference's
    own train/val/test artists, only the songs it validated, and (lyrics) the vocabulary size of word_ids.csv."""
    ref = refpin.load()
    from data.episode import load_npy_corpus, load_sampler_from_config
    rng = np.random.RandomState(2)
    artists = {"AC,DC": 6, "K_s Choice": 5, "Beyoncé": 7, "a b": 5, "tiny": 2, "e": 5, "f": 6, "g": 5, "h": 5, "i%": 5, "j": 5}

    class Loader(object):
        def is_song(self, name):
            return name.endswith(".txt")

        def validate(self, path):
            return "bad" not in os.path.basename(path)

    for a, n in artists.items():
        os.makedirs(tmp_path / a)
        for s in range(n):
            name = ("bad%d.txt" if s == 1 else "song %d.txt") % s
            open(tmp_path / a / name, "w").close()
            np.save(str(tmp_path / a / (name + ".16.npy")), rng.randint(0, 500, size=16).astype(np.int32))
    want = {}
    for split in ("train", "val", "test"):
        meta = ref.dataset.Metadata(str(tmp_path), "few_shot_metadata_lyrics_16")
        ds = ref.dataset.Dataset(str(tmp_path), split, Loader(), meta, split_proportions=(6, 2, 2), min_songs=4, seed=3)
        want[split] = {a.name: sorted(a.songs) for a in ds.artists}
    meta = ref.dataset.Metadata(str(tmp_path), "few_shot_metadata_lyrics_16")
    for i in range(1234):
        meta.write("word_ids.csv", "%d,w%d\n" % (i, i))
    meta.close()
    seen = set()
    for split in ("train", "val", "test"):
        corpus = load_npy_corpus(str(tmp_path), 16, 4, split, props=(6, 2, 2), seed=3, dataset="lyrics")
        assert corpus.artist_names == list(want[split].keys())
        for name, rows in zip(corpus.artist_names, corpus.artists):
            assert len(rows) == len(want[split][name])                    # the invalid song is not in the corpus
            for fname, row in zip(want[split][name], rows):
                assert np.array_equal(row, np.load(str(tmp_path / name / (fname + ".16.npy"))))
        assert corpus.vocab == 1234
        seen |= set(corpus.artist_names)
    assert "tiny" not in seen and len(seen) == 10
    # and without persisted files the same rule is applied to the artists in listdir order
    import shutil
    shutil.rmtree(tmp_path / "few_shot_metadata_lyrics_16")
    for split in ("train", "val", "test"):
        corpus = load_npy_corpus(str(tmp_path), 16, 4, split, props=(6, 2, 2), seed=3, dataset="lyrics")
        assert corpus.artist_names == list(want[split].keys())
    s = load_sampler_from_config(dict(dataset="lyrics", dataset_path=str(tmp_path), split="train", batch_size=2, support_size=2,
                                      query_size=2, max_len=16, train_proportion=6, val_proportion=2, test_proportion=2, dataset_seed=3, seed=0))
    ep = s.get_episode()
    assert ep.support.shape == (2, 2, 16) and ep.query.shape == (2, 2, 16)
