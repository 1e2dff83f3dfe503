"""Generate ``tests/golden/reference_*.{npz,json}`` by RUNNING THE UNMODIFIED REFERENCE (everything of it that runs without
TensorFlow; see tests/refpin.py) in the build container:

    python tests/golden/make_reference_golden.py        # needs /root/reference

The fixtures travel to the GPU box (where /root/reference does not exist) and pin, bit-exactly:
  reference_shift.npz     models/base_model.py:57-86   flatten_first_two_dims + convert_tokens_to_input_and_target
  reference_midi.json     data/midi_loader.py:62-399   MIDILoader.tokenize / detokenize on seeded random songs
  reference_lyrics.json   data/lyrics_loader.py:65-95  word ids, word_ids.csv, detokenize
  reference_dataset.json  data/dataset.py:22-232       split persistence + url-quoted valid_songs.csv
"""
import hashlib
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import refpin  # noqa: E402

N_MIDI_SONGS = 160
MIDI_SEED = 20260917
LYRICS_TEXTS = [
    "Hello , world ! Do n't stop , world",
    "I 'm gon na say it 's fine ; is n't it ? ( yes ) -- `` ok ''",
    "la la la , la . la ! 'cause we 're here",
    "",
]
ARTISTS = {  # directory name -> song files (names exercise the url-quoting: spaces, commas, percent, unicode, apostrophes)
    "K_s Choice": ["ironflowers.mid", "not, an addict.mid", "100% sure.mid"],
    "AC,DC": ["back in black.mid", "t.n.t.mid", "high voltage.mid", "thunderstruck.mid"],
    "Beyoncé": ["déjà vu.mid", "halo.mid", "formation.mid"],
    "tool": ["lateralus.mid", "schism.mid", "46 & 2.mid"],
    "Guns N' Roses": ["november rain.mid", "patience.mid", "don't cry.mid"],
    "a+b=c": ["x.mid", "y.mid", "z.mid"],
    "few": ["only one.mid"],
    "seven": ["1.mid", "2.mid", "3.mid"], "eight": ["1.mid", "2.mid", "3.mid"], "nine": ["1.mid", "2.mid", "3.mid"],
    "ten": ["1.mid", "2.mid", "3.mid"], "eleven": ["1.mid", "2.mid", "3.mid"], "twelve": ["1.mid", "2.mid", "3.mid"],
}


def song_digest(song):
    return hashlib.sha256(json.dumps(song, sort_keys=True).encode()).hexdigest()[:16]


def make_shift(ref):
    rng = np.random.RandomState(3)
    out = {}
    for i, (b, s, t, v) in enumerate([(5, 9, 32, 10000), (1, 1, 1, 7), (3, 4, 2, 50), (5, 4, 128, 4708), (2, 3, 17, 300)]):
        tok = rng.randint(0, v, size=(b, s, t)).astype(np.int32)
        if i == 2:
            tok[:] = 0
        x, y = ref.base_model.convert_tokens_to_input_and_target(tok, start_word=v)
        x2, y2 = ref.base_model.convert_tokens_to_input_and_target(tok)
        out.update({f"tok{i}": tok, f"x{i}": np.asarray(x), f"y{i}": np.asarray(y), f"xn{i}": np.asarray(x2), f"yn{i}": np.asarray(y2),
                    f"flat{i}": ref.base_model.flatten_first_two_dims(tok), f"start{i}": np.asarray(v)})
    np.savez_compressed(HERE / "reference_shift.npz", **out)


def make_midi(ref):
    rng = np.random.RandomState(MIDI_SEED)
    cases = []
    for _ in range(N_MIDI_SONGS):
        song = refpin.random_song(rng)
        toks = refpin.reference_tokenize(ref, song)
        # decode with the reference's detokenize (stub pretty_midi objects): [(program, start, end, pitch, velocity)]
        midi = ref.midi_loader.MIDILoader(max_len=1 << 30).detokenize(np.asarray(toks[:400], dtype=np.int64))
        dec = [[int(inst.program), float(n.start), float(n.end), int(n.pitch), int(n.velocity)] for inst in midi.instruments for n in inst.notes]
        cases.append(dict(digest=song_digest(song), tokens=toks, decoded=dec))
    (HERE / "reference_midi.json").write_text(json.dumps(dict(seed=MIDI_SEED, n=N_MIDI_SONGS, cases=cases), separators=(",", ":")))


def make_lyrics(ref):
    with tempfile.TemporaryDirectory() as d:
        meta = ref.dataset.Metadata(d, "few_shot_metadata_lyrics_50")
        loader = ref.lyrics_loader.LyricsLoader(50, metadata=meta, tokenizer=lambda text: text.split())
        ids = [loader.tokenize(t) for t in LYRICS_TEXTS]
        meta.close()
        csv = open(os.path.join(meta.dir, "word_ids.csv")).read()
        detok = [loader.detokenize(np.asarray(i, dtype=np.int64)) for i in ids]
        # a second loader bootstraps from the persisted file (lyrics_loader.py:36-46)
        again = ref.lyrics_loader.LyricsLoader(50, metadata=ref.dataset.Metadata(d, "few_shot_metadata_lyrics_50"), tokenizer=lambda text: text.split())
        n_tokens = again.get_num_tokens()
    (HERE / "reference_lyrics.json").write_text(json.dumps(dict(texts=LYRICS_TEXTS, ids=ids, word_ids_csv=csv, detokenized=detok,
                                                                num_tokens=n_tokens), indent=1))


class _AlwaysValidLoader(object):
    """Stands in for MIDILoader inside the reference's Dataset: every '.mid' file is a valid song except names starting with 'bad'."""

    def is_song(self, name):
        return name.endswith(".mid")

    def validate(self, path):
        return not os.path.basename(path).startswith("bad")


def make_dataset(ref):
    cases = []
    for props, seed in (((8, 1, 1), 0), ((6, 2, 2), 5), ((1, 1, 1), 123)):
        with tempfile.TemporaryDirectory() as d:
            for artist, songs in ARTISTS.items():
                os.makedirs(os.path.join(d, artist))
                for s in songs + ["bad one.mid", "notes.txt"]:
                    open(os.path.join(d, artist, s), "w").close()
            listing = os.listdir(d)
            splits = {}
            for split in ("train", "val", "test"):
                meta = ref.dataset.Metadata(d, "few_shot_metadata_midi_50")
                ds = ref.dataset.Dataset(d, split, _AlwaysValidLoader(), meta, split_proportions=props, min_songs=3, seed=seed)
                splits[split] = {a.name: sorted(a.songs) for a in ds.artists}
            files = {f: open(os.path.join(d, "few_shot_metadata_midi_50", f)).read() for f in ("train.csv", "val.csv", "test.csv", "valid_songs.csv")}
        cases.append(dict(proportions=list(props), seed=seed, listdir=[a for a in listing if a in ARTISTS], splits=splits, files=files))
    (HERE / "reference_dataset.json").write_text(json.dumps(dict(artists=ARTISTS, min_songs=3, cases=cases), indent=1, ensure_ascii=False))


if __name__ == "__main__":
    ref = refpin.load()
    make_shift(ref)
    make_midi(ref)
    make_lyrics(ref)
    make_dataset(ref)
    for f in sorted(HERE.glob("reference_*")):
        print(f.name, f.stat().st_size, "bytes")
