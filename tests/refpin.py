"""Test infrastructure: load the parts of the UNMODIFIED reference that run without TensorFlow.

The reference's arithmetic (the LSTM graph) lives in TensorFlow 1.x and cannot run here, but everything either side of it
is plain Python and imports fine once the two absent third-party modules are stubbed:

* ``models/base_model.py:57-86``     flatten_first_two_dims / convert_tokens_to_input_and_target (the input/target shift);
* ``data/midi_loader.py:62-399``     the whole MIDI event pipeline after file parsing (``pretty_midi`` stubbed: only
                                     ``PrettyMIDI(path)`` — file parsing — needs the real library);
* ``data/lyrics_loader.py:65-95``    word ids + detokenize (``nltk`` stubbed: only the default word splitter needs it);
* ``data/dataset.py:22-232``         Metadata / Dataset: split persistence and the url-quoted valid_songs.csv;
* ``data/base_loader.py:52-64``      the ``<song>.<max_len>.npy`` cache contract.

The reference and this repo both call their top-level packages ``data`` / ``models``; ``load()`` imports the reference's
copies with the repo's temporarily moved out of ``sys.modules`` and puts everything back afterwards, so both can be used in
one test process.  ``/root/reference`` exists only in the build container: tests that need it skip elsewhere and rely on
the committed fixtures ``tests/golden/reference_*.json|npz`` written by ``tests/golden/make_reference_golden.py``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from types import SimpleNamespace

import numpy as np

REFERENCE_SRC = os.environ.get("FSMG_REFERENCE_SRC", "/root/reference/src")
_SHADOWED = ("data", "models", "train", "config", "evaluation")
_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "models", "base_model.py"))


class _StubNote(object):
    """pretty_midi.Note / ControlChange stand-ins: plain attribute bags (that is all the reference reads)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _stub_modules():
    pm = types.ModuleType("pretty_midi")

    class PrettyMIDI(object):
        def __init__(self, path=None):
            if path is not None:
                raise RuntimeError("pretty_midi is stubbed: MIDI file parsing is not available")
            self.instruments = []

    class Instrument(object):
        def __init__(self, program=0, is_drum=False):
            self.program, self.is_drum, self.notes, self.control_changes = program, is_drum, [], []

    class Note(_StubNote):
        def __init__(self, velocity, pitch, start, end):
            _StubNote.__init__(self, velocity=velocity, pitch=pitch, start=start, end=end)

    class ControlChange(_StubNote):
        def __init__(self, number, value, time):
            _StubNote.__init__(self, number=number, value=value, time=time)

    pm.PrettyMIDI, pm.Instrument, pm.Note, pm.ControlChange = PrettyMIDI, Instrument, Note, ControlChange
    nl = types.ModuleType("nltk")
    nl.word_tokenize = lambda text: text.split()
    return {"pretty_midi": pm, "nltk": nl}


def load():
    """-> namespace(base_model, midi_loader, lyrics_loader, dataset, base_loader, pretty_midi) of reference modules."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_SRC)
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in _SHADOWED or k in ("pretty_midi", "nltk")}
    for k in saved:
        del sys.modules[k]
    stubs = _stub_modules()
    sys.modules.update(stubs)
    sys.path.insert(0, REFERENCE_SRC)
    try:
        mods = {}
        for name in ("models.base_model", "data.base_loader", "data.midi_loader", "data.lyrics_loader", "data.dataset"):
            mods[name.split(".")[1]] = importlib.import_module(name)
        for m in mods.values():
            assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE_SRC)), m.__file__
    finally:
        sys.path.remove(REFERENCE_SRC)
        for k in [k for k in sys.modules if k.split(".")[0] in _SHADOWED or k in ("pretty_midi", "nltk")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cache = SimpleNamespace(pretty_midi=stubs["pretty_midi"], **mods)
    return _cache


# ---- seeded random songs: plain records, convertible to the reference's (stub) pretty_midi objects and to the repo's Note/ControlChange ----
def random_song(rng: np.random.RandomState, max_tracks: int = 5, max_notes: int = 40):
    """{'tracks': [{'program', 'is_drum', 'notes': [[start, end, pitch, velocity]...], 'ccs': [[time, number, value]...]}]}.
    Times sit on a 5 ms grid (quantisation ties at x.xx5), tracks often share a program family (pitch clashes), family 16
    (programs 120..127) appears (the reference's 1-based-family id overlap), gaps above one second chain TIME_SHIFT events,
    pedal (controller 64) events straddle note boundaries, a few other controllers are mixed in."""
    tracks = []
    n_tracks = int(rng.randint(1, max_tracks + 1))
    base_prog = int(rng.randint(0, 128))
    for _ in range(n_tracks):
        mode = rng.randint(0, 4)
        if mode == 0:
            program = base_prog                                   # same program -> same family, guaranteed clashes
        elif mode == 1:
            program = (base_prog // 8) * 8 + int(rng.randint(0, 8))  # same family, other program
        elif mode == 2:
            program = int(rng.randint(120, 128))                  # family 16
        else:
            program = int(rng.randint(0, 128))
        is_drum = bool(rng.rand() < 0.15)
        notes, t = [], 0.0
        for _ in range(int(rng.randint(1, max_notes + 1))):
            t += float(rng.choice([0.0, 0.005, 0.01, 0.02, 0.05, 0.25, 1.3, 2.6], p=[0.2, 0.15, 0.15, 0.15, 0.15, 0.1, 0.07, 0.03]))
            dur = float(rng.choice([0.0, 0.003, 0.005, 0.01, 0.04, 0.1, 0.5, 1.7]))
            pitch = int(rng.choice([55, 60, 64, 67, int(rng.randint(0, 128))]))
            vel = int(rng.randint(1, 128))
            notes.append([round(t, 4), round(t + dur, 4), pitch, vel])
        ccs = []
        if rng.rand() < 0.6:
            tc = 0.0
            for _ in range(int(rng.randint(1, 8))):
                tc += float(rng.choice([0.0, 0.005, 0.03, 0.2, 0.9]))
                number = 64 if rng.rand() < 0.8 else int(rng.choice([1, 7, 10, 66]))
                ccs.append([round(tc, 4), number, int(rng.choice([0, 30, 63, 64, 100, 127]))])
        tracks.append(dict(program=program, is_drum=is_drum, notes=notes, ccs=ccs))
    return dict(tracks=tracks)


def to_reference_midi(ref, song):
    midi = ref.pretty_midi.PrettyMIDI()
    for tr in song["tracks"]:
        inst = ref.pretty_midi.Instrument(program=tr["program"], is_drum=tr["is_drum"])
        inst.notes = [ref.pretty_midi.Note(velocity=v, pitch=p, start=a, end=b) for a, b, p, v in tr["notes"]]
        inst.control_changes = [ref.pretty_midi.ControlChange(number=n, value=v, time=t) for t, n, v in tr["ccs"]]
        midi.instruments.append(inst)
    return midi


def to_repo_notes(song):
    from data import midi_events as M
    notes, ccs = [], []
    for idx, tr in enumerate(song["tracks"]):
        notes += [M.Note(a, b, p, v, program=tr["program"], is_drum=tr["is_drum"], instrument=idx) for a, b, p, v in tr["notes"]]
        ccs += [M.ControlChange(t, n, v, idx) for t, n, v in tr["ccs"]]
    return notes, ccs


def reference_tokenize(ref, song):
    """MIDILoader.tokenize (midi_loader.py:62-84) on the stubbed PrettyMIDI object — the reference's own code, unmodified."""
    return [int(t) for t in ref.midi_loader.MIDILoader(max_len=1 << 30).tokenize(to_reference_midi(ref, song))]
