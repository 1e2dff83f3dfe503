"""CPU tests of the data-format layer next to the hot path: the MIDI event vocabulary and the lyrics
word-id table, re-implemented without pretty_midi / nltk, against hand-computed expectations taken from the reference's
rules (src/data/midi_loader.py, src/data/lyrics_loader.py)."""
import numpy as np

from data import midi_events as M
from data.lyrics_vocab import LyricsVocab, simple_word_tokenize


def test_token_ranges_and_count():
    assert M.NUM_TOKENS == 4708
    assert M.token_of(M.NOTE_ON, 60, 1) == 128 + 60
    assert M.token_of(M.NOTE_OFF, 60, 1) == 2048 + 128 + 60
    assert M.token_of(M.VELOCITY, 32, 1) == 4096 + 32 + 32
    assert M.token_of(M.TIME_SHIFT, 1, 0) == 4608 and M.token_of(M.TIME_SHIFT, 100, 0) == 4707
    # the reference's 1-based families: family 16 spills into the next range, yet every id stays inside the vocabulary
    assert M.token_of(M.NOTE_OFF, 127, 16) == 4223 < M.NUM_TOKENS
    assert M.token_of(M.VELOCITY, 32, 16) == 4640 < M.NUM_TOKENS


def test_two_notes_become_the_expected_event_stream():
    notes = [M.Note(0.0, 0.5, 60, 64, program=0), M.Note(0.5, 3.0, 64, 64, program=0), M.Note(0.0, 1.0, 36, 100, program=0, is_drum=True)]
    toks = M.tokenize_notes(notes)
    fam = 1
    vel = (64 - 1) // 4 + 1
    want = [M.token_of(M.VELOCITY, vel, fam), M.token_of(M.NOTE_ON, 60, fam),
            M.token_of(M.TIME_SHIFT, 50, 0),
            M.token_of(M.NOTE_OFF, 60, fam), M.token_of(M.NOTE_ON, 64, fam),      # same step: note index order, no new velocity event
            M.token_of(M.TIME_SHIFT, 100, 0), M.token_of(M.TIME_SHIFT, 100, 0), M.token_of(M.TIME_SHIFT, 50, 0),   # 250 steps chained
            M.token_of(M.NOTE_OFF, 64, fam)]
    assert toks == want                      # the drum track is gone
    assert max(toks) < M.NUM_TOKENS


def test_quantisation_rounds_half_up_and_keeps_one_step():
    q = M.quantize([M.Note(0.004, 0.0049, 60, 80), M.Note(0.005, 0.0251, 61, 80)])
    assert (q[0][0], q[0][1]) == (0, 1)      # zero-length after rounding -> one step
    assert (q[1][0], q[1][1]) == (1, 3)


def test_same_pitch_clash_first_note_finishes_then_the_rest_of_the_second():
    # two programs of one family (24 and 25: guitars) play pitch 55 overlapping: steps [100,102) and [101,110)
    a, b = M.Note(1.00, 1.02, 55, 80, program=24), M.Note(1.01, 1.10, 55, 80, program=25)
    kept = M.resolve_pitch_clashes(M.quantize([a, b]))
    assert [(s, e) for s, e, _ in kept] == [(100, 102), (102, 110)]
    # a second note that ends inside the first disappears
    c = M.Note(1.00, 1.10, 55, 80, program=24)
    d = M.Note(1.02, 1.05, 55, 80, program=25)
    assert [(s, e) for s, e, _ in M.resolve_pitch_clashes(M.quantize([c, d]))] == [(100, 110)]
    # other families are untouched
    e = M.Note(1.02, 1.05, 55, 80, program=40)
    assert len(M.resolve_pitch_clashes(M.quantize([c, e]))) == 2


def test_sustain_pedal_extends_released_notes_until_pedal_up_or_restrike():
    n1 = M.Note(0.0, 0.2, 60, 80)
    n2 = M.Note(0.1, 0.3, 62, 80)
    n3 = M.Note(0.5, 0.6, 60, 80)            # re-strikes pitch 60 while it still rings under the pedal
    cc = [M.ControlChange(0.05, 64, 127, 0), M.ControlChange(1.0, 64, 0, 0)]
    M.apply_sustain([n1, n2, n3], cc)
    assert n1.end == 0.5                     # cut by the re-strike
    assert n2.end == 1.0 and n3.end == 1.0   # ring until the pedal comes up
    # another controller number is ignored
    n4 = M.Note(0.0, 0.2, 60, 80)
    M.apply_sustain([n4], [M.ControlChange(0.05, 7, 127, 0)])
    assert n4.end == 0.2


def test_decoding_follows_the_reference_detokenizer():
    notes = [M.Note(0.0, 0.5, 60, 64, program=0), M.Note(0.25, 0.75, 67, 100, program=8)]
    dec = M.notes_from_tokens(np.asarray(M.tokenize_notes(notes), dtype=np.int32))
    # families are written 1-based and read 0-based: program family f decodes as family index f (one General-MIDI family up)
    assert [(n.family, n.pitch) for n in dec] == [(1, 60), (2, 67)]
    assert [(round(n.start, 2), round(n.end, 2)) for n in dec] == [(0.0, 0.5), (0.25, 0.75)]
    assert [n.velocity for n in dec] == [((64 - 1) // 4 + 1) * 4, ((100 - 1) // 4 + 1) * 4]
    assert 'pitch  60' in M.describe_tokens(M.tokenize_notes(notes))
    assert M.notes_from_tokens([4608, 4707]) == []


def test_lyrics_vocab_ids_in_order_of_first_appearance_and_csv_round_trip(tmp_path):
    path = tmp_path / 'word_ids.csv'
    v = LyricsVocab(str(path))
    ids = v.tokenize("Hello, world! Don't stop, world")
    assert simple_word_tokenize("Don't stop") == ['Do', "n't", 'stop']
    assert ids == [0, 1, 2, 3, 4, 5, 6, 1, 2]
    assert v.get_num_tokens() == 7
    assert path.read_text().splitlines()[:3] == ['0,Hello', '1,,', '2,world']      # '<id>,<word>', the word may be a comma
    again = LyricsVocab(str(path))
    assert again.word_to_id == v.word_to_id and again.get_num_tokens() == 7
    assert again.tokenize('world of Hello') == [2, 7, 0] and again.get_num_tokens() == 8
    assert v.detokenize(ids) == "Hello, world! Don't stop, world"
