"""MIDI event vocabulary of the reference, without pretty_midi.

The reference turns a parsed MIDI file into a sequence of 4708 event ids (src/data/midi_loader.py:53-84): notes of the
16 General-MIDI program families (8 programs each) are merged per family, times are quantised to 10 ms steps, and the
song becomes NOTE_ON / NOTE_OFF / VELOCITY / TIME_SHIFT events.  Parsing the file needs pretty_midi (absent here); the
whole pipeline AFTER parsing is plain arithmetic over notes and is re-implemented here over light-weight records, so that
a corpus parsed elsewhere — or synthetic note lists — can be tokenised into the `<song>.<max_len>.npy` caches the
samplers read (src/data/base_loader.py:52-64), and generated id sequences can be turned back into notes.

Id layout (midi_loader.py:72-83), F = family, numbered 1..16 by the tokeniser:
    NOTE_ON    F*128 + pitch                      NOTE_OFF   2048 + F*128 + pitch
    VELOCITY   4096 + 32*F + bin                  TIME_SHIFT 4608 + steps - 1          (steps 1..100)
`detokenize` of the reference reads the same ranges with families numbered 0..15 (midi_loader.py:86-128), so a decoded
note lands one family above the one it was encoded from, NOTE_OFF ids of family 16 collide with the first VELOCITY ids and
VELOCITY ids of family 16 with the first TIME_SHIFT ids; every id stays below 4708.  Both directions are kept exactly as
the reference has them (corpora tokenised by the reference must decode the same way here).
"""
from collections import defaultdict, namedtuple

NOTE_ON, NOTE_OFF, TIME_SHIFT, VELOCITY = 1, 2, 3, 4
MAX_SHIFT_STEPS = 100
STEPS_PER_SECOND = 100
PROGRAMS_PER_FAMILY = 8
N_FAMILIES = 16
VELOCITY_BIN_SIZE = 4                      # ceil(127 / 32)
NUM_TOKENS = N_FAMILIES * 128 * 2 + 32 * N_FAMILIES + MAX_SHIFT_STEPS   # 4708 (midi_loader.py:53-60)

ControlChange = namedtuple('ControlChange', 'time number value instrument')
DecodedNote = namedtuple('DecodedNote', 'family start end pitch velocity')   # seconds; velocity 0..127; family 0..15


class Note(object):
    """One note of one instrument track.  start / end in seconds; `instrument` identifies the track within the song."""
    __slots__ = ('start', 'end', 'pitch', 'velocity', 'program', 'is_drum', 'instrument')

    def __init__(self, start, end, pitch, velocity, program=0, is_drum=False, instrument=0):
        self.start, self.end, self.pitch, self.velocity = start, end, pitch, velocity
        self.program, self.is_drum, self.instrument = program, is_drum, instrument

    def __repr__(self):
        return 'Note(%.3f-%.3f p%d v%d prog%d)' % (self.start, self.end, self.pitch, self.velocity, self.program)


def family_of(program):
    """1-based program family, as the tokeniser numbers it (midi_loader.py:166,239)."""
    return program // PROGRAMS_PER_FAMILY + 1


def apply_sustain(notes, control_changes, sustain_number=64):
    """Sustain pedal (controller 64, value >= 64 = down): a key released while its track's pedal is down keeps sounding
    until the pedal comes up, or until the same pitch is struck again on that track; whatever still sounds at the last
    event ends there (midi_loader.py:281-362).  Modifies and returns `notes`."""
    _PEDAL_DOWN, _PEDAL_UP, _KEY_DOWN, _KEY_UP = 0, 1, 2, 3
    timeline = [(n.start, _KEY_DOWN, n.instrument, n) for n in notes]
    timeline += [(n.end, _KEY_UP, n.instrument, n) for n in notes]
    for cc in control_changes:
        if cc.number == sustain_number:
            timeline.append((cc.time, _PEDAL_DOWN if cc.value >= 64 else _PEDAL_UP, cc.instrument, cc))
    timeline.sort(key=lambda item: item[0])            # stable: ties keep key-downs, key-ups, pedal events in that order
    sounding = defaultdict(list)
    pedal = defaultdict(bool)
    now = 0
    for now, kind, track, what in timeline:
        if kind == _PEDAL_DOWN:
            pedal[track] = True
        elif kind == _PEDAL_UP:
            pedal[track] = False
            still = []
            for n in sounding[track]:
                if n.end < now:
                    n.end = now                          # its key came up under the pedal: it rang until now
                else:
                    still.append(n)
            sounding[track] = still
        elif kind == _KEY_DOWN:
            if pedal[track]:
                still = []
                for n in sounding[track]:
                    if n.pitch == what.pitch:            # re-striking a ringing pitch cuts the old note
                        n.end = now                      # (a note cut to zero length stays in the song: the reference's
                    else:                                #  removal looks the Note up in a list of tuples and never finds it)
                        still.append(n)
                sounding[track] = still
            sounding[track].append(what)
        elif not pedal[track] and what in sounding[track]:
            sounding[track].remove(what)
    for ringing in sounding.values():
        for n in ringing:
            n.end = now
    return notes


def quantize(notes, steps_per_second=STEPS_PER_SECOND):
    """[(start_step, end_step, note)]: round half up, at least one step long (midi_loader.py:255-278)."""
    out = []
    for n in notes:
        a, b = int(n.start * steps_per_second + 0.5), int(n.end * steps_per_second + 0.5)
        out.append((a, b + 1 if a == b else b, n))
    return out


def resolve_pitch_clashes(quantized):
    """Tracks of one family are merged; when the same pitch overlaps itself the first note finishes and only the part of
    the later one that outlasts it is kept (midi_loader.py:130-183)."""
    kept = []
    ringing = defaultdict(list)                          # family -> [(pitch, end_step)]
    for a, b, n in sorted(quantized, key=lambda q: (q[0], q[1], q[2].program)):
        fam = family_of(n.program)
        ringing[fam] = [(p, e) for p, e in ringing[fam] if e > a]
        busy_until = max([e for p, e in ringing[fam] if p == n.pitch] or [0])
        if busy_until >= b:
            continue
        a = max(a, busy_until)
        ringing[fam].append((n.pitch, b))
        kept.append((a, b, n))
    return kept


def events_from_notes(quantized):
    """[(event_type, value, family)] in time order: TIME_SHIFT (1..100 steps, longer gaps are chained), VELOCITY when a
    family's velocity bin changes at a note-on, NOTE_ON, NOTE_OFF (midi_loader.py:198-252)."""
    marks = []
    for idx, (a, b, n) in enumerate(quantized):
        marks.append((a, idx, n.program, False))
        marks.append((b, idx, n.program, True))
    marks.sort()
    now = 0
    vel_bin = defaultdict(int)
    events = []
    for step, idx, program, is_off in marks:
        if step > now:
            while step > now + MAX_SHIFT_STEPS:
                events.append((TIME_SHIFT, MAX_SHIFT_STEPS, 0))
                now += MAX_SHIFT_STEPS
            events.append((TIME_SHIFT, step - now, 0))
            now = step
        n = quantized[idx][2]
        fam = family_of(program)
        if is_off:
            events.append((NOTE_OFF, n.pitch, fam))
            continue
        b = (n.velocity - 1) // VELOCITY_BIN_SIZE + 1
        if b != vel_bin[fam]:
            vel_bin[fam] = b
            events.append((VELOCITY, b, fam))
        events.append((NOTE_ON, n.pitch, fam))
    return events


def token_of(event_type, value, family):
    if event_type == NOTE_ON:
        return family * 128 + value
    if event_type == NOTE_OFF:
        return N_FAMILIES * 128 + family * 128 + value
    if event_type == VELOCITY:
        return N_FAMILIES * 128 * 2 + 32 * family + value
    if event_type == TIME_SHIFT:
        return N_FAMILIES * 128 * 2 + 32 * N_FAMILIES + value - 1
    raise ValueError('unknown event type %r' % (event_type,))


def tokenize_notes(notes, control_changes=()):
    """The reference's MIDILoader.tokenize after file parsing (midi_loader.py:62-84): sustain -> quantise -> drop drum
    tracks -> merge families -> events -> ids."""
    notes = apply_sustain(list(notes), list(control_changes))
    quantized = [q for q in quantize(notes) if not q[2].is_drum]
    return [token_of(*ev) for ev in events_from_notes(resolve_pitch_clashes(quantized))]


def notes_from_tokens(tokens):
    """The reference's MIDILoader.detokenize up to (not including) the pretty_midi objects (midi_loader.py:86-128):
    sorted DecodedNote records per family 0..15 (General-MIDI program = family * 8), velocity = bin * 4."""
    now = 0
    velocity = [16] * N_FAMILIES
    open_notes = [[None] * 128 for _ in range(N_FAMILIES)]
    done = [[] for _ in range(N_FAMILIES)]
    on_end, off_end, vel_end = N_FAMILIES * 128, N_FAMILIES * 128 * 2, N_FAMILIES * 128 * 2 + 32 * N_FAMILIES
    for tok in tokens:
        tok = int(tok)
        if tok < on_end:
            open_notes[tok // 128][tok % 128] = (velocity[tok // 128], now)
        elif tok < off_end:
            fam, pitch = (tok - on_end) // 128, (tok - on_end) % 128
            if open_notes[fam][pitch] is not None:
                vel, start = open_notes[fam][pitch]
                done[fam].append((start, now, pitch, vel))
                open_notes[fam][pitch] = None
        elif tok < vel_end:
            velocity[(tok - off_end) // 32] = (tok - off_end) % 32
        else:
            now += tok - vel_end + 1
    out = []
    for fam, rows in enumerate(done):
        for start, end, pitch, vel in sorted(rows):
            out.append(DecodedNote(fam, 0.01 * start, 0.01 * end, pitch, vel * 4))
    return out


def describe_tokens(tokens):
    """Readable dump of a generated id sequence (what `write_seq` stores when no MIDI writer is available)."""
    lines = ['family %2d  %7.2fs - %7.2fs  pitch %3d  velocity %3d' % tuple(n) for n in notes_from_tokens(tokens)]
    return '\n'.join(lines) if lines else '(no complete note in %d events)' % len(list(tokens))
