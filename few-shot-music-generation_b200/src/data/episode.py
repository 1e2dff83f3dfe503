"""Episodic support/query sampler with the reference's API
(reference src/data/episode.py:15-18 `Episode`, :44-80 `EpisodeSampler`, :82-149
`load_sampler_from_config`), re-implemented for two kinds of corpora that need none of the
reference's absent dependencies (nltk / pretty_midi):

* ``synthetic_lyrics`` / ``synthetic_midi`` — seeded token streams generated in memory;
* ``lyrics`` / ``midi`` — a directory tree ``root/<artist>/<song>.<max_len>.npy`` holding the
  token caches the reference's loaders persist (reference src/data/base_loader.py:52-64),
  so a corpus tokenised once by the reference can be trained on directly.

The RNG call sequence matches the reference on old NumPy: ``choice(n, size, replace=False)`` over
artists, then per artist ``choice(n_songs, S+Q, replace=False)`` with the first Q drawn songs
forming the query set (reference :34-41, :62-74).
"""
import os

import numpy as np
import yaml
from numpy.random import RandomState

MIDI_VOCAB = 16 * 128 * 2 + 32 * 16 + 100  # 4708 event ids (reference src/data/midi_loader.py:53-60)


class Episode(object):
    def __init__(self, support, query):
        self.support = support
        self.query = query


class TokenCorpus(object):
    """artists -> list of int32 [max_len] songs, zero-padded (pad id 0 is a real token)."""

    def __init__(self, songs_by_artist, vocab, max_len):
        self.artists = list(songs_by_artist)
        self.vocab = int(vocab)
        self.max_len = int(max_len)

    def __len__(self):
        return len(self.artists)


def _zipf_tokens(rng, shape, vocab):
    pmf = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(pmf / pmf.sum())
    return np.minimum(np.searchsorted(cdf, rng.random_sample(size=shape)), vocab - 1).astype(np.int32)


def make_synthetic_corpus(kind, max_len, vocab, n_artists, songs_per_artist, seed, split):
    rng = RandomState((int(seed) * 7919 + {'train': 0, 'val': 1, 'test': 2}.get(split, 3)) % (2 ** 31))
    shape = (n_artists, songs_per_artist, max_len)
    if kind == 'synthetic_lyrics':
        tok = _zipf_tokens(rng, shape, vocab)
    else:
        tok = rng.randint(0, vocab, size=shape).astype(np.int32)
    return TokenCorpus([tok[a] for a in range(n_artists)], vocab, max_len)


def load_npy_corpus(root, max_len, min_songs, split, props=(8, 1, 1), seed=0, dataset=None):
    """A corpus tokenised by the reference: ``root/<artist>/<song>.<max_len>.npy`` token caches (base_loader.py:52-64) plus,
    when present, the metadata directory its Dataset persisted (data/dataset_meta.py).  The reference's data contract is
    followed: the split comes from ``<split>.csv`` when it exists, else from the reference's own rule (floor counts,
    ``RandomState(seed).shuffle`` over the artists in ``os.listdir`` order, dataset.py:122-182); ``valid_songs.csv`` filters the
    songs; the lyrics vocabulary is ``highest word id + 1`` of ``word_ids.csv`` (every word of every song, not just the
    truncated caches)."""
    from data.dataset_meta import Metadata, metadata_dir_name, read_valid_songs, read_split, split_artists, highest_word_id, VALID_SONGS_FILE
    suffix = '.%s.npy' % max_len
    meta = Metadata(root, metadata_dir_name(dataset, max_len), create=False) if dataset else None
    valid = read_valid_songs(meta) if meta is not None and meta.exists(VALID_SONGS_FILE) else None
    songs_of, eligible = {}, []
    for artist in os.listdir(root):                 # listdir order, like the reference (the shuffle below depends on it)
        adir = os.path.join(root, artist)
        if not os.path.isdir(adir):
            continue
        names = sorted(f[:-len(suffix)] for f in os.listdir(adir) if f.endswith(suffix))
        if valid is not None:
            names = [n for n in names if n in valid.get(artist, ())]
        if names:
            songs_of[artist] = names
            if len(names) >= min_songs:
                eligible.append(artist)
    persisted = read_split(meta, split) if meta is not None else None
    if persisted is not None:
        missing = [a for a in persisted if a not in songs_of]
        if missing:
            raise RuntimeError('%d artists of the persisted %s split have no token caches "*%s" under %s (first: %r)'
                               % (len(missing), split, suffix, root, missing[0]))
        chosen = persisted
    else:
        chosen = split_artists(eligible, props, seed)[split]
    if not chosen:
        raise RuntimeError('no artist under %s has >= %d token caches "*%s" for the %s split (tokenise the corpus with the '
                           "reference's loaders first: nltk / pretty_midi are not available here)" % (root, min_songs, suffix, split))
    artists, vocab = [], 0
    for artist in chosen:
        songs = [np.load(os.path.join(root, artist, n + suffix)).astype(np.int32) for n in songs_of[artist]]
        artists.append(np.stack(songs))
        vocab = max(vocab, int(max(s.max() for s in songs)) + 1)
    corpus = TokenCorpus(artists, vocab, max_len)
    corpus.artist_names = list(chosen)
    corpus.word_ids_path = None
    for cand in ((meta.path('word_ids.csv') if meta is not None else None), os.path.join(root, 'word_ids.csv')):
        if cand and os.path.isfile(cand):
            corpus.word_ids_path = cand
            break
    if meta is not None and meta.exists('word_ids.csv'):
        corpus.vocab = highest_word_id(meta) + 1     # LyricsLoader.get_num_tokens (lyrics_loader.py:63-64)
    return corpus


class EpisodeSampler(object):
    def __init__(self, dataset, batch_size, support_size, query_size, max_len, dtype=np.int32, seed=None):
        self.dataset = dataset
        self.batch_size = batch_size
        self.support_size = support_size
        self.query_size = query_size
        self.max_len = max_len
        self.dtype = dtype
        self.random = RandomState(seed) if seed is not None else np.random

    def get_episode(self):
        b, s, q = self.batch_size, self.support_size, self.query_size
        support = np.zeros((b, s, self.max_len), dtype=self.dtype)
        query = np.zeros((b, q, self.max_len), dtype=self.dtype)
        artists = self.random.choice(len(self.dataset), size=b, replace=False)
        for bi, ai in enumerate(artists):
            songs = self.dataset.artists[ai]
            pick = self.random.choice(len(songs), size=s + q, replace=False)
            query[bi] = songs[pick[:q]]
            support[bi] = songs[pick[q:]]
        return Episode(support, query)

    def get_num_unique_words(self):
        return self.dataset.vocab

    detokenizer = None      # set by load_sampler_from_config: MIDI event decoder / lyrics word table when available

    def detokenize(self, numpy_data):
        """A string for write_seq (reference train.py:17-24): decoded notes for the MIDI vocabulary, words when the corpus
        ships its word_ids.csv, else the space-separated ids (no .mid writer without pretty_midi)."""
        ids = np.asarray(numpy_data).reshape(-1)
        if self.detokenizer is not None:
            return self.detokenizer(ids)
        return ' '.join(str(int(t)) for t in ids)


def load_sampler_from_config(config):
    """Create an EpisodeSampler from a config dict / yaml path (reference :82-149)."""
    if isinstance(config, str):
        config = yaml.safe_load(open(config, 'r'))
    elif not isinstance(config, dict):
        config = yaml.safe_load(config)
    for key in ('dataset_path', 'query_size', 'support_size', 'batch_size', 'max_len', 'dataset', 'split'):
        if key not in config:
            raise RuntimeError('required config key "%s" not found' % key)
    min_songs = config['support_size'] + config['query_size']
    kind = config['dataset']
    if kind in ('synthetic_lyrics', 'synthetic_midi'):
        vocab = int(config.get('synthetic_vocab', 10000 if kind == 'synthetic_lyrics' else MIDI_VOCAB))
        corpus = make_synthetic_corpus(kind, config['max_len'], vocab, int(config.get('synthetic_artists', 64)),
                                       max(int(config.get('synthetic_songs_per_artist', 24)), min_songs),
                                       config.get('dataset_seed', 0), config['split'])
    elif kind in ('lyrics', 'midi'):
        root = config['dataset_path']
        if not os.path.isdir(root):
            raise RuntimeError('required data directory %s does not exist' % root)
        props = (config.get('train_proportion', 8), config.get('val_proportion', 1), config.get('test_proportion', 1))
        corpus = load_npy_corpus(root, config['max_len'], min_songs, config['split'], props, config.get('dataset_seed', 0), dataset=kind)
        if kind == 'midi':
            corpus.vocab = MIDI_VOCAB
    else:
        raise RuntimeError('unknown dataset "%s"' % kind)
    if len(corpus) < config['batch_size']:
        raise RuntimeError('split "%s" has %d artists < batch_size %d' % (config['split'], len(corpus), config['batch_size']))
    if config.get('device_episodes', False):   # corpus resident in HBM, episodes are index sets (data/device_episode.py)
        from data.device_episode import DeviceEpisodeSampler
        sampler = DeviceEpisodeSampler(corpus, config['batch_size'], config['support_size'], config['query_size'],
                                       config['max_len'], seed=config.get('seed', None))
    else:
        sampler = EpisodeSampler(corpus, config['batch_size'], config['support_size'], config['query_size'],
                                 config['max_len'], seed=config.get('seed', None))
    if kind in ('midi', 'synthetic_midi') and corpus.vocab == MIDI_VOCAB:
        from data.midi_events import describe_tokens
        sampler.detokenizer = describe_tokens
    elif kind == 'lyrics' and getattr(corpus, 'word_ids_path', None):
        from data.lyrics_vocab import LyricsVocab
        vocab = LyricsVocab(corpus.word_ids_path, persist=False)
        sampler.detokenizer = lambda ids: vocab.detokenize([t for t in ids if int(t) in vocab.id_to_word])
    return sampler
