"""On-disk metadata of a corpus prepared by the reference: the files its `Dataset` / `Metadata` classes persist
under ``<dataset_path>/few_shot_metadata_<dataset>_<max_len>/`` (reference src/data/episode.py:120-121, src/data/dataset.py:22-47):

    valid_songs.csv     one ``quote(artist),quote(song)`` line per song that passed validation      (dataset.py:103-110, 150-152)
    train.csv, val.csv, test.csv
                        the artists of each split, one RAW (unquoted) directory name per line, no trailing newline
                                                                                                   (dataset.py:112-115, 172-175)
    word_ids.csv        ``<id>,<word>`` in order of first appearance (lyrics only; data/lyrics_vocab.py)

Reading them makes a tree prepared by the reference train here on the reference's own split and vocabulary; writing them
follows the same formats so the reference can consume a tree prepared here.  Pure host code, no third-party dependency.
"""
import os
from urllib.parse import quote, unquote

import numpy as np

VALID_SONGS_FILE = 'valid_songs.csv'
SPLITS = ('train', 'val', 'test')


def metadata_dir_name(dataset, max_len):
    """episode.py:120."""
    return 'few_shot_metadata_%s_%s' % (dataset, max_len)


class Metadata(object):
    """A directory of append-only text files (reference dataset.py:22-47)."""

    def __init__(self, root, name, create=True):
        self.dir = os.path.join(root, name)
        self._open = {}
        if create and not os.path.exists(self.dir):
            os.makedirs(self.dir)

    def path(self, filename):
        return os.path.join(self.dir, filename)

    def exists(self, filename):
        return os.path.exists(self.path(filename))

    def lines(self, filename):
        if self.exists(filename):
            with open(self.path(filename), 'r') as f:
                for line in f:
                    yield line

    def write(self, filename, line):
        if filename not in self._open:
            self._open[filename] = open(self.path(filename), 'a')
        self._open[filename].write(line)

    def close(self):
        for f in self._open.values():
            f.close()
        self._open = {}


def read_valid_songs(metadata, filename=VALID_SONGS_FILE):
    """{artist: set(song file names)} — both fields are url-quoted on disk, split at the FIRST comma (dataset.py:103-110)."""
    valid = {}
    for line in metadata.lines(filename):
        line = line.rstrip('\n')
        if not line:
            continue
        artist, song = line.split(',', 1)
        valid.setdefault(unquote(artist), set()).add(unquote(song))
    return valid


def append_valid_song(metadata, artist, song, filename=VALID_SONGS_FILE):
    """dataset.py:150-152."""
    metadata.write(filename, '%s,%s\n' % (quote(artist), quote(song)))


def read_split(metadata, split):
    """Artists of a persisted split, or None when `<split>.csv` does not exist (dataset.py:112-115)."""
    if not metadata.exists('%s.csv' % split):
        return None
    return [line.rstrip('\n') for line in metadata.lines('%s.csv' % split)]


def split_artists(all_artists, proportions=(8, 1, 1), seed=0):
    """The reference's split of a list of artists (dataset.py:167-182): floor counts for train and val, the rest is test,
    after an in-place `RandomState(seed).shuffle` of the list in the order it was collected.  -> {split: [artists]}"""
    artists = list(all_artists)
    total = sum(proportions)
    train_count = int(float(proportions[0]) / total * len(artists))
    val_count = int(float(proportions[1]) / total * len(artists))
    np.random.RandomState(seed).shuffle(artists)
    return {'train': artists[:train_count], 'val': artists[train_count:train_count + val_count],
            'test': artists[train_count + val_count:]}


def write_splits(metadata, splits):
    """'\\n'.join without a trailing newline, exactly what the reference appends (dataset.py:172-175)."""
    for split in SPLITS:
        metadata.write('%s.csv' % split, '\n'.join(splits[split]))
    metadata.close()


def highest_word_id(metadata, filename='word_ids.csv'):
    """get_num_tokens() - 1 of the reference's LyricsLoader (lyrics_loader.py:36-46, 63-64); -1 without the file."""
    top = -1
    for line in metadata.lines(filename):
        head = line.split(',', 1)[0]
        if head.strip():
            top = max(top, int(head))
    return top
