"""Word-id vocabulary of the reference's lyrics loader and its on-disk format.

The reference assigns ids to words in order of first appearance and appends every new pair to `word_ids.csv` as
`<id>,<word>` (src/data/lyrics_loader.py:36-46, 77-90), so ids stay stable across runs; `get_num_tokens()` is the highest
id + 1 (:63-64) and `detokenize` glues words back with the spacing rules of :92-102.  The word splitter itself is nltk's
`word_tokenize` (absent here): any callable can be plugged in; the default is a small regex splitter that separates
punctuation and the clitics nltk splits off ("n't", "'s", "'re", ...) — corpora tokenised by the reference are consumed
through their `.npy` caches and `word_ids.csv`, which do not depend on it.
"""
import os
import re
import string

_WORD = re.compile(r"n't\b|'(?:s|re|ve|ll|d|m)\b|\w+(?=n't\b)|\w+|[^\w\s]", re.IGNORECASE)


def simple_word_tokenize(text):
    return _WORD.findall(text)


class LyricsVocab(object):
    def __init__(self, path=None, tokenizer=simple_word_tokenize, persist=True):
        self.path = path
        self.tokenizer = tokenizer
        self.persist = persist and path is not None
        self.word_to_id = {}
        self.id_to_word = {}
        self.highest_word_id = -1
        if path is not None and os.path.isfile(path):
            with open(path, 'r') as f:
                for line in f:
                    word_id, word = line.rstrip('\n').split(',', 1)     # the word itself may contain commas
                    word_id = int(word_id)
                    self.word_to_id[word] = word_id
                    self.id_to_word[word_id] = word
                    self.highest_word_id = max(self.highest_word_id, word_id)

    def get_num_tokens(self):
        return self.highest_word_id + 1

    def tokenize(self, raw_lyrics):
        tokens = []
        for word in self.tokenizer(raw_lyrics):
            if word not in self.word_to_id:
                self.highest_word_id += 1
                self.word_to_id[word] = self.highest_word_id
                self.id_to_word[self.highest_word_id] = word
                if self.persist:
                    with open(self.path, 'a') as f:
                        f.write('%s,%s\n' % (self.highest_word_id, word))
            tokens.append(self.word_to_id[word])
        return tokens

    def detokenize(self, ids):
        out = ''
        for token in ids:
            word = self.id_to_word[int(token)]
            if word == "n't":
                out += word
            elif word not in string.punctuation and not word.startswith("'"):
                out += ' ' + word
            else:
                out += word
        return out.strip()
