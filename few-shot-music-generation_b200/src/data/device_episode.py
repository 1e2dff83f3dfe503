"""Episode assembly on the device.

The reference's sampler builds every episode on the host, song by song (src/data/episode.py:62-74, one
`dataset.load` per song), and the model feeds the resulting [B, S+Q, T] arrays to the session.  Here the whole
tokenised split is uploaded ONCE as an int32 [n_songs, T] CUDA tensor; `get_episode()` draws the same artists and
songs with the same RandomState call sequence as `EpisodeSampler.get_episode` and returns an `IndexedEpisode` that
carries only the song row indices.  `LSTMBaseline.train / eval` gather the token batch on the device
(`fsmg_gather_token_rows`), so a step moves B*(S+Q) indices over PCIe instead of B*(S+Q)*T tokens.

`IndexedEpisode.support` / `.query` still materialise the reference's host arrays on demand, so models written
against the reference API keep working unchanged.
"""
import numpy as np

from data.episode import Episode, EpisodeSampler


class IndexedEpisode(Episode):
    """support_ids [B, S] and query_ids [B, Q]: rows of `sampler.corpus_device`."""

    def __init__(self, sampler, support_ids, query_ids):
        self.sampler = sampler
        self.support_ids = support_ids
        self.query_ids = query_ids

    @property
    def support(self):
        return self.sampler.corpus_host[self.support_ids]

    @property
    def query(self):
        return self.sampler.corpus_host[self.query_ids]

    @property
    def corpus_device(self):
        return self.sampler.corpus_device


class DeviceEpisodeSampler(EpisodeSampler):
    """Same constructor and RNG stream as EpisodeSampler; episodes are index sets into a device-resident corpus."""

    def __init__(self, dataset, batch_size, support_size, query_size, max_len, dtype=np.int32, seed=None, device=None):
        super(DeviceEpisodeSampler, self).__init__(dataset, batch_size, support_size, query_size, max_len, dtype, seed)
        counts = [len(songs) for songs in dataset.artists]
        self.artist_offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.corpus_host = np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.int32) for s in dataset.artists], axis=0))
        assert self.corpus_host.shape[1] == max_len
        self._device = device
        self._corpus_device = None

    @classmethod
    def from_sampler(cls, sampler, device=None):
        """Wrap an existing EpisodeSampler; the new sampler CONTINUES its RandomState stream."""
        out = cls(sampler.dataset, sampler.batch_size, sampler.support_size, sampler.query_size, sampler.max_len,
                  sampler.dtype, None, device)
        out.random = sampler.random
        return out

    @property
    def corpus_device(self):
        if self._corpus_device is None:      # one upload per split
            import torch
            dev = self._device if self._device is not None else 'cuda:%d' % torch.cuda.current_device()
            self._corpus_device = torch.from_numpy(self.corpus_host).to(dev)
        return self._corpus_device

    def get_episode(self):
        b, s, q = self.batch_size, self.support_size, self.query_size
        support_ids = np.zeros((b, s), dtype=np.int64)
        query_ids = np.zeros((b, q), dtype=np.int64)
        artists = self.random.choice(len(self.dataset), size=b, replace=False)        # same draws as EpisodeSampler
        for bi, ai in enumerate(artists):
            pick = self.random.choice(len(self.dataset.artists[ai]), size=s + q, replace=False)
            query_ids[bi] = self.artist_offset[ai] + pick[:q]
            support_ids[bi] = self.artist_offset[ai] + pick[q:]
        return IndexedEpisode(self, support_ids, query_ids)
