"""UnigramModel — drop-in for the reference's `models.unigram_model.UnigramModel`
(reference src/models/unigram_model.py:8-78, config/unigram.yaml), resolved through the same
model_module_name / model_class_name registry (reference src/train/train.py:12-15).

The `word_count` variable lives on the GPU; train / eval / sample call the CUDA kernels of
csrc/unigram.cuh through the C-ABI (fsmg_unigram_step, fsmg_unigram_argmax).  No CPU fallback: the
constructor raises without a CUDA device or without libfsmg.so.
"""
import glob
import os
import sys

import numpy as np

from models.base_model import BaseModel, flatten_first_two_dims

try:
    import fsmg  # noqa: F401
except ImportError:  # running from the source tree: <pkg>/src/models/ -> <pkg>/
    sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..')))
    import fsmg  # noqa: F401
from fsmg import _lib
from fsmg._lib import FsmgError


class UnigramModel(BaseModel):
    """Word frequencies of the meta-training set give the word probabilities; evaluation ignores the
    support set and scores only the query set (reference :9-13)."""

    ALPHA = 1.0  # reference :24

    def __init__(self, config):
        super(UnigramModel, self).__init__(config)
        import torch
        if not torch.cuda.is_available():
            raise FsmgError('fsmg requires a CUDA device (B200, sm_100a); there is no CPU fallback')
        self._torch = torch
        self._lib = _lib.load()
        self._input_size = int(config['input_size'])
        self._time_steps = int(config['max_len'])
        self._device = torch.device('cuda:%d' % torch.cuda.current_device())
        self._counts = torch.full((self._input_size,), self.ALPHA, dtype=torch.float32, device=self._device)
        self._scratch = torch.zeros(8, dtype=torch.float32, device=self._device)   # [0..3] kernel scratch, [4] mean NLL
        self._word = torch.zeros(1, dtype=torch.int32, device=self._device)
        self._cap = 0
        self._tok = self._pinned = None
        self._pinned_out = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._global_step = 0    # TFModel's global_step is never incremented by this model (no optimizer)

    # ---- staging ---------------------------------------------------------------------------------
    def _stage(self, tokens):
        torch = self._torch
        tok = np.ascontiguousarray(tokens, dtype=np.int32).reshape(-1, self._time_steps)
        if tok.size and (tok.min() < 0 or tok.max() >= self._input_size):
            raise FsmgError('token id outside [0, input_size=%d)' % self._input_size)
        if tok.size > self._cap:
            self._cap = int(tok.size)
            self._tok = torch.empty(self._cap, dtype=torch.int32, device=self._device)
            self._pinned = torch.empty(self._cap, dtype=torch.int32).pin_memory()
        torch.cuda.current_stream(self._device).synchronize()     # previous H2D done with the pinned buffer
        self._pinned[:tok.size].copy_(torch.from_numpy(tok.reshape(-1)))
        self._tok[:tok.size].copy_(self._pinned[:tok.size], non_blocking=True)
        return tok.shape[0]

    def _step(self, n_rows, col_begin, col_end, update):
        torch = self._torch
        s = torch.cuda.current_stream(self._device)
        _lib.check(self._lib.fsmg_unigram_step(self._counts.data_ptr(), self._input_size, self._tok.data_ptr(), n_rows,
                                               self._time_steps, col_begin, col_end, int(update), self._scratch.data_ptr(),
                                               self._scratch.data_ptr() + 16, s.cuda_stream))
        self._pinned_out[:1].copy_(self._scratch[4:5], non_blocking=True)
        s.synchronize()
        return float(self._pinned_out[0])

    # ---- BaseModel API ---------------------------------------------------------------------------
    def train(self, episode):
        """Concatenate support and query sets; words = tokens[:, :-1] (reference :41-55)."""
        rows = np.concatenate([flatten_first_two_dims(episode.support), flatten_first_two_dims(episode.query)], axis=0)
        n = self._stage(rows)
        return self._step(n, 0, self._time_steps - 1, True)

    def eval(self, episode):
        """Query set only; words = tokens[:, 1:] (reference :57-67)."""
        n = self._stage(flatten_first_two_dims(episode.query))
        return self._step(n, 1, self._time_steps, False)

    def sample(self, support_set, num):
        """argmax of the word distribution, `num` times; the support set is ignored (reference :69-78)."""
        torch = self._torch
        s = torch.cuda.current_stream(self._device)
        _lib.check(self._lib.fsmg_unigram_argmax(self._counts.data_ptr(), self._input_size, self._word.data_ptr(), s.cuda_stream))
        word = int(self._word.cpu()[0])
        return [word for _ in range(int(num))]

    # ---- checkpoints: <checkpt_path>/<name>/<name>-<global_step>.npz, variable '<name>/word_count' --------
    def save(self, checkpt_path):
        directory = os.path.join(checkpt_path, self.name)
        if not os.path.exists(directory):
            os.makedirs(directory)
        path = os.path.join(directory, '%s-%d.npz' % (self.name, self._global_step))
        np.savez(path, **{self.name + '/word_count': self._counts.cpu().numpy(),
                          self.name + '/Variable': np.asarray(self._global_step, dtype=np.int64)})
        with open(os.path.join(directory, 'checkpoint'), 'w') as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
        return path

    def recover_or_init(self, init_path, only_load_trainable_vars=False):
        """word_count is NOT trainable: a trainable-only restore leaves it at alpha (reference tf_model.py:112-129)."""
        self._counts.fill_(self.ALPHA)
        if not init_path or only_load_trainable_vars:
            return
        found = glob.glob(os.path.join(init_path, self.name, self.name + '-*.npz'))
        if not found:
            return
        latest = max(found, key=lambda p: int(p.rsplit('-', 1)[1][:-4]))
        blob = np.load(latest)
        key = self.name + '/word_count'
        if key in blob.files and tuple(blob[key].shape) == (self._input_size,):
            self._counts.copy_(self._torch.from_numpy(np.asarray(blob[key], np.float32)))

    # ---- test hooks -----------------------------------------------------------------------------
    @property
    def word_count(self):
        return self._counts.cpu().numpy()
