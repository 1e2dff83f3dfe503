"""LSTMBaseline — drop-in for the reference's `models.lstm_baseline.LSTMBaseline`
(reference src/models/lstm_baseline.py:8-156), resolved through the same
`model_module_name` / `model_class_name` registry (reference src/train/train.py:12-15,
config/lstm_baseline.yaml:1-3).

Same constructor `(config)`, same five methods, same return conventions; the TensorFlow
session is replaced by the B200 engine (`fsmg.Engine` -> libfsmg.so).  Extensions that default
to reference behaviour: `train` / `eval` also accept a list of episodes (episodes_per_step),
`per_token_nll` exposes the parity quantity, `sample_batch` decodes many songs at once.
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
from fsmg.engine import Engine


def _as_list(episode):
    return list(episode) if isinstance(episode, (list, tuple)) else [episode]


class LSTMBaseline(BaseModel):
    """LSTM language model.  Trained on songs from the meta-training set; during evaluation
    the support set is ignored and only the query set is scored (reference :9-13)."""

    MAX_TO_KEEP = 10  # tf.train.Saver(max_to_keep=10), reference tf_model.py:96-97

    def __init__(self, config):
        super(LSTMBaseline, self).__init__(config)
        # reference _define_placedholders (:18-29)
        self._start_word = config['input_size']
        self._input_size = config['input_size'] + 1
        self._time_steps = config['max_len']
        self._embd_size = config['embedding_size']
        self._hidden_size = config['hidden_size']
        self._n_layers = config['n_layers']
        self._seed = config.get('seed', 1234)
        seqs_per_episode = config.get('batch_size', 5) * (config.get('support_size', 5) + config.get('query_size', 4))
        self._episodes_per_step = int(config.get('episodes_per_step', 1))
        max_seqs = int(config.get('max_seqs', max(seqs_per_episode * self._episodes_per_step,
                                                  config.get('sample_batch', 1))))
        self._engine = Engine(config, max_seqs=max_seqs, flags=int(config.get('fsmg_flags', 0)))
        self._initialized = False
        # summaries: tags of the reference (lstm_baseline.py:106-111,126-131; tf_model.py:84-88)
        self._summary_writer = None
        self._train_calls = 0
        self._eval_calls = 0
        if config.get('checkpt_dir') and config.get('tensorboard', True) and self._rank() == 0:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._summary_writer = SummaryWriter(config['checkpt_dir'])
            except Exception:  # tensorboard is optional plumbing
                self._summary_writer = None

    @staticmethod
    def _rank():
        import torch.distributed as dist
        return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0

    # ---- token assembly (reference :91-96 and :117-118) ------------------------------------------
    @staticmethod
    def _train_tokens(episodes):
        rows = []
        for ep in episodes:  # support rows first, then query rows, per episode
            rows.append(flatten_first_two_dims(ep.support))
            rows.append(flatten_first_two_dims(ep.query))
        return np.concatenate(rows, axis=0)

    @staticmethod
    def _eval_tokens(episodes):
        return np.concatenate([flatten_first_two_dims(ep.query) for ep in episodes], axis=0)

    @staticmethod
    def _indexed(episodes):
        """All episodes of the step index ONE device-resident corpus (data.device_episode.IndexedEpisode)?"""
        first = getattr(episodes[0], 'sampler', None)
        return first is not None and all(getattr(ep, 'sampler', None) is first and hasattr(ep, 'support_ids') for ep in episodes)

    @staticmethod
    def _train_ids(episodes):
        rows = []
        for ep in episodes:  # same row order as _train_tokens
            rows.append(np.asarray(ep.support_ids).reshape(-1))
            rows.append(np.asarray(ep.query_ids).reshape(-1))
        return np.concatenate(rows)

    def _ensure_init(self):
        if not self._initialized:
            self.recover_or_init('')

    # ---- BaseModel API ----------------------------------------------------------------------------
    def train(self, episode):
        """Concatenate support and query sets and take one optimizer step; returns the loss."""
        self._ensure_init()
        episodes = _as_list(episode)
        if self._indexed(episodes):   # device-resident corpus: only the song indices are uploaded
            loss = self._engine.train_indexed(episodes[0].corpus_device, self._train_ids(episodes))
        else:
            blocks = []
            for ep in episodes:     # support rows first, then query rows, per episode (reference :95-96)
                blocks.append(flatten_first_two_dims(ep.support))
                blocks.append(flatten_first_two_dims(ep.query))
            loss = self._engine.train_host_rows(blocks)
        if self._summary_writer:
            self._summary_writer.add_scalar('Train/loss', loss, self._train_calls)
        self._train_calls += 1
        return loss

    def eval(self, episode):
        """Ignore the support set and evaluate only on the query set; returns mean NLL."""
        self._ensure_init()
        episodes = _as_list(episode)
        if self._indexed(episodes):
            ids = np.concatenate([np.asarray(ep.query_ids).reshape(-1) for ep in episodes])
            avg_neg_log = self._engine.eval_indexed(episodes[0].corpus_device, ids)
        else:
            avg_neg_log = self._engine.eval_host(self._eval_tokens(episodes))
        if self._summary_writer is not None:
            self._summary_writer.add_scalar('Eval/Avg_NLL', avg_neg_log, self._eval_calls)
        self._eval_calls += 1
        return avg_neg_log

    def per_token_nll(self, tokens):
        """tokens [..., T] int -> per-token NLL [N, T] (the parity quantity of SURVEY §8c)."""
        self._ensure_init()
        return self._engine.eval_host(np.asarray(tokens), return_nll=True)[1]

    def sample(self, support_set, num):
        """Greedy decode of `num` tokens; the support set is ignored (reference :135-156)."""
        self._ensure_init()
        return [int(w) for w in self._engine.sample_host(1, int(num))[0]]

    def sample_batch(self, n_songs, num):
        self._ensure_init()
        return self._engine.sample_host(int(n_songs), int(num))

    # ---- checkpoints: <checkpt_path>/<name>/<name>-<global_step>.npz (reference tf_model.py:106-129) ---
    def _checkpt_dir(self, checkpt_path):
        directory = os.path.join(checkpt_path, self.name)
        if not os.path.exists(directory):
            os.makedirs(directory)
        return directory

    def save(self, checkpt_path):
        """Parameters, Adam slots and global_step, keyed by the TF variable names (A.1)."""
        if self._rank() != 0:
            return None
        self._ensure_init()
        eng = self._engine
        directory = self._checkpt_dir(checkpt_path)
        blob = {}
        for k, v in eng.export('params').items():
            blob[k] = v
        for k, v in eng.export('adam_m').items():
            blob[k + '/Adam'] = v
        for k, v in eng.export('adam_v').items():
            blob[k + '/Adam_1'] = v
        blob[self.name + '/Variable'] = np.asarray(eng.global_step, dtype=np.int64)
        path = os.path.join(directory, '%s-%d.npz' % (self.name, eng.global_step))
        np.savez(path, **blob)
        kept = sorted(glob.glob(os.path.join(directory, self.name + '-*.npz')),
                      key=lambda p: int(p.rsplit('-', 1)[1][:-4]))
        for old in kept[:-self.MAX_TO_KEEP]:
            os.remove(old)
        with open(os.path.join(directory, 'checkpoint'), 'w') as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(path))
        return path

    def _latest_checkpoint(self, checkpt_path):
        if not checkpt_path:
            return None
        directory = os.path.join(checkpt_path, self.name)
        index = os.path.join(directory, 'checkpoint')
        if os.path.isfile(index):
            line = open(index).readline()
            cand = os.path.join(directory, line.split('"')[1]) if '"' in line else None
            if cand and os.path.isfile(cand):
                return cand
        found = glob.glob(os.path.join(directory, self.name + '-*.npz'))
        return max(found, key=lambda p: int(p.rsplit('-', 1)[1][:-4])) if found else None

    def _recover(self, checkpt_path, only_load_trainable_vars=False):
        latest = self._latest_checkpoint(checkpt_path)
        if latest is None:
            return False
        print('recovering %s from %s' % (self.name, latest))
        blob = np.load(latest)
        eng = self._engine
        # optimistic restore: only variables whose name AND shape match (reference tf_model.py:28-75)
        loaded = eng.load_params({k: blob[k] for k in blob.files}, strict=False)
        if not only_load_trainable_vars:
            import torch
            for which, suffix in (('adam_m', '/Adam'), ('adam_v', '/Adam_1')):
                for name, view in eng.param_views(which).items():
                    key = name + suffix
                    if key in blob.files and tuple(blob[key].shape) == tuple(view.shape):
                        view.copy_(torch.from_numpy(np.asarray(blob[key], np.float32)))
            if self.name + '/Variable' in blob.files:
                eng.global_step = int(blob[self.name + '/Variable'])
        self._restored = set(loaded)
        return True

    def recover_or_init(self, init_path, only_load_trainable_vars=False):
        """Restore what a checkpoint under init_path provides and initialise the rest (reference tf_model.py:112-129: restore,
        then run the initialiser of the variables that are STILL uninitialised).  On a fresh model every variable is first
        drawn Glorot-uniform from config['seed'] (A.1) with zero Adam slots and global_step 0; on a model that already holds
        weights (trained, or restored before) nothing is re-drawn — only the checkpoint, if any, is applied."""
        if not self._initialized:
            self._engine.init_params(int(self._seed))
            self._engine.adam_m.zero_()
            self._engine.adam_v.zero_()
            self._engine.global_step = 0
            self._initialized = True
        self._recover(init_path, only_load_trainable_vars)

    # ---- test / parity hooks ----------------------------------------------------------------------
    def set_params(self, params):
        """Inject weights by TF variable name (used for parity against the oracle)."""
        self._engine.load_params(params)
        self._initialized = True

    def get_params(self):
        return self._engine.export('params')

    @property
    def engine(self):
        return self._engine

    @property
    def global_step(self):
        return self._engine.global_step
