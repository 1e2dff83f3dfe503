"""CPU restatement of the reference's unigram baseline — TEST INFRASTRUCTURE ONLY (imported by tests/ and nothing else).

Follows /root/reference/src/models/unigram_model.py:
  :18-24  words placeholder is [None, max_len - 1]; alpha = 1
  :26-39  word_count = get_variable('word_count', [input_size], constant_initializer(alpha), trainable=False) (fp32);
          train_op = scatter_add(word_count, words.flatten(), ones); prob = gather(word_count, words) / reduce_sum(word_count);
          avg_neg_log = -reduce_mean(log(prob)); prob_all = word_count / sum
  :41-55  train(): X = tokens[:, :-1] of support and of query (convert_tokens_to_input_and_target without a start word,
          base_model.py:63-86), concatenated; returns the loss of sess.run([train_op, avg_neg_log])
  :57-67  eval(): Y = tokens[:, 1:] of the QUERY set only
  :69-78  sample(): argmax(prob_all) `num` times

Parity unpinned in the strict sense (TensorFlow 1.x is not installable here).  The evaluation order of
train_op and avg_neg_log inside one sess.run is unspecified in TF1; this restatement (and the CUDA path) takes the loss on the
counts BEFORE the update."""
import numpy as np


class UnigramOracle:
    def __init__(self, input_size: int, alpha: float = 1.0, dtype=np.float32):
        self.dtype = dtype
        self.word_count = np.full((input_size,), alpha, dtype=dtype)

    def _avg_neg_log(self, words: np.ndarray) -> float:
        total = self.word_count.astype(np.float64).sum()
        prob = self.word_count[words.reshape(-1)].astype(np.float64) / total
        return float(-np.mean(np.log(prob)))

    def train(self, support: np.ndarray, query: np.ndarray) -> float:
        x = np.concatenate([support.reshape(-1, support.shape[-1])[:, :-1], query.reshape(-1, query.shape[-1])[:, :-1]])
        loss = self._avg_neg_log(x)
        np.add.at(self.word_count, x.reshape(-1), self.dtype(1.0))
        return loss

    def eval(self, query: np.ndarray) -> float:
        return self._avg_neg_log(query.reshape(-1, query.shape[-1])[:, 1:])

    def sample(self, num: int):
        return [int(np.argmax(self.word_count))] * num
