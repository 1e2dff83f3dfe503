"""Model plugin contract — the five-method API every model of the reference implements
(reference src/models/base_model.py:4-54) plus its token helpers (:57-86).

Kept so that `train.train` and third-party models written against the reference keep working;
the CUDA LSTMBaseline does the start-word shift on the device instead (csrc: prep_tokens_kernel).
"""
import numpy as np


class BaseModel(object):
    """train / eval / sample / save / recover_or_init + the `name` property."""

    def __init__(self, config):
        self._config = config

    @property
    def name(self):
        return self._config['name']

    def train(self, episode):
        raise NotImplementedError()

    def eval(self, episode):
        raise NotImplementedError()

    def sample(self, support_set, num):
        raise NotImplementedError()

    def save(self, checkpt_path):
        raise NotImplementedError()

    def recover_or_init(self, init_path):
        raise NotImplementedError()


def flatten_first_two_dims(token_array):
    """[B,S,T] -> [B*S,T]."""
    token_array = np.asarray(token_array)
    return token_array.reshape((-1, token_array.shape[-1]))


def convert_tokens_to_input_and_target(token_array, start_word=None):
    """Host version of the input/target split: with a start word X=[start|tok[:-1]], Y=tok;
    without, X=tok[:-1], Y=tok[1:]."""
    flat = flatten_first_two_dims(token_array)
    if start_word is None:
        return flat[:, :-1], flat[:, 1:].copy()
    x = np.empty_like(flat)
    x[:, 0] = start_word
    x[:, 1:] = flat[:, :-1]
    return x, flat.copy()
