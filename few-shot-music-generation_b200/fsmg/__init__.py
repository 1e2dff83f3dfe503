"""fsmg — B200-native engine behind models.lstm_baseline.LSTMBaseline (see include/fsmg.h)."""
from ._lib import FsmgError, FSMG_FLAG_SIMT_GEMM, FSMG_FLAG_SIMT_RECURRENT, LIB_PATH  # noqa: F401


def __getattr__(name):  # lazy: importing fsmg must not require torch+CUDA until the engine is used
    if name in ("Engine", "glorot_uniform_init"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
