import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/few-shot-music-generation_b200'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from oracle import lstm_oracle as O
from fsmg.engine import Engine
for L in (1, 2):
    cfg = dict(name="lstm_baseline", input_size=50, embedding_size=12, hidden_size=10, n_layers=L, max_len=6, lr=5e-3, n_decay=10000, max_grad_norm=5)
    params = O.glorot_init(cfg, 1234)
    eng = Engine(cfg, max_seqs=8, device="cuda:0", flags=int(os.environ.get("FL", "0")))
    eng.load_params(params)
    got = eng.sample_host(3, 8)[0].tolist()
    d = O.greedy_deficits(params, got, np.float64)
    print("L", L, "zero", os.environ.get("FSMG_SAMPLE_ZERO", "0"), "got", got, "max deficit %.2e" % d.max())
    eng.close()
