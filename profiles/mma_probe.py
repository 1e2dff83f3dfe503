"""Prints the TMEM placement of D for tcgen05.mma.cta_group::2 (fsmg_debug_mma_probe): for every CTA and lane, which rows / columns landed where."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "few-shot-music-generation_b200"))
from fsmg import _lib  # noqa: E402

lib = _lib.load()
for M, N in ((256, 64), (128, 64), (128, 128)):
    out = torch.full((2, 128, N), -7.0, device="cuda")
    _lib.check(lib.fsmg_debug_mma_probe(M, N, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    print(f"=== M={M} N={N}")
    for cta in range(2):
        for lane in range(0, 128):
            row = o[cta, lane]
            valid = row >= 0
            if not valid.any():
                desc = "untouched" if (row == -1).all() else f"other {row[:4]}"
            else:
                r = np.unique((row[valid] // 256).astype(int))
                n = (row[valid] % 256).astype(int)
                cols = np.nonzero(valid)[0]
                desc = f"rows {r.tolist()} cols[{cols[0]}..{cols[-1]}] -> n[{n[0]}..{n[-1]}] ({valid.sum()} cells)"
            if lane % 8 == 0 or lane in (1, 63, 65, 127):
                print(f"cta {cta} lane {lane:3d}: {desc}")
