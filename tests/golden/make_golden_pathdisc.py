#!/usr/bin/env python
"""Golden vectors of the REFERENCE's path discretiser: oracle/_ref/libpathdisc_ref.so (the unmodified
src/nmpc_nav_control/PathDiscretizer.cpp, `make -C oracle ref`, needs /root/reference) run on seeded paths.
    python tests/golden/make_golden_pathdisc.py      -> tests/golden/pathdisc.npz"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import pathcases  # noqa: E402
from oracle import pathdisc  # noqa: E402

if __name__ == "__main__":
    assert pathdisc.build_ref(), "the reference tree is needed to build oracle/_ref"
    paths, pid, u0 = pathcases.cases(seed=2024, n_paths=12, B=48)
    out = {}
    for hol in (0, 1):
        for period, num in ((0.025, 81), (1.0, 12)):
            out[f"poses_h{hol}_T{period}_n{num}"] = np.stack([pathdisc.ref(paths[p], u, period, num, bool(hol)) for p, u in zip(pid, u0)])
    off = np.cumsum([0] + [len(p) for p in paths]).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "pathdisc.npz"), segments=np.concatenate(paths), offsets=off, path_id=pid, u0=u0, **out)
    print("wrote pathdisc.npz", {k: v.shape for k, v in out.items()})
