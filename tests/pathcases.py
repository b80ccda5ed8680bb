"""seeded paths and start parameters for the path-discretiser tests (test infrastructure, numpy only)"""
import math

import numpy as np

from nmpc_nav_control_b200 import paths as P


def random_path(rng, n_seg):
    """a C0 chain of lines, arcs and cubic Beziers with per-segment signed speeds"""
    segs = []
    pos = rng.uniform(-1, 1, 2); head = rng.uniform(-math.pi, math.pi)
    for _ in range(n_seg):
        kind = rng.integers(0, 3)
        vel = rng.uniform(0.15, 0.9) * (1.0 if rng.random() < 0.8 else -1.0)
        th0, th1 = rng.uniform(-1, 1, 2)
        if kind == 0:
            L = rng.uniform(0.3, 2.0)
            end = pos + L * np.array([math.cos(head), math.sin(head)])
            segs.append(P.line(pos, end, vel, th0, th1)); pos = end
        elif kind == 1:
            r = rng.uniform(0.3, 1.5); sweep = rng.uniform(0.3, 2.5) * rng.choice([-1.0, 1.0])
            a0 = head - math.copysign(math.pi / 2, sweep)
            c = pos - r * np.array([math.cos(a0), math.sin(a0)])
            segs.append(P.arc(c, r, a0, a0 + sweep, vel, th0, th1))
            pos = c + r * np.array([math.cos(a0 + sweep), math.sin(a0 + sweep)]); head += sweep
        else:
            L = rng.uniform(0.5, 2.0); h1 = head + rng.uniform(-1.2, 1.2)
            p1 = pos + L / 3 * np.array([math.cos(head), math.sin(head)])
            p3 = pos + L * np.array([math.cos(0.5 * (head + h1)), math.sin(0.5 * (head + h1))])
            p2 = p3 - L / 3 * np.array([math.cos(h1), math.sin(h1)])
            segs.append(P.bezier3(pos, p1, p2, p3, vel, th0, th1)); pos = p3; head = h1
    return np.array(segs)


def cases(seed, n_paths, B):
    rng = np.random.default_rng(seed)
    paths = [random_path(rng, int(rng.integers(1, 7))) for _ in range(n_paths)]
    pid = rng.integers(0, n_paths, B).astype(np.int32)
    nseg = np.array([len(paths[p]) for p in pid])
    # start anywhere on the path, some robots almost at its end (padding) and some exactly on a segment boundary
    u0 = rng.uniform(0, 1, B) * nseg
    bnd = rng.random(B) < 0.15
    u0[bnd] = np.floor(u0[bnd])
    end = rng.random(B) < 0.1
    u0[end] = nseg[end] - rng.uniform(0, 0.05, end.sum())
    return paths, pid, u0
