"""Shared helpers for the parity tests: seeded scenes at the shapes oracle/build_ref.py compiled the
reference kernels for, and a shifted cloud at a chosen step of the default camera path."""
import numpy as np

from ken_burns_effect_b200.utils import synthetic
from ken_burns_effect_b200.utils import common as kb


def scene(W, H, focal, extra=0, seed=1234, baseline=120):
    pts, rgb, dep, common = synthetic.scene_cloud(W, H, seed=seed, focal=focal, baseline=baseline, extra_points=extra)
    return pts, rgb, dep, common


def settings(common, W, H, steps, dolly=False):
    zoom = synthetic.default_zoom(W, H, dolly)
    return {'dblSteps': steps, 'objectFrom': zoom['objectFrom'], 'objectTo': zoom['objectTo'], 'dolly': dolly}


def shifted_cloud(pts, common, W, H, step, dolly=False):
    """Points after process_shift at `step` of the default path, computed with the CPU oracle."""
    import oracle
    st = settings(common, W, H, [step], dolly)
    (sh, focal), = kb.kenburns_poses(st, common)
    return oracle.shift_points(pts, sh), sh, focal


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
