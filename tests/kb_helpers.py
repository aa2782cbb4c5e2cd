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


# ---- deterministic, construction-order independent weights (shared with tests/golden/make_golden_cpu.py) ----
def deterministic_state(model):
    """Fill every entry of model.state_dict() from a generator seeded by the entry's NAME, so the reference
    module (in the golden generator) and the mirror (in the tests) get identical weights without relying on
    RNG consumption order."""
    import zlib

    import torch
    sd = model.state_dict()
    new = {}
    for k, v in sd.items():
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        if k.endswith('num_batches_tracked'):
            new[k] = v.clone()
        elif k.endswith('running_var'):
            new[k] = 1.0 + 0.2 * torch.rand(v.shape, generator=g)
        elif k.endswith('running_mean'):
            new[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif v.dim() == 4:                                  # conv weight
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            new[k] = torch.randn(v.shape, generator=g) * (1.6 / fan_in) ** 0.5
        elif v.dim() == 1 and ('moduleMain' in k or 'p_relu' in k or 'moduleContext' in k) and k.endswith('weight') \
                and _is_prelu(model, k):
            new[k] = 0.25 + 0.1 * torch.randn(v.shape, generator=g)
        elif k.endswith('weight'):                          # batch-norm scale
            new[k] = 1.0 + 0.1 * torch.randn(v.shape, generator=g)
        else:                                               # biases
            new[k] = 0.05 * torch.randn(v.shape, generator=g)
    model.load_state_dict(new)
    return model


def _is_prelu(model, key):
    import torch
    mod = model
    for part in key.split('.')[:-1]:
        mod = mod._modules[part]
    return isinstance(mod, torch.nn.PReLU)
