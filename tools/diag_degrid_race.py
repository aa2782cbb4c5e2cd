"""GPU diagnostic: how does the reference's in-place degrid race (utils/common.py:556-567) actually resolve
on this GPU, compared with the race-free (Jacobi) and sequential-raster restatements of the oracle?
Run on the GPU box; prints one JSON line per case and writes golden fixtures into gpurun_out/golden/."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import refgpu  # noqa: E402
from ken_burns_effect_b200.utils import common as kb  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kb_helpers as helpers  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)
oracle.set_threads(0)


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


for (W, H, focal, extra) in [(64, 48, 32.0, 517), (256, 192, 128.0, 4099), (1024, 768, 512.0, 0), (1024, 768, 512.0, 70001)]:
    for step in (0.0, 1.0):
        pts, rgb, dep, common = helpers.scene(W, H, focal, extra)
        shifted, sh, f = helpers.shifted_cloud(pts, common, W, H, step)
        data = np.concatenate([rgb, dep], 0)
        tp, td = torch.from_numpy(shifted[None]).cuda(), torch.from_numpy(data[None]).cuda()
        runs = []
        for rep in range(3):
            r_render, r_exist, r_zraw, r_zdeg, r_out = refgpu.render_pointcloud(tp, td, W, H, focal, 120, stages=True)
            runs.append((r_render.cpu().numpy(), r_exist.cpu().numpy(), r_zdeg.cpu().numpy(), r_out.cpu().numpy()))
        o_render, o_exist, o_zraw, o_zj = oracle.render_pointcloud(shifted[None], data[None], W, H, focal, 120, want_zee=True)
        s_render, s_exist, _, o_zs = oracle.render_pointcloud(shifted[None], data[None], W, H, focal, 120, degrid_mode=1, want_zee=True)
        m_render, m_exist, m_zraw, m_zdeg = kb.render_pointcloud(tp, td, W, H, focal, 120, return_zee=True)
        rr, re_, rz, rout = runs[0]          # every saved array comes from the SAME run of the reference
        info = dict(W=W, H=H, extra=extra, step=step,
                    zraw_bitexact_ref_vs_oracle=bool(np.array_equal(r_zraw.cpu().numpy().view(np.int32), o_zraw.view(np.int32))),
                    zraw_bitexact_ref_vs_mine=bool(torch.equal(r_zraw.view(torch.int32), m_zraw.view(torch.int32))),
                    degrid_updated_jacobi=int((o_zj != o_zraw).sum()),
                    jacobi_vs_raster=int((o_zj != o_zs).sum()),
                    ref_vs_jacobi=int((rz != o_zj).sum()), ref_vs_raster=int((rz != o_zs).sum()),
                    ref_run_to_run=[int((runs[i][2] != rz).sum()) for i in (1, 2)],
                    mine_vs_jacobi=int((m_zdeg.cpu().numpy() != o_zj).sum()),
                    holes_ref_vs_jacobi=int(((re_ == 0) != (o_exist == 0)).sum()),
                    holes_ref_vs_raster=int(((re_ == 0) != (s_exist == 0)).sum()),
                    relL2_render_ref_vs_jacobi=rel(o_render, rr), relL2_render_ref_vs_raster=rel(s_render, rr),
                    relL2_render_ref_vs_mine=rel(m_render.cpu().numpy(), rr),
                    relL2_ref_run_to_run=rel(runs[1][0], rr))
        print(json.dumps(info), flush=True)
        if W <= 256:
            np.savez_compressed(os.path.join(OUT, f"ref_render_{W}x{H}_e{extra}_s{int(step)}.npz"),
                                points=shifted.astype(np.float32), data=data.astype(np.float32), focal=focal, baseline=120,
                                zee_raw=r_zraw.cpu().numpy(), zee_degrid=rz, out=rout,
                                render=rr, existing=re_)
