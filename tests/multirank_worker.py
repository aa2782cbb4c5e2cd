"""Worker of tests/test_gpu_multirank.py (launched by torchrun, one process per GPU): the frames of a sharded effect,
gathered on rank 0, must equal the frames rank 0 renders alone from the same cloud -- byte for byte outside the pixels
that fp32 atomic summation order can move by one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ken_burns_effect_b200.utils import common as kb   # noqa: E402
from ken_burns_effect_b200.utils import shard, synthetic   # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = shard.world()
    W, H, focal = 256, 192, 128.0
    zoom = synthetic.default_zoom(W, H)
    settings = {'dblSteps': np.linspace(0, 1, 11).tolist(), 'objectFrom': zoom['objectFrom'], 'objectTo': zoom['objectTo'],
                'dolly': False}
    common = None
    if rank == 0:
        pts, rgb, dep, common = synthetic.scene_cloud(W, H, focal=focal, extra_points=3001)
        common.update(intWidth=W, intHeight=H,
                      tensorInpaPoints=torch.from_numpy(pts).to(dev).view(1, 3, -1),
                      tensorInpaImage=torch.from_numpy(rgb).to(dev).view(1, 3, -1),
                      tensorInpaDepth=torch.from_numpy(dep).to(dev).view(1, 1, -1))
    cloud = shard.broadcast_cloud(common, dev, src=0)
    poses = kb.kenburns_poses(settings, cloud)
    frames = shard.render_sharded(poses, lambda mine: kb.render_poses(settings, cloud, mine, to_host=False))
    ok = True
    if rank == 0:
        alone = kb.render_poses(settings, cloud, poses, to_host=False)
        d = (frames.short() - alone.short()).abs()
        # same bar as tests/test_gpu_frames.py: fp32 atomic order moves a byte by 1 at a few pixels, and a filled hole copies the
        # FARTHER of two end points, which can flip (a whole colour) when their depths agree to the last ulp
        ok = frames.shape == alone.shape and int((d > 1).sum()) <= max(12, 3e-5 * d.numel()) and float((d > 0).float().mean()) < 1e-3
        print(f"multirank: world {world}, frames {tuple(frames.shape)}, max diff {int(d.max())}, differing {float((d > 0).float().mean()):.2e}")
    # the sync-free exchange (preallocated buffer, host-side header) must deliver the same cloud
    ex = shard.CloudExchange(dev, capacity_points=3 * W * H, src=0)
    for _ in range(2):
        c2 = ex.broadcast(common if rank == 0 else None)
        ok = ok and all(torch.equal(c2[k], cloud[k]) for k in ('tensorInpaPoints', 'tensorInpaImage', 'tensorInpaDepth'))
        ok = ok and c2['objectDepthrange'] == cloud['objectDepthrange']
    # Pipeline.__call__ under torchrun: every rank renders its share into a shared pinned host segment, every rank returns all
    # frames, PNGs are written by the rank that rendered them and the video by rank 0
    import tempfile
    from ken_burns_effect_b200.utils.pipeline import Pipeline
    torch.manual_seed(5)
    Wp, Hp = 384, 320
    img, _ = synthetic.synthetic_scene(Wp, Hp, seed=5)
    t = torch.from_numpy(img).permute(2, 0, 1).contiguous().float().div(255).view(1, 3, Hp, Wp)
    out_dir = [tempfile.mkdtemp(prefix="kb200_mr_") if rank == 0 else None]
    dist.broadcast_object_list(out_dir, src=0)
    pipe = Pipeline(model_paths=None, dolly=True, output_frames=True, frames=9)
    zoom_p = synthetic.default_zoom(Wp, Hp, dolly=True)
    frames_p = pipe(t, zoom_p, output_path=out_dir[0])
    ok = ok and len(frames_p) == 9 and frames_p[0].shape == (Hp, Wp, 3)
    dist.barrier()
    if rank == 0:
        import cv2
        st = {'dblSteps': np.linspace(0, 1, 9).tolist(), 'objectFrom': zoom_p['objectFrom'], 'objectTo': zoom_p['objectTo'], 'dolly': True}
        oc = pipe.objectCommon
        alone = kb.render_poses(st, oc, kb.kenburns_poses(st, oc)).numpy()
        d = np.abs(np.stack(frames_p).astype(np.int16) - alone.astype(np.int16))
        ok = ok and int((d > 1).sum()) <= max(12, 1e-4 * d.size) and float((d > 0).mean()) < 1e-3
        names = sorted(os.listdir(os.path.join(out_dir[0], 'frames')), key=lambda s: int(s.split('.')[0]))
        ok = ok and names == [f"{i}.png" for i in range(9)]
        ok = ok and bool((cv2.imread(os.path.join(out_dir[0], 'frames', '4.png')) == frames_p[4]).all())
        ok = ok and os.path.getsize(os.path.join(out_dir[0], '3d_kbe.mp4')) > 1000
        print(f"multirank pipeline: frames vs single-rank render max diff {int(d.max())}, png/mp4 written, timing {pipe.last_timing}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
