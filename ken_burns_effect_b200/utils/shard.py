"""Frame sharding across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

The frames of a Ken Burns effect are independent given the final point cloud -- the per-frame loop of the reference
(utils/common.py:222-260) reads only tensorInpa{Points,Image,Depth} and a handful of scalars -- so the path shards by
pose with exactly one exchange step: rank `src` runs the CNN stage and the two inpainting passes, then the packed cloud
[xyz(3) | rgb(3) | depth(1)] x N fp32 (28 B/point) and a small header travel to every rank in ONE broadcast each
(NCCL over NVLink on GPUs, gloo in the CPU tests); rank r renders poses r, r+R, r+2R, ... (interleaved: the point count
is common and the hole count varies smoothly along the path, so the shards are balanced) and, when a single writer is
wanted, the uint8 frames are gathered back in pose order.  Nothing here touches the kernels: it is host logic.
"""
import torch
import torch.distributed as dist

HEADER_DOUBLES = 16   # N, H, W, focal, baseline, depth-range min, max, argmin x, y, argmax x, y, dispmin, dispmax, 3 spare


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPU cores nearest to its GPU (NVML's ideal CPU affinity) BEFORE it allocates pinned host
    buffers: with one process per GPU and 2.36 MB per frame going back over PCIe, a staging buffer on the far socket
    halves the copy rate.  Best effort: returns the CPU set, or None when NVML / sched_setaffinity are unavailable."""
    import os
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(int(device_index))
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def world():
    """-> (rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n, rank, world_size):
    """Pose indices rendered by `rank`: interleaved, every index exactly once over the ranks."""
    return list(range(rank, n, world_size))


def pack_cloud(objectCommon):
    """objectCommon (after stage A of process_kenburns) -> (packed [7,N] fp32, header [16] fp64)."""
    pts = objectCommon['tensorInpaPoints']
    N = pts.shape[-1]
    packed = torch.cat([pts.reshape(3, N), objectCommon['tensorInpaImage'].reshape(3, N),
                        objectCommon['tensorInpaDepth'].reshape(1, N)], 0).contiguous().float()
    rng = objectCommon['objectDepthrange']
    hdr = torch.zeros(HEADER_DOUBLES, dtype=torch.float64)
    vals = [N, objectCommon['intHeight'], objectCommon['intWidth'], objectCommon['dblFocal'], objectCommon['dblBaseline'],
            rng[0], rng[1], rng[2][0], rng[2][1], rng[3][0], rng[3][1],
            objectCommon.get('dblDispmin', 0.0), objectCommon.get('dblDispmax', 0.0)]
    hdr[:len(vals)] = torch.tensor(vals, dtype=torch.float64)
    return packed, hdr


def unpack_cloud(packed, hdr):
    """Inverse of pack_cloud: the part of objectCommon the per-frame loop reads."""
    h = hdr.tolist()
    N = int(h[0])
    baseline = h[4]
    return {
        'intHeight': int(h[1]), 'intWidth': int(h[2]), 'dblFocal': h[3],
        'dblBaseline': int(baseline) if float(baseline).is_integer() else baseline,   # the reference keeps an int (pipeline.py:27)
        'objectDepthrange': (h[5], h[6], (int(h[7]), int(h[8])), (int(h[9]), int(h[10]))),
        'dblDispmin': h[11], 'dblDispmax': h[12],
        'tensorInpaPoints': packed[0:3].view(1, 3, N), 'tensorInpaImage': packed[3:6].view(1, 3, N),
        'tensorInpaDepth': packed[6:7].view(1, 1, N),
        'tensorPacked': packed,       # [7,N]: FrameRenderer reads xyz and rgb+depth straight out of it (no copy)
    }


def broadcast_cloud(objectCommon, device, src=0, group=None):
    """The path's one exchange step.  `objectCommon` is only read on rank `src` (may be None elsewhere).
    Returns the unpacked cloud on every rank (on `src` the views alias the packed local buffer)."""
    rank, R = world()
    if R == 1:
        packed, hdr = pack_cloud(objectCommon)
        return unpack_cloud(packed.to(device), hdr)
    hdr = torch.zeros(HEADER_DOUBLES, dtype=torch.float64, device=device)
    packed = None
    if rank == src:
        packed, h = pack_cloud(objectCommon)
        packed = packed.to(device)
        hdr.copy_(h)
    dist.broadcast(hdr, src=src, group=group)          # 128 bytes: sizes the receive buffer
    if rank != src:
        packed = torch.empty(7, int(hdr[0].item()), dtype=torch.float32, device=device)
    dist.broadcast(packed, src=src, group=group)       # 28 * N bytes over NVLink
    return unpack_cloud(packed, hdr.cpu())


def gather_frames(local_frames, n_total, dst=0, group=None):
    """local_frames: uint8 [n_local,H,W,3] holding this rank's poses shard_indices(n_total, rank, R) in order.
    -> on rank `dst`: uint8 [n_total,H,W,3] in pose order; None elsewhere."""
    rank, R = world()
    if R == 1:
        return local_frames
    per = -(-n_total // R)                              # every rank sends `per` frames (last ones padded)
    H, W = local_frames.shape[1:3]
    send = local_frames
    if send.shape[0] < per:
        pad = torch.zeros(per - send.shape[0], H, W, 3, dtype=torch.uint8, device=send.device)
        send = torch.cat([send, pad], 0)
    send = send.contiguous()
    bufs = [torch.empty_like(send) for _ in range(R)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty(n_total, H, W, 3, dtype=torch.uint8, device=send.device)
    for r in range(R):
        idx = shard_indices(n_total, r, R)
        out[idx] = bufs[r][:len(idx)]
    return out


def render_sharded(poses, render_fn, gather=True, dst=0, group=None):
    """Render this rank's interleaved share of `poses` with render_fn(list_of_poses) -> uint8 [n,H,W,3] and, when
    `gather`, reassemble all frames on rank `dst` (other ranks get None); otherwise return (indices, local frames)."""
    rank, R = world()
    idx = shard_indices(len(poses), rank, R)
    local = render_fn([poses[i] for i in idx])
    if not gather:
        return idx, local
    return gather_frames(local, len(poses), dst=dst, group=group)
