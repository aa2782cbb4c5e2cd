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
    return packed, cloud_header(objectCommon)


def cloud_header(objectCommon):
    """The scalars of objectCommon the per-frame loop reads, as 16 doubles (host tensor)."""
    N = objectCommon['tensorInpaPoints'].shape[-1]
    rng = objectCommon['objectDepthrange']
    hdr = torch.zeros(HEADER_DOUBLES, dtype=torch.float64)
    vals = [N, objectCommon['intHeight'], objectCommon['intWidth'], objectCommon['dblFocal'], objectCommon['dblBaseline'],
            rng[0], rng[1], rng[2][0], rng[2][1], rng[3][0], rng[3][1],
            objectCommon.get('dblDispmin', 0.0), objectCommon.get('dblDispmax', 0.0)]
    hdr[:len(vals)] = torch.tensor(vals, dtype=torch.float64)
    return hdr


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


class CloudExchange:
    """The exchange step without a stream drain: preallocated receive buffers, the 128-byte header on a HOST side channel
    (a gloo group: the non-source ranks learn N without reading anything back from their GPU, so the kernels of the previous
    effect keep running while the host already posts the next broadcast), then ONE broadcast of exactly 28*N payload bytes.
    On the source rank the cloud is sent from where it lies when it is already packed ([7,N] contiguous,
    objectCommon['tensorPacked']), else three device copies place it in the buffer (no torch.cat, no fresh allocation).

    On CUDA the broadcast runs on its own stream into one of TWO buffers, so the cloud of effect i+1 can travel while effect i
    still renders out of the other buffer (broadcast_async + wait(cloud) where the cloud is first read; broadcast() = both at
    once).  Call consumed(cloud) after enqueuing the work that reads a cloud: its buffer is rewritten only after that mark.
    """

    def __init__(self, device, capacity_points, src=0, group=None):
        self.device, self.src, self.group = torch.device(device), src, group
        self.cuda = self.device.type == 'cuda'
        nbuf = 2 if self.cuda else 1
        self.bufs = [torch.empty(7 * int(capacity_points), dtype=torch.float32, device=device) for _ in range(nbuf)]
        self.slot = 0
        self.stream = torch.cuda.Stream(self.device) if self.cuda else None
        self.read_done = [None] * nbuf                     # event: the last reader of this buffer has been enqueued and passed
        self.in_use = [False] * nbuf                       # broadcast into, not yet marked consumed
        self.host_group = None
        rank, R = world()
        if R > 1:
            backend = dist.get_backend(group)
            self.host_group = group if backend == 'gloo' else dist.new_group(backend='gloo')

    @property
    def buf(self):
        return self.bufs[0]

    def _fit(self, slot, n):
        if 7 * n > self.bufs[slot].numel():                # rare: a cloud larger than promised -- grow once, keep it
            self.bufs[slot] = torch.empty(7 * n, dtype=torch.float32, device=self.device)

    def consumed(self, cloud):
        """Mark on the current stream: everything enqueued so far has finished reading `cloud` (a result of broadcast*)."""
        slot = cloud.get('_slot')
        if self.cuda and slot is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.read_done[slot] = ev
            self.in_use[slot] = False

    def wait(self, cloud):
        """Make the current stream wait for the arrival of `cloud` (needed once, before its first reader, after broadcast_async)."""
        ev = cloud.get('_ready')
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        return cloud

    def broadcast(self, objectCommon):
        """objectCommon: read on the source rank only (None elsewhere) -> the unpacked cloud on every rank, ordered before
        everything enqueued on the current stream afterwards."""
        return self.wait(self.broadcast_async(objectCommon))

    def broadcast_async(self, objectCommon):
        """Like broadcast(), but the current stream does not wait: call wait(cloud) where the cloud is first read."""
        rank, R = world()
        if R == 1:
            packed, hdr = pack_cloud(objectCommon)
            return unpack_cloud(packed.to(self.device), hdr)
        slot = self.slot
        self.slot = (self.slot + 1) % len(self.bufs)
        hdr = torch.zeros(HEADER_DOUBLES, dtype=torch.float64)
        send = None
        main = torch.cuda.current_stream(self.device) if self.cuda else None
        if self.cuda:
            if self.in_use[slot] or self.read_done[slot] is None:
                self.stream.wait_stream(main)                           # the reader never marked (or first use): stream order
            else:
                self.stream.wait_event(self.read_done[slot])           # the previous content of this buffer is no longer read
        if rank == self.src:
            hdr = cloud_header(objectCommon)
            n = int(hdr[0])
            pk = objectCommon.get('tensorPacked')
            if pk is not None and pk.is_contiguous() and tuple(pk.shape) == (7, n) and pk.device == self.bufs[slot].device:
                send = pk.view(-1)
                if self.cuda:
                    self.stream.wait_stream(main)                       # whatever produced the packed cloud
            else:
                self._fit(slot, n)
                send = self.bufs[slot][:7 * n]
                v = send.view(7, n)
                if self.cuda:
                    self.stream.wait_stream(main)                       # stage A produced the tensors on the caller's stream
                with (torch.cuda.stream(self.stream) if self.cuda else _null()):
                    v[0:3].copy_(objectCommon['tensorInpaPoints'].reshape(3, n))
                    v[3:6].copy_(objectCommon['tensorInpaImage'].reshape(3, n))
                    v[6:7].copy_(objectCommon['tensorInpaDepth'].reshape(1, n))
        dist.broadcast(hdr, src=self.src, group=self.host_group)            # host to host: no GPU involved
        n = int(hdr[0])
        if rank != self.src:
            self._fit(slot, n)
            send = self.bufs[slot][:7 * n]
        with (torch.cuda.stream(self.stream) if self.cuda else _null()):
            dist.broadcast(send, src=self.src, group=self.group)            # 28 * N bytes over NVLink, on the exchange stream
        cloud = unpack_cloud(send.view(7, n), hdr)
        cloud['_slot'] = slot
        cloud['_ready'] = None
        if self.cuda:
            cloud['_ready'] = torch.cuda.Event()
            cloud['_ready'].record(self.stream)
            self.in_use[slot] = True
        return cloud


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class SharedFrames:
    """Host frames of one effect in a POSIX shared-memory segment that every rank of the box maps and page-locks: rank r
    copies ITS frames device-to-host over ITS OWN PCIe link straight into its block, and the rank that writes the video reads
    all blocks from host memory -- no gather through one GPU, no second host copy.  Block r holds the frames of the poses
    r, r+R, r+2R, ... in order (shard_indices)."""

    def __init__(self, n_total, H, W, tag="kb200"):
        import os
        rank, R = world()
        self.rank, self.R, self.n_total, self.H, self.W = rank, R, int(n_total), int(H), int(W)
        self.counts = [len(range(r, self.n_total, R)) for r in range(R)]
        self.offsets = [sum(self.counts[:r]) for r in range(R)]
        self.frame_bytes = self.H * self.W * 3
        nbytes = max(1, self.n_total * self.frame_bytes)
        port = os.environ.get("MASTER_PORT", "0")
        self.path = f"/dev/shm/{tag}_{port}_{self.n_total}x{self.H}x{self.W}"
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        if R > 1:
            dist.barrier()
        self.flat = torch.from_file(self.path, shared=True, size=nbytes, dtype=torch.uint8)
        self.pinned = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.flat.data_ptr(), nbytes, 0)
            self.pinned = int(rc) == 0
        if R > 1:
            dist.barrier()
        if rank == 0:
            try:
                os.unlink(self.path)              # the mappings keep the memory alive; nothing is left behind in /dev/shm
            except OSError:
                pass

    def block(self, r=None):
        """uint8 [n_r,H,W,3] view of rank r's block (default: this rank's)."""
        r = self.rank if r is None else r
        a = self.offsets[r] * self.frame_bytes
        return self.flat[a:a + self.counts[r] * self.frame_bytes].view(self.counts[r], self.H, self.W, 3)

    def frame(self, i):
        """uint8 [H,W,3] numpy view of pose i, wherever it was rendered."""
        r, j = i % self.R, i // self.R
        return self.block(r)[j].numpy()

    def close(self):
        if self.pinned:
            torch.cuda.cudart().cudaHostUnregister(self.flat.data_ptr())
            self.pinned = False


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
