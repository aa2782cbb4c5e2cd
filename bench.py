#!/usr/bin/env python3
"""bench.py -- novel-view frames/sec of the per-frame render loop (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames 150] [--impl b200|reference]

One "step" = one 150-frame Ken Burns effect of a synthetic 1024x768 cloud rendered by this rank through the
fused loop (utils/common.py:222-260 of the reference; here kb_render_frames): process_shift -> z-buffered
splat -> degrid -> accumulate -> normalise -> disocclusion fill -> uint8 -> crop -> resize, 150 poses.
  value  : frames/s, point cloud resident in HBM, frames left in HBM           (kernel-side number)
  e2e    : frames/s through the public FrameRenderer with HOST buffers: H2D of the cloud from pinned memory
           and D2H of every uint8 frame into pinned memory inside the timed region (N = 1: the next effect's cloud
           is uploaded on its own stream while the current effect's frames copy out)
  N > 1  : weak scaling -- every rank renders its own 150-pose shard of a 150*N-pose effect after one NCCL
           broadcast of the packed cloud per step (the path's only exchange step); value = all ranks' frames
           / max-over-ranks device time.
`--impl reference` times the CPU restatement of the reference's kernels (oracle/, all host threads) on the
same workload -- the reference itself has no CPU render path (its kernels are cupy-only).
"""
import argparse
import json
import os

os.environ.setdefault("KB200_RANDOM_VGG", "1")   # synthetic weights: there are no checkpoints offline (explicit opt-in)
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H, FOCAL, BASELINE = 1024, 768, 512.0, 120
# The points the two inpainting passes append (common.py:75-80) come from a numpy stand-in for the CNN
# (synthetic._standin_inpaint): disoccluded pixels of both extreme views, back-projected at background depth.
METRIC = "novel-view frames/sec at 1024x768, 150-frame KBE"


DOLLY = False


def _dtype_cnn():
    """Operand type of the convolution stack in this process (KB200_CONV_F16=1 opts into fp16 operands)."""
    from ken_burns_effect_b200.utils import convstack as cs
    if cs.F16_ENABLED:
        return "f16 operands in the GridNets / Refine (tcgen05 kind::f16, fp32 accumulate; KB200_CONV_F16=1), tf32 elsewhere"
    return "tf32 (tcgen05 kind::tf32, fp32 accumulate)"


def workload_name(frames):
    """config.workload, identical in both arms (the driver compares the strings)."""
    cfg = (2 if DOLLY else 1) if (W, H) == (1024, 768) else 3
    return (f"kbe {'--dolly ' if DOLLY else ''}{W}x{H} -> {frames}-frame 3D KBE (configs[{cfg}]), per-frame render loop "
            "(process_shift..resize, utils/common.py:222-260)")


def build_workload(frames, world=1, rank=0):
    from ken_burns_effect_b200.utils import common as kb
    from ken_burns_effect_b200.utils import synthetic
    # dolly mode skips the inpainting passes (utils/common.py:217-218): the cloud is the raw H*W grid and the focal length
    # changes with every pose (:225-229)
    pts, rgb, dep, common = synthetic.scene_cloud(W, H, seed=1234, focal=FOCAL, baseline=BASELINE,
                                                  inpaint_standin=not DOLLY)
    zoom = synthetic.default_zoom(W, H, dolly=DOLLY)
    steps = np.linspace(0.0, 1.0, frames * world).tolist()[rank::world]
    st = {'dblSteps': steps, 'objectFrom': zoom['objectFrom'], 'objectTo': zoom['objectTo'], 'dolly': DOLLY}
    poses = kb.kenburns_poses(st, common)
    cw = max(zoom['objectFrom']['intCropWidth'], zoom['objectTo']['intCropWidth'])
    ch = max(zoom['objectFrom']['intCropHeight'], zoom['objectTo']['intCropHeight'])
    return pts, rgb, dep, common, poses, (cw, ch)


_CPU_WORKLOAD = None


def cpu_baseline(budget_s=12.0, max_frames=300, threads=0):
    """Time the CPU oracle (port of the reference kernels + its numpy/OpenCV tail) with all host threads on a
    bounded sample of the same workload: poses of the 150-pose path in an evenly spread order until
    `budget_s` seconds of CPU work or `max_frames` frames.  -> (frames/s, threads, seconds, frames)."""
    global _CPU_WORKLOAD
    import oracle
    if _CPU_WORKLOAD is None:
        _CPU_WORKLOAD = build_workload(150)
    pts, rgb, dep, common, poses, (cw, ch) = _CPU_WORKLOAD
    oracle.set_threads(threads if threads > 0 else (os.cpu_count() or 1))
    data = np.concatenate([rgb, dep], 0)
    order = [(i * 37) % len(poses) for i in range(max_frames)]      # 37 is coprime with 150: spread over the path
    oracle.frame(oracle.shift_points(pts, poses[0][0]), data, W, H, poses[0][1], BASELINE, cw, ch)  # warm
    t0 = time.perf_counter()
    n = 0
    for i in order:
        sh, f = poses[i]
        oracle.frame(oracle.shift_points(pts, sh), data, W, H, f, BASELINE, cw, ch)
        n += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, oracle.max_threads(), dt, n


def full_pipeline(steps=3, warmup=3, frames=150):
    """BASELINE configs[1] end to end: one 1024x768 image in HOST memory -> Semantics/Disparity/Refine (tcgen05 convs) ->
    two pointcloud_inpainting passes -> 150 rendered uint8 frames in pinned HOST memory, through Pipeline (the call
    kbe.py makes).  Random-init weights (no checkpoints offline), synthetic image.  -> dict for the JSON line."""
    import torch
    from ken_burns_effect_b200.utils import common as kb
    from ken_burns_effect_b200.utils import synthetic
    from ken_burns_effect_b200.utils.pipeline import Pipeline
    torch.manual_seed(1234)
    img, _ = synthetic.synthetic_scene(W, H, seed=1234)
    t = torch.from_numpy(img).permute(2, 0, 1).float().div(255).view(1, 3, H, W).pin_memory()
    pipe = Pipeline(model_paths=None, dolly=False, frames=frames)
    zoom = synthetic.default_zoom(W, H)
    settings = {'dblSteps': np.linspace(0.0, 1.0, frames).tolist(), 'objectFrom': zoom['objectFrom'],
                'objectTo': zoom['objectTo'], 'boolInpaint': True, 'dolly': False}
    t_cnn = t_render = 0.0
    npts = 0
    for i in range(warmup + steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.estimate_depth(t)
        kb.prepare_cloud(settings, pipe.objectCommon, pipe.moduleInpaint)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        poses = kb.kenburns_poses(settings, pipe.objectCommon)
        out = kb.render_poses(settings, pipe.objectCommon, poses)      # synchronises; frames in pinned host memory
        t2 = time.perf_counter()
        if i >= warmup:
            t_cnn += t1 - t0
            t_render += t2 - t1
        npts = pipe.objectCommon['tensorInpaPoints'].shape[-1]
        assert out.shape == (frames, H, W, 3)
    total = t_cnn + t_render
    return {"value": frames * steps / total, "unit": "frames/s", "ms_per_kbe": 1e3 * total / steps,
            "ms_cnn_and_inpaint_stage": 1e3 * t_cnn / steps, "ms_render_loop": 1e3 * t_render / steps, "points": int(npts),
            "conv_tflop_per_kbe": 2.30, "dtype_cnn": _dtype_cnn() + "; the render loop is f32", "note": "random-init weights: the disparity is noise-like, so the appended point "
            "count and hole statistics are not those of a trained model; CNN forwards replay from CUDA graphs from their 4th call on (captured during warm-up)"}


def _max_over_ranks(ms, world, dev):
    import torch
    import torch.distributed as dist
    if world == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


DEPTH_BATCH = 4      # images whose depth stage runs as one batched forward in throughput mode (Pipeline.estimate_depth_batch)


def throughput_mode(rank, world, dev, images_per_gpu=8, frames=150):
    """BASELINE configs[4]: a batch of images (64 at 8 GPUs), one full 150-frame KBE each, IMAGES sharded over the GPUs: every rank
    runs the whole pipeline (depth CNNs, two inpainting passes, render loop, frames into pinned host memory) on its own images --
    replicas, no collective on the data path, one barrier on each side of the timed region.  Images start in pinned host memory."""
    import torch
    import torch.distributed as dist
    from ken_burns_effect_b200.utils import common as kb
    from ken_burns_effect_b200.utils import synthetic
    from ken_burns_effect_b200.utils.pipeline import Pipeline
    torch.manual_seed(1234)
    imgs = []
    for i in range(images_per_gpu):
        img, _ = synthetic.synthetic_scene(1024, 768, seed=1234 + rank * images_per_gpu + i)
        imgs.append(torch.from_numpy(img).permute(2, 0, 1).contiguous().float().div(255).view(1, 3, 768, 1024).pin_memory())
    pipe = Pipeline(model_paths=None, dolly=False, frames=frames)
    zoom = synthetic.default_zoom(1024, 768)
    settings = {'dblSteps': np.linspace(0.0, 1.0, frames).tolist(), 'objectFrom': zoom['objectFrom'],
                'objectTo': zoom['objectTo'], 'boolInpaint': True, 'dolly': False}

    def one(t):
        pipe.estimate_depth(t)
        kb.prepare_cloud(settings, pipe.objectCommon, pipe.moduleInpaint)
        out = kb.render_poses(settings, pipe.objectCommon, kb.kenburns_poses(settings, pipe.objectCommon))   # synchronises
        return out

    for t in imgs[:4]:                      # warm-up: weight packing, CUDA-graph capture of the forwards, pinned pool
        one(t)
    for _ in range(3):
        pipe.run_many(imgs[:DEPTH_BATCH], zoom, keep=False, depth_batch=DEPTH_BATCH)   # incl. graph capture at batch size
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for t in imgs:                          # (a) one image after the other, the way kbe.py would be called in a loop
        one(t)
    torch.cuda.synchronize()
    ms_serial = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    pipe.run_many(imgs, zoom, keep=False, depth_batch=DEPTH_BATCH)   # (b) Pipeline.run_many: batched depth stage, CNN stage of image i+1 overlaps the frame loop of image i
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    if world > 1:
        dist.barrier()
    ms = _max_over_ranks(ms, world, dev)
    ms_serial = _max_over_ranks(ms_serial, world, dev)
    n_img = images_per_gpu * world
    return {"what": f"configs[4]: {n_img} images of 1024x768, one {frames}-frame KBE each, {images_per_gpu} images per GPU, full pipeline "
                    f"per rank (image in pinned host memory -> frames in pinned host memory), tcgen05 convolutions, depth stage in batches of {DEPTH_BATCH}",
            "images_per_s": n_img / (ms / 1e3), "frames_per_s": n_img * frames / (ms / 1e3), "ms_per_image_per_gpu": ms / images_per_gpu,
            "one_image_at_a_time": {"images_per_s": n_img / (ms_serial / 1e3), "ms_per_image_per_gpu": ms_serial / images_per_gpu},
            "images": n_img, "dtype_cnn": _dtype_cnn(), "weights": "random-init (no checkpoints offline)"}


def config3_4k(rank, world, dev, frames=300, steps=3, warmup=2):
    """BASELINE configs[3]: 3840x2160 input, 300-frame KBE, frames sharded over the GPUs after ONE broadcast of the cloud
    (10.2 M points, 286 MB).  Render loop only, like the headline metric; device-resident and end to end."""
    import torch
    import torch.distributed as dist
    from ken_burns_effect_b200.utils import common as kb
    from ken_burns_effect_b200.utils import shard, synthetic
    w4, h4, f4 = 3840, 2160, 1920.0
    zoom = synthetic.default_zoom(w4, h4)
    cloud0, packed_host, packed = None, None, None
    if rank == 0:
        pts, rgb, dep, common = synthetic.scene_cloud(w4, h4, seed=1234, focal=f4, baseline=BASELINE, inpaint_standin=True)
        n = pts.shape[1]
        packed_host = torch.from_numpy(np.concatenate([pts, rgb, dep], 0)).pin_memory()
        packed = packed_host.to(dev)
        cloud0 = dict(common)
        cloud0.update(intWidth=w4, intHeight=h4, tensorInpaPoints=packed[0:3].view(1, 3, n), tensorInpaImage=packed[3:6].view(1, 3, n),
                      tensorInpaDepth=packed[6:7].view(1, 1, n), tensorPacked=packed)
    xchg = shard.CloudExchange(dev, 3 * w4 * h4 // 2, src=0)
    cloud = xchg.broadcast(cloud0)
    st = {'dblSteps': np.linspace(0.0, 1.0, frames).tolist(), 'objectFrom': zoom['objectFrom'], 'objectTo': zoom['objectTo'], 'dolly': False}
    poses = kb.kenburns_poses(st, cloud)[rank::world]
    cw, ch = kb.crop_size(st)
    r = kb.FrameRenderer(cloud['tensorInpaPoints'], cloud['tensorInpaImage'], cloud['tensorInpaDepth'], w4, h4, BASELINE, cw, ch, batch=16)
    fdev = torch.empty(len(poses), h4, w4, 3, dtype=torch.uint8, device=dev)
    fhost = torch.empty(len(poses), h4, w4, 3, dtype=torch.uint8).pin_memory()

    def run(dst, h2d):
        for i in range(warmup + steps):
            if i == warmup:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if h2d and rank == 0:
                packed.copy_(packed_host, non_blocking=True)
            c = xchg.broadcast(cloud0)
            r.set_cloud(c['tensorInpaPoints'], c['tensorInpaImage'], c['tensorInpaDepth'])
            r.render_into(poses, dst)
            xchg.consumed(c)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return _max_over_ranks(e0.elapsed_time(e1), world, dev)

    ms_d = run(fdev, False)
    ms_e = run(fhost, True)
    return {"what": f"configs[3]: 3840x2160 -> {frames}-frame KBE, poses rank::{world} per GPU, one NCCL broadcast of the cloud per effect",
            "points": int(cloud['tensorInpaPoints'].shape[-1]), "value": frames * steps / (ms_d / 1e3), "unit": "frames/s",
            "ms_per_effect": ms_d / steps,
            "e2e": {"value": frames * steps / (ms_e / 1e3), "unit": "frames/s", "ms_per_effect": ms_e / steps,
                    "d2h_bytes_per_effect": int(frames * h4 * w4 * 3)}}


def cpu_cnn_stage():
    """The CNN forwards of one KBE on the host cores: the nn.Module mirrors run plain torch fp32 on CPU tensors -- the same
    layers, shapes and arithmetic as the reference's modules (tests/test_models_cpu.py checks them against fixtures made by the
    reference itself).  One timed call each after a small-size warm-up; -> seconds per KBE (Semantics + Disparity + Refine +
    2 x (context + Inpaint.forward)), thread count."""
    import torch
    from ken_burns_effect_b200.models.disparity_estimation import Disparity, Semantics
    from ken_burns_effect_b200.models.disparity_refinement import Refine
    from ken_burns_effect_b200.models.pointcloud_inpainting import Inpaint
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    with torch.no_grad():
        sem, dis, ref, inp = Semantics().eval(), Disparity().eval(), Refine().eval(), Inpaint().eval()
        small = torch.rand(1, 3, 64, 64)
        dis(small, sem(small))                                   # warm the thread pool / oneDNN primitives
        img, half = torch.rand(1, 3, H, W), torch.rand(1, 3, H // 2, W // 2)
        t0 = time.perf_counter()
        d = dis(half, sem(half))
        t1 = time.perf_counter()
        ref(img, d)
        t2 = time.perf_counter()
        mask = (torch.rand(1, 1, H, W) > 0.1).float()
        inp(mask, tensorImage=img * mask, tensorDisparity=torch.rand(1, 1, H, W) * mask)
        t3 = time.perf_counter()
    return (t1 - t0) + (t2 - t1) + 2.0 * (t3 - t2), torch.get_num_threads()


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report it instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # Same process environment as the cpu_baseline leg of the GPU arm: the reference is a PyTorch program, and the OpenMP runtime
    # PyTorch brings along is the one its CPU kernels (and this restatement, a plain `#pragma omp parallel for` per kernel) run on
    # -- measured on the B200 box: 56 frames/s with it, 37 frames/s with the system libgomp loaded first (profiles/).
    import torch  # noqa: F401
    # keep the whole run within a few minutes: ~60 s of CPU work spread over the timed steps
    budget = float(os.environ.get("KB_BENCH_CPU_BUDGET_S", "60"))       # seconds of CPU work over the timed steps
    per_step = max(1.0, budget / max(1, args.steps))
    vals, nframes = [], 0
    for i in range(args.warmup + args.steps):
        fps, cores, dt, n = cpu_baseline(budget_s=per_step if i >= args.warmup else 1.0, max_frames=150)
        if i >= args.warmup:
            vals.append(fps)
            nframes = n
    v = float(np.mean(vals))
    sample = nframes
    desc = (f"~{sample} of the 150 poses per step (spread over the path, {per_step:.0f} s of CPU work per step), "
            f"full 1024x768 cloud incl. stand-in inpainted points")
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sample / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.frames),
                   "note": "reference has no CPU render path (cupy-only kernels); this is the C/OpenMP restatement "
                           "of its kernels + its numpy/OpenCV tail (oracle/kb_oracle.c)"},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "per_step_frames_per_s": [round(x, 2) for x in vals],
    })


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner on communicator creation, for
    one) are sent to stderr for the duration of the run; emit() writes the result to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, line)
    else:
        os.write(_REAL_STDOUT, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=150)
    ap.add_argument("--batch", type=int, default=32, help="poses per launch with frames left on the device (host-bound frames: <= 16)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-pipeline", action="store_true")
    ap.add_argument("--dolly", action="store_true", help="BASELINE configs[2]: the dolly-zoom path (per-pose focal length, no inpainted points)")
    ap.add_argument("--images-per-gpu", type=int, default=8, help="throughput mode (configs[4]): images per GPU, one full KBE each")
    ap.add_argument("--config3", action="store_true", help="also run configs[3] (3840x2160, 300 frames, sharded) at this N; default only at N=8")
    ap.add_argument("--size", default="1024x768", help="WxH of the synthetic input (BASELINE configs[3]: 3840x2160 --frames 300)")
    args = ap.parse_args()
    global W, H, FOCAL, METRIC, DOLLY
    DOLLY = args.dolly
    W, H = (int(v) for v in args.size.lower().split("x"))
    FOCAL = max(W, H) / 2.0          # dblFocal = max side / 2 like pipeline.py:26 for 1024x768
    if (W, H) != (1024, 768) or args.frames != 150 or DOLLY:
        METRIC = f"novel-view frames/sec at {W}x{H}, {args.frames}-frame KBE" + (" (--dolly)" if DOLLY else "")
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from ken_burns_effect_b200 import _native
    from ken_burns_effect_b200.utils import common as kb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the render path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        from ken_burns_effect_b200.utils import shard as _shard
        numa = _shard.bind_to_gpu_numa_node(local)       # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    pts, rgb, dep, common, poses, (cw, ch) = build_workload(args.frames, world, rank)
    N = pts.shape[1]
    F = len(poses)
    from ken_burns_effect_b200.utils import shard
    packed_host = torch.from_numpy(np.concatenate([pts, rgb, dep], 0)).pin_memory()   # [7,N]: xyz | rgb | depth
    packed = packed_host.to(dev)              # only rank 0's copy is ever read when world > 1
    common["intWidth"], common["intHeight"] = W, H

    def cloud_on_rank0():
        c = dict(common)
        c.update(tensorInpaPoints=packed[0:3].view(1, 3, N), tensorInpaImage=packed[3:6].view(1, 3, N),
                 tensorInpaDepth=packed[6:7].view(1, 1, N), tensorPacked=packed)
        return c

    renderer = kb.FrameRenderer(packed[0:3], packed[3:6], packed[6:7], W, H, BASELINE, cw, ch, batch=args.batch)
    frames_dev = torch.empty(F, H, W, 3, dtype=torch.uint8, device=dev)
    frames_host = torch.empty(F, H, W, 3, dtype=torch.uint8).pin_memory()

    xchg = shard.CloudExchange(dev, N, src=0) if world > 1 else None

    pending = [None]

    def exchange():
        # the path's one exchange step (ken_burns_effect_b200/utils/shard.py): the cloud travels over NVLink once per effect;
        # preallocated receive buffers, header on a host side channel -> no stream drain, no allocation, no re-pack.  Effects are
        # software-pipelined: the broadcast of the NEXT effect's cloud is posted (on the exchange stream, into the other buffer)
        # before this effect renders, so one broadcast and one render run per step and the wire time hides behind the kernels.
        if pending[0] is None:
            pending[0] = xchg.broadcast_async(cloud_on_rank0() if rank == 0 else None)
        c = pending[0]
        pending[0] = xchg.broadcast_async(cloud_on_rank0() if rank == 0 else None)
        xchg.wait(c)
        renderer.set_cloud(c['tensorInpaPoints'], c['tensorInpaImage'], c['tensorInpaDepth'])
        return c

    def step_device():
        c = exchange() if world > 1 else None
        renderer.render_into(poses, frames_dev[:len(poses)])
        if c is not None:
            xchg.consumed(c)

    # single GPU, end to end: the cloud of the NEXT effect is uploaded on its own stream into a second device buffer while the frames
    # of the current effect still copy out (PCIe is full duplex); every step still moves its own 27 MB in and its 354 MB out
    up = {"stream": torch.cuda.Stream(dev), "bufs": [packed, torch.empty_like(packed)], "ev_up": [torch.cuda.Event(), torch.cuda.Event()],
          "ev_done": [None, None], "i": 0} if world == 1 else None

    def step_e2e():
        if world == 1:
            j = up["i"] & 1
            up["i"] += 1
            buf = up["bufs"][j]
            main = torch.cuda.current_stream(dev)
            with torch.cuda.stream(up["stream"]):
                if up["ev_done"][j] is not None:
                    up["stream"].wait_event(up["ev_done"][j])          # the effect before last rendered from this buffer
                buf.copy_(packed_host, non_blocking=True)
                up["ev_up"][j].record(up["stream"])
            main.wait_event(up["ev_up"][j])
            renderer.set_cloud(buf[0:3], buf[3:6], buf[6:7])
            renderer.render_into(poses, frames_host[:len(poses)])
            up["ev_done"][j] = torch.cuda.Event()
            up["ev_done"][j].record(main)
            return
        if rank == 0:
            packed.copy_(packed_host, non_blocking=True)
        c = exchange() if world > 1 else None
        renderer.render_into(poses, frames_host[:len(poses)])
        if c is not None:
            xchg.consumed(c)

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile:
            _native.profile_enable(True)
            _native.profile_read()
        n0 = _native.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = _native.launch_count() - n0
        stages = None
        if profile:
            stages, calls = _native.profile_read()
            _native.profile_enable(False)
            stages = {k: v / max(calls, 1) for k, v in stages.items()}
            stages["calls"] = calls
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, stages

    # 1) kernel-side number: un-instrumented timed region
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches, _ = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop()
    # 2) per-stage durations (CUDA events around every kernel, on the launching stream), separate pass
    _, _, stages = timed(step_device, max(2, args.steps // 2), 1, profile=True)
    # 3) end to end with host buffers
    ms_e2e, _, _ = timed(step_e2e, args.steps, args.warmup)
    if world == 1:
        renderer.set_cloud(packed[0:3], packed[3:6], packed[6:7])

    total_frames = F * world * args.steps
    value = total_frames / (ms / 1000.0)
    e2e_value = total_frames / (ms_e2e / 1000.0)

    # 4) the metric's own configuration at N GPUs: ONE args.frames-pose effect split N ways (strong scaling) -- the same
    #    exchange, every rank renders poses rank::world of the single path
    strong = None
    if world > 1:
        all_poses = build_workload(args.frames, 1, 0)[4]
        weak_poses = poses
        poses = all_poses[rank::world]
        ms_s, _, _ = timed(step_device, args.steps, args.warmup)
        ms_se, _, _ = timed(step_e2e, args.steps, args.warmup)
        poses = weak_poses
        strong = {"what": f"one {args.frames}-pose effect split over {world} GPUs (poses rank::{world}), cloud broadcast per effect",
                  "value": args.frames * args.steps / (ms_s / 1000.0), "unit": "frames/s", "ms_per_effect": ms_s / args.steps,
                  "e2e": {"value": args.frames * args.steps / (ms_se / 1000.0), "unit": "frames/s", "ms_per_effect": ms_se / args.steps}}

    P = W * H
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    # algorithmic bytes per pose: splat_min 12N+4P, degrid 8P, accumulate 28N+24P, post 36P
    bytes_stage = {"splat_min": 12 * N + 4 * P, "degrid": 8 * P, "splat_accum": 28 * N + 24 * P, "resolve": 36 * P}
    calls_per_step = -(-F // renderer.batch)
    avg_k = F / calls_per_step
    dom = max(bytes_stage, key=lambda k: stages[k])
    achieved = bytes_stage[dom] * avg_k / (stages[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        # per-launch dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the last `ncu --set full` capture of this
        # very command (tools/ncu_summary.py --traffic; K = 16 poses per launch like here)
        kernel_of = {"splat_min": "kf_splat_min", "degrid": "kf_degrid4", "splat_accum": "kf_accum", "resolve": "kf_resolve"}
        tj = json.load(open(tpath))
        traffic = tj.get(kernel_of[dom])
        if traffic is not None:      # captured at tj["_poses_per_launch"] poses per launch; this run averages avg_k
            traffic = int(traffic * avg_k / tj.get("_poses_per_launch", 16))
    render_ms_per_call = sum(stages[k] for k in bytes_stage)
    whole = (40 * N + 72 * P) * avg_k / (render_ms_per_call * 1e-3) / 1e9

    out = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.frames),
                   "frames_per_step_per_gpu": F, "points": N, "pixels": P, "poses_per_launch": renderer.batch, "poses_per_launch_e2e": min(renderer.batch, kb.FRAME_BATCH_TO_HOST),
                   "parallelism": f"frame-shard x{world}" + (" + one NCCL broadcast of the cloud per step, posted one effect ahead on its own stream (double-buffered)" if world > 1 else ""),
                   "cpu_affinity": (f"rank 0 bound to {len(numa)} cores near its GPU (NVML)" if numa else "unbound"),
                   "l2": "no explicit flush: each step streams ~0.8 GB of z-buffers/accumulators/frames (> 126 MB L2)"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(packed_host.numel() * 4) if rank == 0 else 0,
                "d2h_bytes_per_step": int(frames_host.numel()),
                "overlap": ("the cloud of step i+1 is uploaded (own stream, second device buffer) while the frames of step i copy out; "
                            "every step moves its own cloud in and its own frames out" if world == 1 else
                            "rank 0 uploads the cloud, then the NCCL broadcast, then every rank's frames copy out")},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": int(bytes_stage[dom] * avg_k),
                     "avg_launch_ms": stages[dom],
                     "render_path_all_kernels": {"achieved": whole, "frac": whole / peak,
                                                 "bytes_per_frame": 40 * N + 72 * P}},
        "stage_ms_per_launch": stages,
    }
    if strong is not None:
        out["strong_scaling"] = strong
    del renderer, frames_dev, frames_host
    torch.cuda.empty_cache()
    if not args.no_full_pipeline:
        try:
            tm = throughput_mode(rank, world, dev, args.images_per_gpu, args.frames)
        except Exception as e:
            tm = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0:
            out["throughput_mode"] = tm
    if (world == 8 or args.config3) and (W, H) == (1024, 768) and not DOLLY:
        try:
            c3 = config3_4k(rank, world, dev)
        except Exception as e:
            c3 = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0:
            out["config3_4k"] = c3
    if rank == 0 and world == 1 and not args.no_full_pipeline:
        try:
            out["kbe_full_pipeline"] = full_pipeline()
        except Exception as e:   # the headline numbers above must survive a failure of this extra leg
            out["kbe_full_pipeline"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            fps, cores, dt, n = cpu_baseline()
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                   "sample": f"{n} frames of the same 150-pose path (spread over it), same cloud, "
                                             f"{dt:.1f} s of CPU work, all host threads"}
            full = out.get("kbe_full_pipeline")
            if isinstance(full, dict) and "error" not in full:
                try:
                    cnn_s, threads = cpu_cnn_stage()
                    total_s = cnn_s + args.frames / fps
                    full["cpu_baseline"] = {"value": args.frames / total_s, "unit": "frames/s", "cores": threads, "kind": "port",
                                            "s_per_kbe": total_s, "s_cnn_stage": cnn_s,
                                            "sample": "one timed CPU forward of each network at the KBE's shapes (torch fp32, all "
                                                      "threads; Inpaint counted twice) + the render loop at the CPU frames/s above"}
                except Exception as e:
                    full["cpu_baseline"] = {"error": f"{type(e).__name__}: {e}"}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
