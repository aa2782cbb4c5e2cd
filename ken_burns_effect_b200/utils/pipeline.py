"""Pipeline -- mirror of utils/pipeline.py:23-134: one image in, the frames of a 3D Ken Burns effect out.

Same constructor and __call__ signature.  Differences, all additive or dead-code removal:
  * the Mask R-CNN the reference builds and never uses (pipeline.py:36) is not built;
  * `model_paths=None` keeps the freshly initialised weights (tests, benchmarks -- no checkpoints offline);
  * `frames` (default 75 like pipeline.py:104,113) selects the number of rendered poses;
  * the mp4 is written with cv2.VideoWriter (moviepy is not installed here), same 25 fps ping-pong sequence, and PNGs / video
    are encoded by an asynchronous sink (utils/sink.py) while the frames still render;
  * the per-frame loop runs fused on the device (utils.common.process_kenburns).
"""
import os

import cv2
import numpy as np
import torch

from ..models.disparity_estimation import Disparity, Semantics
from ..models.disparity_refinement import Refine
from ..models.disparity_refinement_pretrained import Refine as RefineP
from ..models.partial_inpainting import Inpaint as PartialInpaint
from ..models.pointcloud_inpainting import Inpaint
from . import shard
from .common import depth_to_points, kenburns_poses, prepare_cloud, render_poses
from .utils import device, load_models, resize_image


class Pipeline():
    def __init__(self, model_paths=None, partial_inpainting=False, dolly=False, output_frames=False, pretrain=False,
                 d2=False, frames=75):
        self.objectCommon = {'dblFocal': 1024.0 / 2, 'dblBaseline': 120}
        self.partial_inpainting = partial_inpainting
        self.dolly = dolly
        self.output_frames = output_frames
        self.d2 = d2
        self.frames = frames

        if model_paths is None:                  # freshly initialised everything: the VGG trunk too (explicit opt-in)
            from ..models import disparity_estimation
            disparity_estimation.ALLOW_RANDOM_VGG = True
        self.moduleSemantics = Semantics().to(device).eval()
        self.moduleDisparity = Disparity().to(device).eval()
        self.moduleRefine = (RefineP() if pretrain else Refine()).to(device).eval()
        self.moduleInpaint = (PartialInpaint() if partial_inpainting else Inpaint()).to(device).eval()

        models_list = [{'model': self.moduleDisparity, 'type': 'disparity'},
                       {'model': self.moduleRefine, 'type': 'refine'},
                       {'model': self.moduleInpaint, 'type': 'inpaint'}]
        if model_paths is not None and len(model_paths) == 4:
            self.moduleInpaintDepth = Inpaint().to(device).eval()
            models_list.append({'model': self.moduleInpaintDepth, 'type': 'inpaint'})
        if model_paths is not None:
            load_models(models_list, model_paths)

    @torch.no_grad()
    def estimate_depth(self, tensorImage):
        """The one-time depth stage, pipeline.py:61-100: fills objectCommon with the raw point cloud."""
        oc = self.objectCommon
        tensorImage = tensorImage.to(device).contiguous()
        oc['intWidth'], oc['intHeight'] = tensorImage.size(3), tensorImage.size(2)
        tensorResized = resize_image(tensorImage, max_size=int(max(oc['intWidth'], oc['intHeight']) / 2))
        tensorDisparity = self.moduleDisparity(tensorResized, self.moduleSemantics(tensorResized))
        if self.d2:
            tensorDisparity = torch.ones_like(tensorDisparity)
        tensorDisparity = self.moduleRefine(tensorImage, tensorDisparity)
        if tensorDisparity.min() < 0.0:
            tensorDisparity = tensorDisparity - tensorDisparity.min()
        tensorDisparity = tensorDisparity / tensorDisparity.max() * oc['dblBaseline']
        tensorDepth = (oc['dblFocal'] * oc['dblBaseline']) / (tensorDisparity + 1e-7)
        tensorPoints = depth_to_points(tensorDepth, oc['dblFocal'])
        oc['dblDispmin'] = tensorDisparity.min().item()
        oc['dblDispmax'] = tensorDisparity.max().item()
        oc['objectDepthrange'] = cv2.minMaxLoc(src=tensorDepth[0, 0, 128:-128, 128:-128].detach().cpu().numpy(), mask=None)
        oc['tensorRawPoints'] = tensorPoints.view(1, 3, -1)
        oc['tensorRawImage'] = tensorImage
        oc['tensorRawDisparity'] = tensorDisparity
        oc['tensorRawDepth'] = tensorDepth
        return oc

    @torch.no_grad()
    def estimate_depth_batch(self, tensorImages):
        """The one-time depth stage (pipeline.py:61-100) for B images of one size at once: Semantics, Disparity and Refine run ONE
        forward with batch B (the reference asserts B == 1, pipeline.py:63; its modules are batch-generic and so are the tcgen05
        executors) -- the 24x32 ... 6x8 maps of the Disparity GridNet then fill the GPU -- followed by the per-image normalisation.
        -> list of B objectCommon dicts, each what estimate_depth() leaves for that image alone (the convolutions work image by
        image; torch's per-sample mean / std reductions sum a [B, n] tensor in another order than a [1, n] one, so the result
        agrees to fp32 reassociation noise, ~1e-6, not bit for bit)."""
        base = self.objectCommon
        imgs = tensorImages.to(device).contiguous()
        B, _, H, W = imgs.shape
        resized = resize_image(imgs, max_size=int(max(W, H) / 2))
        disparity = self.moduleDisparity(resized, self.moduleSemantics(resized))
        if self.d2:
            disparity = torch.ones_like(disparity)
        disparity = self.moduleRefine(imgs, disparity)
        lo = disparity.reshape(B, -1).min(1)[0].view(B, 1, 1, 1)
        disparity = torch.where(lo < 0.0, disparity - lo, disparity)                      # pipeline.py:79-80, per image
        disparity = disparity / disparity.reshape(B, -1).max(1)[0].view(B, 1, 1, 1) * base['dblBaseline']
        depth = (base['dblFocal'] * base['dblBaseline']) / (disparity + 1e-7)
        points = depth_to_points(depth, base['dblFocal'])
        crop = depth[:, 0, 128:-128, 128:-128].detach().cpu().numpy()                    # one D2H for the batch
        dmin = disparity.reshape(B, -1).min(1)[0].tolist()
        dmax = disparity.reshape(B, -1).max(1)[0].tolist()
        out = []
        for b in range(B):
            out.append({'dblFocal': base['dblFocal'], 'dblBaseline': base['dblBaseline'], 'intWidth': W, 'intHeight': H,
                        'dblDispmin': dmin[b], 'dblDispmax': dmax[b], 'objectDepthrange': cv2.minMaxLoc(src=crop[b], mask=None),
                        'tensorRawPoints': points[b:b + 1].view(1, 3, -1), 'tensorRawImage': imgs[b:b + 1],
                        'tensorRawDisparity': disparity[b:b + 1], 'tensorRawDepth': depth[b:b + 1]})
        return out

    @torch.no_grad()
    def __call__(self, tensorImage, zoom_settings, output_path=None, inpaint_depth=False, pretrained_estim=False):
        """-> list of uint8 [H,W,3] frames (every rank gets the whole list under torchrun).  With output_path the frames go to an
        asynchronous sink (utils/sink.py) while they render: <out>/frames/<i>.png when output_frames, <out>/3d_kbe.mp4 always --
        the files of pipeline.py:120-134.  self.last_timing holds the wall-clock breakdown of the call."""
        import time
        from .sink import FrameSink
        t_start = time.perf_counter()
        settings = {
            'dblSteps': np.linspace(0.0, 1.0, self.frames).tolist(),
            'objectFrom': zoom_settings['objectFrom'],
            'objectTo': zoom_settings['objectTo'],
            'boolInpaint': True,
            'dolly': self.dolly,
        }
        moduleInpaint = self.moduleInpaint
        if inpaint_depth:                        # colour from the first network, disparity from the second (common.py:50-62)
            if not hasattr(self, 'moduleInpaintDepth'):
                raise ValueError("inpaint_depth=True needs a fourth checkpoint (model_paths[3]: the disparity inpainting network)")
            moduleInpaint = [self.moduleInpaint, self.moduleInpaintDepth]
        rank, world = shard.world()
        timing = {}
        dev = torch.device(device)

        def mark(name):
            if dev.type == 'cuda':
                torch.cuda.synchronize(dev)
            timing[name] = time.perf_counter() - t_start

        if world == 1:
            self.estimate_depth(tensorImage)
            mark('t_depth_stage_s')
            prepare_cloud(settings, self.objectCommon, moduleInpaint)
            mark('t_inpaint_stage_s')
            sink = None
            if output_path is not None:
                sink = FrameSink(output_path, self.frames, write_frames=self.output_frames, write_video=True,
                                 rgb_to_bgr=pretrained_estim, t0=t_start)
            poses = kenburns_poses(settings, self.objectCommon)
            frames = render_poses(settings, self.objectCommon, poses, sink=sink).numpy()
            mark('t_frames_in_host_memory_s')
            numpyResult = [frames[i] for i in range(len(poses))]
            if sink is not None:
                timing.update(sink.close())
            timing['t_total_s'] = time.perf_counter() - t_start
            self.last_timing = timing
            return numpyResult

        # One process per GPU (torchrun): rank 0 runs the CNN stage, the cloud is broadcast once, every rank renders its
        # interleaved share of the poses and copies it over ITS OWN PCIe link into a shared host segment (shard.SharedFrames);
        # PNGs are written by the rank that rendered them, the video by rank 0 from the shared segment.
        if rank == 0:
            self.estimate_depth(tensorImage)
            prepare_cloud(settings, self.objectCommon, moduleInpaint)
            mark('t_cnn_stages_s')
        H, W = int(tensorImage.size(2)), int(tensorImage.size(3))
        ex = getattr(self, '_exchange', None)
        if ex is None or ex.buf.numel() < 7 * 3 * H * W:
            ex = self._exchange = shard.CloudExchange(dev, 3 * H * W, src=0)       # each inpainting pass appends at most H*W points
        cloud = ex.broadcast(self.objectCommon if rank == 0 else None)
        poses = kenburns_poses(settings, cloud)
        key = (len(poses), H, W)
        if getattr(self, '_shared_key', None) != key:
            if getattr(self, '_shared', None) is not None:
                self._shared.close()
            self._shared, self._shared_key = shard.SharedFrames(len(poses), H, W), key
        shared = self._shared
        mine = shard.shard_indices(len(poses), rank, world)
        sink = None
        if output_path is not None and self.output_frames:
            sink = FrameSink(output_path, len(mine), write_frames=True, write_video=False, rgb_to_bgr=pretrained_estim,
                             frame_indices=mine, t0=t_start)
        render_poses(settings, cloud, [poses[i] for i in mine], sink=sink, out=shared.block())
        ex.consumed(cloud)
        mark('t_own_frames_in_host_memory_s')
        torch.distributed.barrier()                      # every block of the shared segment is complete
        numpyResult = [shared.frame(i) for i in range(len(poses))]
        if sink is not None:
            timing.update(sink.close())
        if output_path is not None and rank == 0:
            video = FrameSink(output_path, len(poses), write_frames=False, write_video=True, rgb_to_bgr=pretrained_estim, t0=t_start)
            video.submit(0, _as_batch(numpyResult))
            timing.update({'video_' + k: v for k, v in video.close().items()})
        timing['t_total_s'] = time.perf_counter() - t_start
        self.last_timing = timing
        return numpyResult


    @torch.no_grad()
    def run_many(self, images, zoom_settings, consume=None, keep=True, trace=None, depth_batch=1):
        """Throughput mode (BASELINE configs[4]: many images, one effect each, per GPU): the same stages as __call__, software-
        pipelined over the images on two CUDA streams -- the CNN stage of image i+1 (tensor-core bound: depth networks + two
        inpainting passes) runs on one stream while a helper thread renders the frames of image i on another (HBM / PCIe bound:
        the fused frame loop and the device-to-host copies into pinned memory).  The reference processes one image at a time
        (utils/pipeline.py:63 asserts B == 1); each image here is still its own B == 1 pass, bit-for-bit the single-image result.

        images: iterable of [1,3,H,W] tensors in [0,1] (host, ideally pinned); zoom_settings: one dict or one per image;
        consume(i, frames): called from the render thread with the uint8 [n,H,W,3] pinned tensor of image i;
        depth_batch: images whose depth stage runs as one batched forward (estimate_depth_batch; same size required);
        trace: optional list that receives (image, stage, t_begin, t_end) host timestamps (perf_counter);
        -> list of those tensors (None entries when keep is False)."""
        import queue
        import threading
        import time
        dev = torch.device(device)
        # the two streams live with the Pipeline: the caching allocator keeps one pool per stream, and fresh streams would send
        # the first image of every call through cudaMalloc again (measured: 90 ms instead of 21)
        if getattr(self, '_streams', None) is None:
            self._streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        s_cnn, s_render = self._streams
        images = list(images)
        zooms = zoom_settings if isinstance(zoom_settings, (list, tuple)) else [zoom_settings] * len(images)
        results = [None] * len(images)
        q = queue.Queue(maxsize=2)                      # at most two clouds wait for the renderer
        err = []

        def render_loop():
            try:
                torch.cuda.set_device(dev)
                with torch.cuda.stream(s_render):
                    while True:
                        item = q.get()
                        if item is None:
                            return
                        i, oc, settings, ready = item
                        t0 = time.perf_counter()
                        torch.cuda.current_stream().wait_event(ready)
                        out = render_poses(settings, oc, kenburns_poses(settings, oc))      # synchronises its own stream only
                        if trace is not None:
                            trace.append((i, 'render', t0, time.perf_counter()))
                        if consume is not None:
                            consume(i, out)
                        if keep:
                            results[i] = out
            except Exception as e:                       # surfaced by the caller
                err.append(e)
                while q.get() is not None:
                    pass

        th = threading.Thread(target=render_loop, daemon=True)
        th.start()
        s_cnn.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s_cnn):
            depth_batch = max(1, int(depth_batch))
            ahead = []                                   # objectCommon dicts of the current depth batch, in order
            for i, img in enumerate(images):
                settings = {'dblSteps': np.linspace(0.0, 1.0, self.frames).tolist(), 'objectFrom': zooms[i]['objectFrom'],
                            'objectTo': zooms[i]['objectTo'], 'boolInpaint': True, 'dolly': self.dolly}
                t0 = time.perf_counter()
                if depth_batch == 1:
                    self.estimate_depth(img)
                    oc = self.objectCommon
                else:
                    if not ahead:
                        chunk = images[i:i + depth_batch]
                        ahead = self.estimate_depth_batch(torch.cat([c.to(device, non_blocking=True) for c in chunk], 0))
                    oc = ahead.pop(0)
                prepare_cloud(settings, oc, self.moduleInpaint)
                ready = torch.cuda.Event()
                ready.record()
                if trace is not None:
                    trace.append((i, 'cnn', t0, time.perf_counter()))
                q.put((i, dict(oc), settings, ready))
                if err:
                    break
        q.put(None)
        th.join()
        if err:
            raise err[0]
        return results


class _as_batch:
    """A list of equally shaped frames behind the [k,H,W,3] indexing FrameSink.submit uses (no copy)."""

    def __init__(self, frames):
        self.frames = frames
        self.shape = (len(frames),) + tuple(frames[0].shape)

    def __getitem__(self, i):
        return self.frames[i]
