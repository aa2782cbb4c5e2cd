"""Pipeline -- mirror of utils/pipeline.py:23-134: one image in, the frames of a 3D Ken Burns effect out.

Same constructor and __call__ signature.  Differences, all additive or dead-code removal:
  * the Mask R-CNN the reference builds and never uses (pipeline.py:36) is not built;
  * `model_paths=None` keeps the freshly initialised weights (tests, benchmarks -- no checkpoints offline);
  * `frames` (default 75 like pipeline.py:104,113) selects the number of rendered poses;
  * the mp4 is written with cv2.VideoWriter (moviepy is not installed here), same 25 fps ping-pong sequence;
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
from .common import depth_to_points, kenburns_poses, prepare_cloud, process_kenburns, render_poses
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
    def __call__(self, tensorImage, zoom_settings, output_path=None, inpaint_depth=False, pretrained_estim=False):
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
        if world == 1:
            self.estimate_depth(tensorImage)
            numpyResult = process_kenburns(settings, self.objectCommon, moduleInpaint)
        else:
            # one process per GPU (torchrun): rank 0 runs the CNN stage, the cloud is broadcast once, every rank
            # renders its interleaved share of the poses, rank 0 gathers and writes
            dev = torch.device('cuda', torch.cuda.current_device())
            if rank == 0:
                self.estimate_depth(tensorImage)
                prepare_cloud(settings, self.objectCommon, moduleInpaint)
            cloud = shard.broadcast_cloud(self.objectCommon if rank == 0 else None, dev, src=0)
            poses = kenburns_poses(settings, cloud)
            frames = shard.render_sharded(poses, lambda mine: render_poses(settings, cloud, mine, to_host=False))
            if rank != 0:
                return None
            frames = frames.cpu().numpy()
            numpyResult = [frames[i] for i in range(len(poses))]

        if self.output_frames and output_path is not None:
            frames_dir = os.path.join(output_path, 'frames')
            os.makedirs(frames_dir, exist_ok=True)
            for idx, frame in enumerate(numpyResult):
                if pretrained_estim:
                    frame = cv2.cvtColor(frame, cv2.COLOR_RGB2BGR)
                cv2.imwrite(os.path.join(frames_dir, str(idx) + '.png'), frame)

        if output_path is not None:
            os.makedirs(output_path, exist_ok=True)
            seq = numpyResult + list(reversed(numpyResult))[1:]          # ping-pong, pipeline.py:132-134
            h, w = seq[0].shape[:2]
            vw = cv2.VideoWriter(os.path.join(output_path, '3d_kbe.mp4'), cv2.VideoWriter_fourcc(*'mp4v'), 25, (w, h))
            for f in seq:
                # moviepy expects RGB; the reference flips BGR tensors with [:, :, ::-1] unless pretrained_estim.
                # cv2.VideoWriter expects BGR, i.e. the tensor's own channel order in the default case.
                vw.write(np.ascontiguousarray(f if not pretrained_estim else f[:, :, ::-1]))
            vw.release()
        return numpyResult
