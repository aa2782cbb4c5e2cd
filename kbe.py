#!/usr/bin/env python3
"""kbe.py -- the reference's inference CLI (kbe.py:22-181) on top of ken_burns_effect_b200.

Same long options, defaults, crop-window rules and asserts; additive options: --partial-conv (build the
PartialInpaint network, as train.py of the reference can), --frames N (rendered poses, default 75 like the
reference).  Multi-GPU frame sharding is launched with torchrun (see README / bench.py)."""
import getopt
import math
import os
import sys

import cv2
import torch
import torch.distributed

from ken_burns_effect_b200.utils.pipeline import Pipeline

torch.set_grad_enabled(False)
torch.backends.cudnn.enabled = True

OPTIONS = ['in=', 'out=', 'dolly', 'write-frames', 'inpaint-path=', 'refine-path=', 'estim-path=', 'startU=', 'startV=',
           'endU=', 'endV=', 'startW=', 'startH=', 'endW=', 'endH=', 'pretrained-refine', 'pretrained-estim',
           'inpaint-depth=', '2d', 'partial-conv', 'frames=', 'random-weights']


def parse(argv):
    cfg = dict(input_path='images/doublestrike.jpg', output_path='images/kbe', dolly=False, output_frames=False,
               pretrained_estim=False, pretrained_refine=False, inpaint_depth=False, d2=False, partial=False,
               frames=75, random_weights=False,
               inpaint_path='./models/trained/inpainting-color.tar', refine_path='./models/trained/disparity-refinement.tar',
               estim_path='./models/trained/disparity-estimation-no-mask.tar',
               inpaint_depth_path='./models/trained/inpainting-depth.tar',
               startU=None, startV=None, startW=None, startH=None, endU=None, endV=None, endW=None, endH=None)
    flags = {'--dolly': 'dolly', '--write-frames': 'output_frames', '--pretrained-refine': 'pretrained_refine',
             '--pretrained-estim': 'pretrained_estim', '--2d': 'd2', '--partial-conv': 'partial',
             '--random-weights': 'random_weights'}
    paths = {'--in': 'input_path', '--out': 'output_path', '--inpaint-path': 'inpaint_path', '--refine-path': 'refine_path',
             '--estim-path': 'estim_path'}
    for opt, arg in getopt.getopt(argv, '', OPTIONS)[0]:
        if opt in flags:
            cfg[flags[opt]] = True
        elif opt in paths and arg != '':
            cfg[paths[opt]] = arg
        elif opt == '--inpaint-depth' and arg != '':
            cfg['inpaint_depth'], cfg['inpaint_depth_path'] = True, arg
        elif opt == '--frames' and arg != '':
            cfg['frames'] = int(arg)
        elif opt[2:] in ('startU', 'startV', 'startW', 'startH', 'endU', 'endV', 'endW', 'endH') and arg != '':
            cfg[opt[2:]] = int(arg)
    return cfg


def load_image(path, pretrained_estim):
    """kbe.py:96-114: BGR uint8 -> ToTensor/Normalize(.5,.5) -> crop H, W to multiples of 4 -> (x+1)/2."""
    img = cv2.imread(filename=path, flags=cv2.IMREAD_COLOR)
    if img is None:
        raise FileNotFoundError(path)
    if pretrained_estim:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    t = torch.from_numpy(img).permute(2, 0, 1).float().div(255)          # transforms.ToTensor()
    t = (t - 0.5) / 0.5                                                   # transforms.Normalize
    h, w = t.size(1), t.size(2)
    if w % 4 != 0:
        t = t[:, :, :-(w % 4)]
    if h % 4 != 0:
        t = t[:, :-(h % 4), :]
    return t


def crop_windows(cfg, imgWidth, imgHeight):
    """kbe.py:117-146."""
    sU, sV, sW, sH = cfg['startU'], cfg['startV'], cfg['startW'], cfg['startH']
    eU, eV, eW, eH = cfg['endU'], cfg['endV'], cfg['endW'], cfg['endH']
    if eH is not None and eW is None:
        eW = int(imgWidth * eH / imgHeight)
    if eW is not None and eH is None:
        eH = int(imgHeight * eW / imgWidth)
    if sH is not None and sW is None:
        sW = int(imgWidth * sH / imgHeight)
    if sW is not None and sH is None:
        sH = int(imgHeight * sW / imgWidth)
    if None in [sU, sV, sW, sH, eU, eV, eW, eH] and not cfg['dolly']:
        print('At least one of the cropping parameters was not defined, using default ones for 3D kbe.')
        sU, sV = imgWidth / 2.15, imgHeight / 2.15
        sW, sH = int(math.floor(0.90 * imgWidth)), int(math.floor(0.90 * imgHeight))
        eU, eV = imgWidth / 1.85, imgHeight / 1.85
        eW, eH = int(math.floor(0.85 * imgWidth)), int(math.floor(0.85 * imgHeight))
    elif None in [sU, sV, sW, sH, eU, eV, eW, eH] and cfg['dolly']:
        print('At least one of the cropping parameters was not defined, using default ones for dolly effect.')
        sU, sV = imgWidth / 2, imgHeight / 2
        sW, sH = int(math.floor(0.8 * imgWidth)), int(math.floor(0.8 * imgHeight))
        eU, eV = imgWidth / 2, imgHeight / 2
        eW, eH = int(math.floor(0.3 * imgWidth)), int(math.floor(0.3 * imgHeight))
    assert imgHeight >= sV + sH / 2 and sV - sH / 2 >= 0, 'Start window too tall compared to given center'
    assert imgWidth >= sU + sW / 2 and sU - sW / 2 >= 0, 'Start window too tall compared to given center'
    assert imgHeight >= eV + eH / 2 and eV - eH / 2 >= 0, 'End window too tall compared to given center'
    assert imgWidth >= eU + eW / 2 and eU - eW / 2 >= 0, 'End window too tall compared to given center'
    return {'objectFrom': {'dblCenterU': sU, 'dblCenterV': sV, 'intCropWidth': sW, 'intCropHeight': sH},
            'objectTo': {'dblCenterU': eU, 'dblCenterV': eV, 'intCropWidth': eW, 'intCropHeight': eH}}


def init_distributed():
    """Under torchrun (WORLD_SIZE > 1): one process per GPU, NCCL; frames are sharded by Pipeline.__call__."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1 and not torch.distributed.is_initialized():
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        from ken_burns_effect_b200.utils import shard
        shard.bind_to_gpu_numa_node(local)
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    return world


def main(argv):
    cfg = parse(argv)
    init_distributed()
    print('Number of threads used: ', torch.get_num_threads())
    if torch.cuda.is_available():
        # decode on the host (cv2), everything after it on the device: 3 bytes per pixel cross PCIe and the result is the
        # tensor load_image() + (x + 1) / 2 builds, bit for bit (tests/test_gpu_pipeline.py)
        from ken_burns_effect_b200.utils.utils import image_to_tensor
        img = cv2.imread(filename=cfg['input_path'], flags=cv2.IMREAD_COLOR)
        if img is None:
            raise FileNotFoundError(cfg['input_path'])
        image01 = image_to_tensor(img, cfg['pretrained_estim'])
        imgHeight, imgWidth = image01.size(2), image01.size(3)
    else:
        tensorImage = load_image(cfg['input_path'], cfg['pretrained_estim'])
        imgHeight, imgWidth = tensorImage.size(1), tensorImage.size(2)
        image01 = (tensorImage.view(1, 3, imgHeight, imgWidth) + 1) / 2
    zoom_settings = crop_windows(cfg, imgWidth, imgHeight)
    paths = None
    if not cfg['random_weights']:
        paths = [cfg['estim_path'], cfg['refine_path'], cfg['inpaint_path']]
        if cfg['inpaint_depth']:
            paths.append(cfg['inpaint_depth_path'])
    pipe = Pipeline(model_paths=paths, partial_inpainting=cfg['partial'], dolly=cfg['dolly'],
                    output_frames=cfg['output_frames'], pretrain=cfg['pretrained_refine'], d2=cfg['d2'], frames=cfg['frames'])
    with torch.no_grad():
        return pipe(image01, zoom_settings, cfg['output_path'], inpaint_depth=cfg['inpaint_depth'],
                    pretrained_estim=cfg['pretrained_estim'])


if __name__ == '__main__':
    main(sys.argv[1:])
