"""Checkpoint I/O and image resize -- mirrors of utils/utils.py:60-73 (resize_image), :190-198 (save_model),
:202-217 (load_models).  The .tar format is kept byte-compatible: torch.save of
{nb_iter, model_state_dict, optimizer_<type>_state_dict, scheduler_<type>_state_dict}."""
import os

import cv2
import torch
import torch.nn.functional as F

cuda = torch.cuda.is_available()
# the reference pins cuda:0 (utils/utils.py:17-18); under torchrun every process owns the GPU of its LOCAL_RANK
device = ("cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))) if cuda else "cpu"


def resize_image(tensorImage, max_size=512):
    if tensorImage.size(0) == 0:
        return None
    w, h = tensorImage.size(3), tensorImage.size(2)
    ratio = float(w) / float(h)
    new_w = min(int(max_size * ratio), max_size)
    new_h = min(int(max_size / ratio), max_size)
    return F.interpolate(input=tensorImage, size=(new_h, new_w), mode='bilinear', align_corners=False)


def image_to_tensor(numpyImage, pretrained_estim=False, device_=None):
    """kbe.py:96-114 + :181 on the device: the uint8 [H,W,3] array cv2.imread returned -> the [1,3,H',W'] float tensor in [0,1]
    kbe.py hands to Pipeline.__call__ (ToTensor, Normalize(.5,.5), crop to multiples of 4, (x+1)/2), bit-identical to the host
    path (kbe.load_image) but with 3 bytes per pixel crossing PCIe instead of 12 and no float images built on the host."""
    import ctypes
    from .. import _native as nat
    dev = torch.device(device_ or device)
    if dev.type != 'cuda':
        raise RuntimeError("image_to_tensor needs a CUDA device (there is no CPU fallback)")
    src = torch.from_numpy(numpyImage)
    if src.dtype != torch.uint8 or src.dim() != 3 or src.shape[2] != 3:
        raise RuntimeError(f"image_to_tensor: expected uint8 [H,W,3], got {src.dtype} {tuple(src.shape)}")
    H, W = src.shape[0], src.shape[1]
    src = src.contiguous().to(dev, non_blocking=True)
    out = torch.empty(1, 3, H - H % 4, W - W % 4, device=dev, dtype=torch.float32)
    nat.check(nat.lib().kb_image_front_end(ctypes.c_void_p(src.data_ptr()), H, W, 1 if pretrained_estim else 0,
                                           ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
              "kb_image_front_end")
    return out


def save_model(models_dict, nb_iter, path='models/trained'):
    for model_type, model in models_dict.items():
        payload = {'nb_iter': nb_iter, 'model_state_dict': model['model'].state_dict()}
        if model.get('opt') is not None:
            payload['optimizer_' + model_type + '_state_dict'] = model['opt'].state_dict()
        if model.get('schedule') is not None:
            payload['scheduler_' + model_type + '_state_dict'] = model['schedule'].state_dict()
        torch.save(payload, path + '/' + model_type + '-' + model['save_name'] + '.tar')


def load_models(models_list, models_paths, continue_training=False):
    """Load `.tar` checkpoints (or bare state_dicts, the reference's fallback for the paper's pretrained
    files) strictly into the given modules.  Unlike the reference, a failing strict load is reported
    instead of being swallowed by a bare `except`."""
    iter_nb = 0
    print('Loading models parameters...')
    for idx, path in enumerate(models_paths):
        ckpt = torch.load(path, map_location='cpu', weights_only=False)
        entry = models_list[idx]
        if isinstance(ckpt, dict) and 'model_state_dict' in ckpt:
            entry['model'].load_state_dict(ckpt['model_state_dict'])
            if continue_training:
                entry['opt'].load_state_dict(ckpt['optimizer_' + entry['type'] + '_state_dict'])
                entry['schedule'].load_state_dict(ckpt['scheduler_' + entry['type'] + '_state_dict'])
                iter_nb = ckpt['nb_iter']
            print('Model ' + entry['type'] + ' loaded succesfully.')
        else:
            entry['model'].load_state_dict(ckpt)
            print('Pre-trained model ' + entry['type'] + ' loaded succesfully.')
    return iter_nb


# ---------------------------------------------------------------------------------------------------------
# batched (B > 1) callers of the render operators -- utils/utils.py:221-300, :370-377 of the reference (training-time view
# synthesis: the only callers of generate_mask and of render_pointcloud with B > 1)
# ---------------------------------------------------------------------------------------------------------

def get_item_in_dict(dict_in, idx):
    """utils/utils.py:370-377: sample `idx` of a (nested) dict of batched values."""
    return {k: (get_item_in_dict(v, idx) if isinstance(v, dict) else v[idx]) for k, v in dict_in.items()}


def get_tensor_shift(objectCommon):
    """utils/utils.py:221-245: the camera shift of the END pose (dblStep = 1) of objectCommon['zoomSettings'] -> [1,3,1].
    Same scalars as the reference's process_shift call; the clone of the whole cloud it makes and drops is not made."""
    from . import common as kb
    zoom = objectCommon['zoomSettings']
    st, focal = kb._pose_settings({'objectFrom': zoom['objectFrom'], 'objectTo': zoom['objectTo'], 'dolly': False}, objectCommon, 1.0)
    sx, sy, sz = kb._shift_scalars(st, objectCommon, objectCommon['dblFocal'])
    return torch.FloatTensor([sx, sy, sz]).view(1, 3, 1).to(objectCommon['tensorRawPoints'].device)


def get_masks(tensorImage, tensorDisparity, tensorDepth, zoom_settings, camera, AFromB=True, tensorContext=None):
    """utils/utils.py:248-300: per-sample camera shifts from batched zoom settings, then either the disocclusion masks of the
    shifted views (AFromB: generate_mask, B >= 1 in one launch group) or the shifted renders themselves (render_pointcloud with
    B >= 1, C = 4 or 68).  Same return tuples as the reference."""
    from . import common as kb
    B = tensorImage.shape[0]
    dblFocal, dblBaseline = camera['focal'], camera['baseline']
    intWidth, intHeight = tensorImage.shape[3], tensorImage.shape[2]
    tensorValid = (kb.spatial_filter(tensorDisparity / tensorDisparity.max(), 'laplacian').abs() < 0.03).float()
    tensorPoints = kb.depth_to_points(tensorDepth * tensorValid, dblFocal)
    shiftList, objectList = [], []
    depth_host = tensorDepth[:, 0, 128:-128, 128:-128].detach().cpu().numpy()      # one D2H for the whole batch
    dmin = tensorDisparity.reshape(B, -1).min(1)[0].tolist()
    dmax = tensorDisparity.reshape(B, -1).max(1)[0].tolist()
    for idx in range(B):
        oc = {'dblFocal': dblFocal, 'dblBaseline': dblBaseline, 'intWidth': intWidth, 'intHeight': intHeight,
              'tensorRawImage': tensorImage[idx], 'tensorRawDisparity': tensorDisparity[idx],
              'dblDispmin': dmin[idx], 'dblDispmax': dmax[idx],
              'objectDepthrange': cv2.minMaxLoc(src=depth_host[idx], mask=None),
              'tensorRawPoints': tensorPoints[idx].view(1, 3, -1),
              'zoomSettings': get_item_in_dict(zoom_settings, idx)}
        shiftList.append(get_tensor_shift(oc))
        objectList.append(oc)
    tensorShift = torch.cat(shiftList)
    pts = tensorPoints.view(B, 3, -1)
    if AFromB:
        return kb.generate_mask(pts, tensorShift, intWidth, intHeight, dblFocal, dblBaseline), tensorShift, objectList
    chans = [tensorImage, tensorDisparity] + ([tensorContext] if tensorContext is not None else [])
    data = torch.cat(chans, 1)
    tensorRender, tensorMasks = kb.render_pointcloud(pts + tensorShift, data.view(B, data.shape[1], -1), intWidth, intHeight,
                                                     dblFocal, dblBaseline)
    return tensorRender, (tensorMasks > 0.0).float(), pts, tensorShift, objectList
