"""Checkpoint I/O and image resize -- mirrors of utils/utils.py:60-73 (resize_image), :190-198 (save_model),
:202-217 (load_models).  The .tar format is kept byte-compatible: torch.save of
{nb_iter, model_state_dict, optimizer_<type>_state_dict, scheduler_<type>_state_dict}."""
import os

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
