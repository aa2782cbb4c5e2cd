"""CPU: kbe.py option parsing, image preparation and crop-window rules (kbe.py:42-169 of the reference)."""
import importlib.util
import os

import cv2
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("kbe_cli", os.path.join(ROOT, "kbe.py"))
kbe = importlib.util.module_from_spec(spec)
spec.loader.exec_module(kbe)


def test_defaults_match_reference():
    cfg = kbe.parse([])
    assert cfg['input_path'] == 'images/doublestrike.jpg' and cfg['output_path'] == 'images/kbe'
    assert cfg['inpaint_path'] == './models/trained/inpainting-color.tar'
    assert cfg['estim_path'] == './models/trained/disparity-estimation-no-mask.tar'
    assert cfg['frames'] == 75 and not cfg['dolly'] and not cfg['partial']


def test_flags():
    cfg = kbe.parse(['--in', 'a.png', '--out', 'o', '--dolly', '--write-frames', '--startU', '10', '--endH', '40',
                     '--partial-conv', '--frames', '150', '--2d', '--pretrained-refine'])
    assert cfg['input_path'] == 'a.png' and cfg['dolly'] and cfg['output_frames'] and cfg['startU'] == 10
    assert cfg['endH'] == 40 and cfg['partial'] and cfg['frames'] == 150 and cfg['d2'] and cfg['pretrained_refine']


def test_default_windows_1024x768():
    z = kbe.crop_windows(kbe.parse([]), 1024, 768)
    assert z['objectFrom'] == {'dblCenterU': 1024 / 2.15, 'dblCenterV': 768 / 2.15, 'intCropWidth': 921, 'intCropHeight': 691}
    assert z['objectTo'] == {'dblCenterU': 1024 / 1.85, 'dblCenterV': 768 / 1.85, 'intCropWidth': 870, 'intCropHeight': 652}
    z = kbe.crop_windows(kbe.parse(['--dolly']), 1024, 768)
    assert z['objectFrom']['intCropWidth'] == 819 and z['objectTo']['intCropWidth'] == 307 and z['objectTo']['intCropHeight'] == 230


def test_window_asserts():
    cfg = kbe.parse(['--startU', '10', '--startV', '10', '--startW', '400', '--startH', '300', '--endU', '512', '--endV',
                     '384', '--endW', '100', '--endH', '80'])
    with pytest.raises(AssertionError):
        kbe.crop_windows(cfg, 1024, 768)


def test_load_image_crops_to_multiple_of_4(tmp_path):
    img = np.random.default_rng(0).integers(0, 256, (37, 50, 3), dtype=np.uint8)
    p = str(tmp_path / "x.png")
    cv2.imwrite(p, img)
    t = kbe.load_image(p, False)
    assert t.shape == (3, 36, 48)
    ref = (torch.from_numpy(img[:36, :48]).permute(2, 0, 1).float() / 255 - 0.5) / 0.5
    assert torch.equal(t, ref)
    assert float(((t + 1) / 2).min()) >= 0.0


def test_pinned_pool_never_recycles_frames_a_caller_still_holds(monkeypatch):
    """process_kenburns returns numpy views of a pooled pinned buffer: while any of them is alive the buffer must not be handed
    out again (tensor.numpy() hangs a new tensor wrapper on the array, so only the storage's use count tells)."""
    import torch
    from ken_burns_effect_b200.utils import common as kb
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)      # no CUDA here: plain host memory stands in
    monkeypatch.setattr(kb, "_PINNED", {})
    monkeypatch.setattr(kb, "_PINNED_ORDER", [])
    a = kb.pinned_frames((2, 4, 4, 3))
    ptr_a = a.data_ptr()
    frames = a.numpy()
    views = [frames[i] for i in range(2)]
    del a, frames
    b = kb.pinned_frames((2, 4, 4, 3))
    assert b.data_ptr() != ptr_a
    del views
    held = b                                   # a caller that keeps the tensor itself (Pipeline.run_many's result list)
    ptr_b = b.data_ptr()
    del b
    c = kb.pinned_frames((2, 4, 4, 3))
    assert c.data_ptr() == ptr_a and c.data_ptr() != ptr_b
    del held
    # byte cap: idle buffers of other shapes are released, oldest first
    monkeypatch.setattr(kb, "PINNED_POOL_BYTES", 200)
    del c
    for n in range(3, 8):
        kb.pinned_frames((n, 4, 4, 3))
    assert sum(e[0].numel() for bufs in kb._PINNED.values() for e in bufs) <= 200 + 7 * 48


def test_frame_sink_writes_pngs_and_pingpong_video(tmp_path):
    """utils/sink.py against the files utils/pipeline.py:120-134 of the reference produces: <out>/frames/<i>.png (lossless) and
    a 25 fps clip of forward + backward-without-the-turning-frame (2n-1 frames), fed in batches like the renderer does."""
    import cv2
    import numpy as np
    from ken_burns_effect_b200.utils.sink import FrameSink
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 255, (7, 48, 64, 3), dtype=np.uint8)
    sink = FrameSink(str(tmp_path / "o"), 7, write_frames=True, write_video=True, frame_indices=[10, 11, 12, 13, 14, 15, 16])
    sink.submit(0, frames[0:3])
    sink.submit(3, frames[3:7])
    st = sink.close()
    assert st['frames'] == 7 and st['t_all_written_s'] >= st['t_first_frame_ready_s']
    for j, i in enumerate(range(10, 17)):
        assert (cv2.imread(str(tmp_path / "o" / "frames" / f"{i}.png")) == frames[j]).all()
    cap = cv2.VideoCapture(str(tmp_path / "o" / "3d_kbe.mp4"))
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == 13 and abs(cap.get(cv2.CAP_PROP_FPS) - 25) < 1e-6
    # the colour flip of `pretrained_estim` (pipeline.py:125): PNGs hold the frame converted RGB -> BGR
    sink = FrameSink(str(tmp_path / "p"), 2, write_frames=True, write_video=False, rgb_to_bgr=True)
    sink.submit(0, frames[0:2])
    sink.close()
    assert (cv2.imread(str(tmp_path / "p" / "frames" / "1.png")) == frames[1][:, :, ::-1]).all()
    import pytest
    with pytest.raises(RuntimeError):
        FrameSink(str(tmp_path / "q"), 3, write_video=False).close()       # frames missing
