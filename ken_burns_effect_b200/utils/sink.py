"""Frame sink -- the tail of Pipeline.__call__ (utils/pipeline.py:120-134 of the reference: a PNG loop and a moviepy clip, both
synchronous and single-threaded, after ALL frames have been rendered).

On a B200 the 75..150 frames of an effect are rendered and copied to pinned host memory in a few milliseconds, so encoding is
what a caller of kbe.py waits for (mp4v: ~240 frames/s, PNG: ~30 frames/s per host core at 1024x768).  B200 has no NVENC, so the
sink is host code: frames are consumed batch by batch WHILE the GPU still renders the next ones (an event per batch, see
FrameRenderer.render_into), PNGs are encoded on a thread pool (cv2.imwrite releases the GIL), the mp4 by one ordered writer
thread; the ping-pong clip of the reference (forward, then backward without the turning frame, 25 fps, pipeline.py:132-134) is
written from the pinned buffer, which stays alive until close().
"""
import os
import queue
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np


class FrameSink:
    def __init__(self, output_path, n_frames, write_frames=False, write_video=True, rgb_to_bgr=False, fps=25, workers=None,
                 frame_indices=None, t0=None):
        """output_path: directory (created); n_frames: frames this sink will receive, in order; write_frames: <out>/frames/<i>.png
        (pipeline.py:120-127); write_video: <out>/3d_kbe.mp4; rgb_to_bgr: the `pretrained_estim` colour flip of the reference
        (pipeline.py:125, :131-134); frame_indices: global pose index of every local frame (multi-GPU shards name their PNGs by
        it); t0: perf_counter() of the moment the caller started, for the latency figures."""
        self.out = output_path
        self.n = int(n_frames)
        self.write_frames, self.write_video = bool(write_frames), bool(write_video)
        self.flip = bool(rgb_to_bgr)
        self.fps = fps
        self.indices = list(frame_indices) if frame_indices is not None else list(range(self.n))
        self.t0 = time.perf_counter() if t0 is None else t0
        self.stats = {'frames': self.n, 't_first_frame_ready_s': None, 't_first_png_s': None, 't_last_frame_ready_s': None}
        os.makedirs(self.out, exist_ok=True)
        self._frames = [None] * self.n
        self._got = 0
        self._err = None
        self._pool = None
        self._futs = []
        if self.write_frames:
            os.makedirs(os.path.join(self.out, 'frames'), exist_ok=True)
            self._pool = ThreadPoolExecutor(max_workers=workers or max(2, min(32, (os.cpu_count() or 4) - 2)))
        self._vq = None
        self._vthread = None
        if self.write_video:
            self._vq = queue.Queue()
            self._vthread = threading.Thread(target=self._video_loop, daemon=True)
            self._vthread.start()

    # ---- producer side ------------------------------------------------------------------------------------
    def submit(self, start, frames):
        """frames: uint8 [k,H,W,3] (numpy view of the pinned buffer), the local frames start .. start+k-1, complete."""
        now = time.perf_counter() - self.t0
        if self.stats['t_first_frame_ready_s'] is None:
            self.stats['t_first_frame_ready_s'] = now
        self.stats['t_last_frame_ready_s'] = now
        for i in range(frames.shape[0]):
            f = frames[i]
            self._frames[start + i] = f
            if self._pool is not None:
                self._futs.append(self._pool.submit(self._png, self.indices[start + i], f))
            if self._vq is not None:
                self._vq.put(f)
        self._got += frames.shape[0]

    def _png(self, idx, frame):
        if self.flip:
            frame = cv2.cvtColor(frame, cv2.COLOR_RGB2BGR)
        ok = cv2.imwrite(os.path.join(self.out, 'frames', str(idx) + '.png'), frame)
        if not ok:
            raise IOError(f"cv2.imwrite failed for frame {idx}")
        if self.stats['t_first_png_s'] is None:
            self.stats['t_first_png_s'] = time.perf_counter() - self.t0

    def _video_loop(self):
        vw = None
        try:
            while True:
                f = self._vq.get()
                if f is None:
                    break
                if vw is None:
                    h, w = f.shape[:2]
                    vw = cv2.VideoWriter(os.path.join(self.out, '3d_kbe.mp4'), cv2.VideoWriter_fourcc(*'mp4v'), self.fps, (w, h))
                    if not vw.isOpened():
                        raise IOError("cv2.VideoWriter could not open 3d_kbe.mp4")
                # moviepy expects RGB and the reference flips the BGR tensor's frames with [:, :, ::-1] unless pretrained_estim
                # (pipeline.py:131-134); cv2.VideoWriter expects BGR, i.e. the tensor's own channel order in the default case
                vw.write(np.ascontiguousarray(f[:, :, ::-1]) if self.flip else f)
        except Exception as e:      # surfaced by close()
            self._err = e
        finally:
            if vw is not None:
                vw.release()

    # ---- end of the effect --------------------------------------------------------------------------------
    def close(self):
        """Queue the backward half of the ping-pong clip, wait for every writer, return the timing record."""
        if self._got != self.n:
            raise RuntimeError(f"FrameSink: got {self._got} of {self.n} frames")
        if self._vq is not None:
            for f in reversed(self._frames[:-1]):            # numpyResult + list(reversed(numpyResult))[1:]
                self._vq.put(f)
            self._vq.put(None)
            self._vthread.join()
        if self._pool is not None:
            for fu in self._futs:
                fu.result()
            self._pool.shutdown()
        if self._err is not None:
            raise self._err
        self.stats['t_all_written_s'] = time.perf_counter() - self.t0
        self._frames = None
        return self.stats


class BatchFeeder:
    """Hands the batches FrameRenderer.render_into finishes to a FrameSink from a helper thread: render_into records one CUDA
    event per batch after its device-to-host copy; the thread waits on the event (the GPU keeps rendering the next batch) and
    submits the finished frames."""

    def __init__(self, sink, out_frames):
        self.sink, self.out = sink, out_frames
        self.q = queue.Queue()
        self.err = None
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def __call__(self, start, count, event):
        self.q.put((start, count, event))

    def _loop(self):
        try:
            arr = self.out.numpy()
            while True:
                item = self.q.get()
                if item is None:
                    return
                start, count, event = item
                event.synchronize()
                self.sink.submit(start, arr[start:start + count])
        except Exception as e:
            self.err = e

    def finish(self):
        self.q.put(None)
        self.t.join()
        if self.err is not None:
            raise self.err
