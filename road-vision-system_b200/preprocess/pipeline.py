"""PreprocessPipeline -- drop-in for /root/reference/src/preprocess/pipeline.py:7-45.

Same constructor (`enabled`, `chain`, `auto_gate`), same `pipeline(image, ts=None)` call.
Differences are internal: an adjacent CLAHEDehaze -> MedianDerain pair is dispatched as ONE
fused GPU pass (the ops stay individually callable), and `process_batch` is the new
multi-frame entry (no reference counterpart; the reference loop keeps one frame in flight,
main_preview.py:88-142).
"""
from typing import Any, Dict

import numpy as np

from .._native import DeviceArray, Params, default_context
from .base import as_bgr_u8
from .ops import CLAHEDehaze, MedianDerain
from .ops import clahe_dehaze as _clahe
from .ops import median_derain as _median
from .registry import get_op_class


def _is_cuda_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


class PreprocessPipeline:
    def __init__(self, config: Dict[str, Any], context=None):
        """`config`: the reference's `preprocess:` dict (pipeline.py:13-22).  `context` (new, optional): the rvb200.Context
        this pipeline runs on -- give every camera stream / worker thread its own (a context is not thread-safe);
        without it the process-wide default context of `config["device"]` is used."""
        self._context = context
        self._snap = None
        self.enabled = bool(config.get("enabled", True))
        self.chain_cfg = config.get("chain", []) or []
        self.auto_gate_cfg = config.get("auto_gate", {}) or {}
        self.device = config.get("device")          # optional new key; the stock YAML does not set it
        self.ops = []
        for node in self.chain_cfg:
            name = node.get("name")
            params = node.get("params", {})
            cls = get_op_class(name)
            self.ops.append(cls(**params))

    # -- helpers -------------------------------------------------------------------------
    def _ctx(self):
        return self._context if self._context is not None else default_context(self.device)

    def _gate(self):
        enable = bool(self.auto_gate_cfg.get("enable_low_contrast_gate", False))
        return enable, float(self.auto_gate_cfg.get("contrast_thresh", 20.0))

    def _low_contrast(self, image) -> bool:
        """pipeline.py:24-30: gray span < thresh (gray = cv2 BGR2GRAY fixed point, computed on the GPU)."""
        span = int(self._ctx().gray_span(as_bgr_u8(image)[None])[0])
        return span < self._gate()[1]

    def _segments(self):
        """Fold the op list into fused GPU passes: [(Params | op)].  Params are re-read on every call, like the reference
        (clahe_dehaze.py:14-17, median_derain.py:11-13: mutating `op.params` between frames takes effect); the folded result
        is reused for as long as the op objects and their params dicts compare equal to the snapshot it was built from."""
        snap = self._snap
        ops = self.ops
        if snap is not None and len(snap[0]) == len(ops):
            for (op0, params0), op in zip(snap[0], ops):
                if op0 is not op or params0 != op.params:
                    break
            else:
                for sg in snap[1]:
                    if isinstance(sg, Params):
                        sg.gate_enable = 0
                return snap[1]
        segs = self._fold()
        self._snap = ([(op, dict(op.params)) for op in ops], segs)
        return segs

    def _fold(self):
        segs, i = [], 0
        while i < len(self.ops):
            op = self.ops[i]
            nxt = self.ops[i + 1] if i + 1 < len(self.ops) else None
            if type(op).__call__ is CLAHEDehaze.__call__ and nxt is not None and type(nxt).__call__ is MedianDerain.__call__:
                space, clip, grid = _clahe.coerce(op.params)
                segs.append(Params.make(space, clip, grid, _median.coerce(nxt.params), clahe=True))
                i += 2
            elif type(op).__call__ is CLAHEDehaze.__call__:
                space, clip, grid = _clahe.coerce(op.params)
                segs.append(Params.make(space, clip, grid, 0, clahe=True))
                i += 1
            elif type(op).__call__ is MedianDerain.__call__:
                segs.append(Params.make(ksize=_median.coerce(op.params), clahe=False))
                i += 1
            else:
                segs.append(op)         # foreign operator: called as in the reference
                i += 1
        return segs

    # -- per-frame contract --------------------------------------------------------------
    def __call__(self, image: np.ndarray, ts: float = None) -> np.ndarray:
        if not self.enabled or not self.ops:
            return image
        segs = self._segments()
        if self._gate()[0]:
            if len(segs) == 1 and isinstance(segs[0], Params):
                # one fused pass: the gate's gray span comes out of the histogram pass of the same upload
                seg = segs[0]
                seg.gate_enable, seg.gate_thresh = 1, self._gate()[1]
                flag = np.ones(1, np.int32)
                out = self._ctx().chain(as_bgr_u8(image)[None], seg, processed=flag)[0]
                return out if flag[0] else image          # skipped: the input object itself, as pipeline.py:39-40
            if not self._low_contrast(image):
                return image
        out = image
        for seg in segs:
            if isinstance(seg, Params):
                out = self._ctx().chain(as_bgr_u8(out)[None], seg)[0]
            else:
                out = seg(out)
        return out

    # -- new: batched multi-frame entry --------------------------------------------------
    def process_batch(self, frames: np.ndarray, out: np.ndarray = None) -> np.ndarray:
        """(B,H,W,3) uint8 -> (B,H,W,3) uint8; equals B per-frame calls bit for bit.

        `frames` / `out` may be pinned arrays from `Context.pinned_empty` (fastest: H2D, kernels
        and D2H of consecutive chunks overlap).  Frames the low-contrast gate skips are copied
        through unchanged.  `out="device"` keeps the result on the GPU (a `DeviceArray`, usable from torch / cupy through
        `__cuda_array_interface__`): nothing is copied back, which is what a detector on the same GPU wants
        (main_preview.py:99).
        """
        if _is_cuda_tensor(frames):
            return self._process_batch_device(frames, out)
        if isinstance(out, str):
            if out != "device":
                raise ValueError('out must be an array, None or "device"')
            return self._process_batch_keep(frames)
        if not isinstance(frames, np.ndarray) or frames.ndim != 4 or frames.shape[3] != 3 or frames.dtype != np.uint8:
            raise ValueError("frames must be a (B,H,W,3) uint8 numpy array (or a torch CUDA uint8 tensor of that shape)")
        if not self.enabled or not self.ops:
            if out is None:
                return frames
            np.copyto(out, frames)
            return out
        ctx = self._ctx()
        gate, thresh = self._gate()
        cur = frames
        segs = self._segments()
        if gate and not (len(segs) == 1 and isinstance(segs[0], Params)):
            # general case: decide per frame up front, run the chain on the selected frames only
            span = ctx.gray_span(np.ascontiguousarray(frames))
            sel = np.flatnonzero(span < thresh)
            res = np.array(frames, copy=True) if out is None else out
            if out is not None:
                np.copyto(out, frames)
            if len(sel):
                sub = np.ascontiguousarray(frames[sel])
                for seg in segs:
                    sub = ctx.chain(sub, seg) if isinstance(seg, Params) else np.stack([seg(f) for f in sub])
                res[sel] = sub
            return res
        for idx, seg in enumerate(segs):
            last = idx == len(segs) - 1
            if isinstance(seg, Params):
                if gate:
                    seg.gate_enable, seg.gate_thresh = 1, thresh
                cur = ctx.chain(cur, seg, out=out if last else None)
            else:
                cur = np.stack([seg(f) for f in cur])
                if last and out is not None:
                    np.copyto(out, cur)
                    cur = out
        return cur

    def _process_batch_keep(self, frames):
        """Host frames in, result left on the GPU (no D2H at all)."""
        if not isinstance(frames, np.ndarray) or frames.ndim != 4 or frames.shape[3] != 3 or frames.dtype != np.uint8:
            raise ValueError("frames must be a (B,H,W,3) uint8 numpy array")
        ctx = self._ctx()
        frames = np.ascontiguousarray(frames)
        segs = self._segments() if (self.enabled and self.ops) else []
        if not segs or self._gate()[0] or not all(isinstance(sg, Params) for sg in segs):
            # identity, gated or foreign operators: compute on the host path, then upload
            res = self.process_batch(frames)
            dev = DeviceArray(ctx, res.shape)
            ctx._ck(ctx._lib.rv_memcpy(ctx._h, dev.ptr, np.ascontiguousarray(res).ctypes.data, dev.nbytes, 0))
            return dev
        cur = frames
        for sg in segs:
            dst = DeviceArray(ctx, frames.shape)
            ctx.submit_io(cur, sg, shape=frames.shape[:3], out=dst)
            ctx.wait()
            cur = dst
        return cur

    def _process_batch_device(self, frames, out=None):
        """Device-resident batch: a contiguous torch CUDA uint8 tensor (B,H,W,3) in, a tensor of the same kind out.
        Work is enqueued on torch's current stream (no host synchronisation): use the result with torch as usual."""
        import torch
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_contiguous():
            raise ValueError("frames must be a contiguous (B,H,W,3) uint8 CUDA tensor")
        if out is not None and not (_is_cuda_tensor(out) and out.dtype == torch.uint8 and tuple(out.shape) == tuple(frames.shape)
                                    and out.device == frames.device and out.is_contiguous()):
            raise ValueError("out must be a contiguous uint8 CUDA tensor of the input's shape on the input's device")
        if not self.enabled or not self.ops:
            return frames if out is None else out.copy_(frames)
        if self._gate()[0]:
            raise ValueError("the low-contrast gate is not supported for device-resident batches; use numpy frames")
        ctx = self._context if (self._context is not None and self._context.device == frames.device.index) \
            else default_context(frames.device.index)
        b, h, w, _ = frames.shape
        segs = self._segments()
        if not all(isinstance(sg, Params) for sg in segs):
            raise ValueError("device-resident batches support only CLAHEDehaze / MedianDerain chains")
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        cur = frames
        for idx, sg in enumerate(segs):
            dst = out if (idx == len(segs) - 1 and out is not None) else torch.empty_like(frames)
            if stream == 0:        # legacy default stream: run on the context's stream, synchronously
                torch.cuda.current_stream(frames.device).synchronize()
                ctx.chain_device(cur.data_ptr(), dst.data_ptr(), b, h, w, sg)
            else:
                ctx.submit_device(cur.data_ptr(), dst.data_ptr(), b, h, w, sg, stream=stream)
            cur = dst
        return cur

    # -- new: chain fused with the detector-input stage ------------------------
    def process_batch_to_tensor(self, frames: np.ndarray, size: int = 640, pad_value: int = 114, want_frames: bool = False,
                                out=None, padding_present: bool = False):
        """(B,H,W,3) uint8 -> ((B,3,size,size) float16 network input, processed frames or None).

        `out`: optional (B,3,size,size) float16 array for the tensor (pinned, from `Context.pinned_empty(shape, np.float16)`, for
        the overlapped H2D / kernels / D2H pipeline), or "device" to leave the tensor on the GPU (a `DeviceArray`): with pinned
        `frames` the only PCIe traffic is then the upload of the frames.
        `padding_present=True`: `out` already holds the letterbox padding rows (`Context.fill_tensor_padding`, done once per buffer
        of a ring), so only the image rows come back over PCIe -- 56 % of the tensor for 1080p frames.

        The tensor is what the detector stage builds from `proc` right after the chain (main_preview.py:99 ->
        yolo_ultralytics.py:28-35: letterbox, BGR->RGB, HWC->CHW, /255, half).  When the chain is the stock
        CLAHEDehaze -> MedianDerain pair and the down-scale is an exact integer (1080p -> 640), the resize happens inside
        the chain kernel and the full-resolution frames are only written if `want_frames` is set.
        """
        if not isinstance(frames, np.ndarray) or frames.ndim != 4 or frames.shape[3] != 3 or frames.dtype != np.uint8:
            raise ValueError("frames must be a (B,H,W,3) uint8 numpy array")
        ctx = self._ctx()
        segs = self._segments() if (self.enabled and self.ops) else []
        keep = isinstance(out, str)
        if keep and out != "device":
            raise ValueError('out must be an array, None or "device"')
        if out is not None and not keep and (out.shape != (frames.shape[0], 3, size, size) or out.dtype != np.float16
                                             or not out.flags.c_contiguous):
            raise ValueError("out must be a C-contiguous (B,3,size,size) float16 array")
        if self._gate()[0] or len(segs) != 1 or not isinstance(segs[0], Params):
            proc = self.process_batch(frames)
            t = ctx.letterbox_f16(proc, size, pad_value)
            if keep:
                dev = DeviceArray(ctx, t.shape, np.float16)
                ctx._ck(ctx._lib.rv_memcpy(ctx._h, dev.ptr, t.ctypes.data, dev.nbytes, 0))
                t = dev
            elif out is not None:
                np.copyto(out, t)
                t = out
            return t, (proc if want_frames else None)
        if keep:
            frames = np.ascontiguousarray(frames)
            dev = DeviceArray(ctx, (frames.shape[0], 3, size, size), np.float16)
            full = ctx._pooled_pinned(frames.shape) if (want_frames and frames.nbytes <= (64 << 20)) else \
                (np.empty_like(frames) if want_frames else None)
            ctx.submit_io(frames, segs[0], out=full, tensor=dev, size=size, pad_value=pad_value)
            ctx.wait()
            return dev, full
        return ctx.chain_letterbox(frames, segs[0], size, pad_value, want_full=want_frames, out=out,
                                   padding_present=bool(padding_present and out is not None))
