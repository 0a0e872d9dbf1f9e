"""Checkpoint / mask I/O around the hot path in the reference's on-disk formats, without stalling
the GPU.

The reference writes everything with a synchronous ``torch.save`` of freshly materialised tensors:
  masks         {name: int64 0/1 tensor}                 Classification/generate_mask.py:82 (CUDA tensors),
                                                         DDPM/runners/diffusion.py:1039, SD generate_mask.py:108 (CPU tensors)
  checkpoints   {"state_dict", "evaluation_result"}      Classification/unlearn/impl.py:21-30
                [model_sd, optim_sd, step]               DDPM/runners/diffusion.py:598-610
For SD that is 6.9 GB per mask and 3.4 GB per checkpoint, written while the GPU idles.  Here:

  StreamingSaver   snapshots the tensors into pinned host buffers on a side stream (the compute stream only waits for the
                   snapshot's READS to be ordered, not for the file), then pickles and writes on a worker thread; the
                   file appears atomically (write to ``<path>.tmp`` + rename).  ``cuda_tensors=True`` keeps the reference's
                   Classification quirk -- mask files that hold CUDA tensors, because the consumer multiplies without
                   ``.to()`` (unlearn/RL.py:14) -- by saving device clones from the worker thread on its own stream.
  packed side-car  ``<path>.bits``: the same mask at 1 bit per parameter (uint32 words in ``named_parameters()`` order and
                   PyTorch layout) + names / shapes: 64x smaller than the int64 dict.  ``load_mask`` prefers it (and checks
                   names / shapes), so the masked optimizers start from 1.4 MB (ResNet-18) / 107 MB (SD) instead of
                   parsing 89 MB / 6.9 GB; without a side-car it falls back to the reference dict file.
"""
from __future__ import annotations

import math
import os
import queue
import threading
from collections import OrderedDict
from typing import Dict, Optional

import torch

SIDECAR_SUFFIX = ".bits"


def _tree_map(fn, obj):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, dict):
        return type(obj)((k, _tree_map(fn, v)) for k, v in obj.items())
    if isinstance(obj, (list, tuple)):
        return type(obj)(_tree_map(fn, v) for v in obj)
    return obj


class StreamingSaver:
    """``torch.save`` that does not serialise the GPU: ``save()`` returns as soon as the snapshot copies are enqueued."""

    def __init__(self, device=None):
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self._stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._q: "queue.Queue" = queue.Queue()
        self._errors = []
        self._thread = threading.Thread(target=self._worker, daemon=True)
        self._thread.start()

    def _worker(self):
        while True:
            item = self._q.get()
            if item is None:
                self._q.task_done()
                return
            obj, path, event, on_device = item
            try:
                if event is not None:
                    event.synchronize()
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                tmp = path + ".tmp"
                if on_device and self._stream is not None:
                    with torch.cuda.stream(self._stream):   # torch.save's device->host reads stay off the compute stream
                        torch.save(obj, tmp)
                else:
                    torch.save(obj, tmp)
                os.replace(tmp, path)
            except Exception as e:  # surfaced by wait()
                self._errors.append((path, e))
            finally:
                self._q.task_done()

    def save(self, obj, path: str, cuda_tensors: bool = False):
        """Snapshot `obj` (nested dict / list of tensors) now, write it to `path` in the background.
        cuda_tensors=False: tensors are staged to pinned host memory and saved as CPU tensors (DDPM / SD masks, all
        checkpoints).  cuda_tensors=True: device clones are saved, so the file holds CUDA tensors (Classification masks)."""
        event = None
        if self._stream is None:
            snap = _tree_map(lambda t: t.detach().clone(), obj)
        else:
            cur = torch.cuda.current_stream(self.device)
            self._stream.wait_stream(cur)
            with torch.cuda.stream(self._stream):
                if cuda_tensors:
                    snap = _tree_map(lambda t: t.detach().clone() if t.is_cuda else t.detach().clone(), obj)
                else:
                    def stage(t):
                        if not t.is_cuda:
                            return t.detach().clone()
                        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                        h.copy_(t.detach(), non_blocking=True)
                        t.record_stream(self._stream)
                        return h
                    snap = _tree_map(stage, obj)
                event = torch.cuda.Event()
                event.record(self._stream)
        self._q.put((snap, path, event, cuda_tensors))

    def wait(self):
        """block until every queued file is on disk; raises the first error a background write hit"""
        self._q.join()
        if self._errors:
            path, e = self._errors.pop(0)
            raise RuntimeError(f"background save of {path} failed: {e!r}") from e

    def close(self):
        self.wait()
        self._q.put(None)
        self._thread.join(timeout=10)


_default_saver: Optional[StreamingSaver] = None


def default_saver() -> StreamingSaver:
    global _default_saver
    if _default_saver is None:
        _default_saver = StreamingSaver()
    return _default_saver


# ---------------------------------------------------------------------------------------------------------------------
# packed-bit side-car
# ---------------------------------------------------------------------------------------------------------------------
def sidecar_payload(bits_torch_order: torch.Tensor, shapes: "OrderedDict[str, tuple]", ratio=None, key_prefix: str = "") -> dict:
    """bits: uint32 words (int32 storage) of the 0/1 mask in named_parameters() order and PyTorch layout, bit i of word w =
    element 32 w + i (salun_pack_mask)."""
    n = sum(math.prod(s) for s in shapes.values())
    if bits_torch_order.numel() != (n + 31) // 32:
        raise ValueError("bits length does not match the parameter table")
    return {"format": "salun-mask-bits-v1", "n": n, "ratio": ratio, "names": [key_prefix + k for k in shapes],
            "shapes": [tuple(s) for s in shapes.values()], "bits": bits_torch_order}


def save_sidecar(saver: StreamingSaver, path: str, bits_torch_order, shapes, ratio=None, key_prefix: str = ""):
    saver.save(sidecar_payload(bits_torch_order, shapes, ratio, key_prefix), path + SIDECAR_SUFFIX)


def load_mask(path: str, shapes: "OrderedDict[str, tuple]", ctx, device, key_prefix: str = "", prefer_sidecar: bool = True):
    """-> packed mask bits (device, named_parameters() order, PyTorch layout) for `shapes`.  Reads ``<path>.bits`` when it
    exists and matches the parameter table, else the reference's int64 dict at `path` (any of its key conventions:
    bare, ``module.``-prefixed, SD's ``model.diffusion_model.``-stripped)."""
    side = path + SIDECAR_SUFFIX
    names = list(shapes)
    if prefer_sidecar and os.path.exists(side):
        d = torch.load(side, map_location="cpu")
        strip = lambda k: k[len(key_prefix):] if key_prefix and k.startswith(key_prefix) else (k[7:] if k.startswith("module.") else k)
        if d.get("format") == "salun-mask-bits-v1" and [strip(k) for k in d["names"]] == names and \
                [tuple(s) for s in d["shapes"]] == [tuple(s) for s in shapes.values()]:
            return d["bits"].to(device).contiguous()
    m = torch.load(path, map_location="cpu")
    flat = []
    for k, shp in shapes.items():
        t = None
        for cand in (k, key_prefix + k, "module." + k, k.split("model.diffusion_model.")[-1]):
            if cand in m:
                t = m[cand]
                break
        if t is None:
            raise KeyError(f"mask file {path} has no entry for parameter {k}")
        if tuple(t.shape) != tuple(shp):
            raise ValueError(f"mask entry {k}: shape {tuple(t.shape)} != parameter shape {tuple(shp)}")
        flat.append(t.reshape(-1).to(torch.int64))
    return ctx.pack_mask(torch.cat(flat).to(device).contiguous())
