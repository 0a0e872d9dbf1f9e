"""io.py on the CPU: the streaming saver writes the reference's file formats atomically from a worker thread, and the
packed-bit side-car round-trips to the same mask as the int64 dict."""
import os
from collections import OrderedDict

import numpy as np
import torch

from oracle import tail as OT
from unlearn_saliency_b200 import io as IO


class _Ctx:   # stands in for SalunContext.pack_mask on a GPU-less box (oracle packing = the kernel's bit order)
    def pack_mask(self, m64):
        return torch.from_numpy(OT.pack_mask(m64.numpy()).view(np.int32))


def _mask(shapes, seed=0):
    g = torch.Generator().manual_seed(seed)
    return OrderedDict((k, (torch.rand(s, generator=g) < 0.5).to(torch.int64)) for k, s in shapes.items())


def test_streaming_saver_roundtrip_and_atomic_rename(tmp_path):
    sv = IO.StreamingSaver(device="cpu")
    shapes = OrderedDict(a=(3, 5), b=(7,), c=(2, 2, 3, 3))
    m = _mask(shapes)
    p = str(tmp_path / "sub" / "with_0.5.pt")
    sv.save(m, p)
    m["a"].zero_()                               # the snapshot was taken at save(): later writes must not leak into the file
    states = [{"w": torch.arange(5.0)}, {"exp_avg": {"w": torch.zeros(5)}, "step": 3}, 17]
    sv.save(states, str(tmp_path / "ckpt.pth"))
    sv.wait()
    got = torch.load(p)
    want = _mask(shapes)
    assert list(got.keys()) == list(want.keys()) and all(torch.equal(got[k], want[k]) and got[k].dtype == torch.int64 for k in want)
    assert not os.path.exists(p + ".tmp")
    s2 = torch.load(str(tmp_path / "ckpt.pth"))
    assert s2[2] == 17 and s2[1]["step"] == 3 and torch.equal(s2[0]["w"], torch.arange(5.0))
    sv.close()


def test_sidecar_equals_dict_and_is_preferred(tmp_path):
    sv = IO.StreamingSaver(device="cpu")
    shapes = OrderedDict([("conv.weight", (4, 3, 3, 3)), ("bn.weight", (4,)), ("fc.weight", (10, 4))])
    m = _mask(shapes, seed=3)
    flat = torch.cat([v.flatten() for v in m.values()])
    bits = _Ctx().pack_mask(flat)
    p = str(tmp_path / "with_0.5.pt")
    sv.save(m, p)
    IO.save_sidecar(sv, p, bits, shapes, ratio=0.5)
    sv.wait()
    assert os.path.getsize(p + IO.SIDECAR_SUFFIX) < os.path.getsize(p)
    via_side = IO.load_mask(p, shapes, _Ctx(), "cpu")
    via_dict = IO.load_mask(p, shapes, _Ctx(), "cpu", prefer_sidecar=False)
    assert torch.equal(via_side, bits) and torch.equal(via_dict, bits)
    # a side-car that does not match the parameter table is ignored, the dict is read
    other = OrderedDict([("conv.weight", (4, 3, 3, 3)), ("bn.weight", (4,)), ("fc.weight", (4, 10))])
    try:
        IO.load_mask(p, other, _Ctx(), "cpu")
        assert False, "shape mismatch must be reported"
    except ValueError:
        pass
    # DataParallel-prefixed keys (DDPM masks, runners/diffusion.py:1039)
    pm = str(tmp_path / "ddpm.pt")
    torch.save({"module." + k: v for k, v in m.items()}, pm)
    assert torch.equal(IO.load_mask(pm, shapes, _Ctx(), "cpu"), bits)
    sv.close()
