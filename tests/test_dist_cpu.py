"""Host-side data-parallel logic on CPU: world_size 2, gloo backend, 127.0.0.1 rendezvous.

The CUDA engine is replaced by a tiny CPU stand-in exposing the same surface (forward_backward / grads / ctx), so what
is exercised is the sharding, shard-size weighting and all-reduce logic of
unlearn_saliency_b200.classification.{generate_mask.accumulate_saliency, unlearn.steps.masked_step}.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


class FakeCtx:
    def saliency_accumulate_flat(self, g, acc):
        acc += g

    def abs_(self, a):
        return a.abs_()


class FakeEngine:
    """linear softmax classifier with the engine's host-visible surface"""

    def __init__(self):
        g = torch.Generator().manual_seed(0)
        self.device = torch.device("cpu")
        self.W = torch.randn(10, 3 * 4 * 4, generator=g) * 0.1
        self.params = self.W.flatten().clone()
        self.grads = torch.zeros_like(self.params)
        self.ctx = FakeCtx()
        self._loss = torch.zeros(1)

    def eval(self):
        return self

    def forward_backward(self, x, y, loss_sign=1.0, want_logits=False, train=None):
        w = self.params.view(10, -1).clone().requires_grad_(True)
        logits = x.flatten(1) @ w.t()
        loss = loss_sign * F.cross_entropy(logits, y)
        (g,) = torch.autograd.grad(loss, w)
        self.grads.copy_(g.flatten())
        self._loss = loss.detach().reshape(1)
        return self._loss, (logits.detach() if want_logits else None)


class PlainSGD:
    def __init__(self, eng, lr):
        self.eng, self.lr = eng, lr

    def step(self):
        self.eng.params -= self.lr * self.eng.grads


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data(n=22):
    g = torch.Generator().manual_seed(1)
    return torch.rand(n, 3, 4, 4, generator=g), torch.randint(0, 10, (n,), generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from unlearn_saliency_b200.classification.generate_mask import accumulate_saliency
    from unlearn_saliency_b200.classification.unlearn.steps import masked_step
    x, y = _data()
    loader = [(x[i:i + 8], y[i:i + 8]) for i in range(0, 22, 8)]  # 3 batches, last one partial (6)
    eng = FakeEngine()
    acc = accumulate_saliency(eng, loader)
    eng2 = FakeEngine()
    opt = PlainSGD(eng2, 0.1)
    for bx, by in loader[::-1]:  # includes a 6-sample batch split 3/3 and 8-sample batches split 4/4
        masked_step(eng2, opt, bx, by)
    xb, yb = x[:5], y[:5]  # odd batch: shards of 3 and 2 -> the shard-size weighting matters
    masked_step(eng2, opt, xb, yb)
    q.put((rank, acc.clone(), eng2.params.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    from unlearn_saliency_b200.classification.generate_mask import accumulate_saliency
    from unlearn_saliency_b200.classification.unlearn.steps import masked_step
    x, y = _data()
    loader = [(x[i:i + 8], y[i:i + 8]) for i in range(0, 22, 8)]
    eng = FakeEngine()
    acc = accumulate_saliency(eng, loader)
    eng2 = FakeEngine()
    opt = PlainSGD(eng2, 0.1)
    for bx, by in loader[::-1]:
        masked_step(eng2, opt, bx, by)
    masked_step(eng2, opt, x[:5], y[:5])
    for rank, acc_r, params_r in res:
        torch.testing.assert_close(acc_r, acc, rtol=1e-5, atol=1e-6)       # one all-reduce(sum) of the accumulator
        torch.testing.assert_close(params_r, eng2.params, rtol=1e-5, atol=1e-6)  # sharded batch == global-batch mean grad
    assert torch.equal(res[0][2], res[1][2])  # replicas stay bit-identical


def test_dp_shard_tiles_the_arena():
    """salun_dp_shard (host-only): the shards of the fused DP kernels tile [0, n) exactly, in whole mask words / vectors"""
    import ctypes as C
    from unlearn_saliency_b200 import _lib
    lib = _lib.lib()
    for n in (0, 1, 127, 128, 129, 11173962, 38632323):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for rank in range(world):
                lo, hi = C.c_int64(), C.c_int64()
                assert lib.salun_dp_shard(n, rank, world, C.byref(lo), C.byref(hi)) == 0
                assert lo.value == prev and lo.value <= hi.value <= n
                assert lo.value % 128 == 0 or lo.value == n
                prev = hi.value
            assert prev == n
    lo, hi = C.c_int64(), C.c_int64()
    assert lib.salun_dp_shard(10, 2, 2, C.byref(lo), C.byref(hi)) != 0      # rank out of range


def _adam_worker(rank, world, port, q):
    """The exchange DistMaskedAdam's two kernels implement (salun_dp_grad_reduce_sumsq -> barrier ->
    salun_dp_masked_adam_step), restated with gloo collectives on the CPU: reduce the owned shard of the averaged gradient,
    exchange the shard norms, clip with the GLOBAL norm, mask, Adam on the shard with shard-sized moments, all-gather."""
    import ctypes as C
    from unlearn_saliency_b200 import _lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = _lib.lib()
    n = 1000
    g0 = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=g0)
    mask = (torch.rand(n, generator=g0) < 0.5).float()
    lo, hi = C.c_int64(), C.c_int64()
    lib.salun_dp_shard(n, rank, world, C.byref(lo), C.byref(hi))
    lo, hi = lo.value, hi.value
    m1, m2 = torch.zeros(hi - lo), torch.zeros(hi - lo)
    for step in range(1, 4):
        grad = torch.randn(n, generator=torch.Generator().manual_seed(100 * step + rank)) * (0.2 if step == 2 else 0.01)
        grads = [torch.empty(n) for _ in range(world)]
        dist.all_gather(grads, grad)                      # stands in for the peer-mapped gradient arenas
        gs = sum(g_[lo:hi] for g_ in grads) / world       # phase 1: owned shard of the averaged gradient ...
        norms = [torch.empty(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(norms, gs.double().square().sum().reshape(1))   # ... and its sum of squares, visible to all
        total = float(torch.stack(norms).sum().sqrt())    # phase 2: global pre-clip norm
        coef = min(1.0, 1.0 / (total + 1e-6))
        gi = gs * coef * mask[lo:hi]
        m1 = m1 + (gi - m1) * (1 - 0.9)
        m2 = 0.999 * m2 + (1 - 0.999) * gi * gi
        denom = m2.sqrt() / (1 - 0.999 ** step) ** 0.5 + 1e-8
        new = p[lo:hi] - 1e-3 / (1 - 0.9 ** step) * m1 / denom
        per = (n + world - 1) // world
        per = (per + 127) // 128 * 128
        pad = torch.zeros(per)
        pad[: hi - lo] = new
        allp = [torch.empty(per) for _ in range(world)]
        dist.all_gather(allp, pad)                        # the new weights land in every replica
        p = torch.cat(allp)[:n].clone()
    q.put((rank, p))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_clip_mask_adam_exchange_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_adam_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    # single process: the reference's statements on the DDP-averaged gradient (runners/diffusion.py:582-593)
    n = 1000
    g0 = torch.Generator().manual_seed(0)
    p = torch.nn.Parameter(torch.randn(n, generator=g0))
    mask = (torch.rand(n, generator=g0) < 0.5).float()
    opt = torch.optim.Adam([p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for step in range(1, 4):
        gr = [torch.randn(n, generator=torch.Generator().manual_seed(100 * step + r)) * (0.2 if step == 2 else 0.01)
              for r in range(world)]
        p.grad = sum(gr) / world
        torch.nn.utils.clip_grad_norm_([p], 1.0)
        p.grad *= mask
        opt.step()
    for rank, pr_ in res:
        torch.testing.assert_close(pr_, p.detach(), rtol=1e-5, atol=1e-7)
    assert torch.equal(res[0][1], res[1][1])
