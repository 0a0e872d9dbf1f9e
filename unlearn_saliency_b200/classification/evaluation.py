"""Evaluation passes on the engine's forward kernels: mirrors of

  validate(val_loader, model, criterion, args)          Classification/trainer/val.py:6-72
  collect_prob(data_loader, model)                      Classification/evaluation/SVC_MIA.py:25-50
  entropy / m_entropy / SVC_fit_predict / SVC_MIA       SVC_MIA.py:8-22, 53-148

The per-batch work -- eval-mode forward, cross-entropy, top-1, softmax -- runs in libsalun (salun_resnet_forward +
salun_eval_logits); loss and hit counts accumulate on the device and are read once at the end of the pass (the
reference syncs twice per batch).  The SVC attack itself is scikit-learn on the CPU, as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .._lib import check
from ..tail import _ptr, _stream
from .common import as_engine, check_criterion
from .unlearn.steps import unpack_batch


def _eval_batch(engine, image, target, want_probs=False, acc=None):
    image = image.to(engine.device, non_blocking=True).float().contiguous()
    logits = engine.forward(image)
    n, K = logits.shape
    probs = torch.empty_like(logits) if want_probs else None
    if target is not None:
        target = target.to(engine.device, non_blocking=True).long().contiguous()
    loss_sum, correct = (acc if acc is not None else (None, None))
    check(engine._lib.salun_eval_logits(engine.ctx.handle, _ptr(logits), _ptr(target) if loss_sum is not None else None, n, K,
                                        _ptr(probs), _ptr(loss_sum), _ptr(correct), _stream(engine.device)),
          "salun_eval_logits")
    return logits, probs, target


def validate(val_loader, model, criterion, args):
    """Run evaluation (val.py:6-72): returns top1.avg in percent.  Accepts the reference's nn.Module or an engine."""
    check_criterion(criterion)
    engine = as_engine(model, args)
    was_training = engine.training
    engine.eval()                                                  # val.py:13
    loss_sum = torch.zeros(1, dtype=torch.float64, device=engine.device)
    correct = torch.zeros(1, dtype=torch.int64, device=engine.device)
    seen = 0
    for i, data in enumerate(val_loader):
        image, target = unpack_batch(data, args)
        _eval_batch(engine, image, target, acc=(loss_sum, correct))
        seen += int(image.shape[0])
        if i % args.print_freq == 0:                               # val.py:62-69 (running averages; one sync per print)
            print("Test: [{0}/{1}]\t" "Loss ({2:.4f})\t" "Accuracy ({3:.3f})".format(
                i, len(val_loader), float(loss_sum.item()) / max(1, seen), 100.0 * float(correct.item()) / max(1, seen)))
    top1 = 100.0 * float(correct.item()) / max(1, seen)
    print("valid_accuracy {top1:.3f}".format(top1=top1))
    engine.train(was_training)
    validate.last_loss = float(loss_sum.item()) / max(1, seen)
    return top1


def collect_prob(data_loader, model):
    """softmax probabilities and targets of every sample of the loader (SVC_MIA.py:25-50), on the engine's device"""
    if data_loader is None:
        return torch.zeros([0, 10]), torch.zeros([0])
    engine = as_engine(model)
    was_training = engine.training
    engine.eval()
    prob, targets = [], []
    for data in data_loader:
        image, target = unpack_batch(data)
        _, p, _ = _eval_batch(engine, image, None, want_probs=True)
        prob.append(p)
        targets.append(target.to(engine.device))
    engine.train(was_training)
    return torch.cat(prob), torch.cat(targets)


def entropy(p, dim=-1, keepdim=False):
    return -torch.where(p > 0, p * p.log(), p.new([0.0])).sum(dim=dim, keepdim=keepdim)


def m_entropy(p, labels, dim=-1, keepdim=False):
    """SVC_MIA.py:12-22 as written (both branches of the reference take log(p); the column indexing by `labels` too)"""
    log_prob = torch.where(p > 0, p.log(), torch.tensor(1e-30).to(p.device).log())
    reverse_prob = 1 - p
    log_reverse_prob = torch.where(p > 0, p.log(), torch.tensor(1e-30).to(p.device).log())
    modified_probs = p.clone()
    modified_probs[:, labels] = reverse_prob[:, labels]
    modified_log_probs = log_reverse_prob.clone()
    modified_log_probs[:, labels] = log_prob[:, labels]
    return -torch.sum(modified_probs * modified_log_probs, dim=dim, keepdim=keepdim)


def SVC_fit_predict(shadow_train, shadow_test, target_train, target_test):
    from sklearn.svm import SVC
    n_st, n_ste, n_tt, n_tte = shadow_train.shape[0], shadow_test.shape[0], target_train.shape[0], target_test.shape[0]
    X_shadow = torch.cat([shadow_train, shadow_test]).cpu().numpy().reshape(n_st + n_ste, -1)
    Y_shadow = np.concatenate([np.ones(n_st), np.zeros(n_ste)])
    clf = SVC(C=3, gamma="auto", kernel="rbf")
    clf.fit(X_shadow, Y_shadow)
    accs = []
    if n_tt > 0:
        accs.append(clf.predict(target_train.cpu().numpy().reshape(n_tt, -1)).mean())
    if n_tte > 0:
        accs.append(1 - clf.predict(target_test.cpu().numpy().reshape(n_tte, -1)).mean())
    return np.mean(accs)


def SVC_MIA(shadow_train, target_train, target_test, shadow_test, model):
    """the five membership-inference scores of SVC_MIA.py:76-148 (correctness, confidence, entropy, m_entropy, prob)"""
    st_p, st_y = collect_prob(shadow_train, model)
    ste_p, ste_y = collect_prob(shadow_test, model)
    tt_p, tt_y = collect_prob(target_train, model)
    tte_p, tte_y = collect_prob(target_test, model)
    corr = lambda p, y: (torch.argmax(p, axis=1) == y.to(p.device)).int()
    conf = lambda p, y: torch.gather(p, 1, y.to(p.device).long()[:, None])
    feats = {
        "correctness": [corr(st_p, st_y), corr(ste_p, ste_y), corr(tt_p, tt_y), corr(tte_p, tte_y)],
        "confidence": [conf(st_p, st_y), conf(ste_p, ste_y), conf(tt_p, tt_y), conf(tte_p, tte_y)],
        "entropy": [entropy(st_p), entropy(ste_p), entropy(tt_p), entropy(tte_p)],
        "m_entropy": [m_entropy(st_p, st_y.long()), m_entropy(ste_p, ste_y.long()),
                      m_entropy(tt_p, tt_y.long()) if target_train is not None else entropy(tt_p),
                      m_entropy(tte_p, tte_y.long()) if target_test is not None else entropy(tte_p)],
        "prob": [st_p, ste_p, tt_p, tte_p],
    }
    return {k: SVC_fit_predict(*v) for k, v in feats.items()}
