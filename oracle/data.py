"""TEST INFRASTRUCTURE -- numpy restatement of the reference's input transform and per-batch evaluation arithmetic.

  augment      RandomCrop(32, padding=4) + RandomHorizontalFlip + ToTensor on explicit (left, top, flip) decisions
               (Classification/dataset.py:549-555; torchvision.transforms.functional.pad / crop / hflip / to_tensor)
  eval_logits  summed cross-entropy, top-1 hits and softmax of a batch of logits (trainer/val.py:44-61,
               evaluation/SVC_MIA.py:44-46)
tests/test_data_eval_cpu.py pins `augment` to the torchvision functional ops the reference's transforms call.
"""
import numpy as np


def augment(images_hwc_u8, index, crop_xy=None, flip=None, pad=4):
    imgs = np.asarray(images_hwc_u8)
    n, (H, W) = len(index), imgs.shape[1:3]
    out = np.zeros((n, 3, H, W), dtype=np.float32)
    for i, src in enumerate(index):
        padded = np.zeros((H + 2 * pad, W + 2 * pad, 3), dtype=np.uint8)
        padded[pad:pad + H, pad:pad + W] = imgs[int(src)]
        left, top = (pad, pad) if crop_xy is None else (int(crop_xy[i][0]), int(crop_xy[i][1]))
        c = padded[top:top + H, left:left + W]
        if flip is not None and flip[i]:
            c = c[:, ::-1]
        out[i] = (c.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)
    return out


def eval_logits(logits, labels):
    z = np.asarray(logits, dtype=np.float64)
    m = z.max(axis=1, keepdims=True)
    lse = np.log(np.exp(z - m).sum(axis=1)) + m[:, 0]
    ce = lse - z[np.arange(len(z)), labels]
    probs = np.exp(z - m) / np.exp(z - m).sum(axis=1, keepdims=True)
    return float(ce.sum()), int((z.argmax(axis=1) == labels).sum()), probs.astype(np.float32)
