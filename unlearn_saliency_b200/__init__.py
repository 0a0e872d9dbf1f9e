"""unlearn_saliency_b200 -- B200-native (sm_100a) engine for the SalUn hot path.

Drop-in for the two data-parallel hot paths of OPTML-Group/Unlearn-Saliency:
  (i)  weight-saliency mask generation   (Classification/generate_mask.py:14-82 and siblings)
  (ii) the masked unlearning step        (Classification/unlearn/RL.py:123-159 and siblings)
All arithmetic on those paths runs in libsalun.so (hand-written CUDA, C ABI in include/salun.h);
this package is the host-side mirror of the reference's Python interface.
"""
__version__ = "0.1.0"
