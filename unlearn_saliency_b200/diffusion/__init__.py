"""DDPM side of SalUn.

  engine.py   UNetEngine: host handle of the sm_100a U-Net engine (salun_unet_* in libsalun.so): forward / backward of
              Conditional_Model on flat fp32 arenas, tcgen05 convolutions / attention GEMMs, GroupNorm + swish kernels
  runner.py   Diffusion (mirror of the reference runner's generate_mask / saliency_unlearn), DDPMEngineUnlearner (the loop
              bodies on the engine + fused clip / mask / Adam / accumulate / top-k tail), beta schedule, q-sample, eps loss
  config.py   cifar10_config and the named_parameters() table of Conditional_Model (pure Python)
The PyTorch restatement of the network (the checker's model) is test infrastructure: oracle/unet.py.
"""
