"""DDPM side of SalUn.

  engine.py   UNetEngine: host handle of the sm_100a U-Net engine (salun_unet_* in libsalun.so): forward / backward of
              Conditional_Model on flat fp32 arenas, tcgen05 convolutions / attention GEMMs, GroupNorm + swish kernels
  runner.py   Diffusion (mirror of the reference runner's generate_mask / saliency_unlearn), DDPMEngineUnlearner (the loop
              bodies on the engine + fused clip / mask / Adam / accumulate / top-k tail), DDPMUnlearner (same loop around
              a torch.nn.Module for architectures the engine does not serve), beta schedule, q-sample, eps loss
  unet.py     ConditionalUNet: PyTorch restatement of the architecture (parameter names / order of the reference);
              shape descriptor for the engine, model of the DDPMUnlearner path, and the checker's network (oracle/ddpm.py)
"""
