"""DDPM side of SalUn: the class-conditional U-Net, the eps-prediction loss and the
generate_mask / saliency_unlearn loop bodies of DDPM/runners/diffusion.py, with the HBM-bound tail (clip, mask (.) grad,
Adam, saliency accumulate, top-k) on the sm_100a kernels.  The U-Net forward/backward itself still runs through
PyTorch (cuDNN/cuBLAS) this round; DESIGN.md section 8 lists it as the next tensor-core port."""
