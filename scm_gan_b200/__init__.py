"""scm_gan_b200: B200-native (sm_100a) implementation of the scm-gan world-model training step.

Layout: csrc/ (CUDA kernels + C ABI -> lib/libscmgan.so), _lib.py (ctypes binding), kernels.py (tensor-level
wrappers), ops.py (torch.library custom ops with autograd), dropin/ (models.py & friends mirroring the reference
module interface), train_step.py (host mirror of reference main.py:143-296), dp.py (data-parallel gradient exchange).
"""
__version__ = "0.1.0"
