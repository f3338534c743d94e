"""B200-native (sm_100a) implementation of the MiViT training hot path:
synthetic fluorescence image-sequence rendering + the motion-informed ViT regressor's
training step, behind the reference's own Python API.

    from moleculardiffusion_mivit_b200.helpersGeneration import *   # renderer
    from moleculardiffusion_mivit_b200.models import *              # ViT

All compute runs in hand-written CUDA kernels reached through the C ABI declared in
include/mivit.h; there is no CPU fallback."""
__version__ = "0.1.0"
