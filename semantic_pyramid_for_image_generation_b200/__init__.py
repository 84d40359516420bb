"""B200-native (sm_100a) implementation of the Semantic-Pyramid GAN training step.

Public surface mirrors the reference repository's modules:
    models          Generator, Discriminator, VGG16 (+ the block classes)
    lossfunction    SemanticReconstructionLoss, DiversityLoss, LSGANGeneratorLoss, LSGANDiscriminatorLoss
    model_wrapper   ModelWrapper (train / validate / inference entry points)
    misc            mask builders, Logger
    optim           FusedAdam (torch.optim.Adam semantics on one fused kernel)
    distributed     one-process-per-GPU gradient all-reduce (replaces nn.DataParallel)
All arithmetic runs in libspyramid_b200.so (include/spyramid_b200.h); importing this package does not need a GPU,
using it does.
"""
from . import _native  # noqa: F401

__all__ = ["models", "lossfunction", "model_wrapper", "misc", "optim", "distributed"]
