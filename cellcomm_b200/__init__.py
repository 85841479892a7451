"""cellcomm_b200 — B200-native implementation of mikemey/cellcomm's BiGAN train + encode hot path.

Python surface mirrors the reference's `src/` modules (bigan_basic, bigan_classify, bigan_cont,
cell_type_training, intercepts); the arithmetic runs in libcellcomm_b200.so (hand-written
sm_100a CUDA behind a C ABI, see include/cellcomm_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1.0"
