"""csl_gan_b200: B200-native DP discriminator-update hot path for twosixlabs/csl-gan.

Public surface (mirrors what reference train.py imports from the opacus fork, train.py:13-14):
    PrivacyEngine, ISPrivacyEngine, calc_sample_norms,
    calc_penalty / calc_WGAN_GP_penalty / calc_lipschitz_penalty_WRT, l2_clip, row_l2_norm
The arithmetic lives in libcslgan_b200.so (hand-written sm_100a CUDA, include/cslgan_b200.h);
there is no CPU or eager-PyTorch fallback.
"""
from ._lib import CslGanCudaError, LIB_PATH  # noqa: F401
from .privacy_engine import PrivacyEngine, calc_sample_norms, GradSampleView  # noqa: F401
from .is_engine import ISPrivacyEngine  # noqa: F401
from .functional import (calc_penalty, calc_WGAN_GP_penalty, calc_lipschitz_penalty_WRT, l2_clip,  # noqa: F401
                         row_l2_norm, vec_max)
from .backprop_clip import BackpropClipper  # noqa: F401
from .mean_sampler import DeviceMeanSampler  # noqa: F401
from . import discriminators, accountant  # noqa: F401

__version__ = "0.1.0"
