"""clsurvey_b200 -- B200-native engine for the data-parallel hot path of Mattdl/CLsurvey.

Host side (model definition, task scheduling, the Method / train_model plugin surface) is Python/PyTorch like the
reference; the hot path (conv/linear fwd+bwd, pooling, loss head, penalised SGD / SI step, Fisher / MAS accumulators,
GEM dots/Gram/QP/projection) is hand-written CUDA for sm_100a behind the C ABI of include/clb.h, loaded with ctypes.
"""
__version__ = "0.1.0"
