"""qpnet_b200 -- B200-native (sm_100a) replacement for the hot path of bigpon/QPNet.

``from qpnet_b200.qpnet import QPNet, initialize, encode_mu_law, decode_mu_law`` mirrors
``from qpnet import ...`` of the reference (``src/bin/qpnet_train.py:35-37``).
The CUDA library (``qpnet_b200/csrc`` -> ``libqpnet_b200.so``) is loaded lazily by
``qpnet_b200._lib``; there is no CPU fallback.
"""
__version__ = "0.1.0"
