"""gwfast_b200 -- B200-native Fisher/SNR engine behind the gwfast API.

Drop-in for ONE hot path of CosmoStatGW/gwfast: per-event waveform generation on the frequency grid, detector
projection and the PSD-weighted inner products that give the SNR and the Fisher matrix
(``waveforms.WaveFormModel`` subclasses, ``signal.GWSignal.SNRInteg/FisherMatr``, ``network.DetNet.SNR/FisherMatr``).
All numerical work runs in hand-written sm_100a CUDA kernels (``csrc/``) behind a C ABI (``include/gwfast_b200.h``);
there is no JAX, no Triton and no CPU fallback.
"""
from . import gwfastGlobals, gwfastUtils, waveforms, signal, network  # noqa: F401

__version__ = '0.1.0'
