"""kaldi_fp16_b200 -- B200-native (sm_100a) replacement for the FP16 GEMM hot path of
djeday123/kaldi-fp16: libkaldi_fp16.so (C ABI, include/*.h) + a host mirror of the reference's
Go operator packages.  There is no CPU / PyTorch fallback: importing the operator modules fails
if the CUDA library has not been built (python -m kaldi_fp16_b200.build)."""
__version__ = "0.1.0"
