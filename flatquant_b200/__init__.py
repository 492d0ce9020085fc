"""flatquant_b200 -- B200 (sm_100a) native implementation of FlatQuant's fused W4A4 FlatLinear hot path.

Host code keeps the reference's operator API (FlatQuantizedLinear, quant_utils, trans_utils,
deploy.nn.{OnlineTrans,Quantizer,Linear4bit}, deploy.{matmul,sym_quant,sym_dequant}); compute is
hand-written CUDA behind the C ABI of include/flatq_b200.h.  CUDA-only: no CPU fallback.
"""
__version__ = "0.1.0"
