from .quantization import (asym_dequant, asym_quant, asym_quant_dequant, get_minq_maxq, pack_i4,  # noqa: F401
                           sym_dequant, sym_quant, sym_quant_dequant, two_compl, unpack_i4)
from .online_trans import get_decompose_dim, kronecker_matmul, quant  # noqa: F401
from .hadamard import get_hadK, is_pow2, matmul_hadU_cuda  # noqa: F401
