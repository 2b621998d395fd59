"""nsparse-b200: B200-native (sm_100a) re-implementation of the two nsparse hot paths --
row-binned hash SpGEMM and AMB SpMV -- behind nsparse's own entry points.

Python here is plumbing (device memory via torch tensors, ctypes calls into
lib/libnsparse_b200.so); the product is the CUDA library and its C ABI
(include/nsparse_b200.h, include/nsparse.h).
"""
from .context import Context, default_context          # noqa: F401
from .csr import CSR, DeviceCSR64                        # noqa: F401
from .spgemm import (get_spgemm_flop, spgemm_kernel_hash, spgemm_numeric,   # noqa: F401
                     spgemm_symbolic)
from .amb import AMB, Plan, csr2amb, spmv_amb             # noqa: F401
from .multi_gpu import (PeerBuffers, allgatherv_csr, partition_rows_by_cost, partition_rows_by_ip,   # noqa: F401
                        partition_rows_by_ip_device, partition_rows_by_measured, partition_rows_minmax, row_block,
                        spgemm_kernel_hash_mgpu)
from ._lib import NsparseError, load as load_library     # noqa: F401
