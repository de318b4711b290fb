"""cdmft_lanc_ed_b200 -- B200-native Hamiltonian-times-vector / Lanczos hot path of
QcmPlab/CDMFT-LANC-ED behind the reference's build_Hv_sector / spHtimesV_p /
delete_Hv_sector contract.  The compute lives in csrc/ (hand-written sm_100a CUDA
behind the C ABI of include/cdmft_b200.h); this package is the thin host mirror."""
from . import models  # noqa: F401
