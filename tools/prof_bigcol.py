import sys, torch
sys.path.insert(0, '/root/repo')
from cdmft_lanc_ed_b200 import models
from cdmft_lanc_ed_b200 import ed_hamiltonian as E
mdl = models.hm_ns18()
E.ed_init(0); E.set_stream(torch.cuda.current_stream().cuda_stream); E.ed_set_model(mdl)
n = E.build_Hv_sector(models.get_sector(mdl.ns, 9, 4), True)
v = torch.randn(n, dtype=torch.complex128, device="cuda"); hv = torch.empty_like(v)
for _ in range(3): E.spHtimesV_p(n, v, hv)
torch.cuda.synchronize(); print("done")
E.delete_Hv_sector(); E.ed_finalize()
