"""Ns=18 model, sector (9,ndw): DimUp=48620 (778 KB columns) x DimDw=C(18,ndw) on one GPU: column pass block-split vs generic."""
import sys, json, time, torch
sys.path.insert(0, '/root/repo')
from cdmft_lanc_ed_b200 import models
from cdmft_lanc_ed_b200 import ed_hamiltonian as E
mdl = models.hm_ns18()
E.ed_init(0); E.set_stream(torch.cuda.current_stream().cuda_stream); E.ed_set_model(mdl)
isec = models.get_sector(mdl.ns, 9, int(sys.argv[1]) if len(sys.argv) > 1 else 2)
res = {}
for var in (6, 1):
    E.set_option("colpass_variant", var)
    t0 = time.time(); n = E.build_Hv_sector(isec, True); tb = time.time() - t0
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    v = torch.view_as_complex(torch.randn(n, 2, dtype=torch.float64, device="cuda", generator=g)); hv = torch.empty_like(v)
    for _ in range(3): E.spHtimesV_p(n, v, hv)
    E.set_option("profile", 1); E.profile_query(0); E.profile_query(1)
    for _ in range(10): E.spHtimesV_p(n, v, hv)
    torch.cuda.synchronize()
    col = E.profile_query(0)[0] / 10; row = E.profile_query(1)[0] / 10
    E.set_option("profile", 0)
    res[var] = hv.clone()
    # real-mode Krylov
    E.set_option("profile", 1); E.profile_query(0); E.profile_query(1); E.profile_query(4)
    v0 = torch.ones(n, dtype=torch.complex128, device="cuda")
    E.sp_lanc_tridiag(v0, 10)
    lc = E.profile_query(0)[0] / 10
    E.set_option("profile", 0)
    print(json.dumps(dict(colpass_variant=var, n=n, build_s=round(tb, 3), col_ms=round(col, 3), row_ms=round(row, 3), lanc_real_col_ms=round(lc, 3))), flush=True)
    E.delete_Hv_sector()
print("relerr colblk vs generic", float((res[6] - res[1]).abs().max() / res[1].abs().max()))
E.ed_finalize()
