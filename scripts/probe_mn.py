import sys, torch
sys.path.insert(0, ".")
from view_fusion_b200 import _lib
lib = _lib.require_device()
torch.manual_seed(0)
A = torch.randn(64, 128).to(torch.bfloat16)
B = torch.randn(192, 64).to(torch.bfloat16)
Ad, Bd = A.cuda(), B.cuda()
for lbo, sbo in ((8192, 1024), (1024, 8192), (16, 1024), (8192, 2048)):
    for shift in (0, 8, 1, 3, 17):
        out = torch.zeros(128, 64, device="cuda")
        try:
            _lib.check(_lib.load_probes().vf_debug_umma_mn(Ad.data_ptr(), 64, Bd.data_ptr(), 192, shift, lbo, sbo, out.data_ptr(), _lib.stream_handle()), "mn")
            torch.cuda.synchronize()
        except Exception as e:
            print("ERR", lbo, sbo, shift, str(e)[:80]); break
        ref = A.float().t() @ B.float()[shift:shift + 64]
        err = float((out.cpu() - ref).norm() / ref.norm())
        print(f"lbo {lbo:5d} sbo {sbo:5d} shift {shift:2d}: rel err {err:.4f}", flush=True)
