import torch, sys
sys.path.insert(0, '/root/repo')
from partdistillation_b200 import functional as fn
g = torch.Generator().manual_seed(1)
for M, N, K in [(128,128,64),(128,128,256),(128,128,320),(128,128,384),(128,128,448),(128,128,512),(128,128,2048),(256,128,256),(512,128,256),(128*160,128,64),(128*160,128,256),(128*300,256,128)]:
    a = torch.randn(M, K, generator=g).cuda().to(torch.bfloat16); b = torch.randn(N, K, generator=g).cuda().to(torch.bfloat16)
    ref = a.double() @ b.double().t()
    out = fn.gemm_bf16(a, b, None, 0, torch.float32)
    d = (out.double() - ref).abs()
    print(M, N, K, "rel", float(d.max() / ref.abs().max()), "bad rows", int((d.max(1)[0] > 1e-2).sum()), "bad cols", int((d.max(0)[0] > 1e-2).sum()))
