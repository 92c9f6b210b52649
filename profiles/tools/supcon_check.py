import os, sys, torch, numpy as np
sys.path.insert(0,'.')
from phoneme_contrast_b200 import ops
def t(f,n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
for N in (1024, 4096, 8192):
    F=torch.nn.functional.normalize(torch.randn(N,128,device='cuda'),dim=1).contiguous(); y=torch.randint(0,38,(N,),device='cuda')
    for nrows in (N, N//8):
        os.environ["PC_SUPCON_TC"]="0"; a=t(lambda: ops.supcon_fwd(F,y,None,0.15,0.07,0,nrows))
        os.environ["PC_SUPCON_TC"]="1"; b=t(lambda: ops.supcon_fwd(F,y,None,0.15,0.07,0,nrows))
        print(f"supcon fwd N={N} rows={nrows}: SIMT {a:.0f} us  TC {b:.0f} us  ({2*nrows*N*128/b/1e6:.1f} TFLOP/s)")
        st,_=ops.supcon_fwd(F,y,None,0.15,0.07,0,nrows)
        coef=(0.15/0.07)/N
        sa = st if nrows==N else torch.cat([st, st.new_zeros(N-nrows,4)+1])
        os.environ["PC_SUPCON_TC"]="0"; c=t(lambda: ops.supcon_bwd(F,y,None,0.15,coef,None,sa,0,nrows))
        os.environ["PC_SUPCON_TC"]="1"; e=t(lambda: ops.supcon_bwd(F,y,None,0.15,coef,None,sa,0,nrows))
        print(f"   bwd SIMT {c:.0f} us  TC {e:.0f} us  ({4*nrows*N*128/e/1e6:.1f} TFLOP/s)")
