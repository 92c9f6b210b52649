import sys, torch
sys.path.insert(0,'.')
from phoneme_contrast_b200 import ops
B,H,W,Cout,k=256,40,101,64,7
g=ops.conv_geom(B,H,W,1,Cout,k,1,3)
x=torch.randn(B,H,W,1,device='cuda'); w=torch.randn(Cout,1,k,k,device='cuda')*0.1; bias=torch.randn(Cout,device='cuda')
stats=torch.zeros(2,Cout,device='cuda',dtype=torch.float64)
def t(f,n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
print("stem fwd SIMT us", t(lambda: ops.conv_fwd(x,w,bias,g,None,stats,0)))
y=ops.conv_fwd(x,w,bias,g,None,None,0)
ref=torch.nn.functional.conv2d(x.permute(0,3,1,2).double(),w.double(),bias.double(),padding=3).permute(0,2,3,1)
print("err", float((y.double()-ref).abs().max()/ref.abs().max()))
