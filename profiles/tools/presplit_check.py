import sys, torch
sys.path.insert(0,'.')
from phoneme_contrast_b200 import ops
def t(f,n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n*1e3
def run(B,H,W,Cin,Cout,k,stride,pad):
    g=ops.conv_geom(B,H,W,Cin,Cout,k,stride,pad)
    x=torch.randn(B,H,W,Cin,device='cuda'); w=torch.randn(Cout,Cin,k,k,device='cuda')*0.05; bias=torch.randn(Cout,device='cuda')
    sc=torch.rand(Cin,device='cuda')+0.5; sh=torch.randn(Cin,device='cuda')*0.1
    drop=((torch.rand(B,Cin,device='cuda')>0.2).float()/0.8).contiguous()
    cw=ops.ConvWeights(w,g,3); st=torch.zeros(2,Cout,device='cuda',dtype=torch.float64)
    xf=dict(scale=sc,shift=sh,relu=True,drop=drop)
    a=t(lambda: ops.conv_fwd(x,cw.wf,bias,g,xf,st,cw.prec_f))
    b=t(lambda: ops.conv_fwd(x,cw.wf,bias,g,None,st,cw.prec_f))
    planes=ops.bn_act_split(x,sc,sh,drop,relu=True)
    c=t(lambda: ops.conv_fwd(planes,cw.wf,bias,g,dict(presplit=True),st,cw.prec_f))
    d=t(lambda: ops.bn_act_split(x,sc,sh,drop,relu=True))
    print(f"conv B{B} {H}x{W} {Cin}->{Cout} k{k}s{stride}: xform gather {a:.0f}us | raw gather {b:.0f}us | presplit {c:.0f}us (+ split pass {d:.0f}us)")
run(256,20,51,64,64,3,1,1)
run(256,10,26,128,128,3,1,1)
run(256,5,13,256,256,3,1,1)
run(256,3,7,512,512,3,1,1)
run(256,20,51,64,128,3,2,1)
run(256,20,51,64,128,1,2,0)
