import sys, torch
sys.path.insert(0,'.')
from phoneme_contrast_b200 import ops
def run(B,H,W,Cin,Cout,k,stride,pad):
    g=ops.conv_geom(B,H,W,Cin,Cout,k,stride,pad)
    x=torch.randn(B,H,W,Cin,device='cuda'); dy=torch.randn(B,g.Ho,g.Wo,Cout,device='cuda')*1e-6
    sc=torch.rand(Cin,device='cuda')+0.5; sh=torch.randn(Cin,device='cuda')*0.1
    xf=dict(scale=sc,shift=sh,relu=True)
    amax=dy.abs().max().reshape(1)
    a=torch.relu(x.double()*sc.double()+sh.double()).permute(0,3,1,2)
    wr=torch.zeros(Cout,Cin,k,k,device='cuda',dtype=torch.float64,requires_grad=True)
    torch.nn.functional.conv2d(a,wr,None,stride=stride,padding=pad).backward(dy.permute(0,3,1,2).double())
    for prec in (1,3):
        for _ in range(2): dw,db=ops.conv_wgrad(x,dy,g,xf,prec=prec,dy_amax=amax if prec==3 else None)
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): dw,db=ops.conv_wgrad(x,dy,g,xf,prec=prec,dy_amax=amax if prec==3 else None)
        e1.record(); torch.cuda.synchronize()
        t=e0.elapsed_time(e1)/5
        err=float((dw.double()-wr.grad).abs().max()/wr.grad.abs().max())
        fl=2*B*g.Ho*g.Wo*Cout*Cin*k*k/1e9
        print(f"wgrad B{B} {H}x{W} {Cin}->{Cout} k{k}s{stride} prec{prec}: {t*1e3:.1f}us {fl/t:.0f}TF/s err {err:.2e}")
run(256,20,51,64,64,3,1,1)
run(256,20,51,64,128,3,2,1)
run(256,10,26,128,128,3,1,1)
run(256,5,13,256,256,3,1,1)
run(256,3,7,512,512,3,1,1)
run(256,20,51,64,128,1,2,0)
