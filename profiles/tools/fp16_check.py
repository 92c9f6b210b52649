import sys, torch, numpy as np
sys.path.insert(0,'.')
from phoneme_contrast_b200 import ops, _lib as L
torch.manual_seed(0)
def gemm(M,N,K):
    a=torch.randn(M,K,device='cuda'); b=torch.randn(N,K,device='cuda')*0.05; bias=torch.randn(N,device='cuda')
    ref=(a.double()@b.double().T+bias.double())
    for prec in (1,3):
        c=ops.tc_gemm(a,b,bias,prec)
        print(f"gemm {M}x{N}x{K} prec{prec}: err/max {float((c.double()-ref).abs().max()/ref.abs().max()):.2e}")
gemm(256,128,128); gemm(1000,64,576); gemm(4096,256,2304)
def conv(B,H,W,Cin,Cout,k,stride,pad):
    g=ops.conv_geom(B,H,W,Cin,Cout,k,stride,pad)
    x=torch.randn(B,H,W,Cin,device='cuda'); w=torch.randn(Cout,Cin,k,k,device='cuda')*0.05; bias=torch.randn(Cout,device='cuda')
    sc=torch.rand(Cin,device='cuda')+0.5; sh=torch.randn(Cin,device='cuda')*0.1
    xform=dict(scale=sc,shift=sh,relu=True)
    a=torch.relu(x*sc+sh)
    ref=torch.nn.functional.conv2d(a.permute(0,3,1,2).double(),w.double(),bias.double(),stride=stride,padding=pad).permute(0,2,3,1)
    dy=torch.randn(B,g.Ho,g.Wo,Cout,device='cuda')*1e-6
    dref=torch.nn.grad.conv2d_input((B,Cin,H,W),w.double(),dy.permute(0,3,1,2).double(),stride=stride,padding=pad).permute(0,2,3,1)
    amax=dy.abs().max().reshape(1)
    for prec in (1,3):
        cw=ops.ConvWeights(w,g,prec)
        stats=torch.zeros(2,Cout,device='cuda',dtype=torch.float64)
        for _ in range(3): y=ops.conv_fwd(x,cw.wf,bias,g,xform,stats,cw.prec_f)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): y=ops.conv_fwd(x,cw.wf,bias,g,xform,stats,cw.prec_f)
        e1.record(); torch.cuda.synchronize()
        tf=e0.elapsed_time(e1)/5
        err=float((y.double()-ref).abs().max()/ref.abs().max())
        for _ in range(2): dx=ops.conv_dgrad(dy,cw.wd,g,prec=cw.prec_d,dy_amax=amax if prec==3 else None)
        e0.record()
        for _ in range(5): dx=ops.conv_dgrad(dy,cw.wd,g,prec=cw.prec_d,dy_amax=amax if prec==3 else None)
        e1.record(); torch.cuda.synchronize()
        td=e0.elapsed_time(e1)/5
        derr=float((dx.double()-dref).abs().max()/dref.abs().max())
        fl=2*B*g.Ho*g.Wo*Cout*Cin*k*k/1e9
        print(f"conv B{B} {H}x{W} {Cin}->{Cout} k{k}s{stride} prec{prec}(f{cw.prec_f},d{cw.prec_d}): fwd {tf*1e3:.1f}us {fl/tf:.0f}TF/s err {err:.2e} | dgrad {td*1e3:.1f}us err {derr:.2e}")
conv(256,20,51,64,64,3,1,1)
conv(256,20,51,64,128,3,2,1)
conv(256,10,26,128,128,3,1,1)
conv(256,5,13,256,256,3,1,1)
conv(256,3,7,512,512,3,1,1)
conv(256,20,51,64,128,1,2,0)
