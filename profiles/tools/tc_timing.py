import ctypes, sys, torch, numpy as np
sys.path.insert(0,'.')
from phoneme_contrast_b200 import ops, _lib as L
lib=L.lib()
lib.pc_tc_set_debug.argtypes=[ctypes.c_void_p]; lib.pc_tc_set_debug.restype=None
def run(B,H,W,Cin,Cout,k,stride,pad,prec=1,xf=True,presplit=False):
    g=ops.conv_geom(B,H,W,Cin,Cout,k,stride,pad)
    x=torch.randn(B,H,W,Cin,device='cuda'); w=torch.randn(Cout,Cin,k,k,device='cuda')*0.05; bias=torch.zeros(Cout,device='cuda')
    sc=torch.ones(Cin,device='cuda'); sh=torch.zeros(Cin,device='cuda')
    cw=ops.ConvWeights(w,g,prec)
    stats=torch.zeros(2,Cout,device='cuda',dtype=torch.float64)
    M=B*g.Ho*g.Wo; ntiles=((M+127)//128)*max(1,(Cout+127)//128 if Cout>64 else 1)
    dbg=torch.zeros(ntiles*4,16,device='cuda',dtype=torch.int64)
    xform=dict(scale=sc,shift=sh,relu=True) if xf else None
    if presplit:
        x=ops.bn_act_split(x,sc,sh,None,relu=True); xform=dict(presplit=True)
    for it in range(3):
        y=ops.conv_fwd(x,cw.wf,bias,g,xform,stats,cw.prec_f)
    torch.cuda.synchronize()
    lib.pc_tc_set_debug(ctypes.c_void_p(dbg.data_ptr()))
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); y=ops.conv_fwd(x,cw.wf,bias,g,xform,stats,cw.prec_f); e1.record(); torch.cuda.synchronize()
    lib.pc_tc_set_debug(None)
    d=dbg.cpu().numpy()[:((M+127)//128)]
    d=d[d[:,0]>0]
    t0=d[:,0:1]
    rel=(d-t0)
    names=['start','setup_done','prod_loop_start','prod_kc0_arrive','prod_kc3_arrive','prod_loop_end','acc_full_seen','epi_end','mma_kc0_full','mma_kc1_full','mma_kc4_full','mma_last_commit','exit']
    print(f"conv B{B} {H}x{W} {Cin}->{Cout} k{k} s{stride}: {e0.elapsed_time(e1)*1e3:.1f} us, tiles {len(d)}, flops {2*M*Cout*Cin*k*k/1e9:.2f} G -> {2*M*Cout*Cin*k*k/ (e0.elapsed_time(e1)*1e-3)/1e12:.1f} TF/s")
    med=np.median(rel,axis=0)
    for i,n in enumerate(names): print(f"   {n:18s} median {med[i]:9.0f} cyc")
    span=(d[:,12].max()-d[:,0].min())
    print("   whole-kernel span cycles", span, " CTAs/SM waves ~", len(d)/148)
print('=== presplit fp16x2')
run(256,20,51,64,64,3,1,1,prec=3,presplit=True)
run(256,5,13,256,256,3,1,1,prec=3,presplit=True)
run(256,3,7,512,512,3,1,1,prec=3,presplit=True)
