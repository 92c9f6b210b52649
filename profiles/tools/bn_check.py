import sys, torch
sys.path.insert(0,'.')
from phoneme_contrast_b200 import ops, _lib as L
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
class BN: pass
def run(B,H,W,C,pool):
    y=torch.randn(B,H,W,C,device='cuda')
    bn=torch.nn.BatchNorm2d(C).cuda()
    st=torch.zeros(2,C,device='cuda',dtype=torch.float64)
    st[0]=y.double().sum((0,1,2)); st[1]=(y.double()**2).sum((0,1,2))
    co=ops.bn_finalize(st,B*H*W,bn,True)
    out,argmax=ops.bn_act_fwd(y,co,pool,None)
    dout=torch.randn_like(out)
    t_f=timeit(lambda: ops.bn_act_fwd(y,co,pool,None))
    # time reduce+apply together and separately via profile hooks
    t_b=timeit(lambda: ops.bn_act_bwd(dout,y,co,pool,None,argmax))
    mb=y.numel()*4/1e6
    print(f"bn_act B{B} {H}x{W} C{C} pool{pool}: y {mb:.0f} MB  fwd {t_f:.0f}us  bwd(reduce+apply+zeros) {t_b:.0f}us")
run(256,40,101,64,3)
run(256,20,51,64,0)
run(256,10,26,128,0)
run(256,5,13,256,0)
run(256,3,7,512,0)
