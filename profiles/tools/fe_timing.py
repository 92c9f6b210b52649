import ctypes, sys, torch, numpy as np
sys.path.insert(0,'.')
from phoneme_contrast_b200 import _lib as L
from phoneme_contrast_b200.datasets import MFCCExtractor, build_augmentation_pipeline, build_view_descriptors, pack_view_descs
import bench
lib=L.lib(); lib.pc_fe_set_debug.argtypes=[ctypes.c_void_p]; lib.pc_fe_set_debug.restype=None
n=16384
ext=MFCCExtractor(); pipe=build_augmentation_pipeline(bench.AUG_CFG)
recs,_=build_view_descriptors(range(512),2,40,101,pipe); recs=np.tile(recs,n//512); recs["clip"]=np.repeat(np.arange(n),2)
views=pack_view_descs(recs,'cuda'); wave=bench.frontend_inputs(n,'cuda')
for _ in range(3): out=ext.forward_views(wave,views,2*n)
torch.cuda.synchronize()
dbg=torch.zeros(n,8,device='cuda',dtype=torch.int64)
lib.pc_fe_set_debug(ctypes.c_void_p(dbg.data_ptr()))
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record(); out=ext.forward_views(wave,views,2*n); e1.record(); torch.cuda.synchronize()
lib.pc_fe_set_debug(None)
d=dbg.cpu().numpy().astype(np.float64)
names=['load+window','radix8','radix5 x2','post(power)','mel+log','max','view epilogues(DCT+aug)']
tot=d[:,:7].sum(1)
print(f"{n} clips: {e0.elapsed_time(e1):.2f} ms -> {n/e0.elapsed_time(e1)*1e3/1e6:.2f} M clips/s; per-CTA total cycles median {np.median(tot):.0f}")
for i,nm in enumerate(names): print(f"  {nm:28s} {np.median(d[:,i]):9.0f} cyc  {100*np.median(d[:,i])/np.median(tot):5.1f}%")
