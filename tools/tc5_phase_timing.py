import torch,sys
sys.path.insert(0,".")
from wave_mamba_b200 import ops, _cabi
ops.set_conv_impl("tcgen05")
lib=_cabi.load()
x=torch.randn(1,64,1080,1920,device="cuda"); w3=torch.randn(64,64,3,3,device="cuda")*0.1; w1=torch.randn(64,64,1,1,device="cuda")*0.1; b=torch.zeros(64,device="cuda"); w4=torch.randn(32,64,3,3,device="cuda")*0.1
t=ops.conv3x3(x,w3,gate_w=w1,gate_b=b); y=ops.conv3x3(t,w4); torch.cuda.synchronize()
names=["waitX","xlo","chunks","drain","issue_next","epilogue"]
for label,fn in (("gate 64->64",lambda: ops.conv3x3(x,w3,gate_w=w1,gate_b=b)),("k4 64->32",lambda: ops.conv3x3(t,w4))):
    dbg=torch.zeros(148*6,dtype=torch.int64,device="cuda")
    lib.wm_conv3x3_debug_timing(dbg.data_ptr())
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True); s.record(); fn(); e.record(); torch.cuda.synchronize()
    lib.wm_conv3x3_debug_timing(None)
    d=dbg.view(148,6).double().mean(0)
    tiles=9300/148
    print(label, f"{s.elapsed_time(e):.3f} ms;", "per-tile cycles:", {n:int(v/tiles) for n,v in zip(names,d.tolist())}, "sum", int(d.sum()/tiles))
