"""Isolated-process probes of the tcgen05 kernels (a device trap kills only the probe that caused it)."""
import subprocess
import sys

PROBES = {
    "fprop3x3_c64": "x=[rnd(2,64,64,64)]; w=rnd(64,64,3,3,scale=.05); pw=mk(w); y=ops.conv_fprop([nhwc(x[0])],pw); print('rel',rel(nchw(y),F.conv2d(x[0],w,padding=1)))",
    "fprop3x3_c16": "x=[rnd(2,16,64,64)]; w=rnd(16,16,3,3,scale=.1); pw=mk(w); y=ops.conv_fprop([nhwc(x[0])],pw); print('rel',rel(nchw(y),F.conv2d(x[0],w,padding=1)))",
    "fprop3x3_c32": "x=[rnd(2,32,64,64)]; w=rnd(32,32,3,3,scale=.1); pw=mk(w); y=ops.conv_fprop([nhwc(x[0])],pw); print('rel',rel(nchw(y),F.conv2d(x[0],w,padding=1)))",
    "fprop1x1_c128": "x=[rnd(2,128,16,16)]; w=rnd(256,128,1,1,scale=.1); pw=mk(w); y=ops.conv_fprop([nhwc(x[0])],pw); print('rel',rel(nchw(y),F.conv2d(x[0],w)))",
    "dgrad3x3_c64": "dy=rnd(2,64,64,64); w=rnd(64,64,3,3,scale=.05); pw=mk(w); dx=ops.conv_dgrad(nhwc(dy),pw)[0]; print('rel',rel(nchw(dx),torch.nn.grad.conv2d_input((2,64,64,64),w,dy,padding=1)))",
    "wgrad3x3_c128": "x=rnd(2,128,32,32); dy=rnd(2,128,32,32); w=rnd(128,128,3,3); pw=mk(w); dw=ops.conv_wgrad([nhwc(x)],nhwc(dy),pw); print('rel',rel(dw,torch.nn.grad.conv2d_weight(x,w.shape,dy,padding=1)))",
    "wgrad3x3_c64": "x=rnd(2,64,64,64); dy=rnd(2,64,64,64); w=rnd(64,64,3,3); pw=mk(w); dw=ops.conv_wgrad([nhwc(x)],nhwc(dy),pw); print('rel',rel(dw,torch.nn.grad.conv2d_weight(x,w.shape,dy,padding=1)))",
    "wgrad3x3_c16": "x=rnd(2,16,64,64); dy=rnd(2,16,64,64); w=rnd(16,16,3,3); pw=mk(w); dw=ops.conv_wgrad([nhwc(x)],nhwc(dy),pw); print('rel',rel(dw,torch.nn.grad.conv2d_weight(x,w.shape,dy,padding=1)))",
    "wgrad3x3_c32_64": "x=rnd(2,32,64,64); dy=rnd(2,64,64,64); w=rnd(64,32,3,3); pw=mk(w); dw=ops.conv_wgrad([nhwc(x)],nhwc(dy),pw); print('rel',rel(dw,torch.nn.grad.conv2d_weight(x,w.shape,dy,padding=1)))",
    "convt_fwd": "x=rnd(2,32,32,32); w=rnd(32,16,2,2,scale=.1); pw=mk(w,True); y=ops.convt_fprop(nhwc(x),pw); print('rel',rel(nchw(y),F.conv_transpose2d(x,w,stride=2)))",
    "convt_dgrad": "dy=rnd(2,16,64,64); w=rnd(32,16,2,2,scale=.1); pw=mk(w,True); dx=ops.convt_dgrad(nhwc(dy),pw); print('rel',rel(nchw(dx),F.conv2d(dy,w,stride=2)))",
    "convt_wgrad": "x=rnd(2,32,32,32); dy=rnd(2,16,64,64); w=rnd(32,16,2,2); pw=mk(w,True); dw=ops.convt_wgrad(nhwc(x),nhwc(dy),pw); xr=x.clone().requires_grad_(True); wr=w.clone().requires_grad_(True); F.conv_transpose2d(xr,wr,stride=2).backward(dy); print('rel',rel(dw,wr.grad))",
}
PRE = """
import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.')
import __graft_entry__ as g; g.load_package()
from smsut_b200 import ops
torch.manual_seed(0)
bf=lambda t: t.to(torch.bfloat16)
rnd=lambda *s, scale=1.0: bf(torch.randn(*s, device='cuda')*scale).float()
nhwc=lambda t: bf(t).permute(0,2,3,1).contiguous()
nchw=lambda t: t.float().permute(0,3,1,2).contiguous()
rel=lambda a,b: ((a.float()-b.float()).norm()/(b.float().norm()+1e-12)).item()
def mk(w, tr=False):
    pw=ops.PackedWeight(w, transposed=tr); ops.PackTable([pw]).refresh(); return pw
"""
names = sys.argv[1:] or list(PROBES)
for n in names:
    r = subprocess.run([sys.executable, "-c", PRE + PROBES[n] + "\ntorch.cuda.synchronize()"], capture_output=True,
                       text=True, timeout=300)
    tail = (r.stdout.strip().splitlines() or [""])[-1]
    err = "" if r.returncode == 0 else " | " + " ".join(r.stderr.strip().splitlines()[-3:])[:400]
    print(f"[probe] {n}: exit {r.returncode} {tail}{err}", flush=True)
