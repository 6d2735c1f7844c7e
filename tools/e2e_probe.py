import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from figdraw_b200 import scenes_synth as ss
from figdraw_b200.cuda_context import CudaContext
tr = ss.config_trace(5)
W, H = tr.width, tr.height
dev = torch.device("cuda", 0)
# raw PCIe
a = torch.empty(31_363_456, dtype=torch.uint8).pin_memory(); b = torch.empty_like(a, device=dev)
c = torch.empty(33_177_600, dtype=torch.uint8, device=dev); d = torch.empty(33_177_600, dtype=torch.uint8).pin_memory()
def timeit(fn, n=20):
    fn(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
print("H2D 31MB ms", timeit(lambda: b.copy_(a, non_blocking=True)))
print("D2H 33MB ms", timeit(lambda: d.copy_(c, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): b.copy_(a, non_blocking=True)
    with torch.cuda.stream(s2): d.copy_(c, non_blocking=True)
print("both concurrently ms", timeit(both))
def mk():
    ctx = CudaContext(atlasSize=tr.atlas_size)
    for _i, k, img in tr.images: ctx.putImage(k, img)
    calls = torch.from_numpy(tr.calls.view(np.uint8).reshape(-1,128).copy()).pin_memory()
    out = torch.empty((H,W,4), dtype=torch.uint8).pin_memory()
    return ctx, calls.numpy().view(tr.calls.dtype).reshape(-1), out.numpy(), calls, out
A, B = mk(), mk()
def submit(x):
    t=time.perf_counter(); x[0].beginFrame((W,H), clearMain=True); t1=time.perf_counter(); x[0].submitCalls(x[1]); t2=time.perf_counter(); x[0].endFrame(); t3=time.perf_counter()
    return (t1-t)*1e3, (t2-t1)*1e3, (t3-t2)*1e3
def read(x):
    t=time.perf_counter(); x[0].readPixels((0,0,W,H), out=x[2]); return (time.perf_counter()-t)*1e3
for _ in range(3): submit(A); read(A); submit(B); read(B)
print("single: begin/submit/end", submit(A), "read", read(A))
print("single: begin/submit/end", submit(A), "read", read(A))
pair=[A,B]
submit(pair[0]); rows=[]
t0=time.perf_counter()
for k in range(12):
    cur, nxt = pair[k&1], pair[(k+1)&1]
    s = submit(nxt); r = read(cur); rows.append((s, r))
print("pipelined per step ms", (time.perf_counter()-t0)/12*1e3)
for s, r in rows[-4:]: print("  submit b/s/e", [round(v,3) for v in s], "read", round(r,3))
