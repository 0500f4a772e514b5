import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
from esrganplus_b200 import _lib, conv as K
lib = _lib.load()
dev = 'cuda'
n, h, w = 32, 9, 9
x = torch.randn(n, h, w, 512, device=dev).to(torch.bfloat16)
wt = torch.randn(64, 512, 3, 3, device=dev) * 0.02
wp = K.pack_conv3x3_weights(wt, 64, 64, [64 * i for i in range(8)], rows=64)
out = torch.zeros(n, h, w, 64, device=dev)
call = K.ConvCall(n=n, h=h, w=w, srcs=[x], kc=64, chunks=[(0, 64 * i) for i in range(8)], bn=64, cout=64, w_packed=wp, out_f32=out)
d = call.desc()
st = torch.cuda.current_stream().cuda_stream
for _ in range(10): lib.esrp_conv3x3_nhwc(C.byref(d), st)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): lib.esrp_conv3x3_nhwc(C.byref(d), st)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host per call us", (t1 - t0) / 2000 * 1e6, "incl device drain", (t2 - t0) / 2000 * 1e6)
# descriptor patch cost
t0 = time.perf_counter()
for _ in range(2000):
    d.src[0] = x.data_ptr(); d.out_f32 = out.data_ptr()
t1 = time.perf_counter()
print("patch us", (t1 - t0) / 2000 * 1e6)
