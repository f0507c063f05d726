"""Is a torch-pinned buffer seen as pinned by cudaMemcpyAsync when passed as a numpy pointer?"""
import ctypes as C, time, torch, numpy as np
rt = C.CDLL("libcudart.so.12")
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
hn = h.numpy()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
class Attr(C.Structure):
    _fields_ = [("type", C.c_int), ("device", C.c_int), ("devicePointer", C.c_void_p), ("hostPointer", C.c_void_p)]
a = Attr()
print("attr rc", rt.cudaPointerGetAttributes(C.byref(a), C.c_void_p(hn.ctypes.data)), "type", a.type)
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = rt.cudaMemcpyAsync(d.data_ptr(), hn.ctypes.data, n, 1, s.cuda_stream)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(rc, "call %.3f ms, total %.3f ms, %.1f GB/s" % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, n / (t2 - t0) / 1e9))
