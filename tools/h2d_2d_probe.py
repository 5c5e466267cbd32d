"""probe: pinned-host -> device copy of only the ACTIVE arm's 6 joints of an array-of-State batch (cudaMemcpy2DAsync, pitch 96 B, width 48 B)
against the contiguous 96 B/sample upload -- decides whether the PCS compact entry point should upload half rows"""
import ctypes, time, sys
import torch
rt = ctypes.CDLL("libcudart.so.12")
n = 1_000_000
h = torch.empty((n, 12), dtype=torch.float64).pin_memory(); h.uniform_(-3, 3)
d = torch.empty((n, 12), dtype=torch.float64, device="cuda")
d6 = torch.empty((n, 6), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
H2D = 1
def t(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def full(): rt.cudaMemcpyAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(n * 96), H2D, ctypes.c_void_p(st))
def half2d(): rt.cudaMemcpy2DAsync(ctypes.c_void_p(d6.data_ptr()), ctypes.c_size_t(48), ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(96), ctypes.c_size_t(48), ctypes.c_size_t(n), H2D, ctypes.c_void_p(st))
def half2d_strided_dst(): rt.cudaMemcpy2DAsync(ctypes.c_void_p(d.data_ptr()), ctypes.c_size_t(96), ctypes.c_void_p(h.data_ptr()), ctypes.c_size_t(96), ctypes.c_size_t(48), ctypes.c_size_t(n), H2D, ctypes.c_void_p(st))
a = t(full); b = t(half2d); c = t(half2d_strided_dst)
print("contiguous 96 B x 1e6: %.3f ms (%.1f GB/s) | 2D 48 of 96 B -> packed: %.3f ms | 2D 48 of 96 B -> pitched: %.3f ms" % (1e3 * a, 0.096 / a, 1e3 * b, 1e3 * c))
half2d(); torch.cuda.synchronize()
print("2D copy correct:", bool((d6.cpu() == h[:, :6]).all()))
