"""Scratch: does any block of another kernel become resident next to a persistent Winograd-GEMM CTA
(10 warps x 168 registers, 193 KB of shared memory per SM)?  A spin kernel of a given block size and register
footprint is launched on a second (higher-priority) stream while the full-size GEMM runs; every spin block records
its SM and its start / end in globaltimer ns; stamp kernels bracket the GEMM."""
import ctypes as C, os, subprocess, sys
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from scanpaths_b200 import _lib
so = os.path.join(HERE, "corun_probe.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "corun_probe.cu")])
P = C.CDLL(so)
P.launch_stamp.argtypes = [C.c_void_p, C.c_void_p]
P.launch_spin.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
lib = _lib.load()
dev = torch.device("cuda")
rows_pad, cols = 38400, 2048
u_hi = torch.zeros((24, rows_pad, 512), dtype=torch.float16, device=dev); u_lo = torch.zeros_like(u_hi)
u_hi.normal_(); u_lo.normal_()
w_hi = torch.randn((24 * cols, 512), device=dev).half(); w_lo = torch.randn((24 * cols, 512), device=dev).half()
out = torch.empty((12, cols // 128, rows_pad, 128), dtype=torch.float32, device=dev)
main = torch.cuda.Stream(device=dev)
side = torch.cuda.Stream(device=dev, priority=-1)
stamps = torch.zeros(2, dtype=torch.int64, device=dev)
sink = torch.zeros(1024, dtype=torch.float32, device=dev)


def gemm():
    _lib.check(lib.spb_wino_gemm(_lib.ptr(u_hi), _lib.ptr(u_lo), _lib.ptr(w_hi), _lib.ptr(w_lo), _lib.ptr(out), rows_pad,
                                 cols, 1.0, 0, C.c_void_p(main.cuda_stream)), "gemm")


for _ in range(2):
    gemm()
torch.cuda.synchronize()
print("block shape | regs/thread | blocks launched | blocks that started inside the GEMM | on how many SMs | spin blocks finished inside")
for threads, nr in ((32, 16), (32, 64), (32, 128), (64, 16), (64, 64), (128, 16), (128, 64), (256, 16)):
    blocks = 148 * 4
    rec = torch.zeros((blocks, 3), dtype=torch.int64, device=dev)
    P.launch_stamp(C.c_void_p(stamps.data_ptr()), C.c_void_p(main.cuda_stream))
    gemm()
    P.launch_stamp(C.c_void_p(stamps.data_ptr() + 8), C.c_void_p(main.cuda_stream))
    # the spin kernel is enqueued right behind: by the time it is dispatched the GEMM's CTAs are resident
    import time; time.sleep(0.0005)
    P.launch_spin(nr, blocks, threads, C.c_void_p(rec.data_ptr()), C.c_void_p(sink.data_ptr()), 100000, C.c_void_p(side.cuda_stream))
    torch.cuda.synchronize()
    g0, g1 = [int(x) for x in stamps.cpu()]
    r = rec.cpu().numpy()
    inside = (r[:, 1] > g0 + 200000) & (r[:, 1] < g1 - 200000)
    done_inside = inside & (r[:, 2] < g1)
    print("%4d threads | %3d | %4d | %4d | %3d | %4d   (GEMM %.2f ms)" % (threads, P.spin_regs(nr), blocks, int(inside.sum()),
          len(set(r[inside, 0].tolist())), int(done_inside.sum()), (g1 - g0) / 1e6), flush=True)
