"""dense attention kernel time (works with any build of the library: only lcasr_attention is called through ctypes)"""
import ctypes, os, sys
import torch
lib = ctypes.CDLL(os.environ["LCASR_LIB_PATH"])
vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
lib.lcasr_attention.argtypes = [vp, vp, vp, i32, i32, i64, i32, i32, i32, i64, vp, i32, vp]
dev = torch.device("cuda", 0)
for (B, N, H, Dh) in [(1, 16384, 24, 32), (1, 16384, 6, 128)]:
    q, k, v = (torch.randn(B, N, H, Dh, device=dev).bfloat16() for _ in range(3))
    out = torch.empty(B, N, H * Dh, device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: lib.lcasr_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), 1, B, N, H, Dh, 0, 0, out.data_ptr(), 0, st)
    for _ in range(5):
        assert f() == 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(os.path.basename(os.environ["LCASR_LIB_PATH"]), B, N, H, Dh, f"{us:.1f} us", f"{4.0 * B * H * N * N * Dh / us / 1e6:.1f} TFLOP/s")
