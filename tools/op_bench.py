"""Micro-benchmarks of single kernels through the C ABI (for ncu captures and quick A/B timing).
usage: python tools/op_bench.py attn|gemm [reps]"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fireredtts2_b200 import _native as N

lib = N.load()
P = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
S = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def attn(reps, B=64, H=16, T=3000):
    E = H * 64
    qkv = torch.randn(B, T, 3 * E, device="cuda").half()
    q, k, v = [qkv[..., i * E:(i + 1) * E].contiguous() for i in range(3)]
    out = torch.empty(B, T, E, device="cuda", dtype=torch.half)
    ms = timeit(lambda: N.check(lib.frt2_op_attention(0, P(q), P(k), P(v), P(out), B, H, 64, T, T, 0, 1, S())), reps)
    pairs = sum(8 * min(T, (qq | 7) + 1) for qq in range(0, T, 8))
    fl = 4.0 * 64 * pairs * B * H
    print(f"attention_tc B={B} H={H} T={T}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s")


def gemm(reps, M=192000, K=1024, Nn=1024, act=0, resid=False, out32=False):
    A = torch.randn(1, M, K, device="cuda").half()
    W = (torch.randn(Nn, K, device="cuda") / 32).half()
    bias = torch.randn(Nn, device="cuda")
    o32 = torch.zeros(1, M, Nn, device="cuda") if (out32 or resid) else None
    o16 = None if (out32 or resid) else torch.empty(1, M, Nn, device="cuda", dtype=torch.half)
    ms = timeit(lambda: N.check(lib.frt2_op_gemm(0, P(A), P(W), 1, M, K, 1, Nn, 1.0, P(bias), act,
                                                 P(o32) if resid else None, P(o32), P(o16), S())), reps)
    print(f"gemm M={M} K={K} N={Nn} act={act} resid={resid} out32={out32}: {ms:.3f} ms  {2.0 * M * K * Nn / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    what = sys.argv[1]
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    if what == "attn":
        attn(reps)
    elif what == "gemm":
        gemm(reps, K=1024, Nn=3072)
        gemm(reps, K=1024, Nn=1024, resid=True)
        gemm(reps, K=1024, Nn=4096, act=1)
        gemm(reps, K=4096, Nn=1024, resid=True)
        gemm(reps, K=1024, Nn=1024, out32=True)


def skinny(reps):
    for (K, Nn, taps) in [(1024, 1024, 1), (1024, 3072, 1), (1024, 4096, 1), (4096, 1024, 1), (1024, 1024, 3), (1024, 1024, 7)]:
        A = torch.randn(1, 8, K, device="cuda").half()
        W = (torch.randn(Nn, taps * K, device="cuda") / 32).half()
        bias = torch.randn(Nn, device="cuda")
        o16 = torch.empty(1, 8, Nn, device="cuda", dtype=torch.half)
        for impl in (2, 0):
            ms = timeit(lambda: N.check(lib.frt2_op_gemm(impl, P(A), P(W), 1, 8, K, taps, Nn, 1.0, P(bias), 0, None, None, P(o16), S())), reps)
            print(f"impl={impl} M=8 K={K}x{taps} N={Nn}: {ms * 1e3:.2f} us/launch  {Nn * taps * K * 2 / ms / 1e6:.0f} GB/s")
    x = torch.randn(8, 1024, device="cuda"); g = torch.ones(1024, device="cuda"); b = torch.zeros(1024, device="cuda")
    o = torch.empty(8, 1024, device="cuda", dtype=torch.half)
    ms = timeit(lambda: N.check(lib.frt2_op_layer_norm(P(x), 8, 1024, P(g), P(b), 1e-5, 0, P(o), S())), reps)
    print(f"layer_norm 8 rows: {ms * 1e3:.2f} us/launch")
    E = 1024
    for Tk in (8, 128, 1000):
        q = torch.randn(1, 8, E, device="cuda").half(); k = torch.randn(1, Tk, E, device="cuda").half(); v = torch.randn(1, Tk, E, device="cuda").half()
        out = torch.empty(1, 8, E, device="cuda", dtype=torch.half)
        ms = timeit(lambda: N.check(lib.frt2_op_attention(1, P(q), P(k), P(v), P(out), 1, 16, 64, 8, Tk, Tk - 8, 0, S())), reps)
        print(f"attention_warp Tq=8 Tk={Tk}: {ms * 1e3:.2f} us/launch")
    # empty-ish kernel launch cost through the same path
    ms = timeit(lambda: torch.cuda._sleep(1), reps)
    print(f"torch._sleep(1) launch: {ms * 1e3:.2f} us")


if __name__ == "__main__" and sys.argv[1] == "skinny":
    skinny(int(sys.argv[2]) if len(sys.argv) > 2 else 200)


def ln(reps, rows=192000, C=1024):
    x = torch.randn(rows, C, device="cuda"); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    o = torch.empty(rows, C, device="cuda", dtype=torch.half)
    ms = timeit(lambda: N.check(lib.frt2_op_layer_norm(P(x), rows, C, P(g), P(b), 1e-5, 0, P(o), S())), reps)
    print(f"layer_norm rows={rows} C={C}: {ms:.3f} ms  {rows * C * 6 / ms / 1e6:.0f} GB/s")


if __name__ == "__main__" and sys.argv[1] == "ln":
    ln(int(sys.argv[2]) if len(sys.argv) > 2 else 10)


def resample_bench(reps, B=64, n=720000):
    from fireredtts2_b200.codec import resample
    x = torch.randn(B, n, device="cuda") * 0.1
    ms = timeit(lambda: resample(x, 24000, 16000), reps)
    by = B * n * 4 * (1 + 2 / 3)
    print(f"resample 24k->16k B={B} n={n}: {ms:.3f} ms  {by / ms / 1e6:.0f} GB/s (algorithmic: 4 B in + 8/3 B out per input sample)")


if __name__ == "__main__" and sys.argv[1] == "resample":
    resample_bench(int(sys.argv[2]) if len(sys.argv) > 2 else 10)
