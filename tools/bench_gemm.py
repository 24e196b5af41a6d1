"""Micro-benchmark of the tcgen05 GEMM (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtvaf_b200 import ops, lib as Lb

def timeit(fn, iters=20, warm=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]

def main():
    T = int(os.environ.get("T", 32768))
    res = []
    for name, (M, N, K) in {"qkv": (T, 2304, 768), "attn_out": (T, 768, 768), "ffn1": (T, 3072, 768), "ffn2": (T, 768, 3072)}.items():
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.linear_fwd(x, w, None, out=out))
        ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out))
        fl = 2.0 * M * N * K
        # dgrad / wgrad
        dy = torch.randn(M, N, device="cuda").bfloat16()
        dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
        ms_d = timeit(lambda: ops.gemm(dy, w, b_mn=True, M=M, N=K, K=N, out=dx))
        dw = torch.zeros(N, K, device="cuda")
        ms_w = timeit(lambda: ops.linear_wgrad(dy, x, dw))
        r = dict(name=name, M=M, N=N, K=K, fwd_ms=ms, fwd_tflops=fl / ms / 1e9, cublas_ms=ms_t, cublas_tflops=fl / ms_t / 1e9,
                 dgrad_ms=ms_d, dgrad_tflops=fl / ms_d / 1e9, wgrad_ms=ms_w, wgrad_tflops=fl / ms_w / 1e9)
        print(json.dumps(r)); res.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)

if __name__ == "__main__":
    main()
