"""Micro-benchmark of the tcgen05 GEMMs (CUDA events, L2 flushed between iterations), per implementation
('single' = 128x256 single-CTA tiles, 'auto' = CTA-pair cta_group::2 kernel) next to cuBLAS (torch.matmul)."""
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
    impls = os.environ.get("IMPLS", "single,auto").split(",")
    res = []
    shapes = {"qkv": (T, 2304, 768), "attn_out": (T, 768, 768), "ffn1": (T, 3072, 768), "ffn2": (T, 768, 3072)}
    for name, (M, N, K) in shapes.items():
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        out2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        res_in = torch.randn(M, N, device="cuda").bfloat16()
        dy = torch.randn(M, N, device="cuda").bfloat16()
        dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
        dw = torch.zeros(N, K, device="cuda")
        fl = 2.0 * M * N * K
        ms_t = timeit(lambda: torch.matmul(x, w.t(), out=out))
        r = dict(name=name, M=M, N=N, K=K, cublas_tflops=round(fl / ms_t / 1e9, 1))
        for impl in impls:
            ops.set_gemm_impl(impl)
            t = {}
            t["fwd"] = timeit(lambda: ops.linear_fwd(x, w, None, out=out))
            t["fwd_bias_gelu"] = timeit(lambda: ops.linear_fwd(x, w, bias, out=out, mode=Lb.EPI_GELU, out2=out2))
            t["fwd_resid_drop"] = timeit(lambda: ops.linear_fwd(x, w, bias, out=out, mode=Lb.EPI_RESID, aux=res_in,
                                                                 p_drop=0.1, seed=5))
            t["dgrad"] = timeit(lambda: ops.gemm(dy, w, b_mn=True, M=M, N=K, K=N, out=dx))
            pre = torch.randn(M, K, device="cuda").bfloat16()
            t["dgrad_dgelu"] = timeit(lambda: ops.gemm(dy, w, b_mn=True, M=M, N=K, K=N, out=dx, mode=Lb.EPI_MUL_DGELU, aux=pre))
            t["fwd_gelu_grad"] = timeit(lambda: ops.linear_fwd(x, w, bias, out=out, mode=Lb.EPI_GELU_GRAD, out2=out2))
            t["dgrad_mulaux"] = timeit(lambda: ops.gemm(dy, w, b_mn=True, M=M, N=K, K=N, out=dx, mode=Lb.EPI_MUL_AUX, aux=pre))
            t["dgrad_resid"] = timeit(lambda: ops.gemm(dy, w, b_mn=True, M=M, N=K, K=N, out=dx, mode=Lb.EPI_RESID, aux=pre))
            t["wgrad"] = timeit(lambda: ops.linear_wgrad(dy, x, dw))
            for k, v in t.items():
                r["%s_%s_tflops" % (impl, k)] = round(fl / v / 1e9, 1)
        ops.set_gemm_impl("auto")
        print(json.dumps(r), flush=True)
        res.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
