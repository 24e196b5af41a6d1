"""Runs every hot kernel variant of the training step at the bench shape (B=256 or env B=512, L=128, P=16, bf16) twice, so one

  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'gemm_bf16_tc2|attn_|layernorm_bwd' -o gpurun_out/hot python tools/profile_hot_kernels.py

captures each of them warm, once.  Prints the launch order (one name per kernel) so reports can be matched."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mtvaf_b200 import ops, lib as Lb


def main():
    B, Lq, P, nh, d = int(os.environ.get("B", 256)), 128, 16, 12, 64      # B=512 = the bench batch
    H = nh * d
    T = B * Lq
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)

    def rn(*s, scale=1.0):
        return (torch.randn(*s, device=dev, generator=g) * scale).bfloat16()

    x = rn(T, H)
    xi = rn(T, 4 * H)
    w_qkv, w_o, w_1, w_2 = rn(3 * H, H, scale=0.03), rn(H, H, scale=0.03), rn(4 * H, H, scale=0.03), rn(H, 4 * H, scale=0.03)
    b_h, b_i, b_qkv = torch.randn(H, device=dev), torch.randn(4 * H, device=dev), torch.randn(3 * H, device=dev)
    o_h, o_i, o_i2, o_qkv = torch.empty_like(x), torch.empty_like(xi), torch.empty_like(xi), rn(T, 3 * H)
    dw = torch.zeros(4 * H, H, device=dev)
    kp, vp = rn(B, nh, P, d), rn(B, nh, P, d)
    lens = torch.randint(8, Lq + 1, (B,), device=dev)
    key_mask = (torch.arange(Lq, device=dev).unsqueeze(0) < lens.unsqueeze(1)).long()
    dkp = torch.zeros(B, nh, P, d, device=dev)
    dvp = torch.zeros(B, nh, P, d, device=dev)
    gam, bet = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    dbias_i = torch.zeros(4 * H, device=dev)
    dgam, dbet, dbias = torch.zeros(H, device=dev), torch.zeros(H, device=dev), torch.zeros(H, device=dev)

    state = {}

    def attn_f():
        state["ctx"], state["lse"], _ = ops.attention_fwd(o_qkv, kp, vp, key_mask, B, Lq, nh, d, p_drop=0.1, seed=3)

    def attn_b():
        if "ctx" not in state:              # ONLY=attn_bwd: the backward needs the forward's outputs
            attn_f()
        ops.attention_bwd(x, o_qkv, kp, vp, key_mask, state["ctx"], state["lse"], B, Lq, nh, d, dkp=dkp, dvp=dvp,
                          p_drop=0.1, seed=3)

    def ln_b():
        _, mean, rstd = state["ln"]
        ops.layernorm_bwd(x, o_h, gam, mean, rstd, dgam, dbet, d_bias=dbias, p_drop=0.1, seed=9)

    state["ln"] = ops.layernorm_fwd(o_h, gam, bet, 1e-5)
    variants = [
        ("qkv_fwd_store", lambda: ops.linear_fwd(x, w_qkv, b_qkv, out=o_qkv)),
        ("ffn1_fwd_gelu", lambda: ops.linear_fwd(x, w_1, b_i, out=o_i, mode=Lb.EPI_GELU, out2=o_i2)),
        ("attn_out_fwd_resid", lambda: ops.linear_fwd(x, w_o, b_h, out=o_h, mode=Lb.EPI_RESID, aux=x, p_drop=0.1, seed=5)),
        ("ffn2_fwd_resid", lambda: ops.linear_fwd(xi, w_2, b_h, out=o_h, mode=Lb.EPI_RESID, aux=x, p_drop=0.1, seed=6)),
        ("ffn2_dgrad_dgelu", lambda: ops.gemm(x, w_2, b_mn=True, M=T, N=4 * H, K=H, out=o_i, mode=Lb.EPI_MUL_DGELU, aux=o_i2)),
        ("ffn2_dgrad_dgelu_colsum", lambda: ops.gemm(x, w_2, b_mn=True, M=T, N=4 * H, K=H, out=o_i, mode=Lb.EPI_MUL_DGELU,
                                                     aux=o_i2, colsum=dbias_i)),
        ("ffn1_fwd_gelu_grad", lambda: ops.linear_fwd(x, w_1, b_i, out=o_i, mode=Lb.EPI_GELU_GRAD, out2=o_i2)),
        ("ffn2_dgrad_mulaux_colsum", lambda: ops.gemm(x, w_2, b_mn=True, M=T, N=4 * H, K=H, out=o_i, mode=Lb.EPI_MUL_AUX,
                                                      aux=o_i2, colsum=dbias_i)),
        ("ffn1_dgrad_resid", lambda: ops.gemm(xi, w_1, b_mn=True, M=T, N=H, K=4 * H, out=o_h, mode=Lb.EPI_RESID, aux=x)),
        ("attn_out_dgrad_store", lambda: ops.gemm(x, w_o, b_mn=True, M=T, N=H, K=H, out=o_h)),
        ("ffn1_wgrad", lambda: ops.linear_wgrad(xi, x, dw)),
        ("attn_fwd", attn_f),
        ("attn_bwd", attn_b),
        ("layernorm_bwd", ln_b),
    ]
    only = os.environ.get("ONLY")
    if only:
        variants = [v for v in variants if v[0] in only.split(",")]
    for rep in range(2):
        if rep == 1:
            torch.cuda.profiler.start()          # ncu --profile-from-start off: capture the warm pass only
        for name, fn in variants:
            fn()
            torch.cuda.synchronize()
            if rep == 1:
                print(name, flush=True)
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
