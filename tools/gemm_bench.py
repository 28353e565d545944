"""Micro-benchmark of the tcgen05 GEMM on the MaPLe shapes (CUDA events, L2 flushed between launches)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops

dev = "cuda"
BF16, F32 = torch.bfloat16, torch.float32
M = int(os.environ.get("GB_M", 6368))
TN = int(os.environ.get("GB_TILE_N", 0))
CUBLAS = int(os.environ.get("GB_CUBLAS", 1))
SHAPES = [  # name, M, N, K, epilogue
    ("qkv_fwd", M, 2304, 768, "bias16"),
    ("out_fwd", M, 768, 768, "bias_res32"),
    ("fc_fwd", M, 3072, 768, "gelu"),
    ("proj_fwd", M, 768, 3072, "bias_res32"),
    ("proj_dgrad", M, 3072, 768, "dgelu"),
    ("fc_dgrad", M, 768, 3072, "plain16"),
    ("out_dgrad", M, 768, 768, "plain16"),
    ("qkv_dgrad", M, 768, 2304, "plain16"),
    ("fc_wgrad", 3072, 768, M, "plain32"),
    ("text_qkv", 100, 1536, 512, "bias16"),
    ("big_plain", 8192, 8192, 8192, "plain16"),
]
only = sys.argv[1:]
WS = ops.splitk_workspace(dev) if os.environ.get("GB_WS") else None
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
res = {}
for name, m, n, k, epi in SHAPES:
    if only and name not in only:
        continue
    kp = (k + 7) // 8 * 8
    a = torch.randn(m, kp, device=dev).to(BF16); b = (torch.randn(n, kp, device=dev) * k ** -0.5).to(BF16)
    bias = torch.randn(n, device=dev)
    resid = torch.randn(m, n, device=dev)
    o16 = torch.empty(m, n, device=dev, dtype=BF16); o16b = torch.empty(m, n, device=dev, dtype=BF16)
    o32 = torch.empty(m, n, device=dev)
    aux = torch.randn(m, n, device=dev).to(BF16)
    def run():
        if epi == "bias16": ops.gemm(a, b, bias=bias, out_bf16=o16, k=k, tile_n=TN, ws=WS)
        elif epi == "plain16": ops.gemm(a, b, out_bf16=o16, k=k, tile_n=TN, ws=WS)
        elif epi == "plain32": ops.gemm(a, b, out_f32=o32, k=k, tile_n=TN, ws=WS)
        elif epi == "bias_res32": ops.gemm(a, b, bias=bias, residual=resid, out_f32=o32, k=k, tile_n=TN, ws=WS)
        elif epi == "gelu": ops.gemm(a, b, bias=bias, act=1, out_bf16=o16, out_pre=o16b, k=k, tile_n=TN, ws=WS)
        elif epi == "dgelu": ops.gemm(a, b, act=2, aux=aux, out_bf16=o16, k=k, tile_n=TN, ws=WS)
    for _ in range(3): run()
    ts = []
    for _ in range(10):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); run(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    tf = 2.0 * m * n * k / t / 1e6
    # cuBLAS reference point (library call, for context only)
    tt = [0.0]
    for _ in range(5 if CUBLAS else 0):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); torch.matmul(a[:, :k], b[:, :k].t()); e.record(); torch.cuda.synchronize()
        tt.append(s.elapsed_time(e) * 1e3)
    tt.sort()
    res[name] = dict(us=round(t, 1), tflops=round(tf, 1), cublas_us=round(tt[len(tt) // 2], 1))
    print(f"{name:12s} M={m:5d} N={n:5d} K={k:5d} {epi:10s} {t:8.1f} us  {tf:7.1f} TFLOP/s   (torch.matmul plain {tt[len(tt)//2]:.1f} us)")
json.dump(res, open("gpurun_out/gemm_bench.json", "w"), indent=1)
