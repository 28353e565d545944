"""Cycle-level phase trace of the tcgen05 GEMM (clock64 stamps of every CTA's producer / MMA issuer / first epilogue
warp) on the MaPLe shapes: where does a launch spend its time — ramp-up, main loop, epilogue, tail?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from federated_multi_modal_b200 import ops, _lib

dev = "cuda"
BF16 = torch.bfloat16
M = int(os.environ.get("GB_M", 6368))
TN = int(os.environ.get("GB_TILE_N", 0))
SHAPES = [("qkv_fwd", M, 2304, 768, "bias16"), ("out_fwd", M, 768, 768, "bias_res32"), ("fc_fwd", M, 3072, 768, "gelu"),
          ("proj_fwd", M, 768, 3072, "bias_res32"), ("proj_dgrad", M, 3072, 768, "dgelu"),
          ("fc_dgrad", M, 768, 3072, "plain16"), ("out_dgrad", M, 768, 768, "plain16"),
          ("qkv_dgrad", M, 768, 2304, "plain16")]
only = sys.argv[1:]
WS = ops.splitk_workspace(dev) if os.environ.get("GB_WS") else None
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
out = {}
for name, m, n, k, epi in SHAPES:
    if only and name not in only:
        continue
    a = torch.randn(m, k, device=dev).to(BF16); b = (torch.randn(n, k, device=dev) * k ** -0.5).to(BF16)
    bias = torch.randn(n, device=dev); resid = torch.randn(m, n, device=dev)
    o16 = torch.empty(m, n, device=dev, dtype=BF16); o16b = torch.empty(m, n, device=dev, dtype=BF16)
    o32 = torch.empty(m, n, device=dev); aux = torch.randn(m, n, device=dev).to(BF16)

    def run():
        if epi == "bias16": ops.gemm(a, b, bias=bias, out_bf16=o16, tile_n=TN, ws=WS)
        elif epi == "plain16": ops.gemm(a, b, out_bf16=o16, tile_n=TN, ws=WS)
        elif epi == "bias_res32": ops.gemm(a, b, bias=bias, residual=resid, out_f32=o32, tile_n=TN, ws=WS)
        elif epi == "gelu": ops.gemm(a, b, bias=bias, act=1, out_bf16=o16, out_pre=o16b, tile_n=TN, ws=WS)
        elif epi == "dgelu": ops.gemm(a, b, act=2, aux=aux, out_bf16=o16, tile_n=TN, ws=WS)
    for _ in range(3): run()
    buf = torch.zeros(148 * 3 * 64, device=dev, dtype=torch.int64)
    if not os.environ.get("GB_NOFLUSH"):
        flush.zero_()
    _lib.call("mfk_debug_set_gemm_trace", buf)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); run(); e.record()
    torch.cuda.synchronize()
    _lib.call("mfk_debug_set_gemm_trace", None)
    t = buf.cpu().reshape(148, 3, 64)
    us = s.elapsed_time(e) * 1e3
    print(f"=== {name} M={m} N={n} K={k} {epi}: {us:.1f} us (traced launch)")
    gt = t[:, 0, 63].clone(); gt0 = int(gt[gt > 0].min())
    ends = []
    for cta in (0, 1, 7, 8, 73, 147):
        p, mm, ep = t[cta, 0], t[cta, 1], t[cta, 2]
        base = int(p[0])
        if base == 0:
            continue
        rel = lambda x: int(x) - base
        ntile = sum(1 for i in range(20) if int(mm[3 * i + 2]) > 0)
        line = [f"CTA {cta:3d}: start +{(int(gt[cta]) - gt0)} ns; pdl_wait done @{rel(p[1])}"]
        for it in range(ntile):
            line.append(f"  tile{it}: prod first/last TMA @{rel(p[2 + 2 * it])}/{rel(p[3 + 2 * it])} | mma acc-free @{rel(mm[3 * it])} "
                        f"data @{rel(mm[3 * it + 1])} issued @{rel(mm[3 * it + 2])} (loop {rel(mm[3 * it + 2]) - rel(mm[3 * it + 1])}) | "
                        f"epi ready @{rel(ep[2 * it])} done @{rel(ep[2 * it + 1])} (epi {rel(ep[2 * it + 1]) - rel(ep[2 * it])})")
        if int(ep[40]) > 0:
            line.append("  split-K unit: partial written @%d | fenced+barrier @%d | all slices seen @%d | slices summed @%d | "
                        "combined @%d | finished @%d" % tuple(rel(ep[i]) for i in range(40, 46)))
        if int(ep[48]) > 0:
            line.append("  tile1 chunks (start / tmem ready / chunk done): " + " | ".join(
                "%d/%d/%d" % tuple(rel(ep[48 + 3 * k + j]) for j in range(3)) for k in range(4) if int(ep[48 + 3 * k]) > 0))
        if int(ep[56]) > 0:
            line.append("  last chunk: enter finish %d | bias+store-wait done %d | staged %d | fenced %d" % tuple(rel(ep[i]) for i in range(56, 60)))
        line.append(f"  stores drained @{rel(ep[62])}")
        print("\n".join(line))
    # whole-grid summary in cycles
    base = t[:, 0, 0]
    ok = base > 0
    end = (t[:, 2, 62] - base)[ok].float()
    ok = ok & (t[:, 1, 1] > 0)
    first_data = (t[:, 1, 1] - base)[ok].float()
    end = (t[:, 2, 62] - base)[ok].float()
    print(f"  grid: CTAs {int(ok.sum())}, first data mean {first_data.mean():.0f} cyc, CTA end min/mean/max "
          f"{end.min():.0f}/{end.mean():.0f}/{end.max():.0f} cyc; start skew max {(int(gt[ok].max()) - gt0)} ns")
    out[name] = dict(us=us, end_mean=float(end.mean()), end_max=float(end.max()), first_data=float(first_data.mean()))
json.dump(out, open("gpurun_out/gemm_trace.json", "w"), indent=1)
