"""In-situ phase trace of the vision GEMMs of one transformer layer during a real (eager, PDL-chained) training step:
warm L2, real operands, the real neighbours before and after each launch. Prints per-role cycle stamps."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from federated_multi_modal_b200 import ops, _lib

LAYER = int(os.environ.get("GT_LAYER", 5))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
t = bench.make_trainer(dev, graph=False)
b = bench.host_batches(1, bench.B_PER_GPU, 0)[0]
img, lab = b["img"].to(dev), b["label"].to(dev)
for _ in range(3):
    t.step_async(img, lab)
torch.cuda.synchronize()

M = 32 * 199
bufs = []
orig = ops.call
state = {"n": 0}


def traced(name, *a, **kw):
    if name == "mfk_gemm_bf16" and a[4] == M:
        i = state["n"]
        state["n"] += 1
        # forward: 4 GEMMs per layer (layers 0..10 full), backward runs layers in reverse with 4 dgrads each
        fwd_lo, fwd_hi = 4 * LAYER, 4 * LAYER + 4
        if fwd_lo <= i < fwd_hi or state.get("bwd_lo", 10 ** 9) <= i < state.get("bwd_lo", 10 ** 9) + 4:
            buf = torch.zeros(148 * 3 * 64, device=dev, dtype=torch.int64)
            bufs.append((i, a[5], a[6], a[8], a[11] is not None, buf))
            _lib.call("mfk_debug_set_gemm_trace", buf)
            r = orig(name, *a, **kw)
            _lib.call("mfk_debug_set_gemm_trace", None)
            return r
    return orig(name, *a, **kw)


# count the M-row GEMMs of one step first, to locate the backward of LAYER
ops.call = lambda name, *a, **kw: (state.__setitem__("n", state["n"] + (name == "mfk_gemm_bf16" and a[4] == M)), orig(name, *a, **kw))[1]
t.step_async(img, lab)
torch.cuda.synchronize()
total = state["n"]
nfwd = 4 * 11 + 1          # 11 full layers + in_proj of the last
nbwd = total - nfwd        # 1 (last in_proj dgrad) + 4 per full layer
state["n"] = 0
state["bwd_lo"] = nfwd + 1 + 4 * (10 - LAYER)
print(f"M-row GEMMs per step: {total} (fwd {nfwd}, bwd {nbwd}); tracing layer {LAYER}")
ops.call = traced
torch.cuda._sleep(int(4e7))  # let the host queue the whole step ahead
t.step_async(img, lab)
torch.cuda.synchronize()
ops.call = orig

res = []
for i, N, K, act, has_res, buf in bufs:
    tr = buf.cpu().reshape(148, 3, 64)
    base = tr[:, 0, 0]
    ok = base > 0
    end = (tr[:, 2, 62] - base)[ok].float()
    fd = (tr[:, 1, 1] - tr[:, 0, 1])[ok].float()         # pdl_wait done -> first data
    wait = (tr[:, 0, 1] - base)[ok].float()               # entry -> pdl_wait done (waiting for the previous kernel)
    print(f"=== gemm #{i} N={N} K={K} act={act} res={has_res}: pdl wait mean {wait.mean():.0f} cyc, first data +{fd.mean():.0f}, "
          f"CTA end (from entry) min/mean/max {end.min():.0f}/{end.mean():.0f}/{end.max():.0f}; "
          f"busy (from pdl_wait) mean/max {(end - wait).mean():.0f}/{(end - wait).max():.0f}")
    for cta in (0, 5, 73, 147):
        p, mm, ep = tr[cta, 0], tr[cta, 1], tr[cta, 2]
        b0 = int(p[1])
        rel = lambda x: int(x) - b0
        ntile = sum(1 for j in range(20) if int(mm[3 * j + 2]) > 0)
        for it in range(ntile):
            print(f"  CTA{cta:3d} tile{it}: mma acc-free @{rel(mm[3 * it])} data @{rel(mm[3 * it + 1])} issued @{rel(mm[3 * it + 2])} "
                  f"(loop {rel(mm[3 * it + 2]) - rel(mm[3 * it + 1])}) | epi ready @{rel(ep[2 * it])} done @{rel(ep[2 * it + 1])} "
                  f"(epi {rel(ep[2 * it + 1]) - rel(ep[2 * it])})")
        if int(ep[40]) > 0:
            print("  CTA%3d split-K unit: partial written @%d | fenced+barrier @%d | all slices seen @%d | slices summed @%d | "
                  "combined @%d | finished @%d | re-armed @%d" % ((cta,) + tuple(rel(ep[i]) for i in range(40, 47))))
        if int(ep[48]) > 0:  # chunk-level stamps of warp 4 in its second tile: (before tmem ld, after ld, chunk done) x 4
            print("  CTA%3d tile1 chunks: " % cta + " | ".join(
                "ld @%d +%d fin +%d" % (rel(ep[48 + 3 * j]), int(ep[49 + 3 * j] - ep[48 + 3 * j]), int(ep[50 + 3 * j] - ep[49 + 3 * j]))
                for j in range(4) if int(ep[48 + 3 * j]) > 0))
            if int(ep[32]) > 0:
                print("  CTA%3d tile1 chunk0 inside finish_chunk: entry @%d | staging free +%d | math+st.shared +%d | fence+syncwarp +%d | stores issued +%d"
                      % (cta, rel(ep[32]), int(ep[33] - ep[32]), int(ep[34] - ep[33]), int(ep[35] - ep[34]), int(ep[36] - ep[35])))
        print(f"  CTA{cta:3d} stores drained @{rel(ep[62])}")
    res.append(dict(i=i, N=N, K=K, act=act, res=has_res, wait=float(wait.mean()), busy_mean=float((end - wait).mean()),
                    busy_max=float((end - wait).max())))
json.dump(res, open("gpurun_out/gemm_trace_step.json", "w"), indent=1)
