"""One launch of every kernel class of the step on its real operands, between cudaProfilerStart/Stop, for
`ncu --set full --profile-from-start off` (counter-level evidence for attention, LayerNorm / splice,
head, FedAvg and the optimiser — not only the GEMMs).

    python tools/ncu_targets.py [B] [which,comma,separated]

Operands are the saved activations of one real training step at the bench workload (B=32, C=10), so every
launch reads / writes what it does inside a step.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from federated_multi_modal_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.B_PER_GPU
which = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else None
on = lambda name: which is None or name in which
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
t = bench.make_trainer(dev, graph=False)
b = bench.host_batches(1, B, 0)[0]
img, lab = b["img"].to(dev), b["label"].to(dev)
for _ in range(2):
    t.step_async(img, lab)
torch.cuda.synchronize()
eng = t.model.engine
tw = eng.vis
ws, w = tw.ws, tw.w[3]
l = 3
st = ws["stat"][l]
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def cold():
    flush.zero_()  # > 126 MB: the target's operands come from HBM like in a step whose working set is 1.6 GB


torch.cuda.synchronize()
torch.cuda.profiler.start()
if on("attn"):
    cold(); ops.attn_fwd(ws["qkv"][l], ws["att"][l], ws["lse"][l], tw.N, tw.T, tw.heads, False)
    cold(); ops.attn_bwd(ws["qkv"][l], ws["att"][l], ws["dh"], ws["lse"][l], ws["delta"], ws["dqkv"], tw.N, tw.T,
                         tw.heads, False)
if on("ln"):
    cold(); ops.layernorm_fwd(ws["x1"][l], w["ln_1.g"], w["ln_1.b"], y_bf16=ws["h"], mean=st[0], rstd=st[1])
    cold(); ops.layernorm_fwd(ws["x1"][l], w["ln_1.g"], w["ln_1.b"], y_bf16=ws["h"], mean=st[0], rstd=st[1],
                              splice=(eng.deep_vis[l - 1], tw.T, tw.T - eng.n, eng.n))
    pre = f"image_encoder.transformer.resblocks.{l}."
    cold(); ops.layernorm_bwd(ws["dh"], ws["x1"][l], st[0], st[1], w["ln_1.g"], g_in=ws["g"], g_out=ws["g"],
                              g_out_bf16=ws["g16"], dgamma=eng.g[pre + "ln_1.weight"], dbeta=eng.g[pre + "ln_1.bias"],
                              partial_ws=ws["lnp_all"][0], defer=True)
    tok = eng._bufs["vis.tok"][: B * eng.P * tw.D].view(B * eng.P, tw.D)
    cold(); ops.vis_assemble_lnpre(tok, eng.cls, eng.vpos, eng.shared, eng.p["image_encoder.ln_pre.weight"],
                                   eng.p["image_encoder.ln_pre.bias"], eng.vx0, ws["x1"][0], eng.vstat0[0],
                                   eng.vstat0[1], B, eng.Tv, eng.n)
if on("head"):
    fi, ft = eng.last["image_features"], eng.last["text_features"]
    logits = torch.empty(B, eng.C, device=dev)
    loss = torch.empty(1, device=dev)
    hws = torch.empty(ops.head_workspace_floats(B, eng.C, eng.E), device=dev)
    ops.head_forward_backward(fi, ft, eng.logit_scale, lab, logits, loss, eng.last["dfi"], eng.last["dft"], hws)
if on("opt"):
    n = eng.n_update
    norm = torch.empty(1, device=dev)
    pws = torch.empty(296, device=dev)
    hyper = torch.tensor([0.0026, 0.9, 0.0, 5e-4, 1.0, 0.0, 0.0], device=dev)
    cold(); ops.grad_norm(eng.grads[:n], pws, norm)
    cold(); ops.sgd_step(eng.params[:n], eng.grads[:n], eng.momentum[:n], hyper, norm, n)
    eng.repack_trainable()
if on("fedavg"):
    K = 8
    n = eng.n_update
    rows = torch.randn(K, n, device=dev)
    ptrs = torch.tensor([rows[k].data_ptr() for k in range(K)], dtype=torch.int64, device=dev)
    o32, o16 = torch.empty(n, device=dev), torch.empty(n, device=dev, dtype=torch.float16)
    cold(); ops.fedavg_reduce(ptrs, None, float(K), K, n, False, o32, o16, None)
    flag = torch.zeros(1, device=dev, dtype=torch.int32)
    cold(); ops.check_finite(rows[0], flag)
if on("gemm"):
    gw = tw.gemm_ws
    cold(); ops.gemm(ws["h"], w["attn.in_proj.w"], bias=w["attn.in_proj.b"], out_bf16=ws["qkv"][l], ws=gw)
    cold(); ops.gemm(ws["att"][l], w["attn.out_proj.w"], bias=w["attn.out_proj.b"], residual=ws["x1"][l],
                     out_f32=ws["x2"][l], ws=gw)
    cold(); ops.gemm(ws["h2"], w["mlp.c_fc.w"], bias=w["mlp.c_fc.b"], act=1, out_bf16=ws["act"], out_pre=ws["u"][l],
                     ws=gw)
    cold(); ops.gemm(ws["act"], w["mlp.c_proj.w"], bias=w["mlp.c_proj.b"], residual=ws["x2"][l],
                     out_f32=ws["x1"][l + 1], ws=gw)
    cold(); ops.gemm(ws["g16"], w["mlp.c_proj.wT"], act=2, aux=ws["u"][l], out_bf16=ws["du"], ws=gw)
    cold(); ops.gemm(ws["du"], w["mlp.c_fc.wT"], out_bf16=ws["dh"], ws=gw)
    cold(); ops.gemm(ws["g16"], w["attn.out_proj.wT"], out_bf16=ws["dh"], ws=gw)
    cold(); ops.gemm(ws["dqkv"], w["attn.in_proj.wT"], out_bf16=ws["dh"], ws=gw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ncu_targets done")
