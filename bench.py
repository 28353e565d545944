"""bench.py — MaPLe ViT-B/16 training throughput on N B200 GPUs of one node (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (the reference's CPU implementation, rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic input: CustomCLIP forward + backward for
the reference's trainable set + clip_grad_norm_(1.0) + SGD(momentum, wd) — BASELINE config 2: random-init
ViT-B/16, EuroSAT-shaped (10 classes), batch 32 per GPU (one federated client per GPU), bf16 tensor cores.
`value` = images/s with inputs already resident in HBM (CUDA-event timed, max over ranks); `e2e` = the same
metric through the public trainer API (MaPLe.run_epoch: the Dassl epoch loop, one forward_backward per batch) with
pinned HOST batches: H2D copy of every step's images (on a side stream, under the previous step) and D2H read of every
step's loss inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "MaPLe ViT-B/16 train images/sec"
UNIT = "images/s"
B_PER_GPU, N_CLS, N_CTX, DEPTH = 32, 10, 2, 9
WORKLOAD = ("MaPLe ViT-B/16 random-init, EuroSAT-shaped (10 classes), batch 32 per GPU, bf16, n_ctx=2 depth=9, "
            "1 federated client per GPU, fwd+bwd (reference trainable set: prompts+LN+resblocks.11) + clip + SGD")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1392.1), d.get("hbm_gbs", 6531.9), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        """ONE nvidia-smi process (started by rank 0 only) sampling every GPU of the run."""
        self.indices, self.proc, self.lines = list(indices), None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ",".join(str(i) for i in self.indices),
                                          f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], 0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1])); pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "gpus_sampled": len(self.indices),
                "sm_mhz_min": sm[0] if sm else None}


def make_trainer(device, graph=True):
    from federated_multi_modal_b200 import synth
    from federated_multi_modal_b200.trainers import MaPLe
    cfg = synth.make_cfg(n_ctx=N_CTX, depth=DEPTH, prec="bf16")
    cfg.USE_CUDA_GRAPH = graph
    t = MaPLe(cfg, client_id=int(os.environ.get("RANK", 0)), classnames=synth.synthetic_classnames(N_CLS))
    t.model.train()
    return t


def host_batches(n, B, seed):
    from federated_multi_modal_b200 import synth
    out = []
    for i in range(n):
        img, lab = synth.make_batch(B, N_CLS, seed + i)
        out.append({"img": img.pin_memory(), "label": lab.pin_memory()})
    return out


def cpu_baseline(sample_B=4, steps=2, threads=None):
    """The oracle port (oracle/maple_cpu.py: fp32 torch-CPU restatement of the reference step) timed on the
    host cores, on a bounded sample of the same workload."""
    from federated_multi_modal_b200 import synth
    from oracle.maple_cpu import MapleOracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import customclip_state_dict
    if threads:
        torch.set_num_threads(threads)
    sd, tok = customclip_state_dict(N_CLS)
    orc = MapleOracle(sd, tok)
    img, lab = synth.make_batch(sample_B, N_CLS, 7)
    orc.forward_backward(img, lab)  # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.forward_backward(img, lab)
    dt = (time.perf_counter() - t0) / steps
    return {"value": sample_B / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} fwd+bwd steps of {sample_B} images x {N_CLS} classes (fp32, oracle/maple_cpu.py), "
                      f"{dt:.2f} s/step"}, dt


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the step on the host cores. The unmodified
    reference is used when /root/reference is present (build container); on the GPU box it cannot travel
    (pure-Python tree outside the repo), so the oracle port — pinned to it by the golden fixtures — is timed."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from oracle import ref_harness as rh
    from federated_multi_modal_b200 import synth
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample_B = 8
    if rh.available():
        kind = "reference"
        cfg = synth.make_cfg()
        model = rh.build_reference_customclip(synth.random_clip_state_dict(0), synth.synthetic_classnames(N_CLS), cfg,
                                              synth.random_prompt_learner_state(1), fp32=True)
        model.train()
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.SGD(params, lr=0.0026, momentum=0.9, weight_decay=5e-4)
        img, lab = synth.make_batch(sample_B, N_CLS, 7)

        def step():
            loss = model(img, lab)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            return loss.item()
    else:
        kind = "port"
        from oracle.maple_cpu import MapleOracle
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import customclip_state_dict
        sd, tok = customclip_state_dict(N_CLS)
        orc = MapleOracle(sd, tok)
        img, lab = synth.make_batch(sample_B, N_CLS, 7)

        mom = {}

        def step():
            # the reference's whole step (trainers/maple.py:588-598): fwd + bwd + clip_grad_norm_(1.0) + SGD(momentum
            # 0.9, wd 5e-4) on the oracle's fp32 parameters
            out = orc.forward_backward(img, lab)
            G = out["grads"]
            coef = min(1.0, 1.0 / (float(torch.sqrt(sum((g.double() ** 2).sum() for g in G.values()))) + 1e-6))
            for k, g in G.items():
                d = g * coef + 5e-4 * orc.P[k]
                mom[k] = d.clone() if k not in mom else mom[k].mul_(0.9).add_(d)
                orc.P[k].sub_(0.0026 * mom[k])
            return out["loss"].item()
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    k = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    dt = (time.perf_counter() - t0) / k
    val = sample_B / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
            "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm: each step is a bounded sample of 8 of the 32 images "
                                                    "(fwd + bwd + clip_grad_norm_ + SGD on all host threads)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"{k} steps of {sample_B} images x {N_CLS} classes, fp32, fwd+bwd+clip+SGD"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def gemm_probe(eng, reps=5):
    """The dominant kernel timed on its own, on the launching stream, with the step's real operands: the 8
    tensor-core GEMMs of every vision layer (QKV, out-proj+residual, c_fc+QuickGELU, c_proj+residual and their four
    dgrads), 12 layers back to back (operands of consecutive launches differ; the 12-layer set is ~1.5 GB >> L2)."""
    from federated_multi_modal_b200 import ops
    tw = eng.vis
    ws, M, D, gw = tw.ws, tw.M, tw.D, tw.gemm_ws

    def run():
        for l in range(tw.L):
            w = tw.w[l]
            ops.gemm(ws["h"], w["attn.in_proj.w"], bias=w["attn.in_proj.b"], out_bf16=ws["qkv"][l], ws=gw)
            ops.gemm(ws["att"][l], w["attn.out_proj.w"], bias=w["attn.out_proj.b"], residual=ws["x1"][l],
                     out_f32=ws["x2"][l], ws=gw)
            ops.gemm(ws["h2"], w["mlp.c_fc.w"], bias=w["mlp.c_fc.b"], act=1, out_bf16=ws["act"], out_pre=ws["u"][l], ws=gw)
            ops.gemm(ws["act"], w["mlp.c_proj.w"], bias=w["mlp.c_proj.b"], residual=ws["x2"][l],
                     out_f32=ws["x1"][l + 1], ws=gw)
            ops.gemm(ws["g16"], w["mlp.c_proj.wT"], act=2, aux=ws["u"][l], out_bf16=ws["du"], ws=gw)
            ops.gemm(ws["du"], w["mlp.c_fc.wT"], out_bf16=ws["dh"], ws=gw)
            ops.gemm(ws["g16"], w["attn.out_proj.wT"], out_bf16=ws["dh"], ws=gw)
            ops.gemm(ws["dqkv"], w["attn.in_proj.wT"], out_bf16=ws["dh"], ws=gw)
    run()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        run()
    e.record()
    torch.cuda.synchronize()
    launches = reps * tw.L * 8
    flops = reps * tw.L * 2.0 * M * D * D * 24.0
    t = s.elapsed_time(e) * 1e-3
    return {"tflops": flops / t / 1e12, "us_per_launch": t / launches * 1e6, "launches": launches,
            "flops_per_launch": flops / launches}


def run_ours(args):
    import torch.distributed as dist
    from federated_multi_modal_b200 import _lib
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly ONE JSON line: everything else any library prints (NCCL's version banner appears on
    # stdout at communicator creation) is sent to stderr; the JSON goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    trainer = make_trainer(dev, graph=not args.no_graph)
    eng = trainer.model.engine
    B, K, W = B_PER_GPU, args.steps, max(args.warmup, 3)
    pool = host_batches(4, B, 1000 * rank)
    dev_pool = [(b["img"].to(dev), b["label"].to(dev)) for b in pool]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        t = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() * 1e-3

    # ---- device-resident leg (`value`)
    step_dev = lambda i: trainer.step_async(*dev_pool[i % len(dev_pool)])
    for i in range(W):
        step_dev(i)
    k0 = _lib.kernel_count
    step_dev(0)
    kernels_per_step_eager = _lib.kernel_count - k0  # 0 when replaying a graph: counted at capture instead
    sampler = ClockSampler(range(world)) if rank == 0 else None   # one poller for the whole run, not one per rank
    if sampler:
        sampler.start()
    t_dev = timed(step_dev, K)
    clocks = sampler.stop() if sampler else None
    loss_now, _, _ = trainer.read_step_result()

    # ---- end-to-end leg (`e2e`): public trainer API, pinned host batches, H2D + D2H every step
    # The call a user makes is the Dassl epoch loop: `MaPLe.run_epoch` over the client's train loader, one
    # `forward_backward` per batch (loss returned to the host every step). The loader here yields the pinned host
    # batches; every step's input crosses PCIe inside the timed region (run_epoch copies batch k+1 on a side stream
    # while step k computes) and every step ends with the 12-byte D2H read of (loss, grad norm, validity flag).
    from types import SimpleNamespace

    class _PinnedLoader:
        def __init__(self, n):
            self.n = n
        def __len__(self):
            return self.n
        def __iter__(self):
            for i in range(self.n):
                yield dict(pool[i % len(pool)])

    def epoch_e2e(n):
        trainer.dm = SimpleNamespace(train_loader=_PinnedLoader(n), test_loader=None)
        trainer.run_epoch(0)
    epoch_e2e(3)
    t_e2e = timed(lambda i: epoch_e2e(K), 1)

    # ---- FedAvg round-end exchange of the trainable arena (all clients of all ranks), timed separately
    from federated_multi_modal_b200.fed import FedAvgExchange
    # strict: a failed NVLink (symmetric-memory) setup raises instead of silently measuring the all-gather transport
    ex = FedAvgExchange(eng.n_update, 1, dev, transport=args.fed_transport, strict_transport=True)
    def fedavg(i):
        ex.publish(0, eng.params)
        rows = ex.gather()
        ex.reduce(rows)
    fedavg(0)
    t_fed = timed(fedavg, 5) / 5

    # ---- multi-GPU correctness, outside any timed region: every rank compares what the exchange produced from the
    # clients' REAL trained arenas with the CPU oracle's fixed-order mean of the same gathered rows (checker only)
    fed_check = fedavg_bitexact_check(ex, eng, dev, world)

    # ---- one federated round of this rank's client: 16 local steps + round-end FedAvg exchange + broadcast
    # (load averaged arena, refresh bf16 copies, drop optimiser state) — BASELINE's "FedAvg round time"
    ROUND_STEPS = 16
    def fed_round(i):
        for s_ in range(ROUND_STEPS):
            trainer.step_async(*dev_pool[s_ % len(dev_pool)])
        ex.publish(0, eng.params)
        mean32, mean16, _, _ = ex.reduce(ex.gather())
        eng.params[: eng.n_update].copy_(mean16)
        eng.repack_trainable()
        eng.reset_optimizer_state()
    fed_round(0)
    t_round = timed(fed_round, 2) / 2

    # ---- BASELINE config 4 shape: 4 co-located clients per GPU (21 classes, batch 64, 16 local steps each), round-end
    # exchange over all 4 x N clients
    c4 = None if args.no_c4 else c4_round(dev, world, rank, timed, args.fed_transport)

    if rank == 0:
        peak_tf, peak_hbm, peak_src = peaks()
        probe = gemm_probe(eng)
        # kernels per step: count one eager step's launches
        k0 = _lib.kernel_count
        trainer._use_graph, sg = False, trainer._use_graph
        trainer.step_async(*dev_pool[0]); torch.cuda.synchronize()
        trainer._use_graph = sg
        kernels_per_step = _lib.kernel_count - k0
        step_flops = eng.flops_per_step(B)
        achieved = probe["tflops"]
        # CPU legs: rank 0 at N = 1 only (under torchrun every rank is pinned to OMP_NUM_THREADS=1)
        cpu = cpu_baseline()[0] if world == 1 else None
        fed_roof = fedavg_roofline(eng, dev, peak_hbm)
        fed_cpu = cpu_fedavg_baseline(eng.n_update) if world == 1 else None
        value = world * B * K / t_dev
        img_bytes = pool[0]["img"].numel() * 4 + pool[0]["label"].numel() * 8
        traffic = roofline_traffic()
        burst = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops") \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": t_dev / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "classes": N_CLS,
                       "text_rows_per_class": eng.Te, "parallelism": f"{world} independent clients (dp{world}), "
                       "FedAvg exchange at round end only",
                       "cuda_graph": not args.no_graph,
                       "l2": "per-step working set (saved activations ~1.6 GB) >> 126 MB L2; no explicit flush"},
            "e2e": {"value": world * B * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": img_bytes,
                    "d2h_bytes_per_step": 12, "ms_per_step": t_e2e / K * 1e3},
            "gpu_launches": kernels_per_step * K,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         "frac_of_burst_peak": (achieved / burst) if burst else None,
                         "peak_kind": "sustained bf16 (cuBLAS 8192^3 back to back for 4 s); the probe is a 12 ms chain of "
                                      "480 launches, so frac_of_burst_peak (best-of-10 cuBLAS) is the stricter reading",
                         "traffic": traffic["bytes"] if traffic else None,
                         "traffic_note": traffic["note"] if traffic else "no ncu --set full capture committed",
                         "kernel": "gemm_bf16_tn_kernel (tcgen05/TMEM/TMA)", "peak_source": peak_src,
                         "how": "96 real-operand vision GEMM launches x 5, CUDA events on the launching stream",
                         "us_per_launch": probe["us_per_launch"], "flops_per_launch": probe["flops_per_launch"],
                         "step_flops": step_flops,
                         "step_frac": step_flops / (t_dev / K) / 1e12 / peak_tf,
                         "kernel_shares": "per-kernel shares of the step: profiles/ (ncu launch list + "
                                          "tools/launch_summary.py); not re-derived here — CUDA-event intervals on the "
                                          "two concurrent tower streams overlap and do not sum to the step"},
            "cpu_baseline": cpu,
            "fedavg_exchange_ms": t_fed * 1e3,
            "fedavg_transport": ex.transport,
            "fedavg_bitexact": fed_check["bitexact"],
            "fedavg_check": fed_check,
            "fedavg_exchange_nvlink_gbs_per_rank": ((world - 1) * eng.n_update * 4 / t_fed / 1e9) if world > 1 else None,
            "fedavg_roofline": fed_roof,
            "fedavg_cpu_baseline": fed_cpu,
            "fedavg_round_s": t_round,
            "fedavg_round_config": f"{world} clients (1 per GPU), {ROUND_STEPS} local steps of batch {B}, all-rank "
                                   f"exchange of the {eng.n_update}-element fp32 trainable arena + broadcast",
            "fed_round_c4": c4,
            "loss": loss_now,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not fed_check["bitexact"] or (c4 is not None and not c4.get("fedavg_bitexact", True)):
        raise SystemExit("bench.py: FedAvg exchange output differs from the fixed-order oracle")


def fedavg_bitexact_check(ex, eng, dev, world):
    """Every rank: publish the real arena (perturbed per rank so the clients differ), exchange, and compare with
    oracle/maple_cpu.fedavg_oracle on the gathered rows; also that all ranks hold the same bytes. Returns a dict
    with the AND over ranks. The oracle is only the checker here (never timed, never on the product path)."""
    import torch.distributed as dist
    from oracle.maple_cpu import fedavg_oracle
    rank = dist.get_rank() if world > 1 else 0
    mine = eng.params[: eng.n_update].clone()
    mine.mul_(1.0 + 0.03125 * rank).add_(1e-3 * rank)          # distinct, deterministic per-rank client tensors
    ex.publish(0, mine, ok=True, n_samples=10 + rank)
    rows = ex.gather()
    rows_cpu = [r.clone().cpu() for r in rows]
    ok = True
    for weighted in (False, True):
        m32, m16, valid, _ = ex.reduce(rows, weighted=weighted)
        r32, r16 = fedavg_oracle(rows_cpu, [10.0 + k for k in range(len(rows_cpu))] if weighted else None)
        ok = ok and len(valid) == len(rows_cpu) and torch.equal(m32.cpu(), r32) and torch.equal(m16.cpu(), r16)
        if world > 1:
            # the sharded transport stores results straight into every peer's output buffers: a faster rank must not
            # start the next reduce while a slower one is still copying this one's result to the host (a round of the
            # trainer is ordered by its publish + status all-gather; two reduces in a row, as here, are not)
            dist.barrier()
    chk = m32.double().sum().reshape(1).clone()
    flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
    same = True
    if world > 1:
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        same = all(torch.equal(x, lst[0]) for x in lst)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"bitexact": bool(int(flag.item())) and same, "identical_on_all_ranks": same, "clients": len(rows_cpu),
            "elements": int(eng.n_update), "modes": ["uniform", "sample-count weighted"], "transport": ex.transport,
            "requested_transport": ex.requested_transport}


def fedavg_roofline(eng, dev, peak_hbm):
    """mfk_fedavg_reduce alone on K client arenas resident in local HBM (K = 2 / 8 / 32): algorithmic bytes
    (K reads + fp32 write + fp16 write) x n x 4 B over the kernel time (CUDA events, 10 launches; the K x 55 MB
    inputs exceed L2 from K = 4)."""
    from federated_multi_modal_b200 import ops
    n = eng.n_update
    out = {}
    base = eng.params[:n]
    for K in (2, 8, 32):
        rows = torch.empty(K, n, device=dev, dtype=torch.float32)
        rows.copy_(base.unsqueeze(0).expand(K, n))
        rows.mul_(torch.linspace(0.5, 1.5, K, device=dev).unsqueeze(1))
        ptrs = torch.tensor([rows[k].data_ptr() for k in range(K)], dtype=torch.int64, device=dev)
        o32 = torch.empty(n, device=dev, dtype=torch.float32)
        o16 = torch.empty(n, device=dev, dtype=torch.float16)
        for _ in range(2):
            ops.fedavg_reduce(ptrs, None, float(K), K, n, False, o32, o16, None)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s.record()
        for _ in range(10):
            ops.fedavg_reduce(ptrs, None, float(K), K, n, False, o32, o16, None)
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) * 1e-3 / 10
        byt = (K * 4 + 4 + 2) * n
        out[f"K{K}"] = {"ms": t * 1e3, "bytes": byt, "achieved_gbs": byt / t / 1e9, "frac": byt / t / 1e9 / peak_hbm}
        del rows
    out["bound"], out["peak_gbs"], out["kernel"] = "hbm", peak_hbm, "fedavg_kernel<float> (fixed-order reduce)"
    return out


def cpu_fedavg_baseline(n, threads=None):
    """safe_average_weights (reference trainers/maple_fed.py:309-315) on the host cores: the oracle's restatement on
    the trainable arena (n fp32 elements per client) at K = 2 / 8 / 32 — the CPU side of "FedAvg round time"."""
    from oracle.maple_cpu import fedavg_oracle
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(0)
    base = torch.randn(n, generator=g)
    out = {"cores": torch.get_num_threads(), "kind": "port", "elements": int(n),
           "what": "oracle/maple_cpu.fedavg_oracle (fp32 cast, nan_to_num, fixed-order sum, /K, .half()) on K tensors of "
                   "the trainable arena's size"}
    for K in (2, 8, 32):
        rows = [base * (1.0 + 0.01 * k) for k in range(K)]
        fedavg_oracle(rows[:2])
        t0 = time.perf_counter()
        fedavg_oracle(rows)
        out[f"K{K}_ms"] = (time.perf_counter() - t0) * 1e3
    return out


def roofline_traffic():
    """DRAM bytes of ONE launch of the dominant kernel from the newest committed ncu --set full capture
    (profiles/roofline_traffic.json, written by tools/ncu_summary.py from the raw CSV) — not measurable live."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return {"bytes": d["dram_bytes_per_launch"], "note": d.get("note", "")}


def c4_round(dev, world, rank, timed, transport):
    """BASELINE config 4 on this box: 4 clients per GPU (21 classes, batch 64, 16 local steps per client and round)
    sharing one frozen CLIP copy and one activation workspace, round-end FedAvg over all 4 x N clients, broadcast."""
    import torch.distributed as dist
    from federated_multi_modal_b200 import synth
    from federated_multi_modal_b200.fed import FedAvgExchange
    from federated_multi_modal_b200.trainers import MaPLe
    from oracle.maple_cpu import fedavg_oracle
    C, Bc, KL, STEPS = 21, 64, 4, 16
    cfg = synth.make_cfg(n_ctx=N_CTX, depth=DEPTH, prec="bf16")
    names = synth.synthetic_classnames(C)
    trainers, share = [], None
    for j in range(KL):
        t = MaPLe(cfg, client_id=rank * KL + j, classnames=names, share_engine=share)
        t.model.train()
        share = share or t.model.engine
        trainers.append(t)
    pool = [tuple(x.to(dev) for x in synth.make_batch(Bc, C, 4000 + 10 * rank + i)) for i in range(2)]
    n = share.n_update
    ex = FedAvgExchange(n, KL, dev, transport=transport, strict_transport=True)
    def one_round(i):
        for j, t in enumerate(trainers):
            for s_ in range(STEPS):
                t.step_async(*pool[(s_ + j) % 2])
            ex.publish(j, t.model.engine.params)
        rows = ex.gather()
        m32, m16, valid, _ = ex.reduce(rows)
        for t in trainers:
            e = t.model.engine
            e.params[:n].copy_(m16)
            e.repack_trainable()
            e.reset_optimizer_state()
    one_round(0)                                   # captures the 4 CUDA graphs
    rows = ex.gather()
    rows_cpu = [r.clone().cpu() for r in rows]
    m32 = ex.reduce(rows)[0]
    ok = torch.equal(m32.cpu(), fedavg_oracle(rows_cpu)[0])
    flag = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    t_round = timed(one_round, 2) / 2
    def ex_only(i):
        ex.reduce(ex.gather())
    t_ex = timed(ex_only, 5) / 5
    return {"round_s": t_round, "clients": KL * world, "clients_per_gpu": KL, "classes": C, "batch": Bc,
            "steps_per_client": STEPS, "images_per_s": world * KL * STEPS * Bc / t_round,
            "exchange_ms": t_ex * 1e3, "transport": ex.transport, "fedavg_bitexact": bool(int(flag.item())),
            "note": "co-located clients run one after another on shared frozen weights / workspace (device-resident "
                    "batches, CUDA-graph steps); timed with CUDA events, max over ranks"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="skip the config-4-shaped federated round (extra key)")
    ap.add_argument("--fed-transport", default="auto", choices=["auto", "p2p", "p2p_sharded", "nccl"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the MaPLe hot path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    run_ours(args)


if __name__ == "__main__":
    main()
