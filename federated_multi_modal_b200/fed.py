"""Round-end exchange for FedAvg over one 8xB200 NVLink/NVSwitch box.

One process per GPU. Clients are assigned to ranks in contiguous blocks (rank r owns clients
[r*k, (r+1)*k)), so rank-major gather order == client order and the fixed-order reduction
(`mfk_fedavg_reduce`) gives bit-identical results on every rank, for any number of GPUs.

Transports for the client tensors (flat fp32 arenas of the trainable parameters):
  "p2p"  — symmetric-memory buffers: every rank's reduce kernel loads the peers' rows straight over
           NVLink (peer pointers in the kernel's pointer table), i.e. transfer and weighted reduction are
           ONE kernel; a symmetric-memory barrier on each side orders it against the producers.
  "p2p_sharded" — the all-reduce-shaped form of the same thing (the default on NCCL since round 2; cfg.FED.TRANSPORT
           selects another, cfg.FED.STRICT_TRANSPORT forbids the collective downgrade): rank r reduces only elements
           [r*n/W, (r+1)*n/W) of the K rows (same per-element order => bit-identical) and pushes the fp32 / fp16 result
           into every rank's output buffer with peer stores (`mfk_fedavg_reduce_scatter`): (W-1)/W * 10 bytes per
           element over NVLink instead of (W-1) * 4.
  "nccl" — torch.distributed all_gather_into_tensor, then the same reduce kernel on local rows
           (also the path used with the gloo backend in CPU tests of the host logic).
No ring/tree all-reduce is used: its summation order differs from the reference's (not bit-exact).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def dist_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def clients_of_rank(num_clients: int, rank: int, world: int) -> List[int]:
    """Contiguous block assignment; requires num_clients % world == 0 (BASELINE configs 3/4: 8/8, 32/{2,4,8})."""
    if num_clients % world != 0:
        raise ValueError(f"num_clients={num_clients} must be a multiple of world size {world}")
    k = num_clients // world
    return list(range(rank * k, (rank + 1) * k))


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Elements [lo, hi) reduced by `rank` in the sharded exchange: equal shards rounded up to a multiple of 4 (the
    kernel works on float4 groups), the last ranks' shards clipped at n (possibly empty). The shards of all ranks
    tile [0, n) exactly."""
    per = ((n + world - 1) // world + 3) // 4 * 4
    return min(n, rank * per), min(n, (rank + 1) * per)


class FedAvgExchange:
    def __init__(self, n: int, k_local: int, device, transport: str = "auto", strict_transport: bool = False):
        """``strict_transport``: raise instead of downgrading to the NCCL all-gather when the requested NVLink
        transport cannot be set up (bench.py and the trainer's cfg.FED.STRICT_TRANSPORT use it)."""
        self.rank, self.world = dist_info()
        self.n, self.k_local, self.K = n, k_local, k_local * self.world
        self.dev = torch.device(device)
        self.transport = transport
        self._symm = None
        cuda = self.dev.type == "cuda"
        if transport == "auto":
            # measured on 8 x B200 (profiles/r02_exchange_n8_v18.json, 13.86 M fp32 per client): sharded 0.55 ms,
            # p2p 1.01 ms, nccl all-gather 1.23 ms at 1 client per GPU; 1.33 / 3.48 / 3.97 ms at 4 clients per GPU
            self.transport = "p2p_sharded" if (cuda and self.world > 1 and dist.get_backend() == "nccl") else "nccl"
        self._sharded = None
        if self.transport == "p2p_sharded" and self.world == 1:
            self.transport = "p2p"
        # rows start on 16-byte boundaries whatever n is (the reduce kernels vectorise by 4 elements)
        self.n_pad = n_pad = (n + 3) // 4 * 4
        self.requested_transport = self.transport
        if self.transport in ("p2p", "p2p_sharded") and self.world > 1:
            err = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._send_flat = symm_mem.empty(k_local * n_pad, device=self.dev, dtype=torch.float32)
                self._symm = symm_mem.rendezvous(self._send_flat, dist.group.WORLD)
                self.send = self._send_flat.view(k_local, n_pad)[:, :n]
                if self.transport == "p2p_sharded":
                    # outputs live in symmetric memory too: every rank's kernel stores its shard into all of them
                    o32 = symm_mem.empty(n_pad, device=self.dev, dtype=torch.float32)
                    h32 = symm_mem.rendezvous(o32, dist.group.WORLD)
                    o16 = symm_mem.empty(n_pad, device=self.dev, dtype=torch.float16)
                    h16 = symm_mem.rendezvous(o16, dist.group.WORLD)
                    p32 = [h32.get_buffer(r, (n_pad,), torch.float32) for r in range(self.world)]
                    p16 = [h16.get_buffer(r, (n_pad,), torch.float16) for r in range(self.world)]
                    lo, hi = shard_range(n, self.rank, self.world)
                    self._sharded = dict(
                        out32=o32[:n], out16=o16[:n], keep=(h32, h16, p32, p16),
                        p32=torch.tensor([t.data_ptr() for t in p32], dtype=torch.int64, device=self.dev),
                        p16=torch.tensor([t.data_ptr() for t in p16], dtype=torch.int64, device=self.dev),
                        lo=lo, hi=hi)
            except Exception as e:  # noqa: BLE001 — decided collectively below
                err = e
            # The transport is chosen by ALL ranks together: a rank that falls back on its own would sit in an NCCL
            # all_gather while its peers wait in the symmetric-memory barrier (deadlock at the first round end).
            # A rank whose rendezvous raised while others are still inside it cannot be helped here — symmetric-memory
            # rendezvous is itself collective and fails on every rank or none in practice.
            ok = torch.tensor([0 if err is not None else 1], device=self.dev, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if strict_transport:
                    raise RuntimeError(f"FedAvgExchange: transport '{self.transport}' unavailable on at least one rank "
                                       f"(this rank: {type(err).__name__ if err else 'ok'}: {err}); refusing the silent "
                                       "downgrade to nccl all_gather (strict_transport=True)")
                print(f"[fed] symmetric memory unavailable on at least one rank ({type(err).__name__ if err else 'peer'}"
                      f": {err}); all ranks use nccl all_gather")
                self.transport, self._symm, self._sharded = "nccl", None, None
        if self._symm is None:
            self._send_flat = torch.zeros(k_local * n_pad, device=self.dev, dtype=torch.float32)
            self.send = self._send_flat.view(k_local, n_pad)[:, :n]
        self._gathered_flat = None if self._symm is not None else torch.zeros(self.K * n_pad, device=self.dev,
                                                                               dtype=torch.float32)
        self.gathered = None if self._symm is not None else self._gathered_flat.view(self.K, n_pad)[:, :n]
        # per client: [ok, n_samples, NaN/Inf flag word of its published tensor (1 = NaN, 2 = Inf)]
        self.status_local = torch.zeros(k_local, 3, device=self.dev, dtype=torch.float32)
        self.status = torch.zeros(self.K, 3, device=self.dev, dtype=torch.float32)
        self.flags_local = torch.zeros(k_local, device=self.dev, dtype=torch.int32)
        if self._sharded is not None:
            self.out32, self.out16 = self._sharded["out32"], self._sharded["out16"]
        elif cuda:
            self.out32 = torch.zeros(n, device=self.dev, dtype=torch.float32)
            self.out16 = torch.zeros(n, device=self.dev, dtype=torch.float16)

    # -------------------------------------------------------------------- stage 1: publish + gather
    def publish(self, j: int, arena: torch.Tensor, ok: bool = True, n_samples: float = 1.0):
        self.send[j].copy_(arena[: self.n])
        self.status_local[j, 0] = 1.0 if ok else 0.0
        self.status_local[j, 1] = float(n_samples)

    def _scan_local(self):
        """check_weights_valid (trainers/maple_fed.py:317-325) of this rank's OWN clients, on the device and out of
        local HBM; the flag words travel with the status all-gather, so no rank ever scans a peer's tensor."""
        self.flags_local.zero_()
        if self.dev.type == "cuda":
            from . import ops
            for j in range(self.k_local):
                ops.check_finite(self.send[j], self.flags_local[j:j + 1])
        else:  # host logic under the gloo tests
            for j in range(self.k_local):
                self.flags_local[j] = int(torch.isnan(self.send[j]).any()) | (int(torch.isinf(self.send[j]).any()) << 1)
        self.status_local[:, 2] = self.flags_local.to(torch.float32)

    def gather(self) -> List[torch.Tensor]:
        """Returns the K client rows in client order (views; peer memory for the p2p transport)."""
        self._reduced_since_gather = False
        self._scan_local()
        if self.world == 1:
            self.status.copy_(self.status_local)
            return [self.send[j] for j in range(self.k_local)]
        dist.all_gather_into_tensor(self.status.view(-1), self.status_local.view(-1))
        if self._symm is not None:
            self._symm.barrier()
            rows = []
            for r in range(self.world):
                peer = self._symm.get_buffer(r, (self.k_local, self.n_pad), torch.float32)
                rows += [peer[j, : self.n] for j in range(self.k_local)]
            return rows
        dist.all_gather_into_tensor(self._gathered_flat, self._send_flat)
        return [self.gathered[k] for k in range(self.K)]

    # -------------------------------------------------------------------- stage 2: fixed-order reduce (CUDA)
    def reduce(self, rows: Sequence[torch.Tensor], weighted: bool = False):
        """-> (mean fp32 [n], mean fp16 [n], valid client ids, per-client NaN/Inf flags[K]). Invalid clients
        (failed locally, or NaN/Inf in their tensors — check_weights_valid, trainers/maple_fed.py:271-277)
        are excluded; the divisor is the number (or sample count) of the valid ones.
        The two means are the exchange's own output buffers, which the sharded transport fills by PEER stores: they stay
        valid until this rank calls `gather()` or `reduce` again. A round of the trainer is ordered by the next `gather()`
        (status all-gather: every rank's earlier stream work, including its reads of the means, is complete before any
        rank gets past it); a second `reduce` on the same rows starts with a symmetric-memory barrier for the same reason."""
        from . import ops
        status = self.status.cpu()  # the one host synchronisation of the exchange
        bad = status[:, 2].to(torch.int32)  # validity scans ran on the owners' GPUs (gather)
        valid = [k for k in range(self.K) if status[k, 0] > 0 and int(bad[k]) == 0]
        if not valid:
            return None, None, valid, bad
        ptrs = torch.tensor([rows[k].data_ptr() for k in valid], dtype=torch.int64, device=self.dev)
        if weighted:
            w = torch.tensor([float(status[k, 1]) for k in valid], dtype=torch.float32, device=self.dev)
            div = float(sum(float(status[k, 1]) for k in valid))
        else:
            w, div = None, float(len(valid))
        if self._sharded is not None:
            sh = self._sharded
            if getattr(self, "_reduced_since_gather", False):
                # a second reduce on the same gathered rows: no status all-gather has ordered the peers' reads of the
                # previous result (device copies or synchronous host copies made before this call) against the peer
                # stores of this one, so every rank waits here until all of them have arrived
                self._symm.barrier()
            self._reduced_since_gather = True
            ops.fedavg_reduce_scatter(ptrs, w, div, len(valid), self.n, sh["lo"], sh["hi"], sh["p32"], sh["p16"],
                                      self.world)
        else:
            ops.fedavg_reduce(ptrs, w, div, len(valid), self.n, False, self.out32, self.out16, None)
        if self._symm is not None:
            self._symm.barrier()  # peers may overwrite their send buffers only after everyone has read them
        return self.out32, self.out16, valid, bad
