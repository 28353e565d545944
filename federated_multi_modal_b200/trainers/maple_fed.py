"""Drop-in mirror of ``trainers/maple_fed.py`` (MaPLeFederated) — one process per GPU.

Reference behaviour kept (file:line in /root/reference/trainers/maple_fed.py):
  round loop, failure bookkeeping ........ train(), 228-303 / nan_stats 38-42
  safe_average_weights(local_dicts, valid)  309-315  per key: fp32 -> nan_to_num -> mean over clients -> .half()
  check_weights_valid(state_dict) ......... 317-325  NaN / Inf scan
  broadcast_weights(global_sd) ............ 327-339  load_state_dict(strict) + optimiser-state reset + scheduler rebuild
  save_model / load_model ................. 367-411  Dassl ``model.pth.tar-<epoch>`` under MultiModalPromptLearner_Aggregator/

What changes: clients are sharded over the ranks of ``torch.distributed`` (contiguous blocks), co-located
clients share one frozen CLIP copy and one activation workspace, and the round-end average moves only the
trainable tensors (one flat fp32 arena per client; the frozen tensors are identical on every client, so their
average is the identity) through ``fed.FedAvgExchange`` + the fixed-order ``mfk_fedavg_reduce`` kernel.
The reference's ``.half()`` of the averaged tensors is reproduced (``cfg.FED.REFERENCE_FP16_CAST``, default
on), including its one-time fp16 rounding of the fp32 frozen tensors after the first aggregation.
"""
from __future__ import annotations

import os.path as osp
import time
from typing import Dict, List, Optional

import torch

from .. import ops
from ..dassl_compat import (TRAINER_REGISTRY, TrainerX, build_lr_scheduler, load_checkpoint, mkdir_if_missing,
                            save_checkpoint)
from ..fed import FedAvgExchange, clients_of_rank, dist_info
from .maple import MaPLe

F32 = torch.float32


def _fed(cfg, name, default):
    return getattr(getattr(cfg, "FED", None), name, default)


@TRAINER_REGISTRY.register()
class MaPLeFederated(TrainerX):
    def __init__(self, cfg, client_data_managers=None, classnames=None):
        self.lab2cname = {}
        self.cfg = cfg
        self.num_clients = cfg.FED.NUM_CLIENTS
        self.num_rounds = cfg.FED.NUM_ROUNDS
        self.local_epochs = cfg.FED.LOCAL_EPOCHS
        self.clients: List[MaPLe] = []
        self.global_weights = None
        self.nan_stats = {"total_updates": 0, "failed_clients": [], "skipped_rounds": 0}
        self.round_times: List[Dict[str, float]] = []
        self._injected = (client_data_managers, classnames)
        super().__init__(cfg)

    # ------------------------------------------------------------------ A) data
    def build_data_loader(self):
        dms, classnames = self._injected
        if dms is None:
            raise NotImplementedError(
                "MaPLeFederated: pass client_data_managers=[ClientDataManager,...] and classnames=[...]. The "
                "reference's PatternNet/UCMerced/EuroSAT readers and label-space union (trainers/maple_fed.py:48-159) "
                "are dataset I/O outside the hot path (SURVEY.md §2 #8).")
        if len(dms) != self.num_clients:
            raise ValueError(f"expected {self.num_clients} client data managers, got {len(dms)}")
        self.client_data_managers = dms
        self.global_classnames = list(classnames)
        self.lab2cname = {i: n for i, n in enumerate(self.global_classnames)}
        self.num_classes = len(self.global_classnames)
        self.dm = dms[0]

    # ------------------------------------------------------------------ B) local trainers
    def build_model(self):
        self.rank, self.world = dist_info()
        self.local_ids = clients_of_rank(self.num_clients, self.rank, self.world)
        self.clients = []
        share = None
        for cid in self.local_ids:
            t = MaPLe.__new__(MaPLe)
            t.dm = self.client_data_managers[cid]
            MaPLe.__init__(t, self.cfg, client_id=cid, classnames=self.global_classnames, share_engine=share)
            t.dm = self.client_data_managers[cid]
            share = share or t.model.engine
            self.clients.append(t)
        eng = self.clients[0].model.engine
        self.exchange = FedAvgExchange(eng.n_update, len(self.clients), self.device,
                                       transport=_fed(self.cfg, "TRANSPORT", "auto"),
                                       strict_transport=bool(_fed(self.cfg, "STRICT_TRANSPORT", False)))
        self.global_arena = eng.params[: eng.n_update].clone()
        self.global_weights = self.clients[0].model.state_dict()
        self._frozen_rounded = False

    # ------------------------------------------------------------------ C) rounds
    def train(self):
        for round_idx in range(self.num_rounds):
            t0 = time.perf_counter()
            print(f"\n--- Federated Round {round_idx + 1}/{self.num_rounds} ---")
            if not self._arena_valid(self.global_arena):
                print("Invalid global weights detected! Skipping round.")
                self.nan_stats["skipped_rounds"] += 1
                continue
            self._broadcast_arena(self.global_arena)
            round_losses = []
            for j, trainer in enumerate(self.clients):
                trainer.epoch = round_idx * self.local_epochs
                trainer.max_epoch = (round_idx + 1) * self.local_epochs
                ok, last = True, 0.0
                try:
                    for ep in range(trainer.epoch, trainer.max_epoch):
                        last = trainer.run_epoch(ep).get("avg_loss", 0.0)
                # the reference catches RuntimeError only (trainers/maple_fed.py:262-265): "NaN/Inf in total loss" drops
                # the client for the round, while the ValueError of a NaN / Inf INPUT (check_tensor_validity,
                # trainers/maple.py:525-535) escapes and ends the run; cfg.FED.DROP_ON_INPUT_ERROR also drops those
                except ((RuntimeError, ValueError) if _fed(self.cfg, "DROP_ON_INPUT_ERROR", False) else RuntimeError) as e:
                    print(f"Client {trainer.client_id} failed training: {e}")
                    self.nan_stats["failed_clients"].append(trainer.client_id)
                    ok = False
                if ok:
                    round_losses.append(last)
                eng = trainer.model.engine
                n_samples = len(trainer.dm.train_x_list) if hasattr(trainer.dm, "train_x_list") else 1
                self.exchange.publish(j, eng.params, ok=ok, n_samples=n_samples)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            if round_losses:
                print(f"[Round {round_idx + 1}] Avg local training loss = {sum(round_losses) / len(round_losses):.4f}")
            valid = self._aggregate()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if valid:
                self.nan_stats["total_updates"] += 1
            else:
                print("All clients failed! Reverting to previous global model.")
                self.nan_stats["skipped_rounds"] += 1
            if self._arena_valid(self.global_arena):
                self._broadcast_arena(self.global_arena)
                if self.rank == 0 and getattr(self.clients[0].dm, "test_loader", None) is not None:
                    res = self.clients[0].test()
                    print(f"[Round {round_idx + 1}] Test accuracy (client 0) = {res.get('accuracy', 0):.2f}%")
            else:
                print("Global weights invalid after aggregation, skipping test.")
            torch.cuda.synchronize()
            self.round_times.append({"local_s": t1 - t0, "fedavg_s": t2 - t1, "round_s": time.perf_counter() - t0})
        self.finalize_training()

    def _aggregate(self) -> List[int]:
        """Round-end FedAvg of the trainable arena over all clients of all ranks (fixed order)."""
        rows = self.exchange.gather()
        mean32, mean16, valid, bad = self.exchange.reduce(rows, weighted=bool(_fed(self.cfg, "WEIGHTED", False)))
        for k in range(self.exchange.K):
            if int(bad[k]):
                print(f"Client {k} produced invalid weights, skipping aggregation")
        if not valid:
            return valid
        self.last_mean_fp32 = mean32
        if _fed(self.cfg, "REFERENCE_FP16_CAST", True):
            self.global_arena.copy_(mean16)  # `.half()` of every averaged tensor (trainers/maple_fed.py:314)
            if not self._frozen_rounded:
                for t in self.clients:
                    t.model.engine.round_frozen_to_fp16()
                self._frozen_rounded = True
        else:
            self.global_arena.copy_(mean32)
        return valid

    def _arena_valid(self, arena) -> bool:
        flag = torch.zeros(1, device=arena.device, dtype=torch.int32)
        ops.check_finite(arena, flag)
        return int(flag) == 0

    def _broadcast_arena(self, arena):
        """broadcast_weights on the trainable arena: copy into every local client, drop optimiser state, rebuild
        the LR scheduler with last_epoch = epoch - 1 (trainers/maple_fed.py:327-339)."""
        for t in self.clients:
            eng = t.model.engine
            eng.params[: eng.n_update].copy_(arena)
            eng.repack_trainable()
            eng.reset_optimizer_state()
            t.model._arena_newer = True
            t.sched = build_lr_scheduler(t.optim, t.cfg.OPTIM)
            if hasattr(t, "epoch"):
                t.sched.last_epoch = t.epoch - 1

    # ------------------------------------------------------------------ D) reference-signature utilities
    def safe_average_weights(self, local_dicts, valid_clients):
        """Drop-in for trainers/maple_fed.py:309-315 on whole state_dicts (any mix of fp16/fp32 CUDA tensors):
        one fixed-order reduction kernel per key, fp16 output for every key like the reference."""
        avg_state = {}
        K = len(local_dicts)
        for key in local_dicts[0].keys():
            ts = [sd[key] for sd in local_dicts]
            t0 = ts[0]
            if not t0.is_cuda:
                raise RuntimeError("safe_average_weights: state_dict tensors must be on the CUDA device")
            if t0.dtype not in (torch.float16, torch.float32):
                ts = [t.float() for t in ts]
            ts = [t.contiguous() for t in ts]
            n = ts[0].numel()
            out = torch.empty(ts[0].shape, device=t0.device, dtype=torch.float16)
            ptrs = torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=t0.device)
            ops.fedavg_reduce(ptrs, None, float(K), K, n, ts[0].dtype == torch.float16, None, out, None)
            self._keepalive = ts
            avg_state[key] = out
        return avg_state

    def check_weights_valid(self, state_dict):
        dev = self.device
        flag = torch.zeros(1, device=dev, dtype=torch.int32)
        for name, p in state_dict.items():
            if p.is_floating_point() and p.numel() > 0:
                if not p.is_cuda:
                    raise RuntimeError("check_weights_valid: tensors must be on the CUDA device")
                t = p if p.dtype in (torch.float16, torch.float32, torch.bfloat16) else p.float()
                ops.check_finite(t.contiguous(), flag)
        return int(flag) == 0  # one host sync for the whole dict (the reference does 2 per key)

    def broadcast_weights(self, global_sd):
        for t in self.clients:
            t.model.load_state_dict(global_sd, strict=True)
            t.model.engine.reset_optimizer_state()
            t.sched = build_lr_scheduler(t.optim, t.cfg.OPTIM)
            if hasattr(t, "epoch"):
                t.sched.last_epoch = t.epoch - 1
        eng = self.clients[0].model.engine
        self.global_arena = eng.params[: eng.n_update].clone()

    def _global_state_dict(self):
        """The aggregator's ``global_weights`` in the reference's wire layout: all 634 keys; after at least one
        aggregation EVERY tensor is fp16 because ``safe_average_weights`` ends in ``.half()``
        (trainers/maple_fed.py:314) — before that, the model's own dtypes (fp16 weights, fp32 LN / embeddings)."""
        sd = self.clients[0].model.state_dict()
        if _fed(self.cfg, "REFERENCE_FP16_CAST", True) and self.nan_stats["total_updates"] > 0:
            sd = type(sd)((k, v.half() if v.is_floating_point() else v) for k, v in sd.items())
        return sd

    def finalize_training(self):
        print("\nTraining Summary:")
        print(f"Completed Rounds: {self.nan_stats['total_updates']}")
        print(f"Skipped Rounds: {self.nan_stats['skipped_rounds']}")
        fail_rate = len(self.nan_stats["failed_clients"]) / max(1, self.num_clients)
        print(f"Client failure rate: {fail_rate:.2%}")
        self.global_weights = self._global_state_dict()
        out_dir = getattr(self.cfg, "OUTPUT_DIR", "")
        if out_dir and self.rank == 0:
            self.save_model(directory=out_dir)

    def before_save(self):
        for name in self.get_model_names():
            self._models[name].load_state_dict(self.global_weights)

    def save_model(self, epoch=None, directory="", is_best=False, val_result=None):
        directory = directory or self.cfg.OUTPUT_DIR
        target = osp.join(directory, "MultiModalPromptLearner_Aggregator")
        mkdir_if_missing(target)
        ckpt = {"epoch": self.cfg.OPTIM.MAX_EPOCH, "state_dict": {k: v.cpu() for k, v in self.global_weights.items()},
                "optimizer": None, "scheduler": None, "val_result": val_result,
                "cfg": self.cfg.dump() if hasattr(self.cfg, "dump") else None}
        return save_checkpoint(ckpt, target, is_best=is_best)

    def load_model(self, directory, epoch=None):
        if not directory:
            print("Skipping load_model, no pretrained path given")
            return
        name = f"model.pth.tar-{epoch}" if epoch is not None else "model.pth.tar"
        path = osp.join(directory, "MultiModalPromptLearner_Aggregator", name)
        if not osp.exists(path):
            raise FileNotFoundError(f"Model not found at {path}")
        ckpt = load_checkpoint(path)
        self.global_weights = {k: v.to(self.device) for k, v in ckpt["state_dict"].items()}
        if self.check_weights_valid(self.global_weights):
            self.broadcast_weights(self.global_weights)
        else:
            print("Warning: loaded global weights invalid! Skipping broadcast.")

    def test(self):
        return self.clients[0].test()
