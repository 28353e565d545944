"""CPU: the drop-in surface (names, signatures, state_dict layout, C ABI symbols, host logic)."""
import inspect
import os
import re

import pytest
import torch

from helpers import load_golden, REPO
from federated_multi_modal_b200 import synth


def _model(C=10):
    from federated_multi_modal_b200.clip import build_model
    from federated_multi_modal_b200.trainers import CustomCLIP
    dd = {"trainer": "MaPLe", "vision_depth": 0, "language_depth": 0, "vision_ctx": 0, "language_ctx": 0,
          "maple_length": 2}
    clip = build_model(synth.random_clip_state_dict(0), dd)
    return CustomCLIP(synth.make_cfg(), synth.synthetic_classnames(C), clip)


def test_state_dict_layout_matches_reference():
    spec = load_golden("state_dict_spec.pt")
    m = _model()
    sd = torch.nn.Module.state_dict(m)
    ours = [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()]
    assert ours == spec["spec"]  # same 634 keys, same order, shapes and dtypes as the reference in fp16 mode


def test_freeze_policy_matches_reference():
    from federated_multi_modal_b200.trainers.maple import MaPLe
    spec = load_golden("state_dict_spec.pt")
    m = _model()
    # the policy of MaPLe.build_model, applied without moving to a device
    for p in m.parameters():
        p.requires_grad_(False)
    for n, p in m.named_parameters():
        if "prompt_learner" in n or "transformer.resblocks.11" in n:
            p.requires_grad_(True)
    for _, mod in m.named_modules():
        if isinstance(mod, torch.nn.LayerNorm):
            for p in mod.parameters():
                p.requires_grad_(True)
    ours = sorted(n for n, p in m.named_parameters() if p.requires_grad)
    assert ours == spec["trainable"] and len(ours) == 147
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 14250496
    src = inspect.getsource(MaPLe.build_model)
    assert "transformer.resblocks.11" in src and "prompt_learner" in src


def test_signatures_match_reference_surface():
    from federated_multi_modal_b200.clip import model as cm
    from federated_multi_modal_b200.trainers import (ClientDataManager, CustomCLIP, MaPLe, MaPLeFederated,
                                                     MultiModalPromptLearner, TextEncoder, partition_dataset_iid)
    def args(f):
        return list(inspect.signature(f).parameters)
    assert args(CustomCLIP.forward) == ["self", "image", "label", "caption", "return_feature"]
    assert args(CustomCLIP.__init__) == ["self", "cfg", "classnames", "clip_model"]
    assert args(MultiModalPromptLearner.__init__) == ["self", "cfg", "classnames", "clip_model"]
    assert args(TextEncoder.forward) == ["self", "prompts", "tokenized_prompts", "compound_prompts_deeper_text"]
    assert args(cm.VisionTransformer_MaPLe.forward) == ["self", "x", "shared_ctx", "compound_deeper_prompts",
                                                        "clip_embeddings"]
    assert args(cm.ResidualAttentionBlock_MaPLe.__init__) == ["self", "d_model", "n_head", "attn_mask",
                                                              "design_details", "text_layer", "i"]
    assert args(cm.build_model) == ["state_dict", "design_details"]
    assert args(ClientDataManager.__init__)[:5] == ["self", "train_x", "val", "test", "cfg"]
    assert args(partition_dataset_iid) == ["dataset", "num_clients"]
    for name in ("forward_backward", "parse_batch_train", "run_epoch", "update_lr", "test", "load_model",
                 "build_model", "check_cfg", "model_inference"):
        assert hasattr(MaPLe, name)
    for name in ("train", "safe_average_weights", "check_weights_valid", "broadcast_weights", "save_model",
                 "load_model", "finalize_training"):
        assert hasattr(MaPLeFederated, name)
    assert args(MaPLeFederated.safe_average_weights) == ["self", "local_dicts", "valid_clients"]


def test_no_cpu_fallback():
    m = _model().train()
    img, lab = synth.make_batch(2, 10)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(img, lab)
    from federated_multi_modal_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(32, 8, dtype=torch.bfloat16),
                 out_f32=torch.zeros(8, 32))


def test_cabi_library_exports_every_declared_symbol():
    from federated_multi_modal_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(REPO, "include", "mfk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mfk_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mfk_version() >= 100
    assert b"invalid argument" in lib.mfk_error_string(-1)
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    assert lib.mfk_gemm_bf16(None, 0, None, 0, 0, 0, 0, None, 0, None, 0, None, 0, None, 0, None, 0, None, 0, 0,
                             None, 0, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "federated_multi_modal_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("oracle/maple_cpu.py", ""), os.path.join(root, f)


def test_lr_schedule_restates_dassl_constant_warmup_cosine():
    from federated_multi_modal_b200.dassl_compat import ConstantWarmupCosine

    class H:
        lr = 0.0
    h = H()
    s = ConstantWarmupCosine(0.0026, 2, 1, 1e-4, h)
    seq = [h.lr]
    for _ in range(3):
        s.step(); seq.append(h.lr)
    assert seq == pytest.approx([1e-4, 0.0026, 0.0013, 0.0], abs=1e-12)
    # the reference rebuilds the scheduler every round and sets last_epoch = epoch - 1 (maple_fed.py:337-339)
    h = H(); s = ConstantWarmupCosine(0.0026, 2, 1, 1e-4, h); s.last_epoch = 9
    seq = [h.lr]
    for _ in range(2):
        s.step(); seq.append(h.lr)
    assert seq == pytest.approx([1e-4, 0.0013, 0.0], abs=1e-12)


def test_partitioners():
    from types import SimpleNamespace
    from federated_multi_modal_b200.trainers import Datum, partition_dataset_dirichlet, partition_dataset_iid
    from federated_multi_modal_b200.trainers.data_partition import dirichlet_label_split
    items = [Datum(label=i % 7, classname=f"c{i % 7}") for i in range(700)]
    ds = SimpleNamespace(train_x=items, val=[1], test=[2])
    parts = partition_dataset_iid(ds, 3)
    assert [len(p[0]) for p in parts] == [233, 233, 234] and parts[0][1] == [1] and parts[2][2] == [2]
    assert sorted(id(x) for p in parts for x in p[0]) == sorted(id(x) for x in items)
    a = dirichlet_label_split([it.label for it in items], 8, 0.5, seed=0)
    b = dirichlet_label_split([it.label for it in items], 8, 0.5, seed=0)
    assert a == b and sorted(i for p in a for i in p) == list(range(700))
    sizes = [len(p) for p in a]
    assert max(sizes) > 1.5 * min(sizes)  # label skew gives unequal clients
    d = partition_dataset_dirichlet(ds, 8, 0.5, 0)
    assert sum(len(p[0]) for p in d) == 700


def test_rrc_param_sampling_matches_torchvision():
    """sample_rrc_params draws from a torch.Generator exactly what torchvision's RandomResizedCrop.get_params
    draws from the global RNG (same call sequence), so with equal seeds the crop boxes are identical."""
    import torch
    from torchvision.transforms import RandomResizedCrop
    from federated_multi_modal_b200.trainers.client_datamanager import sample_rrc_params
    img = torch.zeros(3, 200, 260)
    g = torch.Generator().manual_seed(5)
    torch.manual_seed(5)
    for _ in range(50):
        want = RandomResizedCrop.get_params(img, scale=[0.08, 1.0], ratio=[3.0 / 4.0, 4.0 / 3.0])
        assert sample_rrc_params(200, 260, g) == tuple(want)


def test_client_datamanager_loader_host_logic():
    """ClientDataManager (reference trainers/client_datamanager.py:15-112): label validation errors, class bookkeeping,
    train loader = shuffled + drop_last with a seeded order, test loader = in order and complete."""
    import pytest
    from federated_multi_modal_b200.trainers import ClientDataManager, Datum
    names = synth.synthetic_classnames(3)
    items = [Datum(impath=f"synthetic://{i}", label=i % 3, classname=names[i % 3], img=torch.full((3, 4, 4), float(i)))
             for i in range(10)]
    cfg = synth.make_cfg()
    cfg.DATALOADER = synth._NS(TRAIN_X=synth._NS(BATCH_SIZE=4), TEST=synth._NS(BATCH_SIZE=4))
    dm = ClientDataManager(items, [], items[:6], cfg)
    assert dm.num_classes == 3 and dm.lab2cname == {0: names[0], 1: names[1], 2: names[2]}
    assert len(dm.train_loader) == 2 and len(dm.test_loader) == 2          # 10 // 4 (drop_last) and ceil(6 / 4)
    ep = [[int(b["img"][j, 0, 0, 0]) for j in range(b["img"].shape[0])] for b in dm.train_loader]
    assert [len(b) for b in ep] == [4, 4] and len({i for b in ep for i in b}) == 8
    dm2 = ClientDataManager(items, [], items[:6], cfg)
    assert ep == [[int(b["img"][j, 0, 0, 0]) for j in range(4)] for b in dm2.train_loader]   # seeded order
    te = [b["label"].tolist() for b in dm.test_loader]
    assert te == [[0, 1, 2, 0], [1, 2]]
    for b in dm.test_loader:
        assert b["img"].dtype == torch.float32 and b["label"].dtype == torch.int64
    class StrLabel:
        classname, label = "a", "0"
    with pytest.raises(TypeError):
        ClientDataManager([StrLabel()], [], [], cfg)
    class NoLabel:
        classname = "a"
    with pytest.raises(ValueError):
        ClientDataManager([NoLabel()], [], [], cfg)


def test_fedavg_exchange_host_side_validity_flags():
    """Single-process exchange on CPU tensors: publish -> gather carries [ok, n_samples, NaN/Inf flag] per client, the
    flag word follows check_weights_valid (trainers/maple_fed.py:317-325): bit 0 = NaN, bit 1 = Inf."""
    from federated_multi_modal_b200.fed import FedAvgExchange
    ex = FedAvgExchange(8, 3, "cpu")
    good = torch.arange(8, dtype=torch.float32)
    nan = good.clone(); nan[2] = float("nan")
    both = good.clone(); both[1] = float("inf"); both[5] = float("nan")
    ex.publish(0, good, ok=True, n_samples=5)
    ex.publish(1, nan, ok=True, n_samples=6)
    ex.publish(2, both, ok=False, n_samples=7)
    rows = ex.gather()
    assert len(rows) == 3 and torch.equal(rows[0], good)
    assert ex.status.tolist() == [[1.0, 5.0, 0.0], [1.0, 6.0, 1.0], [0.0, 7.0, 3.0]]


def test_checkpoint_wire_format_matches_reference(tmp_path):
    """MaPLeFederated.save_model (reference trainers/maple_fed.py:367-386): directory layout, top-level keys and the
    key / shape / dtype list of the saved state_dict equal what the unmodified reference hands to Dassl's
    save_checkpoint after one aggregation (tests/golden/ckpt_spec.pt, recorded from the reference); values that are
    determined by the seeds (ctx from the token embedding, logit_scale) are equal too. File round trip through
    Dassl's ``model.pth.tar-<epoch>`` naming and MaPLe.load_model's key filtering are checked on the host."""
    import types
    from federated_multi_modal_b200.dassl_compat import load_checkpoint, save_checkpoint
    from federated_multi_modal_b200.trainers import MaPLe, MaPLeFederated
    spec = load_golden("ckpt_spec.pt")
    m = _model()
    fed = MaPLeFederated.__new__(MaPLeFederated)
    fed.cfg = types.SimpleNamespace(OUTPUT_DIR=str(tmp_path), VERBOSE=False, OPTIM=types.SimpleNamespace(MAX_EPOCH=2),
                                    dump=lambda: "cfg-dump", FED=types.SimpleNamespace(REFERENCE_FP16_CAST=True))
    fed.clients = [types.SimpleNamespace(model=m)]
    fed.nan_stats = {"total_updates": 0, "failed_clients": [], "skipped_rounds": 0}
    before = fed._global_state_dict()
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in before.items()] == spec["spec_before_aggregation"]
    fed.nan_stats["total_updates"] = 1          # one safe_average_weights happened: every tensor is fp16 from here on
    fed.global_weights = fed._global_state_dict()
    path = fed.save_model(directory=str(tmp_path))
    want_dir = os.path.join(str(tmp_path), os.path.basename(spec["target_dir"]))
    assert os.path.basename(spec["target_dir"]) == "MultiModalPromptLearner_Aggregator"
    assert path == os.path.join(want_dir, f"model.pth.tar-{spec['epoch']}") and os.path.exists(path)
    assert open(os.path.join(want_dir, "checkpoint")).read().strip() == os.path.basename(path)
    ck = load_checkpoint(path)
    assert list(ck.keys()) == spec["top_keys"]
    assert (ck["epoch"], ck["optimizer"], ck["scheduler"], ck["cfg"]) == (spec["epoch"], None, None, "cfg-dump")
    assert [(k, tuple(v.shape), str(v.dtype)) for k, v in ck["state_dict"].items()] == spec["spec_after_aggregation"]
    assert torch.equal(ck["state_dict"]["logit_scale"], spec["logit_scale_after"])
    assert torch.equal(ck["state_dict"]["prompt_learner.ctx"], spec["ctx_after"])
    # a reference-written checkpoint (all fp16) loads into the module with the reference's strict=True semantics,
    # restoring the module's own dtypes
    m2 = _model()
    with torch.no_grad():
        m2.prompt_learner.ctx.add_(1.0)
    m2.load_state_dict(ck["state_dict"], strict=True)
    sd2 = torch.nn.Module.state_dict(m2)
    assert [(k, str(v.dtype)) for k, v in sd2.items()] == [(k, d) for k, _, d in spec["spec_before_aggregation"]]
    assert torch.equal(sd2["prompt_learner.ctx"], ck["state_dict"]["prompt_learner.ctx"])
    # MaPLe.load_model (trainers/maple.py:683-716): per-model sub-directory, token_prefix / token_suffix dropped,
    # non-strict load; missing file -> FileNotFoundError
    t = MaPLe.__new__(MaPLe)
    m3 = _model(C=4)                            # other class list: the buffers must NOT be taken from the checkpoint
    t._models = {"MultiModalPromptLearner_0": m3}
    t.get_model_names = lambda names=None: list(t._models)
    sd = {k: v for k, v in ck["state_dict"].items()}
    save_checkpoint({"epoch": 2, "state_dict": sd}, os.path.join(str(tmp_path), "MultiModalPromptLearner_0"))
    prefix_before = m3.prompt_learner.token_prefix.clone()
    with torch.no_grad():
        m3.prompt_learner.ctx.zero_()
    t.load_model(str(tmp_path), epoch=2)
    assert torch.equal(m3.prompt_learner.ctx, ck["state_dict"]["prompt_learner.ctx"])
    assert torch.equal(m3.prompt_learner.token_prefix, prefix_before) and prefix_before.shape[0] == 4
    with pytest.raises(FileNotFoundError):
        t.load_model(str(tmp_path), epoch=7)
    with pytest.raises(FileNotFoundError):
        fed.load_model(str(tmp_path), epoch=9)
