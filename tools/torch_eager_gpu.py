"""Context number for the headline: the reference's OWN way of running this step on a GPU is stock PyTorch eager
(fp16 CLIP, nn.MultiheadAttention, autograd, clip_grad_norm_, torch.optim.SGD — trainers/maple.py:304-381, 547-627;
clip/model.py:269-352, 478-572). The reference tree does not travel to the GPU box, so this file restates that
module structure in plain torch (test/bench infrastructure, not product code, never imported by the package) and
times it on the same synthetic batch and the same trainable set as bench.py: "reference-on-GPU, torch eager"
(SURVEY.md §8d). Prints one JSON line.

    python tools/torch_eager_gpu.py [--batch 32] [--classes 10] [--steps 20] [--warmup 5] [--dtype fp16|bf16]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn as nn
import torch.nn.functional as F
from federated_multi_modal_b200 import synth
from helpers import customclip_state_dict


class LayerNorm(nn.LayerNorm):  # clip/model.py:153-159 (fp32 inside)
    def forward(self, x):
        return super().forward(x.float()).to(x.dtype)


class QuickGELU(nn.Module):  # clip/model.py:162-164
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class Block(nn.Module):  # clip/model.py:269-352 (MaPLe variant)
    def __init__(self, d, heads, mask, text, i, n_ctx):
        super().__init__()
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_1, self.ln_2 = LayerNorm(d), LayerNorm(d)
        self.mlp = nn.Sequential()
        self.mlp.c_fc, self.mlp.gelu, self.mlp.c_proj = nn.Linear(d, 4 * d), QuickGELU(), nn.Linear(4 * d, d)
        self.mask, self.text, self.i, self.n = mask, text, i, n_ctx

    def forward(self, inp):
        x, deep, counter = inp
        if self.i != 0 and counter < len(deep):  # prompt splice by torch.cat (clip/model.py:320-349)
            ctx = deep[counter].expand(x.shape[1], -1, -1).permute(1, 0, 2).half().to(x.dtype)
            if self.text:
                x = torch.cat([x[:1], ctx, x[1 + self.n:]], dim=0)
            else:
                x = torch.cat([x[:x.shape[0] - self.n], ctx], dim=0)
            counter += 1
        h = self.ln_1(x)
        m = self.mask.to(dtype=x.dtype, device=x.device) if self.mask is not None else None  # clip/model.py:303-304
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=m)[0]
        x = x + self.mlp(self.ln_2(x))
        return [x, deep, counter]


class EagerMaPLe(nn.Module):
    def __init__(self, sd, tok, n_ctx, depth, dtype):
        super().__init__()
        self.n, self.J, self.dt = n_ctx, depth, dtype
        self.register_buffer("tok", tok)
        mask = torch.full((77, 77), float("-inf")).triu_(1)
        self.vblocks = nn.Sequential(*[Block(768, 12, None, False, i, n_ctx) for i in range(12)])
        self.tblocks = nn.Sequential(*[Block(512, 8, mask, True, i, n_ctx) for i in range(12)])
        self.conv1 = nn.Conv2d(3, 768, 16, 16, bias=False)
        self.ln_pre, self.ln_post, self.ln_final = LayerNorm(768), LayerNorm(768), LayerNorm(512)
        g = lambda k: sd[k].clone()
        self.cls, self.vpos, self.vproj = (nn.Parameter(g("image_encoder." + k)) for k in
                                           ("class_embedding", "positional_embedding", "proj"))
        self.tpos, self.tproj = nn.Parameter(g("text_encoder.positional_embedding")), nn.Parameter(g("text_encoder.text_projection"))
        self.logit_scale = nn.Parameter(g("logit_scale"))
        self.ctx = nn.Parameter(g("prompt_learner.ctx"))
        self.register_buffer("prefix", g("prompt_learner.token_prefix"))
        self.register_buffer("suffix", g("prompt_learner.token_suffix"))
        self.l2v = nn.Linear(512, 768)
        self.text_deep = nn.ParameterList([nn.Parameter(g(f"prompt_learner.compound_prompts_text_parameters.{i}")) for i in range((depth - 1 + 1) // 2)])
        self.vis_deep = nn.ParameterList([nn.Parameter(g(f"prompt_learner.visual_deep_prompts_parameters.{i}")) for i in range((depth - 1) // 2)])
        self.projs = nn.ModuleList([nn.Linear(512, 768) if i % 2 == 0 else nn.Linear(768, 512) for i in range(depth - 1)])
        # load the synthetic weights
        own = {}
        for tower, blocks in (("image_encoder", self.vblocks), ("text_encoder", self.tblocks)):
            for i, b in enumerate(blocks):
                pre = f"{tower}.transformer.resblocks.{i}."
                own.update({f"{'v' if tower[0] == 'i' else 't'}blocks.{i}.{k[len(pre):]}": v for k, v in sd.items() if k.startswith(pre)})
        own["conv1.weight"] = sd["image_encoder.conv1.weight"]
        for a, b in (("ln_pre", "image_encoder.ln_pre"), ("ln_post", "image_encoder.ln_post"), ("ln_final", "text_encoder.ln_final")):
            own[a + ".weight"], own[a + ".bias"] = sd[b + ".weight"], sd[b + ".bias"]
        own["l2v.weight"], own["l2v.bias"] = sd["prompt_learner.proj_lang_to_vis.weight"], sd["prompt_learner.proj_lang_to_vis.bias"]
        for i in range(depth - 1):
            own[f"projs.{i}.weight"] = sd[f"prompt_learner.compound_prompt_projections.{i}.weight"]
            own[f"projs.{i}.bias"] = sd[f"prompt_learner.compound_prompt_projections.{i}.bias"]
        missing, unexpected = self.load_state_dict(own, strict=False)
        assert not unexpected, unexpected
        # convert_weights (clip/model.py:726-747): Conv/Linear/MHA/projections to the model dtype, LN stays fp32
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear, nn.MultiheadAttention)):
                m.to(dtype)
        for p_ in (self.vproj, self.tproj, self.ctx):
            p_.data = p_.data.to(dtype)
        # freeze policy (trainers/maple.py:447-479)
        for name, p_ in self.named_parameters():
            train = ("ln_" in name or name.startswith(("ctx", "l2v", "text_deep", "vis_deep", "projs"))
                     or "blocks.11." in name)
            p_.requires_grad_(train)

    def forward(self, img, label):
        C = self.prefix.shape[0]
        prompts = torch.cat([self.prefix.to(self.dt), self.ctx.unsqueeze(0).expand(C, -1, -1), self.suffix.to(self.dt)], 1)
        deep_t, deep_v = [], []
        for i in range(self.J - 1):
            if i % 2 == 0:
                deep_t.append(self.text_deep[i // 2]); deep_v.append(self.projs[i](self.text_deep[i // 2].to(self.dt)))
            else:
                deep_v.append(self.vis_deep[(i - 1) // 2]); deep_t.append(self.projs[i](self.vis_deep[(i - 1) // 2].to(self.dt)))
        shared = self.l2v(self.ctx)
        x = (prompts + self.tpos.to(self.dt)).permute(1, 0, 2)
        x = self.tblocks([x, deep_t, 0])[0].permute(1, 0, 2)
        x = self.ln_final(x).to(self.dt)
        ft = x[torch.arange(C), self.tok.argmax(-1)] @ self.tproj
        v = self.conv1(img.to(self.dt)).flatten(2).permute(0, 2, 1)
        v = torch.cat([self.cls.to(self.dt).expand(v.shape[0], 1, -1), v], 1) + self.vpos.to(self.dt)
        v = torch.cat([v, shared.expand(v.shape[0], -1, -1).half().to(self.dt)], 1)
        v = self.ln_pre(v).permute(1, 0, 2)
        v = self.vblocks([v, deep_v, 0])[0].permute(1, 0, 2)
        fi = self.ln_post(v[:, 0]) @ self.vproj
        fi_n, ft_n = F.normalize(fi, dim=-1, eps=1e-8), F.normalize(ft, dim=-1, eps=1e-8)
        logits = self.logit_scale.exp().clamp(max=100) * fi_n @ ft_n.t()
        loss = F.cross_entropy(logits, label) + 0.5 * (1 - F.cosine_similarity(fi_n, ft_n[label], dim=1).mean())
        if not torch.isfinite(loss):  # trainers/maple.py:375-376 (one host sync, as in the reference)
            raise RuntimeError("NaN/Inf in total loss")
        return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--classes", type=int, default=10)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()
    dev = torch.device(a.device)
    dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
    sd, tok = customclip_state_dict(a.classes)
    model = EagerMaPLe(sd, tok, 2, 9, dt).to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.0035, momentum=0.9, weight_decay=5e-4)
    img, lab = synth.make_batch(a.batch, a.classes, 7)
    img_h, lab_h = (img.pin_memory(), lab.pin_memory()) if dev.type == "cuda" else (img, lab)

    def step():
        x, y = img_h.to(dev, non_blocking=True), lab_h.to(dev, non_blocking=True)
        loss = model(x, y)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss.item()  # trainers/maple.py:615

    for _ in range(a.warmup):
        l = step()
    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    sync()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        l = step()
    sync()
    dt_s = (time.perf_counter() - t0) / a.steps
    print(json.dumps({"impl": "torch-eager restatement of the reference modules (GPU)", "dtype": a.dtype,
                      "metric": "MaPLe ViT-B/16 train images/sec", "value": a.batch / dt_s, "unit": "images/s",
                      "ms_per_step": dt_s * 1e3, "batch": a.batch, "classes": a.classes, "steps": a.steps,
                      "trainable_params": sum(p.numel() for p in params), "loss": l,
                      "note": "77-token text tower, nn.MultiheadAttention, autograd, clip_grad_norm_, torch SGD; "
                              "end to end incl. H2D of the batch and loss.item() per step"}))


if __name__ == "__main__":
    main()
