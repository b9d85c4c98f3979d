"""Drop-in for reference models/kernel_shap_bert.py: the KernelSHAP baseline around the BERT classifier.

The reference delegates coalition sampling, the synthetic-data build and the weighted least squares to the
third-party `shap.KernelExplainer(link="logit")` on the CPU (models/kernel_shap_bert.py:170-185).  Here the
classifier forwards run on the sm_100a kernels and the solve is `agb_kernelshap_solve` (batched Gram +
Cholesky in float64 on the device).  As shap does, the regression only covers the features that VARY between
the explained row and the background (the others get attribution 0), and an under-determined system (fewer
distinct coalitions than varying features) gets the minimum-norm least-squares solution.
NOT reproduced — stated loudly: shap's own RNG stream, and its default l1_reg="auto", which pre-selects features
with LassoLarsIC("aic") whenever the sampled coalitions cover < 20 % of the 2^M space (i.e. always for M >= 14):
this path solves the UN-regularised constrained weighted least squares on all varying features.  Parity with
shap is therefore unpinned (shap is absent from the image; oracle/kernelshap.py).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import pydantic
import torch
from torch import Tensor, nn

from .. import ops
from . import _tree
from .vanilla_bert import VanillaBertClassifier, VanillaBertConfig


class KernelShapBertConfig(pydantic.BaseModel):
    """reference models/kernel_shap_bert.py:15-60 (identical fields)"""

    attention_probs_dropout_prob: float
    explainer_attn_num_layers: int
    explainer_head_hidden_size: int
    explainer_normalize: bool
    hidden_dropout_prob: float
    hidden_size: int
    intermediate_size: int
    layer_norm_eps: float
    max_position_embeddings: int
    num_attention_heads: int
    num_hidden_layers: int
    num_labels: int
    pad_token_id: int
    type_vocab_size: int
    vocab_size: int
    kernel_shap_n_samples: int
    kernel_shap_data_size: int

    @property
    def is_decoder(self) -> bool:
        return False

    def into(self) -> VanillaBertConfig:
        d = self.model_dump()
        d.pop("kernel_shap_n_samples")
        d.pop("kernel_shap_data_size")
        return VanillaBertConfig(**d)


class KernelShapBertClassifier(VanillaBertClassifier):
    """reference models/kernel_shap_bert.py:63-74"""

    def __init__(self, config: KernelShapBertConfig):
        super().__init__(config.into())

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "bert")
        _tree.freeze_model_parameters(self, "bert_pooler")
        _tree.freeze_model_parameters(self, "classifier")
        return self


class KernelShapBertSurrogate(KernelShapBertClassifier):
    pass


class KernelShapBertExplainer(nn.Module):
    """reference models/kernel_shap_bert.py:81-102 — only stores the k-means'd background token ids."""

    def __init__(self, config: KernelShapBertConfig):
        super().__init__()
        self.config = config
        self.Xs_train = nn.Parameter(
            torch.zeros((config.kernel_shap_data_size, config.max_position_embeddings), dtype=torch.long),
            requires_grad=False)

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        return self


class KernelShapBertFinal(nn.Module):
    """reference models/kernel_shap_bert.py:105-127"""

    def __init__(self, config: KernelShapBertConfig):
        super().__init__()
        self.config = config
        self.classifier = KernelShapBertClassifier(config)
        self.explainer = KernelShapBertExplainer(config)

    def train(self, mode: bool = True):
        nn.Module.train(self, mode)
        _tree.freeze_model_parameters(self, "classifier")
        return self

    def forward(self, input_ids: Tensor, attention_mask, token_type_ids: Optional[Tensor] = None) -> Tensor:
        return self.classifier(input_ids, attention_mask, token_type_ids)


def sample_coalitions(d: int, n_samples: int, device, seed: int = 0) -> Tuple[Tensor, Tensor]:
    """KernelSHAP coalition set on the device: every subset of size 1 and d-1 with its exact Shapley-kernel
    weight (when the budget allows), then paired random subsets whose sizes follow the kernel's size
    distribution and share the remaining weight equally.  Returns Z (S, d) int64, w (S,) float64, sum(w) = 1.
    (Input generation for the solve; shap's own enumeration/sampling RNG is not reproduced.)"""
    g = torch.Generator(device=device).manual_seed(seed)
    ks = torch.arange(1, d, device=device, dtype=torch.float64)
    size_w = (d - 1.0) / (ks * (d - ks))
    size_w = size_w / size_w.sum()
    rows, weights = [], []
    mass = 0.0
    lo, hi = 1, d - 1
    if d == 2:        # the only proper subsets: {0} and {1}
        z = torch.eye(2, dtype=torch.int64, device=device).repeat((n_samples + 1) // 2, 1)[:max(n_samples, 2)]
        return z, torch.full((z.shape[0],), 1.0 / z.shape[0], dtype=torch.float64, device=device)
    if n_samples >= 2 * d + 2 and d > 3:
        eye = torch.eye(d, dtype=torch.int64, device=device)
        rows += [eye, 1 - eye]
        w1 = float(size_w[0]) / d
        weights += [torch.full((d,), w1, dtype=torch.float64, device=device),
                    torch.full((d,), float(size_w[d - 2]) / d, dtype=torch.float64, device=device)]
        mass = float(size_w[0] + size_w[d - 2])
        lo, hi = 2, d - 2
    n_left = n_samples - sum(r.shape[0] for r in rows)
    n_pairs = (n_left + 1) // 2
    p = size_w[lo - 1:hi]
    sizes = torch.multinomial(p / p.sum(), n_pairs, replacement=True, generator=g) + lo          # (n_pairs,)
    ranks = torch.rand((n_pairs, d), device=device, generator=g).argsort(dim=1).argsort(dim=1)    # random permutation ranks
    z = (ranks < sizes[:, None]).to(torch.int64)
    zz = torch.stack([z, 1 - z], dim=1).reshape(2 * n_pairs, d)[:n_left]
    rows.append(zz)
    weights.append(torch.full((n_left,), (1.0 - mass) / max(n_left, 1), dtype=torch.float64, device=device))
    return torch.cat(rows, 0), torch.cat(weights, 0)


def _min_norm_wls(Z: Tensor, w: Tensor, probs: Tensor, f_x: Tensor, f_null: Tensor) -> Tensor:
    """Under-determined constrained WLS (rank(E) < M - 1): minimum-norm solution of the sqrt-weighted system, what
    numpy.linalg.lstsq gives shap in that regime.  Z (S, M) {0,1}, w (S,), probs (S, C), f_x (C,), f_null (C,) -> (C, M) fp64.
    Rare fallback on the device (torch.linalg.pinv, fp64)."""
    link = lambda q: torch.log(q / (1.0 - q))   # noqa: E731
    Zd = Z.double()
    y = link(probs.double()) - link(f_null.double())[None, :]
    delta = link(f_x.double()) - link(f_null.double())
    E = Zd[:, :-1] - Zd[:, -1:]
    yt = y - Zd[:, -1:] * delta[None, :]
    sw = w.double().sqrt()[:, None]
    sol = torch.linalg.pinv(sw * E) @ (sw * yt)                     # (M-1, C)
    return torch.cat([sol, (delta - sol.sum(0))[None, :]], 0).t().contiguous()


@torch.no_grad()
def kernel_shap_torch(fw_classifier: Callable[[Tensor], Tensor], Xs_train: Tensor, Xs_explain: Tensor, n_samples: int,
                      batch_size: int, silent: bool = True, seed: int = 0) -> Tensor:
    """reference models/kernel_shap_bert.py:130-200 — same signature.
    fw_classifier: (ids (bs, T) int64) -> probabilities (bs, C) without attention masking;
    Xs_train (data_size, T) background token ids; Xs_explain (bs, T).
    Returns (bs, C, T-1) attributions with the CLS feature dropped, following the ModelRecipe contract
    (recipes/types.py:144-148; the reference's own slicing at l.183-185 depends on the shap version).
    Per explained row, as shap.KernelExplainer does: only the M features whose value differs from at least one background
    row take part (the rest get 0), M = 0 -> all zeros, M = 1 -> that feature gets link(f(x)) - link(E f)."""
    _ = silent
    dev = Xs_explain.device
    assert dev.type == "cuda", "KernelSHAP runs on CUDA only (no CPU path)"
    Xs_train = Xs_train.to(dev)
    K, T = Xs_train.shape
    bs = Xs_explain.shape[0]

    def run(ids: Tensor) -> Tensor:
        outs = [fw_classifier(ids[i:i + batch_size]) for i in range(0, ids.shape[0], batch_size)]
        return torch.cat(outs, 0).double()

    link = lambda q: torch.log(q / (1.0 - q))   # noqa: E731
    f_null = run(Xs_train).mean(dim=0)                          # E_bg f
    f_x = run(Xs_explain)                                        # (bs, C)
    C = f_x.shape[1]
    out = torch.zeros((bs, C, T), dtype=torch.float64, device=dev)
    for i in range(bs):
        x = Xs_explain[i]
        idx = (Xs_train != x[None, :]).any(dim=0).nonzero().reshape(-1)      # varying features
        M = int(idx.numel())
        if M == 0:
            continue
        if M == 1:
            out[i, :, idx[0]] = link(f_x[i]) - link(f_null)
            continue
        Zm, w = sample_coalitions(M, n_samples, dev, seed=seed + i)         # (S, M) over the varying features only
        S = Zm.shape[0]
        Z = torch.zeros((S, T), dtype=torch.int64, device=dev)
        Z[:, idx] = Zm
        # h_x(z): present features from x, absent ones from each background row -> (S*K, T) synthetic ids
        synth = torch.where(Z[:, None, :].bool(), x[None, None, :], Xs_train[None, :, :]).reshape(S * K, T)
        probs = run(synth).reshape(S, K, C).mean(dim=1)
        phi, info = ops.kernelshap_solve(ops.pack_feature_masks(Zm)[None], w[None], probs[None], f_x[i:i + 1], f_null, M,
                                         link_logit=True)
        if int(info.abs().max()) != 0:      # fewer independent coalitions than unknowns: minimum-norm solution
            phi = _min_norm_wls(Zm, w, probs, f_x[i], f_null)[None]
        out[i][:, idx] = phi[0]
    return out[:, :, 1:].float()
