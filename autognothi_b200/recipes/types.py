"""The plugin boundary of the reference, restated: a `ModelRecipe` is a bundle of callables through
which every stage script touches a model (reference recipes/types.py:96-162).  Field names, argument
order and tensor layouts are kept so that a recipe from this package can be registered in the
reference's `get_recipe` table (scripts/resources.py:55-83) unchanged."""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Optional, Tuple, Type, Union

import torch
from torch import Tensor, nn


@dataclasses.dataclass
class ModelRecipe_Training:
    """reference recipes/types.py:42-48"""
    support_classifier: bool
    support_surrogate: bool
    support_explainer: bool
    exp_variant_duo: bool
    exp_variant_kernel_shap: bool


@dataclasses.dataclass
class ModelRecipe_Measurements:
    """reference recipes/types.py:77-93"""
    verify_final_coherency: bool
    allow_accuracy: bool
    allow_faithfulness: bool
    allow_cls_acc: bool
    allow_performance_cls: bool
    allow_performance_srg_exp: bool
    allow_performance_fin: bool
    allow_train_resources: bool
    allow_dual_task_similarity: Union[bool, Any]
    allow_branches_cka: bool


@dataclasses.dataclass
class ModelRecipe:
    """reference recipes/types.py:96-162.

    fw_classifier / fw_surrogate :: (model, Xs (N,...), mask (N, n_players) int64) -> (Ys (N,C), aux)
    fw_explainer :: (model, Xs, mask, surrogate_grand (B,C), surrogate_null (1,C)) -> (phi (B,C,n), aux)
    fw_final     :: (model, Xs) -> (Ys (B,C), phi (B,C,n))
    Extension (additive): masks may be `PackedMasks`; fw_surrogate also accepts Xs (B,...) with a
    (B,S,n) mask or a PackedMasks of B*S rows, evaluating S coalitions per input without replicating Xs.
    """
    id: str
    version: str
    t_config: Type[Any]
    t_classifier: Type[nn.Module]
    t_surrogate: Type[nn.Module]
    t_explainer: Type[nn.Module]
    t_final: Type[nn.Module]
    load_misc: Callable[[Any, Any], Any]
    conv_pretrained_classifier: Callable[[Any, Any], nn.Module]
    conv_classifier_surrogate: Callable[[Any, Any, nn.Module], nn.Module]
    conv_surrogate_explainer: Callable[[Any, Any, nn.Module], nn.Module]
    conv_explainer_final: Callable[[Any, Any, nn.Module, nn.Module, nn.Module], nn.Module]
    n_players: Callable[[Any], int]
    gen_input: Callable[[Any, Any, torch.device], Callable[[Any, Any], Tuple[Tensor, Tensor]]]
    gen_null: Callable[[Any, Any, torch.device], Tensor]
    training: ModelRecipe_Training
    fw_classifier: Callable[[nn.Module, Tensor, Any], Tuple[Tensor, Tensor]]
    fw_surrogate: Callable[[nn.Module, Tensor, Any], Tuple[Tensor, Optional[Tensor]]]
    fw_explainer: Callable[[nn.Module, Tensor, Any, Tensor, Tensor], Tuple[Tensor, Optional[Tensor]]]
    fw_final: Callable[[nn.Module, Tensor], Tuple[Tensor, Tensor]]
    measurements: ModelRecipe_Measurements
