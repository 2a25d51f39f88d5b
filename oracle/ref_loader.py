"""TEST INFRASTRUCTURE ONLY -- loads the *actual* reference from /root/reference.

This module exists only in the authoring container (the GPU box has no
/root/reference).  It is used by `oracle/make_golden.py` to generate the
fixtures under `tests/golden/` and by the `not gpu` tests that pin
`oracle/wsdl_oracle.py` against the real reference functions.  Nothing in the
product package may import it.

The reference does not import as shipped: module-name
typos, missing imports, a notebook that trains at import time.  The shims here
change *no arithmetic*:

* `ExtraUtils` -> `ExtraUtilities` alias            (TraditionalModel/LayerCAM.py:5)
* `resnet50(pretrained=True)` -> `weights=None`     (TraditionalModel/ClassificationModel.py:12; no network)
* `LocalNormalizedCutLoss`, `compute_affinities`, `refine_pseudo_mask` are
  pulled out of the notebook by AST node            (TraditionalModel/AlternatingDirectionCutLoss.py:65-105,612-637,709-767)
* `ConstrainToBoundaryLossSingle.compute_affinities_single` gets the missing
  `@staticmethod`                                    (TraditionalModel/AlternatingDirectionBoundaryLoss.py:29,46)
"""
import ast
import os
import sys

REF_ROOT = "/root/reference"
REF_TM = os.path.join(REF_ROOT, "TraditionalModel")


def available() -> bool:
    return os.path.isdir(REF_TM)


def _exec_nodes(path, wanted, namespace):
    with open(path) as fh:
        tree = ast.parse(fh.read())
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in wanted:
            exec(compile(ast.Module([node], []), path, "exec"), namespace)
    missing = [w for w in wanted if w not in namespace]
    if missing:
        raise RuntimeError(f"reference symbols not found in {path}: {missing}")
    return namespace


_cache = {}


def load():
    """Returns a dict of the reference's hot-path callables."""
    if _cache:
        return _cache
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    import torchvision

    if REF_TM not in sys.path:
        sys.path.insert(0, REF_TM)
    import ExtraUtilities  # noqa: E402  (reference module)

    sys.modules.setdefault("ExtraUtils", ExtraUtilities)
    import ClassificationModel  # noqa: E402
    import LayerCAM  # noqa: E402

    ClassificationModel.resnet50 = lambda pretrained=False, **kw: torchvision.models.resnet50(weights=None, **kw)

    base = {"torch": torch, "nn": nn, "F": F}
    cut = _exec_nodes(
        os.path.join(REF_TM, "AlternatingDirectionCutLoss.py"),
        {"LocalNormalizedCutLoss", "compute_affinities", "refine_pseudo_mask", "LayerCAMGenerator"},
        dict(base, gc=__import__("gc")),
    )
    bnd = _exec_nodes(
        os.path.join(REF_TM, "AlternatingDirectionBoundaryLoss.py"),
        {"ConstrainToBoundaryLossSingle"},
        dict(base),
    )
    Boundary = bnd["ConstrainToBoundaryLossSingle"]
    if not isinstance(Boundary.__dict__["compute_affinities_single"], staticmethod):
        Boundary.compute_affinities_single = staticmethod(Boundary.__dict__["compute_affinities_single"])

    _cache.update(
        LayerCAMGenerator=LayerCAM.LayerCAMGenerator,
        LayerCAMGeneratorVariant=cut["LayerCAMGenerator"],
        FrozenResNetCAM=ClassificationModel.FrozenResNetCAM,
        compute_iou_and_acc=ExtraUtilities.compute_iou_and_acc,
        LocalNormalizedCutLoss=cut["LocalNormalizedCutLoss"],
        compute_affinities=cut["compute_affinities"],
        refine_pseudo_mask=cut["refine_pseudo_mask"],
        ConstrainToBoundaryLossSingle=Boundary,
    )
    return _cache
