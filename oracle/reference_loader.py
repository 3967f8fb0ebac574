"""Import the *actual* reference model classes from ``/root/reference`` on CPU.

TEST INFRASTRUCTURE, build-container only: ``/root/reference`` does not exist
on the GPU box, so nothing that runs there may call this module.  It is used
by ``oracle/make_golden.py`` (to produce ``tests/golden/*.npz``) and by the
``not gpu`` tests that re-check the golden vectors when the reference is
present.

Shims; none of them touches the arithmetic of the
reference files themselves:

1. ``allennlp`` is absent -> stub modules providing ``TextFieldEmbedder`` (unused
   base class) and ``CosineMatrixAttention`` (third-party; restated in
   ``interaction_oracle.cosine_matrix`` -- PARITY UNPINNED for that one op).
2. The constructors hard-require CUDA (``torch.cuda.FloatTensor``,
   ``torch.cuda.LongTensor``) -> mapped to their CPU twins while constructing.
3. ``colbert.py`` / ``bert_dot.py`` cannot be imported under transformers 5.x
   (their config classes are rejected at class-creation time) and need HF
   weights -> imported against a stub ``transformers`` module; instances are
   created without ``__init__`` and ``forward_representation`` is replaced by
   "return the vectors I was given", so ``forward`` runs its own scoring lines
   (colbert.py:68-75, bert_dot.py:62) unmodified.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MATCHMAKER_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "matchmaker", "models", "knrm.py"))


class _CosineMatrixAttention(torch.nn.Module):
    """Stand-in for allennlp's CosineMatrixAttention (third party)."""

    def forward(self, a, b):
        from .interaction_oracle import cosine_matrix
        return cosine_matrix(a, b)


def _stub_modules():
    mods = {}

    def mk(name):
        m = types.ModuleType(name)
        mods[name] = m
        return m

    mk("allennlp")
    mk("allennlp.modules")
    tfe = mk("allennlp.modules.text_field_embedders")
    tfe.TextFieldEmbedder = type("TextFieldEmbedder", (torch.nn.Module,), {})
    mk("allennlp.modules.matrix_attention")
    cma = mk("allennlp.modules.matrix_attention.cosine_matrix_attention")
    cma.CosineMatrixAttention = _CosineMatrixAttention
    dpa = mk("allennlp.modules.matrix_attention.dot_product_matrix_attention")
    dpa.__all__ = []

    tr = mk("transformers")

    class PretrainedConfig:  # noqa: D401 - stub
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

    class PreTrainedModel(torch.nn.Module):
        def __init__(self, cfg=None):
            super().__init__()
            self.config = cfg

    class AutoModel:
        @staticmethod
        def from_pretrained(name):
            raise RuntimeError("no HF weights in this container")

    tr.PretrainedConfig = PretrainedConfig
    tr.PreTrainedModel = PreTrainedModel
    tr.AutoModel = AutoModel
    return mods


@contextlib.contextmanager
def _shimmed():
    stubs = _stub_modules()
    saved = {k: sys.modules.get(k) for k in stubs}
    had_ft = getattr(torch.cuda, "FloatTensor", None)
    had_lt = getattr(torch.cuda, "LongTensor", None)
    sys.modules.update(stubs)
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.LongTensor = lambda n, device=None: torch.LongTensor(n)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        torch.cuda.FloatTensor = had_ft
        torch.cuda.LongTensor = had_lt


def _load(relpath: str, modname: str):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    with _shimmed():
        spec.loader.exec_module(mod)
    return mod


def load_knrm(n_kernels: int = 11):
    mod = _load("matchmaker/models/knrm.py", "_ref_knrm")
    with _shimmed():
        return mod.KNRM(n_kernels)


def load_tk(emb: int, mu, sigma, heads: int = 10, layers: int = 2, ff: int = 300, max_len: int = 200,
            diff_pos: bool = True, mix: bool = True):
    mod = _load("matchmaker/models/published/ecai20_tk.py", "_ref_tk")
    with _shimmed():
        return mod.ECAI20_TK(emb, list(mu), list(sigma), heads, layers, ff, max_len, diff_pos, mix)


def load_tkl(emb: int, mu, sigma, heads: int = 10, layers: int = 2, ff: int = 300, max_len: int = 2000,
             use_pos: bool = True, diff_pos: bool = True, saturation: str = "embedding"):
    mod = _load("matchmaker/models/published/sigir20_tkl.py", "_ref_tkl")
    with _shimmed():
        return mod.TKL_sigir20(emb, list(mu), list(sigma), heads, layers, ff, max_len, use_pos, diff_pos, saturation)


def load_tk_sparse(emb: int, mu, sigma, heads: int = 4, layers: int = 1, proj: int = 16, ff: int = 32, max_len: int = 64,
                   diff_pos: bool = True):
    mod = _load("matchmaker/models/published/cikm20_tk_sparse.py", "_ref_tk_sparse")
    with _shimmed():
        return mod.CIKM20_TK_Sparse(emb, list(mu), list(sigma), heads, layers, proj, ff, max_len, diff_pos)


def load_conv_knrm(emb: int, n_grams: int = 3, n_kernels: int = 11, conv_out: int = 32):
    mod = _load("matchmaker/models/conv_knrm.py", "_ref_conv_knrm")
    with _shimmed():
        return mod.Conv_KNRM(emb, n_grams, n_kernels, conv_out)


class _Passthrough:
    """forward_representation replacement: tokens dict carries the vectors."""

    @staticmethod
    def rep(tokens, sequence_type=None):
        return tokens["vecs"]


def load_colbert():
    """Returns (ColBERT class, instance with __init__ bypassed).  ``forward``
    then executes colbert.py:68-75 on ``query["vecs"]`` / ``document["vecs"]``."""
    mod = _load("matchmaker/models/colbert.py", "_ref_colbert")
    inst = mod.ColBERT.__new__(mod.ColBERT)
    torch.nn.Module.__init__(inst)
    inst.return_vecs = False
    inst.forward_representation = _Passthrough.rep
    return mod.ColBERT, inst


def load_bert_dot():
    mod = _load("matchmaker/models/bert_dot.py", "_ref_bert_dot")
    inst = mod.BERT_Dot.__new__(mod.BERT_Dot)
    torch.nn.Module.__init__(inst)
    inst.return_vecs = False
    inst.use_compressor = False
    inst.forward_representation = _Passthrough.rep
    return mod.BERT_Dot, inst
