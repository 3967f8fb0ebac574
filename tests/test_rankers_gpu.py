"""Drop-in model classes (matchmaker_b200.rankers) against golden vectors produced by the reference's own
classes: load the reference state dict, run forward on the GPU, compare scores and secondary outputs."""
import pytest
import torch

from conftest import assert_close_rel, load_golden
from oracle import interaction_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("tag", ["small", "cfg1"])
def test_knrm_class(tag):
    from matchmaker_b200.rankers import KNRM
    g = load_golden(f"knrm_{tag}")
    m = KNRM.from_config({"knrm_kernels": 11}, g["q"].shape[-1])
    m.load_state_dict({"dense.weight": g["weight"].view(1, -1)})
    m = m.to(DEV)
    args = [g[k].to(DEV) for k in ("q", "d", "q_mask", "d_mask")]
    score = m(*args)
    assert_close_rel(score, g["score"], what="score")
    score2, sec = m(*args, output_secondary_output=True)
    assert set(sec) == {"score", "per_kernel", "query_mean_vector", "cosine_matrix_masked"}
    assert_close_rel(sec["per_kernel"], g["per_kernel"], what="per_kernel")
    assert_close_rel(sec["cosine_matrix_masked"], g["cosine_matrix_masked"], what="cosine")
    assert_close_rel(sec["query_mean_vector"], g["query_mean_vector"], what="qmean")


@pytest.mark.parametrize("tag", ["k11", "k21"])
def test_tk_class_full_model(tag):
    from matchmaker_b200.rankers import ECAI20_TK
    g = load_golden(f"tk_{tag}")
    emb, heads, layers, ff, max_len = [int(x) for x in g["cfg"]]
    cfg = {"tk_kernels_mu": g["mu"].tolist(), "tk_kernels_sigma": g["sigma"].tolist(), "tk_att_heads": heads,
           "tk_att_layer": layers, "tk_att_ff_dim": ff, "max_doc_length": max_len, "tk_use_diff_posencoding": True,
           "tk_mix_hybrid_context": True}
    m = ECAI20_TK.from_config(cfg, emb)
    sd = {k[4:]: v for k, v in g.items() if k.startswith("sd__")}
    assert set(sd) == set(m.state_dict()), "state-dict keys must match the reference"
    # positional buffers computed here must equal the reference's
    assert torch.allclose(m.positional_features_q, sd["positional_features_q"], atol=1e-6)
    assert torch.allclose(m.positional_features_d, sd["positional_features_d"], atol=1e-6)
    m.load_state_dict(sd)
    m = m.to(DEV).eval()
    args = [g[k].to(DEV) for k in ("q", "d", "q_mask", "d_mask")]
    with torch.no_grad():
        score, sec = m(*args, output_secondary_output=True)
    assert_close_rel(score, g["score"], what="score")
    assert_close_rel(sec["per_kernel"], g["per_kernel"], what="per_kernel")
    assert_close_rel(sec["cosine_matrix"], g["cosine_matrix"], rel=2e-3, what="cosine")
    # training step: gradients reach the transformer and the kernel parameters through the CUDA backward
    m.train()
    s = m(*args)
    s.sum().backward()
    assert m.kernel_bin_weights.weight.grad.abs().sum() > 0 and m.kernel_alpha_scaler.grad.abs().sum() > 0
    assert m.mixer.grad is not None and torch.isfinite(m.mixer.grad).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.contextualizer.parameters())


class _TinyEncoder(torch.nn.Module):
    """Stand-in for the HF encoder (no weights on the box): embedding + linear, returns (hidden,)."""

    class _Cfg:
        hidden_size = 48

    def __init__(self, vocab=100):
        super().__init__()
        self.config = self._Cfg()
        self.emb = torch.nn.Embedding(vocab, 48)
        self.lin = torch.nn.Linear(48, 48)

    def forward(self, input_ids=None, attention_mask=None, **kw):
        return (torch.tanh(self.lin(self.emb(input_ids))),)


def _tokens(B, L, g):
    lens = torch.randint(2, L + 1, (B,), generator=g)
    ids = torch.randint(1, 100, (B, L), generator=g)
    mask = (torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)).long()
    return {"input_ids": (ids * mask).to(DEV), "attention_mask": mask.to(DEV)}


def test_colbert_class():
    from matchmaker_b200.rankers.colbert import ColBERT, ColBERTConfig
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    m = ColBERT(ColBERTConfig(bert_model=_TinyEncoder(), compression_dim=64)).to(DEV)
    assert any(k.startswith("bert_model.") for k in m.state_dict()) and "compressor.weight" in m.state_dict()
    q, d = _tokens(6, 12, g), _tokens(6, 40, g)
    score = m(q, d, use_fp16=False)
    qv = m.forward_representation(q)
    dv = m.forward_representation(d)
    ref = O.maxsim_pairs(qv.detach().cpu().clone(), dv.detach().cpu().clone(), q["attention_mask"].cpu(),
                         d["attention_mask"].cpu())
    assert_close_rel(score, ref, what="ColBERT.forward fp32")
    score.sum().backward()
    assert m.compressor.weight.grad.abs().sum() > 0
    # autocast path: vectors in fp16, scores returned in fp16 like the reference
    s16 = m(q, d, use_fp16=True)
    assert s16.dtype == torch.float16
    assert_close_rel(s16.float(), ref, rel=2e-2, what="ColBERT.forward autocast")
    # encode + aggregate path (indexing_heads.py) and the teacher's in-batch path
    qe = m.forward_representation(q, "query_encode")
    de = m.forward_representation(d, "doc_encode")
    agg = m.forward_aggregation(qe, de)
    assert_close_rel(agg, O.maxsim_pairs(qe.detach().cpu().clone(), de.detach().cpu().clone(), None, None), what="aggregate")
    ib = m.forward_inbatch_aggregation(qv, q["attention_mask"], dv, d["attention_mask"])
    assert_close_rel(ib, O.maxsim_allpairs(qv.detach().cpu().clone(), q["attention_mask"].cpu(), dv.detach().cpu().clone(),
                                           d["attention_mask"].cpu()), what="in-batch (reference indexing)")
    m.is_teacher_model = True
    out = m(q, d, use_fp16=False)
    assert isinstance(out, tuple) and len(out) == 3


def test_bert_dot_class():
    from matchmaker_b200.rankers.bert_dot import BERT_Dot, BERT_Dot_Config
    g = load_golden("bert_dot_small")
    from matchmaker_b200 import interaction
    assert_close_rel(interaction.dot_pairs(g["qv"].to(DEV), g["dv"].to(DEV)), g["score"], what="dot golden")
    assert_close_rel(interaction.dot_pairs(g["qv"].half().to(DEV), g["dv"].half().to(DEV)),
                     O.dot_pairs(g["qv"].half().float(), g["dv"].half().float()), what="dot fp16")
    gen = torch.Generator().manual_seed(2)
    m = BERT_Dot(BERT_Dot_Config(bert_model=_TinyEncoder(), compress_dim=32, return_vecs=True)).to(DEV)
    q, d = _tokens(5, 10, gen), _tokens(5, 30, gen)
    m.eval()
    s = m(q, d, use_fp16=False)
    qv, dv = m.forward_representation(q), m.forward_representation(d)
    assert_close_rel(s, O.dot_pairs(qv.detach().cpu(), dv.detach().cpu()), what="BERT_Dot.forward")
    m.train()
    out = m(q, d, use_fp16=False)
    assert isinstance(out, tuple) and out[1].shape == (5, 32)
    out[0].sum().backward()
    assert m.compressor.weight.grad.abs().sum() > 0


def test_model_factory_names():
    from matchmaker_b200 import rankers
    for name in ("knrm", "TK", "TKL", "ColBERT", "bert_dot", "bert_tower"):
        assert rankers.get_model_class(name) is not None
    with pytest.raises(KeyError):
        rankers.get_model_class("drmm")
