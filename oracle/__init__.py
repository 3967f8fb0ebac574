"""CPU oracle for the matchmaker interaction-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``matchmaker_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker or as the timed CPU baseline -- never as the product path.

Parity pinning status ("Oracle"):

* KNRM / TK / TKL / ColBERT / BERT_Dot scoring: PINNED -- the restatements in
  ``interaction_oracle.py`` are checked against golden vectors produced by
  running the reference's own classes (``/root/reference/matchmaker/models``)
  in the build container; script ``make_golden.py``, vectors ``tests/golden``.
* ``CosineMatrixAttention`` (allennlp 2.5.1.dev20210625, third party, absent)
  and ``faiss.IndexFlatIP.search`` (faiss-gpu 1.7.0, third party, absent):
  PARITY UNPINNED -- restated from their published algorithm; anchored on the
  reference's call sites only.
"""
