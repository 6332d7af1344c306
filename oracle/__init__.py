"""oracle/ — TEST INFRASTRUCTURE, not product code.

A CPU restatement (pure torch, fp32 or fp64) of the reference's hot path:
superpoint-graph self-attention + segment pooling + the norms / transforms either
side of it (drprojects/superpoint_transformer @ eb959b6, mounted read-only at
/root/reference in the build container).  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` leg may import it; nothing under
`superpoint_transformer_b200/` does.

Parity status
-------------
* GLUE pinned: `oracle/path.py` (the restatement of src/nn/{attention,pool,unpool,
  norm,transformer,stage}.py, src/models/components/spt.py:760-944 and
  src/transforms/graph.py:1137-1277,1419-1452) is checked against the reference's
  OWN source files executed in the build container (`oracle/reference_shim.py`
  loads them verbatim with importlib) — `oracle/make_golden.py` writes the
  resulting input/output vectors to `tests/golden/*.pt`, and
  `tests/test_oracle_golden.py` replays them anywhere.
* SELECTION pinned the same way: `oracle/select.py` (NAG.select / Data.select /
  Cluster.select / CSRData.__getitem__, src/data/{nag,data,cluster,csr}.py) against
  `tests/golden/select.pt`, which `oracle/make_golden_select.py` produces by running the
  reference's own src/data/*.py and src/utils/{tensor,sparse}.py, loaded verbatim by
  `oracle/reference_data.py` (restated there: the PyG `Data` attribute store and
  `consecutive_cluster`, third-party and absent like the leaves below).
* SAMPLING: `oracle/sampling.py` (sparse_sample, SampleSegments weights) against
  `tests/golden/sampling.pt` from the same loader: counts per segment and the seeded
  SampleSegments output bit-exact; the random draw itself is checked against the sampling law.
* LEAVES unpinned: the arithmetic leaves the reference calls live in third-party
  wheels that are absent from /root/reference and from this image
  (`torch_scatter` unpinned for torch 2.2.0, `torch_geometric==2.3.0`;
  /root/reference/install.sh:99-100).  `oracle/leaves.py` restates their
  published semantics; the reference's tests hold no
  golden vector at that boundary => **parity unpinned at the
  torch_scatter / PyG leaf boundary**.
"""
