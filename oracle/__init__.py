"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy fp64 + torch-CPU autograd) of the pyramid Gatys-loss hot
path of irenemizus/ArtStyleTransfer.  Nothing in the product package
(`artstyletransfer_b200/`) imports this directory; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs do, and there only as the checker or as the timed CPU baseline.

Parity pinning: the reference ships no tests, goldens or known-answer vectors
(SURVEY.md §4, §8c).  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF, imported unmodified from /root/reference in the build
container by `oracle/make_goldens.py`, whose results are committed under
`tests/golden/` and re-checked by `tests/test_oracle_golden.py`.
"""
