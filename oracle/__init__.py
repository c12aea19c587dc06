"""CPU oracle for the DeMethify NMF-deconvolution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `demethify_b200/` may import this
package: it exists so that `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` have an independent
restatement of the reference algorithm to check (and time) against.

Parity status: PINNED.  `oracle.bssmf_numpy` is checked (tests/test_oracle_golden.py)
against (a) the six golden output directories the reference ships under
`test/` and (b) outputs of the live reference imported from `/root/reference`
in the build container, frozen as `tests/golden/*.npz` by
`tests/golden/make_golden.py`.
"""
