"""CPU oracle for the consistent_depth fine-tuning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`consistent_depth_b200/`) imports this directory; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs do, and only as the checker / the timed CPU arm.

The reference is pure Python on PyTorch, so the oracle is a restatement of the
reference algorithm as plain PyTorch-CPU / numpy functions (fp32 or fp64),
each citing the reference file:line it follows.  It is PINNED against the real
reference: `oracle/make_golden.py` imports the unmodified modules from
/root/reference (in the build container, where that tree exists), runs them on
seeded inputs and writes `tests/golden/*.npz`; `tests/test_oracle_golden.py`
checks every oracle function against those vectors.  The reference itself
ships no tests / golden vectors for this path, so these
reference-generated fixtures are the pin.
"""
