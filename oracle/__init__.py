"""CPU oracle: a restatement of the reference's calibration / rounding-finetune arithmetic.

TEST INFRASTRUCTURE ONLY. Nothing under dipoorlet_b200/ may import this package; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
and there only as the checker (or the timed CPU baseline), never as the product.

Every function cites the reference file:line it follows (paths relative to the
ModelTC/Dipoorlet checkout). "Reference expressions evaluated under NumPy 2.x" is the
definition of the oracle (NumPy-version note).

Pinning: oracle/gen_golden.py runs the reference's OWN modules (imported from
/root/reference with stand-ins for onnx / onnxruntime, which are not installable here)
on the same inputs and commits their outputs under tests/golden/; tests/test_oracle_*.py
check this restatement against those fixtures.
"""
