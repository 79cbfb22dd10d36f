#!/bin/bash
# Check of the NumPy stand-in itself (build container only: needs /root/reference): the
# reference's OWN test file for the simulator runs against the reference's OWN sources on
# tools/jax_numpy_shim.  Tests that need PennyLane, jax.grad / jax.tree_util (the batched
# route), shots (jax.random.choice), diffrax / equinox (pulse mode) or pytest-benchmark
# fail with "placeholder ... was called" / "no attribute" - none on a numerical assertion.
cd /tmp && PYTHONPATH=/root/repo/tools/jax_numpy_shim:/root/reference python -c "
import stub_missing; stub_missing.ROOTS.update({'pennylane'}); stub_missing.install()
import pytest, sys
sys.exit(pytest.main(['/root/reference/tests/${1:-test_jaqsi.py}','-q','-p','no:cacheprovider','--rootdir','/tmp','--no-header','-W','ignore','--tb=line']))
"
