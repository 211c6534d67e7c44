"""CPU oracle for the cost-volume hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under
``stereo_toolbox_b200/`` imports it: the product path fails loudly if its CUDA
library is missing instead of routing through here.

Pinning: the reference (xxxupeng/stereo_toolbox @ 6af8685) ships no golden
vectors or known-answer tests for this path, so the oracle is
pinned against outputs of the reference itself, generated in the build
container by ``tests/golden/make_golden.py`` (which imports the Python reference
from /root/reference) and committed as small fixtures under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
