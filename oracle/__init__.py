"""CPU oracle for the MVSNet cost-volume hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package; the product
package ``mvsnet_b200`` never does.

PARITY UNPINNED: the reference (ubiquity6/MVSNet) ships no tests, golden
vectors or fixtures, and its arithmetic lives in tensorflow==1.12.0
(requirements.txt:1), which cannot be installed here.  This oracle is a
line-by-line fp32 restatement of the reference source plus the TF-1.12 op
semantics written down in SURVEY.md Appendix A.
"""
from .mvs_oracle import *  # noqa: F401,F403
