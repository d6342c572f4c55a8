"""CPU oracle for the two-tower hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference repository ships no implementation, tests or golden vectors
for this path, and TensorFlow / TFRS / Keras / FAISS are not
installable here.  Everything in this package restates the published upstream algorithms
(TFRS 0.7.3, Keras 2.15, tf.math.top_k, faiss.IndexFlatIP) recorded in SURVEY.md Appendix A.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package never does.
"""
from .twotower_oracle import *  # noqa: F401,F403
