"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference (cboin1996/avddpg) hot path, used as the parity
checker by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.  Nothing under ``avddpg_b200/`` (the
product) may import, call, link or execute anything in this package: the product path
is CUDA-only and raises when its shared library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * environment / OU noise / reset / replay ring / Polyak / FedAvg: PINNED -- the
    restatements here are checked against golden vectors produced by running the
    reference's own code in the build container (``oracle/make_golden.py`` ->
    ``tests/golden/``).
  * actor/critic forward+backward, TF-Keras Adam: PARITY UNPINNED against real
    TensorFlow 2.4.1 (not installable here, no network).  The restatement follows the
    reference call sites (agent/model.py, workers/trainer.py:472-508) and published
    TF-Keras semantics, and is cross-checked against an independent torch-autograd
    statement of the same networks.
"""
