"""CPU oracle for the PFPN hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, op by op, the reference's TF-1.14 graph for the
particle-mixture policy head (``/root/reference/networks/utils.py:85-236``),
the resampler (``networks/actor_critic/a2c.py:385-474``), the PPO / SAC
head-facing losses and the DPPO learner update, on torch-CPU / numpy.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  Nothing under ``pfpn_b200/`` does:
the product path fails loudly when the CUDA library is missing.

Parity status: the reference ships no tests / golden vectors and TF 1.14 cannot
be installed here.  The restatement is pinned two ways:
  1. ``oracle/tf_shim`` executes the reference's OWN source files
     (``networks/utils.py``, ``networks/actor_critic/a2c.py``) unmodified on an
     eager TF-1 API emulation; the outputs are committed under
     ``tests/golden/`` and the oracle is checked against them;
  2. constants are checked against the lowered graph of the shipped
     checkpoint ``.meta`` (``oracle/metagraph.py``).
Semantics of the un-vendored third-party kernels (TF ``Multinomial``, TFP
``RelaxedOneHotCategorical``) are restated from their published algorithm and
remain "parity unpinned" by the reference itself.
"""
