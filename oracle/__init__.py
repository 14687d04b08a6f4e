"""CPU oracle for the SNNQuantPrune hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement of the reference's quantized + pruned
spiking-layer forward pass (DuQ -> prune -> conv/dense -> BatchNorm -> LIF over
T timesteps, wired as TCJA-SNN ``CextNet``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product (``snnquantprune_b200``) never
does.

PARITY UNPINNED: the reference ships no golden vectors / tests for this path
 and cannot be imported here (jax / flax are not
installed, no network).  The oracle is therefore pinned only against itself:
two independent restatements (``ref_float`` = reference op order in fp32 using
torch-CPU contractions; ``ref_int`` / ``csrc/snn_oracle.c`` = integer
accumulators + folded epilogue) must agree within the north-star tolerances,
and the committed fixtures under ``tests/golden/`` were produced by
``tests/golden/make_golden.py`` from these restatements.
"""
