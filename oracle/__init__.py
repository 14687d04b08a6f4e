"""CPU oracle for the SNNQuantPrune hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement of the reference's quantized + pruned
spiking-layer forward pass (DuQ -> prune -> conv/dense -> BatchNorm -> LIF over
T timesteps, wired as TCJA-SNN ``CextNet``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``parity`` /
``--impl reference`` legs may import it.  The product (``snnquantprune_b200``)
never does.

PARITY PINNED TO THE EXECUTED REFERENCE.  The reference ships no golden vectors
for this path and jax / flax are not installable here,
but it is pure Python: ``tests/golden/make_from_reference.py`` imports the
UNMODIFIED files quant.py, spiking_learning.py, flax_qconv.py, flax_qdense.py
and examples/tcja/models.py through numpy stand-ins for jax / flax.linen /
ml_collections (``tests/golden/ref_shim``), exec's the mask-construction lines
of examples/train_inpt_spikingjelly.py:147-229, and writes
``tests/golden/from_reference_*.npz``.  ``tests/test_from_reference_cpu.py``
holds every function here to those fixtures: DuQ / prune / calibrators / masks /
atan / multi_step_LIF bit for bit, QuantConv / QuantDense to 2e-6, the whole
CextNet (H = 32 and H = 128 / T = 20, 8 / 4 / 2 bit) with 0 flipped spikes,
membranes <= 1.5e-6, identical logits.  What the shim restates (third-party
arithmetic absent from /root/reference): conv / dot with float64 accumulation
rounded once, flax 0.4.0's eval BatchNorm formula, sigmoid, reduce_window.

Two restatements: ``ref_snn`` = the reference's fp32 op order with torch-CPU
contractions (also the bench's CPU baseline); ``ref_int`` / ``ref_net`` /
``csrc/snn_oracle.c`` = exact int32 accumulators + folded fp32 epilogue, the
form the CUDA kernels must match bit for bit.
"""
